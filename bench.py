#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native EDXRaster raster hot path.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # CPU reference arm: the reference's own sources (oracle/_ref) on host cores

Workload (BASELINE.json configs[1], "C2"): 1,000,000 random <=4-px triangles at 1920x1080, depth test
only, fixed submission order. One step = one frame of the hot path (Renderer::RenderMesh,
/root/reference/EDXRaster/Core/Renderer.cpp:100-118) per GPU. Metric: Mtris/s (submitted triangles per
second, whole job); Gpix/s and frames/s ride along as extra keys.

Timing: CUDA events around exactly K steps, barrier + synchronize on both sides, max over ranks. Each GPU
keeps --in-flight (3) independent frames in flight: lanes = edx contexts on their own streams sharing the
meshes; the events sit on a control stream that every lane waits on at the start and that waits on every
lane at the end. The same loop with one frame in flight is reported next to it (`one_frame_in_flight`). Inputs are made larger than L2 by rotating over 4 device copies of the mesh
(4 x 108 MB of SoA streams > 126 MB L2), so every frame reads its geometry from HBM. At N > 1 the
frames are independent (weak scaling: one frame per rank per step) and every finished depth buffer is pushed
into rank 0's memory over NVLink as soon as it is rendered (copy engine, one notification per 4 frames).

Every default run also measures BASELINE.json configs[4] (C5: 256 views of the 10M-triangle mesh, view i on GPU
i mod N, every colour buffer gathered to rank 0 inside the timed region) -> `other_configs.C5`; at N = 1 the other
configs ride along too, each with a same-run CPU baseline.

Rank 0 prints ONE JSON line on stdout; everything else goes to stderr.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

WORKLOADS = {
    # name: (description, algorithmic bytes per frame as a function of (nv, nt, w, h), shaded)
    "C1": "C1: UV sphere 100x100 (20,000 tris), 1280x720, Blinn-Phong + depth",
    "C2": "C2: 1,000,000 random <=4px triangles, 1920x1080, depth test only, fixed order",
    "C3": "C3: 2,000 screen-covering triangles, 3840x2160, Blinn-Phong, perspective-correct interpolation",
    "C4": "C4: 10,000,000-triangle displaced grid, 1920x1080, near/side-plane clipping, Blinn-Phong",
    "C5": "C5: 256 camera views of the C4 mesh (10,000,000 triangles), 1920x1080, view i on GPU i mod N, frames gathered to rank 0",
    "M1": "M1 (stress, not a BASELINE config): 1,000,000 mid-size triangles (vertices over 32..128 px boxes), 1920x1080, depth only",
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def algorithmic_bytes(nv, nt, w, h, shaded):
    """vertex attributes once, indices once, frame buffer written once."""
    return (32 * nv + 12 * nt + 8 * w * h) if shaded else (12 * nv + 12 * nt + 4 * w * h)


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_scene(name, scale):
    from edxraster_b200 import scenes
    if name == "C5":
        sc = scenes.by_name("C4", scale)
        sc["views"] = scenes.config5_views(sc, 256)
        return sc
    return scenes.by_name(name, scale)


# ------------------------------------------------------------------------------------------------
# clocks: NVML sampler thread (the recipe's nvidia-smi line is too coarse for millisecond regions)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples = []          # (t, sm_mhz, reasons_bitmask)
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # honour CUDA_VISIBLE_DEVICES remapping
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except Exception:
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:       # pragma: no cover
            log("clock sampler unavailable:", e)
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, rs))
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()

    def summary(self, t0, t1):
        if not self.nv or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "window": "unavailable"}
        nv = self.nv
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed"
        if len(inside) < 3:
            # a millisecond-scale timed region holds too few ~1 ms NVML samples: use every sample taken under
            # load (the back-to-back warm-up frames and the timed region), dropping the idle start
            busy = [s for s in self.samples if s[0] <= t1]
            inside, window = busy[len(busy) // 4:], "under load: warm-up frames + timed region"
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            getattr(nv, "nvmlClocksEventReasonApplicationsClocksSetting", 0x2): "applications_clocks_setting",
        }
        bits = 0
        for s in inside:
            bits |= s[2]
        reasons = sorted(v for k, v in names.items() if k and (bits & k))
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(inside), "window": window}


# ------------------------------------------------------------------------------------------------
# the `config` block: built the same way by both arms, so the two lines describe the same workload key for key
# ------------------------------------------------------------------------------------------------
def mesh_copies(nv, nt):
    """GPU arm: inputs are kept larger than L2 by rotating over this many device copies of the mesh."""
    return 4 if nv * 32 + nt * 12 < 200e6 else 1          # C4/C5: one 440 MB mesh already exceeds L2


def config_block(name, nv, nt, w, h):
    copies = mesh_copies(nv, nt)
    return {"workload": WORKLOADS[name], "triangles": nt, "vertices": nv, "resolution": [w, h], "frames_per_step_per_gpu": 1,
            "l2": "GPU arm: inputs larger than L2 - round-robin over %d device cop%s of the mesh (%d MB of SoA streams), nothing is "
                  "flushed; CPU arm: host memory" % (copies, "ies" if copies > 1 else "y", copies * (nv * 32 + nt * 12) // 1000000)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own sources compiled against the EDXUtil stand-in (oracle/_ref, kind "reference");
# the restatement (oracle/edx_oracle.cpp, kind "port") only where that library is absent
# ------------------------------------------------------------------------------------------------
def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


class CpuRenderer:
    """One scene on the host cores. Timing builds: rsqrtps + one Newton step, as EDXUtil's Embree-style SSE::Rsqrt."""

    def __init__(self, scene, threads):
        from oracle import ref
        self.scene = scene
        self.kind = "reference" if ref.available() else "port"
        if self.kind == "reference":
            self.r = ref.Reference(scene.width, scene.height, threads, timing=True)
            self.r.set_transform(scene.mv, scene.proj, scene.raster)
            # the reference has Blinn-Phong and the textured Lambert shader only (Shader.h); a depth-only workload
            # runs its default shader - more CPU work per fragment than the GPU arm does, noted in `note`
            self.r.set_shader(scene.shader if scene.shader in (1, 3) else 3)
            self.threads = max(self.r.threads, min(threads, host_threads()))      # OpenMP threads parallel_for really uses
            self.note = ("oracle/_ref: the reference's own Core/*.cpp, Core/*.h, Utils/* compiled unmodified with g++ -O2 -msse4.1 "
                         "-fopenmp against the EDXUtil stand-in (oracle/_ref_shim; SSE::Rsqrt = rsqrtps + Newton step); parallel_for on "
                         "%d OpenMP threads, per-core lists capped at 12 (Tile.h:34)" % host_threads())
            if scene.shader not in (1, 3):
                self.note += "; the reference has no depth-only mode: its default LambertianAlbedo shader runs on every fragment"
        else:
            from oracle import orc
            orc.build()
            self.r = orc.Oracle(scene.width, scene.height, threads, timing=True)
            self.r.set_transform(scene.mv, scene.proj, scene.raster)
            self.r.set_shader(scene.shader)
            self.threads = self.r.threads
            self.note = "oracle/ timing build (CPU restatement of the reference SSE path); oracle/_ref is not present on this box"
        self._tris = None

    def set_view(self, view):
        self.r.set_transform(*view)

    def frame(self, tri_limit=None):
        sc = self.scene
        n = sc.num_tris if tri_limit is None else min(tri_limit, sc.num_tris)
        t = time.perf_counter()
        if self.kind == "reference":
            if self._tris != n:
                self.r.set_mesh(sc.vertices, sc.indices[:n])      # Mesh::LoadMesh copies, outside the timed call
                self._tris = n
                t = time.perf_counter()
            self.r.render()
        else:
            self.r.render(sc.vertices, sc.indices[:n])
        return time.perf_counter() - t, n

    def close(self):
        self.r.close()


def cpu_baseline(scene, cpu_seconds=20.0, max_frames=40, views=None):
    """Bounded sample of `scene` on the host cores: about `cpu_seconds` of CPU work (wall x threads)."""
    c = CpuRenderer(scene, host_threads())
    try:
        c.frame()                                  # cold: page faults, mesh copy
        probe, used = c.frame()
        work = probe * c.threads
        nfr = int(min(max_frames, max(1 if work > cpu_seconds else 3, round(cpu_seconds / max(work, 1e-6)))))
        times = []
        for k in range(nfr):
            if views is not None:
                c.set_view(views[(k * 37) % len(views)])
            times.append(c.frame()[0])
        while sum(times) < 0.5 and len(times) < 2000:       # tiny frames: at least half a second of wall time
            times.append(c.frame()[0])
        v = used * len(times) / sum(times) / 1e6
        what = "full frames" if views is None else "views (every 37th of the 256, extrapolated to all)"
        return {"value": v, "unit": "Mtris/s", "cores": c.threads, "kind": c.kind,
                "sample": "%d %s of the same workload after 1 warm-up (%.3f s each, ~%.0f s of CPU work on %d threads)"
                          % (len(times), what, sum(times) / len(times), sum(times) * c.threads, c.threads),
                "frames_per_s": len(times) / sum(times), "note": c.note}
    finally:
        c.close()


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scene = make_scene(args.workload, args.scale)
    nt = scene.num_tris
    views = scene.get("views")
    c = CpuRenderer(scene, host_threads())     # explicit thread count: torchrun exports OMP_NUM_THREADS=1
    # calibrate: one full frame, then bound the per-step sample so K+W steps end within ~2 minutes
    probe, _ = c.frame()
    budget = 120.0 / max(1, args.steps + args.warmup)
    limit = None
    if probe > budget:
        limit = max(1000, int(nt * budget / probe))
    for _ in range(min(args.warmup, 3)):
        c.frame(limit)
    times, used = [], nt
    for k in range(args.steps):
        if views is not None:
            c.set_view(views[k % len(views)])
        dt, used = c.frame(limit)
        times.append(dt)
    total = sum(times)
    value = used * len(times) / total / 1e6
    sample = ("full frames (%d triangles each)" % used) if limit is None else \
        ("first %d of %d triangles per step (bounded sample, same resolution and state)" % (used, nt))
    line = {
        "impl": "reference", "metric": "Mtris/s", "value": value, "unit": "Mtris/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 depth / i32 28.4 fixed-point coverage",
        "data": "synthetic",
        "config": config_block(args.workload, scene.num_verts, nt, scene.width, scene.height),
        "frames_per_s": len(times) / total, "gpix_per_s": scene.width * scene.height * len(times) / total / 1e9,
        "cpu_baseline": {"value": value, "unit": "Mtris/s", "cores": c.threads, "kind": c.kind, "sample": sample, "note": c.note},
        "e2e": {"value": value, "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    c.close()
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
OPTIONS = []
NVLINK_INGRESS_GBS = 900.0      # NVLink 5 per direction, nominal; measured into one B200 from 7 senders (copy engine): 733 in round 1, 781 in round 2


def time_frames(r, meshes, frames, warm):
    """Device time of `frames` back-to-back frames (CUDA events on the context's stream)."""
    for i in range(warm):
        r.RenderMesh(meshes[i % len(meshes)])
        r.Synchronize()               # (a synchronised frame sizes the internal queues: grow + re-run if it overflowed)
    r.TimerBegin()
    for i in range(frames):
        r.RenderMesh(meshes[i % len(meshes)])
    return r.TimerEnd()


class Farm:
    """K independent frames in flight per GPU (lanes = edx contexts on their own streams sharing the meshes) and,
    at N > 1, the gather of every finished frame into rank 0's memory.

    Exchange (--gather ce, default): rank 0's receive buffer is symmetric memory mapped into every rank over
    NVLink, and each lane's context has its slot of it as frame sink (edx_set_frame_sink): right behind a frame's
    last kernel, on the lane's own stream, the COPY ENGINE pushes the finished buffer to rank 0 (no SM time, no host
    work; nothing waits for a batch to fill, so the only transfer that cannot overlap rendering is the last
    frame's). Behind the push the context stores the number of frames it has pushed into a count word in rank 0's
    memory (edx_set_frame_sink_signal) - what a consumer there polls. No collective and no batches: each lane cycles
    two slots of the store, and its own stream orders a slot's reuse behind the previous push. (Until round 2's last
    change a 4-byte NCCL all-reduce per 4 frames did the notification; its host cost - launch, event, three
    stream waits - made the farm host-bound at 50 us per frame.)
    --gather nccl: dist.gather per batch of G frames. --gather stores: the resolve kernel writes its pixels straight
    into rank 0's memory (measured slower: 32-byte row segments). --gather none: diagnosis only."""

    def __init__(self, torch, dist, R, sc, local, rank, world, in_flight, mode, G, meshes=None):
        self.torch, self.dist, self.R = torch, dist, R
        self.sc, self.rank, self.world, self.G = sc, rank, world, G
        self.dev = dev = torch.device("cuda", local)
        W, H = sc.width, sc.height
        self.shaded = sc.shader != 0
        self.stream = torch.cuda.Stream(device=dev)       # timing events + the exchange
        self.K = K = max(1, in_flight)
        self.lanes = []
        for _ in range(K):
            ls = torch.cuda.Stream(device=dev)
            lr = R.Renderer(local)
            lr.SetStream(ls.cuda_stream)
            lr.Initialize(W, H)
            lr.SetTransform(sc.mv, sc.proj, sc.raster)
            lr.SetPixelShader(sc.shader)
            for kv in OPTIONS:                    # --opt name=value: tuning knobs for experiments (none changes a pixel)
                k, v = kv.split("=")
                lr.SetOption(k, int(v))
            self.lanes.append((ls, lr))
        self.r = self.lanes[0][1]
        views = sc.get("views")
        self.xf = R.PackedTransform(sc.mv, sc.proj, sc.raster)   # marshalled once; edx_set_transform still runs every step
        self.views = None if views is None else [R.PackedTransform(*v) for v in views]
        self.copies = mesh_copies(sc.num_verts, sc.num_tris)
        self.own_meshes = meshes is None
        self.meshes = meshes or [self.r.CreateMesh(sc.vertices, sc.indices) for _ in range(self.copies)]   # read-only while rendering: shared by the lanes
        # render targets are torch tensors, double-buffered batches of G frames
        self.tgt_color = [torch.zeros((G, H, W, 4), dtype=torch.uint8, device=dev) for _ in range(2)]
        self.tgt_depth = [torch.zeros((G, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
        self.result = self.tgt_color if self.shaded else self.tgt_depth
        self.frame_bytes = W * H * 4
        self.peer, self.recv, self.mode = None, None, mode
        self.S = 2                        # --gather ce: slots per lane in rank 0's frame store (a lane's pushes are stream-ordered)
        self.pushed = [0] * K
        if world > 1 and mode in ("stores", "ce"):
            try:
                import torch.distributed._symmetric_memory as symm
                shape = (world, K, self.S) if mode == "ce" else (world, 2, G)
                sbuf = symm.empty(shape + tuple(self.result[0].shape[1:]), dtype=self.result[0].dtype, device=dev)
                hdl = symm.rendezvous(sbuf, dist.group.WORLD)
                self.sbuf = sbuf
                self.peer = hdl.get_buffer(0, sbuf.shape, sbuf.dtype)      # rank 0's buffer, addressable from this GPU
                self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
                if mode == "ce":
                    # rank 0's count words, one per (rank, lane): edx_set_frame_sink_signal stores the number of frames
                    # pushed so far behind every frame's pushes - what a consumer on rank 0 polls; no collective
                    self.counts = symm.empty((world, K), dtype=torch.int32, device=dev)
                    self.counts.zero_()
                    torch.cuda.synchronize()
                    hdl2 = symm.rendezvous(self.counts, dist.group.WORLD)
                    peer_counts = hdl2.get_buffer(0, self.counts.shape, self.counts.dtype)
                    dist.barrier()
                    for l, (_, lr) in enumerate(self.lanes):
                        lr.SetFrameSinkSignal(peer_counts[rank, l].data_ptr())
            except Exception as e:       # pragma: no cover
                log("symmetric memory unavailable (%s): using NCCL gather" % e)
                self.peer = None
        if self.peer is None and mode != "none":
            self.mode = "nccl"
        if world > 1 and rank == 0 and self.peer is None:
            self.recv = [[torch.empty_like(self.result[0]) for _ in range(world)] for _ in range(2)]
        # device addresses, looked up once: tensor indexing costs more host time than a frame takes
        self.cptr = [[self.tgt_color[b][k].data_ptr() for k in range(G)] for b in range(2)]
        self.dptr = [[self.tgt_depth[b][k].data_ptr() for k in range(G)] for b in range(2)]
        self.pptr = None if self.peer is None else [[self.peer[rank, b, k].data_ptr() for k in range(self.peer.shape[2])] for b in range(self.peer.shape[1])]

    # -- exchange ---------------------------------------------------------------------------------
    def notify(self, b, n, works):
        """End of batch buffer b (n frames): works[b] = an event that fires when rank 0 holds the batch. Lanes wait
        on that EVENT before they overwrite buffer b - not on the control stream, which by then also waits for
        later frames (that would drain the frames in flight at every batch boundary)."""
        torch, dist = self.torch, self.dist
        if self.mode == "none" or self.world == 1:
            return
        if self.peer is not None:
            w = dist.all_reduce(self.flag, async_op=True)       # stream-ordered after this rank's copies / frames of the batch
        elif n == self.G:
            w = dist.gather(self.result[b], self.recv[b] if self.rank == 0 else None, dst=0, async_op=True)
        else:                                                   # last, partial batch of the timed region
            part = self.result[b][:n].contiguous()
            rbuf = [torch.empty_like(part) for _ in range(self.world)] if self.rank == 0 else None
            w = dist.gather(part, rbuf, dst=0, async_op=True)
        w.wait()                                                # stream-level: the control stream continues after the exchange
        ev = torch.cuda.Event()
        ev.record(self.stream)
        works[b] = ev

    def step(self, i, works, nl):
        G, world, rank = self.G, self.world, self.rank
        if world > 1 and self.mode == "ce":
            # the copy engine pushes the finished frame into this lane's next slot of rank 0's store and the count word
            # follows it; the lane's stream orders a slot's reuse behind its previous push: no batches, no collective
            l = i % nl
            ls, lr = self.lanes[l]
            dst = self.pptr[l][(i // nl) % self.S]
            if rank == 0 and not os.environ.get("EDX_BENCH_RANK0_SINK"):      # the store is rank 0's own memory: its frames are rendered in place
                lr.SetRenderTarget(dst if self.shaded else 0, 0 if self.shaded else dst)
            else:
                lr.SetFrameSink(dst if self.shaded else 0, 0 if self.shaded else dst)
            if self.views is not None:
                lr.SetTransform(self.views[(i * world + rank) % len(self.views)])    # C5: view v is rendered by rank v mod N
            else:
                lr.SetTransform(self.xf)
            lr.RenderMesh(self.meshes[i % len(self.meshes)])
            self.pushed[l] += 1
            return
        b, k = (i // G) & 1, i % G
        ls, lr = self.lanes[i % nl]
        if k == 0 and works[b] is not None:
            for s2, _ in self.lanes[:nl]:
                s2.wait_event(works[b])           # the exchange that still reads buffer b
            works[b] = None
        if self.mode == "stores" and world > 1:
            dst = self.pptr[b][k]
            lr.SetRenderTarget(dst if self.shaded else self.cptr[b][k], self.dptr[b][k] if self.shaded else dst)
        else:
            lr.SetRenderTarget(self.cptr[b][k], self.dptr[b][k])
            if world > 1 and self.mode == "ce":
                lr.SetFrameSink(self.pptr[b][k] if self.shaded else 0, 0 if self.shaded else self.pptr[b][k])
        if self.views is not None:
            lr.SetTransform(self.views[(i * world + rank) % len(self.views)])    # C5: view v is rendered by rank v mod N
        else:
            lr.SetTransform(self.xf)
        lr.RenderMesh(self.meshes[i % len(self.meshes)])
        if k == G - 1 and world > 1:
            for s2, _ in self.lanes[:nl]:
                self.stream.wait_stream(s2)       # the batch's frames (and their pushes), whichever lane rendered them
            self.notify(b, G, works)

    def flush(self, n_steps, works, nl):
        """exchange the frames of a trailing partial batch, then wait for everything in flight"""
        if n_steps % self.G and self.world > 1 and self.mode != "ce":
            for s2, _ in self.lanes[:nl]:
                self.stream.wait_stream(s2)
            self.notify((n_steps // self.G) & 1, n_steps % self.G, works)
        self.drain(works)

    def drain(self, works):
        for b in (0, 1):
            if works[b] is not None:
                works[b].synchronize()
                works[b] = None

    def sync_lanes(self):
        for _, lr in self.lanes:
            lr.Synchronize()                      # also vets the internal queues of the frames submitted so far

    def timed_run(self, nl, steps, n_warm):
        """n_warm untimed steps, then `steps` timed ones with `nl` frames in flight; returns (ms, t0, t1): device time
        between two events on the control stream, barrier + synchronize on both sides, max over ranks."""
        torch, dist, world = self.torch, self.dist, self.world
        stream = self.stream
        with torch.cuda.stream(stream):
            works = [None, None]
            for i in range(n_warm):
                self.step(i, works, nl)
                if i % 64 == 63:
                    self.drain(works)
                    self.sync_lanes()
            self.flush(n_warm, works, nl)
            self.sync_lanes()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev0.record(stream)
            for s2, _ in self.lanes[:nl]:
                s2.wait_stream(stream)
            for i in range(steps):
                self.step(i, works, nl)
            self.host_submit_ms = (time.perf_counter() - t0) * 1e3      # host time to enqueue the timed steps (not a result: says whether the loop is host-bound)
            self.flush(steps, works, nl)
            for s2, lr in self.lanes[:nl]:
                lr.FlushFrameSink()               # the lane's stream waits for its pushes: they are inside the timed region
                stream.wait_stream(s2)
            ev1.record(stream)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            if world > 1:
                dist.barrier()
            ms = ev0.elapsed_time(ev1)
            self.sync_lanes()
            if world > 1 and self.mode == "ce" and self.peer is not None and self.rank == 0:
                # every rank ran the same schedule: rank 0's count words must show each lane's pushes, of every rank
                got = self.counts.cpu().tolist()
                if any(row != self.pushed for row in got):
                    raise RuntimeError("frame gather incomplete: counts on rank 0 %s, expected %s per rank" % (got, self.pushed))
            if world > 1:
                if os.environ.get("EDX_BENCH_RANK_TIMES"):
                    sys.stderr.write("[bench] rank %d: %.4f ms for %d steps, %d in flight\n" % (self.rank, ms, steps, nl))
                t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
        return ms, t0, t1

    def stage_times(self, frames):
        """per-kernel device time (CUDA events between the kernels, same stream), one frame at a time"""
        r = self.r
        r.SetRenderTarget(0, 0)
        r.SetFrameSink(0, 0)
        r.SetProfiling(True)
        stage = {"geom": 0.0, "clip": 0.0, "tile": 0.0, "total": 0.0}
        for i in range(frames):
            r.RenderMesh(self.meshes[i % len(self.meshes)])
            r.Synchronize()
            st = r.GetStats()
            for k in stage:
                stage[k] += st["stage_ms"][k] / frames
        r.SetProfiling(False)
        return stage, r.GetStats()

    def gather_text(self):
        if self.world == 1:
            return "none (1 GPU)"
        what = "colour" if self.shaded else "depth"
        return {"stores": "every finished %s buffer lands on rank 0 by NVLink peer stores from the resolve kernel itself (symmetric memory), one 4-byte all-reduce per %d frames",
                "ce": "every finished %s buffer is pushed into rank 0's symmetric-memory frame store over NVLink by the copy engine right behind the frame's last kernel (edx_set_frame_sink: one D2D copy per frame on the lane's stream, no SM time, no host work), followed by a 4-byte count in rank 0's memory (edx_set_frame_sink_signal) that a consumer polls: no collective, no batches (%d unused); each lane cycles 2 slots, rank 0 checks the counts after the run",
                "none": "NOT GATHERED (--gather none, diagnosis only) %s %d",
                "nccl": "NCCL gather of every finished %s buffer to rank 0, per batch of %d frames, overlapped with the next batch"}[self.mode] % (what, self.G)

    def nvlink_roofline(self, ms_step):
        """What rank 0's NVLink ingress allows: (N-1) frames of frame_bytes per step."""
        if self.world == 1:
            return None
        inbound = (self.world - 1) * self.frame_bytes
        floor_ms = inbound / (NVLINK_INGRESS_GBS * 1e9) * 1e3
        return {"bytes_into_rank0_per_step": inbound, "ingress_peak_gbs": NVLINK_INGRESS_GBS,
                "peak_source": "NVLink 5 nominal per direction (the most measured into one B200 on this pool, 7 senders: 781 GB/s)",
                "floor_ms_per_step": floor_ms, "achieved_gbs": inbound / (ms_step * 1e-3) / 1e9,
                "frac": floor_ms / ms_step}

    def close(self, release_meshes=True):
        if release_meshes and self.own_meshes:
            for m in self.meshes:
                m.Release()
        for _, lr in self.lanes:
            lr.close()


def secondary_config(name, device, peak, R, with_cpu=True, options=None, sc=None):
    """Quick device-resident measurement of another BASELINE config (rank 0, N=1 only), with its CPU baseline."""
    t0 = time.time()
    sc = sc if sc is not None else make_scene(name, 1.0)
    r = R.Renderer(device)
    r.Initialize(sc.width, sc.height)
    r.SetTransform(sc.mv, sc.proj, sc.raster)
    r.SetPixelShader(sc.shader)
    for k, v in (options or {}).items():
        r.SetOption(k, int(v))
    copies = mesh_copies(sc.num_verts, sc.num_tris)
    meshes = [r.CreateMesh(sc.vertices, sc.indices) for _ in range(copies)]
    frames = 30 if sc.num_tris < 5_000_000 else 10
    ms = time_frames(r, meshes, frames, 3) / frames
    # the same frames through a FrameRing (3 in flight, meshes shared): host wall clock around a synchronised batch
    ring = R.FrameRing(device, depth=3)
    ring.Initialize(sc.width, sc.height)
    ring.SetPixelShader(sc.shader)
    for k, v in (options or {}).items():
        ring.SetOption(k, int(v))
    nring = frames * 4
    xf = R.PackedTransform(sc.mv, sc.proj, sc.raster)      # marshalled once: a 20 us frame leaves no room for numpy conversions
    for lane in ring.lanes:           # size every lane's queues (and let its routing settle) before frames pile up unsynchronised
        lane.SetTransform(xf)
        for _ in range(3):
            lane.RenderMesh(meshes[0])
            lane.Synchronize()
    for rep in range(2):
        ring.Synchronize()
        tw = time.perf_counter()
        for i in range(nring):
            ring.Submit(meshes[i % copies], xf)
        ring.Synchronize()
        ms_ring = (time.perf_counter() - tw) * 1e3 / nring
    ring.close()
    r.SetProfiling(True)
    stage = {"geom": 0.0, "clip": 0.0, "tile": 0.0}
    for i in range(5):
        r.RenderMesh(meshes[i % copies])
        r.Synchronize()
        st = r.GetStats()
        for k in stage:
            stage[k] += st["stage_ms"][k] / 5
    r.SetProfiling(False)
    st = r.GetStats()
    ab = algorithmic_bytes(sc.num_verts, sc.num_tris, sc.width, sc.height, sc.shader != 0)
    out = {"workload": WORKLOADS[name], "ms_per_frame": ms, "mtris_per_s": sc.num_tris / ms / 1e3,
           "ms_per_frame_3_in_flight": ms_ring, "mtris_per_s_3_in_flight": sc.num_tris / ms_ring / 1e3,
           "gpix_per_s": sc.width * sc.height / ms / 1e6, "frames_per_s": 1000.0 / ms,
           "algorithmic_bytes": ab, "hbm_frac_whole_frame": ab / (ms * 1e-3) / 1e9 / peak,
           "hbm_frac_whole_frame_3_in_flight": ab / (ms_ring * 1e-3) / 1e9 / peak,
           "fb_only_frac": (8 if sc.shader != 0 else 4) * sc.width * sc.height / (ms * 1e-3) / 1e9 / peak,
           "stage_ms": stage, "binned_tris": st["binned_tris"], "clipped_tris": st["clipped_tris"], "mid_tris": st["mid_tris"],
           "bin_pairs": st["bin_pairs"], "l2": "rotating %d mesh copies" % copies}
    for m in meshes:
        m.Release()
    r.close()
    log("secondary %s: %.3f ms/frame (%.1fs incl. generation)" % (name, ms, time.time() - t0))
    if with_cpu:
        try:
            out["cpu_baseline"] = cpu_baseline(sc, cpu_seconds=12.0, max_frames=40)
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "error": str(e)}
    return out, sc


def config5(torch, dist, R, sc4, local, rank, world, args, peak):
    """BASELINE.json configs[4]: 256 camera views of the 10M-triangle mesh, view i on GPU i mod N, every finished
    colour buffer gathered to rank 0 inside the timed region. Strong scaling: the 256 views are the job."""
    from edxraster_b200 import scenes
    t0 = time.time()
    sc = sc4 if sc4 is not None else make_scene("C4", args.scale)
    sc = scenes.Scene(sc)
    n_views = 256
    sc["views"] = scenes.config5_views(sc, n_views)
    farm = Farm(torch, dist, R, sc, local, rank, world, args.in_flight, "stores" if args.peer_stores else args.gather, 4)
    per_rank = n_views // world
    ms, _, _ = farm.timed_run(farm.K, per_rank, 8)
    nt, nv, W, H = sc.num_tris, sc.num_verts, sc.width, sc.height
    ab = algorithmic_bytes(nv, nt, W, H, True)
    ms_view = ms / per_rank
    out = {"workload": WORKLOADS["C5"], "views": n_views, "n_gpus": world, "scaling": "strong",
           "ms_total": ms, "ms_per_view_per_gpu": ms_view, "frames_per_s": n_views / (ms * 1e-3),
           "mtris_per_s": n_views * nt / ms / 1e3, "gpix_per_s": n_views * W * H / ms / 1e6,
           "frames_in_flight_per_gpu": farm.K, "gather": farm.gather_text(),
           "gathered_bytes": (world - 1) * per_rank * farm.frame_bytes,
           "algorithmic_bytes_per_view": ab, "hbm_frac_whole_frame": ab / (ms_view * 1e-3) / 1e9 / peak,
           "nvlink": farm.nvlink_roofline(ms_view)}
    if rank == 0 and world == 1:
        stage, st = farm.stage_times(5)
        out["stage_ms"] = stage
        out["binned_tris"], out["clipped_tris"] = st["binned_tris"], st["clipped_tris"]
    farm.close()
    if rank == 0 and world == 1:
        try:
            out["cpu_baseline"] = cpu_baseline(sc, cpu_seconds=12.0, max_frames=8, views=sc["views"])
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "error": str(e)}
    log("C5: %d views on %d GPU(s) in %.2f ms (%.1fs incl. setup)" % (n_views, world, ms, time.time() - t0))
    return out


def ours_arm(args):
    # stdout carries exactly one JSON line: park the real stdout and point fd 1 at stderr so that nothing a
    # library prints (NCCL's version banner, for one) can land next to it
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from edxraster_b200 import renderer as R

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its banner / debug lines to stdout by default; stdout is reserved for the JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    peak, peak_src = measured_hbm_peak()

    sc = make_scene(args.workload, args.scale)
    W, H, nt, nv = sc.width, sc.height, sc.num_tris, sc.num_verts
    shaded = sc.shader != 0
    mode = "stores" if args.peer_stores else args.gather
    G = 4
    farm = Farm(torch, dist, R, sc, local, rank, world, args.in_flight, mode, G)
    K, lanes, stream, meshes, copies, r, xf = farm.K, farm.lanes, farm.stream, farm.meshes, farm.copies, farm.r, farm.xf

    sampler = ClockSampler(local)
    sampler.start()                           # NVML calls take ~1 ms each: sample through warm-up and the timed region
    # warm-up: W steps plus a fixed 1000 more so clocks settle (a fixed count: every rank must issue the
    # same number of exchanges)
    n_warm = ((max(args.warmup, 3) + 1000 + G - 1) // G) * G
    ms_total, t0, t1 = farm.timed_run(K, args.steps, n_warm)
    host_submit_ms = farm.host_submit_ms / args.steps
    sampler.stop()
    launches_per_step = r.LastLaunchCount()
    launch_list = r.LastLaunchList()
    ms_step = ms_total / args.steps
    value = world * nt / ms_step / 1e3        # Mtris/s, whole job
    ms_one = None
    if K > 1:                                 # the same loop with ONE frame in flight, for the record
        ms_one = farm.timed_run(1, args.steps, ((max(args.warmup, 3) + G - 1) // G) * G)[0] / args.steps

    with torch.cuda.stream(stream):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # ---- end-to-end through the C ABI with HOST buffers: upload mesh, render, read result back ----
        hv = torch.from_numpy(np.ascontiguousarray(sc.vertices)).pin_memory()
        hi = torch.from_numpy(np.ascontiguousarray(sc.indices).view(np.int32)).pin_memory()
        houts = [torch.empty((H, W), dtype=torch.float32).pin_memory() if not shaded else None for _ in lanes]
        for _, lr in lanes:
            lr.SetRenderTarget(0, 0)
            lr.SetFrameSink(0, 0)
        # streamed geometry is written by its lane's upload: one private mesh per lane (lane 0 reuses a shared one)
        lane_mesh = [meshes[0]] + [lr.CreateMesh(sc.vertices, sc.indices) for _, lr in lanes[1:]]
        e2e_steps = max(3, min(args.steps, 20 if nt < 5_000_000 else 5))

        def e2e_submit(i, upload):
            _, lr = lanes[i % K]
            if upload:
                lane_mesh[i % K].update(hv.data_ptr(), nv, hi.data_ptr(), nt)
            lr.SetTransform(xf)
            lr.RenderMesh(lane_mesh[i % K] if upload else meshes[i % copies])

        def e2e_read(i):
            _, lr = lanes[i % K]
            if shaded:
                lr.GetBackBuffer()                # D2H into the lane's pinned mirror
            else:
                lr.ReadDepthInto(houts[i % K].data_ptr())

        def e2e_run(n, upload):
            """n frames, K in flight: frame i - K + 1 is read back to the host right after frame i is submitted"""
            for i in range(n + K - 1):
                if i < n:
                    e2e_submit(i, upload)
                if i >= K - 1:
                    e2e_read(i - K + 1)

        e2e = {}
        for label, upload in (("stream", True), ("resident", False)):
            e2e_run(3, upload)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ev0.record(stream)
            for s2, _ in lanes:
                s2.wait_stream(stream)
            e2e_run(e2e_steps, upload)
            for s2, _ in lanes:
                stream.wait_stream(s2)
            ev1.record(stream)
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            e2e[label] = ms / e2e_steps
        h2d = nv * 32 + nt * 12 + 3 * 64
        d2h = W * H * 4

    stage, stats = farm.stage_times(max(5, min(args.steps, 20)))
    for m in lane_mesh[1:]:
        m.Release()
    gather_text, nvlink = farm.gather_text(), farm.nvlink_roofline(ms_step)
    farm.close()

    # ---- BASELINE.json configs[4] at this N (every rank takes part), then the other configs at N = 1 ----
    also = {}
    sc4 = None
    if not args.no_extra and args.workload == "C2":
        if world == 1:
            for name in ("C1", "C3", "C4"):
                try:
                    also[name], scx = secondary_config(name, local, peak, R)
                    if name == "C4":
                        sc4 = scx
                except Exception as e:
                    also[name] = {"error": str(e)}
            # M1: not a BASELINE config - the stress case for long tile-path lists (per-bin lists, stage a7) and for the
            # routing of mid-size triangles; the second line is the same frame with round 1's design (every triangle
            # above 32 px on one shared list that every bin reads, no per-bin lists)
            try:
                also["M1"], scm = secondary_config("M1", local, peak, R, with_cpu=False)
                old, _ = secondary_config("M1", local, peak, R, with_cpu=False, sc=scm,
                                          options={"bin_min": 0, "mid_auto": 0, "mid_max": 0, "small_max": 32})
                also["M1"]["shared_list_round1_routing"] = {k: old[k] for k in ("ms_per_frame", "ms_per_frame_3_in_flight", "stage_ms", "binned_tris", "mid_tris", "bin_pairs")}
                del scm
            except Exception as e:
                also["M1"] = {"error": str(e)}
        try:
            also["C5"] = config5(torch, dist, R, sc4, local, rank, world, args, peak)
        except Exception as e:       # pragma: no cover
            if world > 1:
                raise
            also["C5"] = {"error": str(e)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    ab = algorithmic_bytes(nv, nt, W, H, shaded)
    dom = max(("geom", "clip", "tile"), key=lambda k: stage[k])
    dom_ms = stage[dom]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get(args.workload, {}).get(dom + "_kernel")
        except Exception:
            traffic = None
    achieved = ab / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    line = {
        "metric": "Mtris/s", "value": value, "unit": "Mtris/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 depth / i32 28.4 fixed-point coverage", "data": "synthetic",
        "config": config_block(args.workload, nv, nt, W, H),
        "execution": {"frames_in_flight_per_gpu": K,
                      "in_flight": "each GPU keeps %d independent frames in flight (contexts on their own streams sharing the meshes); "
                                   "every frame is rendered completely and, at N > 1, gathered inside the timed region" % K,
                      "gather": gather_text,
                      "host_submit_ms_per_step": host_submit_ms},        # rank 0's host time to enqueue one step; close to ms_per_step = host-bound
        "frames_per_s": world * 1000.0 / ms_step, "gpix_per_s": world * W * H / ms_step / 1e6,
        "one_frame_in_flight": None if ms_one is None else {"ms_per_step": ms_one, "value": world * nt / ms_one / 1e3},
        "clocks": sampler.summary(t0, t1),
        "e2e": {"value": world * nt / e2e["stream"] / 1e3, "unit": "Mtris/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e["stream"],
                "what": "edx_mesh_update (pinned host vertices+indices -> device) + edx_set_transform + edx_render_mesh + read-back of the frame to pinned host, every step; "
                        "%d frames in flight: frame i-%d is read back right after frame i is submitted, the drain is inside the timed region" % (K, K - 1),
                "resident_mesh_value": world * nt / e2e["resident"] / 1e3, "resident_mesh_ms_per_step": e2e["resident"],
                "resident_mesh_what": "mesh uploaded once (the reference viewer's usage, Main.cpp:42,71-75); per step: transform in, render, frame read back to host"},
        "gpu_launches": launches_per_step * args.steps,
        "kernels_per_step": {k: launch_list.count(k) for k in dict.fromkeys(launch_list)},
        "roofline": {"bound": "hbm", "kernel": dom + "_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ab, "kernel_ms": dom_ms,
                     "stage_ms": stage, "whole_frame_frac": ab / (ms_step * 1e-3) / 1e9 / peak,
                     "fb_only_frac": (8 if shaded else 4) * W * H / (ms_step * 1e-3) / 1e9 / peak,
                     "nvlink": nvlink},
        "path_stats": {"binned_tris": stats["binned_tris"], "clipped_tris": stats["clipped_tris"], "regrows": stats["regrow_count"]},
    }
    if world == 1:
        # CPU baseline on this box's host cores: bounded sample of the same workload
        try:
            line["cpu_baseline"] = cpu_baseline(sc, cpu_seconds=20.0)
        except Exception as e:       # the CPU arm is only a reported baseline
            line["cpu_baseline"] = {"value": None, "error": str(e)}
    if also:
        line["other_configs"] = also
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    real_stdout.write(json.dumps(line) + "\n")
    real_stdout.flush()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--gather", default="ce", choices=["nccl", "ce", "stores", "none"], help="how finished frames reach rank 0 at N > 1")
    ap.add_argument("--in-flight", type=int, default=3, help="independent frames in flight per GPU (1 = one frame at a time)")
    ap.add_argument("--scale", type=float, default=1.0, help="triangle-count scale (debug only; 1.0 = BASELINE size)")
    ap.add_argument("--opt", action="append", default=[], help="experiment: edx_set_option name=value on every context of the headline run")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary C1/C3/C4 measurements")
    ap.add_argument("--peer-stores", action="store_true", help="N > 1: let every rank's resolve kernel store its pixels straight into rank 0's (symmetric) memory over NVLink instead of the NCCL gather; measured slower (32-byte row segments): 88 vs 73 us/step at N = 4")
    args = ap.parse_args()
    OPTIONS.extend(args.opt)
    if args.impl == "reference":
        return reference_arm(args)
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: self-launch one process per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return ours_arm(args)


if __name__ == "__main__":
    sys.exit(main())
