"""Sort-first split of ONE frame across the GPUs of a box (SURVEY.md §8e optional mode), under torchrun:
every rank runs the geometry stages on the whole mesh, rasterises / shades only its interleaved 64x64 bins
(edx_set_screen_partition), and rank 0 gathers the buffers and composites them. Strong scaling of one frame.
    torchrun --nproc-per-node N scripts/sort_first_bench.py --workload C3 --frames 50
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from edxraster_b200 import farm, renderer as R, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C3")
ap.add_argument("--frames", type=int, default=50)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
sc = scenes.by_name(a.workload, 1.0)
stream = torch.cuda.Stream(device=dev)
r = R.Renderer(local)
r.SetStream(stream.cuda_stream)
r.Initialize(sc.width, sc.height)
r.SetTransform(sc.mv, sc.proj, sc.raster)
r.SetPixelShader(sc.shader)
r.SetScreenPartition(rank, world)
m = r.CreateMesh(sc.vertices, sc.indices)
color = torch.zeros((sc.height, sc.width, 4), dtype=torch.uint8, device=dev)
depth = torch.zeros((sc.height, sc.width), dtype=torch.float32, device=dev)
r.SetRenderTarget(color.data_ptr(), depth.data_ptr())
recv = [torch.empty_like(color) for _ in range(world)] if rank == 0 and world > 1 else None
mask = [farm.bin_owner_mask(sc.width, sc.height, p, world, device=dev) for p in range(world)] if rank == 0 else None

def frame():
    r.RenderMesh(m)
    if world > 1:
        dist.gather(color, recv, dst=0)
        if rank == 0:
            out = recv[0]
            for p in range(1, world):
                out[mask[p]] = recv[p][mask[p]]
            return out
    return color

with torch.cuda.stream(stream):
    for _ in range(5):
        frame()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.frames):
        r.RenderMesh(m)
    e1.record(stream)
    torch.cuda.synchronize()
    render_ms = e0.elapsed_time(e1) / a.frames
    e0.record(stream)
    for _ in range(a.frames):
        out = frame()
    e1.record(stream)
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1) / a.frames
    t = torch.tensor([render_ms, total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": a.workload, "n_gpus": world, "render_ms_per_frame_max_rank": float(t[0]), "frame_ms_incl_gather_and_composite": float(t[1]),
                      "frames_per_s": 1000.0 / float(t[1])}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
