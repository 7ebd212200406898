"""bench.py's reference arm runs on CPU: check that it prints exactly one JSON line with the contract's keys
(the GPU arm is exercised on the GPU box). Also the host-side marshalling helper the bench loop relies on."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--scale", "0.02"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mtris/s" and d["unit"] == "Mtris/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("C2")
    from oracle import ref
    # the reference's own compiled sources wherever oracle/_ref/libref.so is present, the restatement elsewhere
    assert d["cpu_baseline"]["kind"] == ("reference" if ref.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    # both arms build `config` with the same function from the same scene: the driver's same_config check holds
    sys.path.insert(0, ROOT)
    import bench
    sc = bench.make_scene("C2", 0.02)
    assert d["config"] == bench.config_block("C2", sc.num_verts, sc.num_tris, sc.width, sc.height)
    assert d["e2e"] == {"value": d["value"], "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_packed_transform_marshals_row_major_float32():
    from edxraster_b200.renderer import PackedTransform
    rng = np.random.default_rng(0)
    mv, p, r = (rng.normal(size=(4, 4)) for _ in range(3))
    t = PackedTransform(mv, p, r)
    for got, want in ((t.mv, mv), (t.proj, p), (t.raster, r)):
        np.testing.assert_array_equal(np.array(list(got), np.float32), want.astype(np.float32).reshape(16))


def test_reference_arm_runs_config5_views():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C5", "--steps", "3", "--warmup", "1", "--scale", "0.005"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip())
    assert d["config"]["workload"].startswith("C5") and d["value"] > 0
