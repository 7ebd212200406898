"""One-off wide fuzz sweep: GPU vs oracle on many random scenes (see tests/test_gpu_parity.py::_fuzz_scene)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity
from test_gpu_parity import _fuzz_scene
from edxraster_b200 import renderer as R
lo, hi = int(sys.argv[1]), int(sys.argv[2])
r = R.Renderer(0)
bad = 0
for seed in range(lo, hi):
    sc, msaa = _fuzz_scene(seed)
    ref = parity.render_oracle(sc, msaa=msaa)
    got = parity.render_gpu(sc, msaa=msaa, stages=False, renderer=r)
    rep = parity.compare(ref, got)
    if not parity.is_parity(rep):
        bad += 1
        print("MISMATCH seed", seed, "msaa", msaa, rep, flush=True)
print("seeds %d..%d: %d mismatches" % (lo, hi, bad))
