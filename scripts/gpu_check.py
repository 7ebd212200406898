"""First-contact GPU check: renders reduced configs on cuda:0 and prints the parity report."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import numpy as np

from edxraster_b200 import scenes
import parity


def main():
    cases = [
        ("C1", scenes.config1()),
        ("C2_100k", scenes.config2(num_tris=100000)),
        ("C3_small", scenes.config3(width=960, height=540, num_tris=200)),
        ("C4_small", scenes.config4(quads_x=500, quads_z=400)),
    ]
    ok_all = True
    for name, sc in cases:
        t = time.time()
        ref = parity.render_oracle(sc)
        t1 = time.time()
        got = parity.render_gpu(sc)
        t2 = time.time()
        rep = parity.compare(ref, got)
        ok = parity.is_parity(rep)
        ok_all &= ok
        print(name, "PARITY" if ok else "MISMATCH", json.dumps(rep), "oracle %.2fs gpu %.2fs" % (t1 - t, t2 - t1), got["stats"], flush=True)
        if not ok:
            bad = np.argwhere(ref["winner"] != got["winner"])
            print("  first winner mismatches (row,col):", bad[:5].tolist())
            for (y, x) in bad[:5]:
                print("   ", y, x, "ref", ref["winner"][y, x], ref["depth"][y, x], "got", got["winner"][y, x], got["depth"][y, x])
    print("ALL OK" if ok_all else "FAILURES")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
