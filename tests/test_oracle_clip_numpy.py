"""A scalar numpy float32 restatement of the polygon clipper written from the reference's formulas alone
(Core/Clipper.h:192-288: predicate, computeT, `v0 * (1 - t) + v1 * t`, snap, weights, plane order L R B T FAR NEAR,
the w <= 0 drop), compared bit for bit with the oracle's clipper on random straddling triangles."""
import numpy as np

from oracle import orc

f32 = np.float32
LEFT, RIGHT, BOTTOM, TOP, FAR, NEAR = 1, 2, 4, 8, 16, 32

PLANES = [   # bit, inside(v), t(v0, v1), snap(v)   -- Clipper.h:237-278
    (LEFT, lambda v: v[0] >= -v[3], lambda a, b: f32(f32(a[3] + a[0]) / f32(f32(a[0] + a[3]) - f32(b[0] + b[3]))), lambda v: v.__setitem__(0, -v[3])),
    (RIGHT, lambda v: v[0] <= v[3], lambda a, b: f32(f32(-a[3] + a[0]) / f32(f32(a[0] - a[3]) - f32(b[0] - b[3]))), lambda v: v.__setitem__(0, v[3])),
    (BOTTOM, lambda v: v[1] >= -v[3], lambda a, b: f32(f32(a[3] + a[1]) / f32(f32(a[1] + a[3]) - f32(b[1] + b[3]))), lambda v: v.__setitem__(1, -v[3])),
    (TOP, lambda v: v[1] <= v[3], lambda a, b: f32(f32(-a[3] + a[1]) / f32(f32(a[1] - a[3]) - f32(b[1] - b[3]))), lambda v: v.__setitem__(1, v[3])),
    (FAR, lambda v: v[2] <= v[3], lambda a, b: f32(f32(-a[3] + a[2]) / f32(f32(a[2] - a[3]) - f32(b[2] - b[3]))), lambda v: v.__setitem__(2, v[3])),
    (NEAR, lambda v: v[2] >= 0, lambda a, b: f32(a[2] / f32(a[2] - b[2])), lambda v: v.__setitem__(2, f32(0.0))),
]


def code(v):    # Clipper.h:48-68
    c = 0
    if v[0] < -v[3]: c |= LEFT
    if v[0] > v[3]: c |= RIGHT
    if v[1] < -v[3]: c |= BOTTOM
    if v[1] > v[3]: c |= TOP
    if v[2] > v[3]: c |= FAR
    if v[2] < 0: c |= NEAR
    return c


def lerp(a, b, t):              # `a * (1 - t) + b * t`, component-wise in float32
    s = f32(f32(1.0) - t)
    return (a * s).astype(f32) + (b * t).astype(f32)


def clip(tri):
    poly = [(tri[k].copy(), np.eye(3, dtype=f32)[k]) for k in range(3)]
    c = [code(v) for v in tri]
    planes = (c[0] ^ c[1]) | (c[1] ^ c[2]) | (c[2] ^ c[0])        # Clipper.h:119
    with np.errstate(all="ignore"):
        for bit, inside, tfun, snap in PLANES:
            if not planes & bit:
                continue
            out = []
            for i in range(len(poly)):
                (v0, w0), (v1, w1) = poly[i], poly[(i + 1) % len(poly)]
                if inside(v0):
                    if inside(v1):
                        out.append((v1, w1))
                    else:
                        t = tfun(v0, v1); p = lerp(v0, v1, t).astype(f32); snap(p); out.append((p, lerp(w0, w1, t).astype(f32)))
                elif inside(v1):
                    t = tfun(v0, v1); p = lerp(v0, v1, t).astype(f32); snap(p); out.append((p, lerp(w0, w1, t).astype(f32)))
                    out.append((v1, w1))
            poly = out
    if any(v[3] <= 0 for v, _ in poly):                           # Clipper.h:280-287
        return []
    return poly


def test_numpy_clipper_matches_the_oracle_bit_for_bit():
    rng = np.random.default_rng(7)
    checked = multi = dropped = 0
    for it in range(3000):
        tri = (rng.normal(size=(3, 4)) * np.array([2.0, 2.0, 1.5, 1.0]) + np.array([0, 0, 0.5, 1.2])).astype(f32)
        if it % 5 == 0:
            tri[:, 3] = np.abs(tri[:, 3]) + f32(0.2)              # all in front: side / far planes only
        c = [code(v) for v in tri]
        pos, wt = orc.clip_triangle(tri)
        if not (c[0] | c[1] | c[2]):
            assert pos is None
            continue
        if c[0] & c[1] & c[2]:
            assert pos is not None and len(pos) == 0              # Clipper.h:109: trivially rejected
            continue
        want = clip(tri)
        assert len(pos) == len(want), (it, tri, len(pos), len(want))
        for k, (p, w) in enumerate(want):
            assert p.view(np.uint32).tolist() == pos[k].view(np.uint32).tolist(), (it, k, p, pos[k])
            assert np.asarray(w, f32).view(np.uint32).tolist() == wt[k].view(np.uint32).tolist(), (it, k, w, wt[k])
        checked += 1
        multi += bin((c[0] ^ c[1]) | (c[1] ^ c[2]) | (c[2] ^ c[0])).count("1") > 1
        dropped += len(want) == 0
    assert checked > 1500 and multi > 500 and dropped > 20
