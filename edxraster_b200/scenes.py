"""Procedural workloads: the five BASELINE.json configs at any scale.

No scene assets exist offline and the reference's own generators live in the absent EDXUtil
(`ObjMesh::LoadSphere/LoadPlane`, Utils/Mesh.cpp:36-70), so every mesh here is generated. The
vertex layout is the reference's submission format (Utils/InputBuffer.h:16-28): 32 bytes per vertex,
position(3) normal(3) texcoord(2), and uint32 x 3 indices (InputBuffer.h:148-194).

All generators are deterministic (fixed seeds, numpy PCG64) so the oracle, the CUDA path and the
committed golden fixtures see identical bytes.
"""
import math

import numpy as np

from . import camera as cam

SHADER_DEPTH_ONLY, SHADER_BLINN_PHONG, SHADER_LAMBERT, SHADER_LAMBERT_ALBEDO = 0, 1, 2, 3


class Scene(dict):
    """dict with attribute access: name,width,height,vertices,indices,mv,proj,raster,shader."""
    __getattr__ = dict.__getitem__

    @property
    def num_tris(self):
        return int(self["indices"].shape[0])

    @property
    def num_verts(self):
        return int(self["vertices"].shape[0])


def _pack(pos, nrm, uv):
    v = np.empty((pos.shape[0], 8), np.float32)
    v[:, 0:3], v[:, 3:6], v[:, 6:8] = pos, nrm, uv
    return v


# ---------------------------------------------------------------------------------------------
# C1: UV sphere as the viewer shows it (RealtimeViewer/Main.cpp:39,42)
# ---------------------------------------------------------------------------------------------
def uv_sphere(radius=1.2, slices=100, stacks=100):
    theta = np.linspace(0.0, math.pi, stacks + 1)
    phi = np.linspace(0.0, 2.0 * math.pi, slices + 1)
    t, p = np.meshgrid(theta, phi, indexing="ij")
    n = np.stack([np.sin(t) * np.cos(p), np.cos(t), np.sin(t) * np.sin(p)], -1).reshape(-1, 3)
    uv = np.stack([p / (2.0 * math.pi), t / math.pi], -1).reshape(-1, 2)
    i, j = np.meshgrid(np.arange(stacks), np.arange(slices), indexing="ij")
    a = (i * (slices + 1) + j).reshape(-1)
    b = a + 1
    c = a + (slices + 1)
    d = c + 1
    idx = np.stack([np.stack([a, b, c], -1), np.stack([b, d, c], -1)], 1).reshape(-1, 3)
    return _pack(n * radius, n, uv), idx.astype(np.uint32)


def config1(width=1280, height=720, slices=100, stacks=100):
    v, i = uv_sphere(1.2, slices, stacks)
    c = cam.Camera((0.0, 0.0, -5.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), width, height, 65.0, 0.01, 100.0)
    return Scene(name="C1_sphere", width=width, height=height, vertices=v, indices=i,
                 mv=c.view, proj=c.proj, raster=c.raster, shader=SHADER_BLINN_PHONG)


# ---------------------------------------------------------------------------------------------
# C2: random small triangles, depth only, fixed order
# ---------------------------------------------------------------------------------------------
def config2(width=1920, height=1080, num_tris=1_000_000, max_px=4.0, seed=0xEDD5A57E2, min_px=None, name="C2_small_tris"):
    rng = np.random.default_rng(seed)
    centre = rng.random((num_tris, 1, 2)) * np.array([width, height])
    size = max_px if min_px is None else min_px + (max_px - min_px) * rng.random((num_tris, 1, 1))
    off = (rng.random((num_tris, 3, 2)) - 0.5) * size
    xy = centre + off                                   # raster-space pixels, y down
    z = 0.1 + 0.8 * rng.random((num_tris, 3))
    # force the orientation the reference keeps (det > 0, RasterTriangle.h:49-51)
    a = xy[:, 0] - xy[:, 2]
    b = xy[:, 1] - xy[:, 2]
    flip = (a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]) < 0
    xy[flip, 0], xy[flip, 1] = xy[flip, 1].copy(), xy[flip, 0].copy()
    z[flip, 0], z[flip, 1] = z[flip, 1].copy(), z[flip, 0].copy()
    pos = np.empty((num_tris, 3, 3))
    pos[..., 0] = 2.0 * xy[..., 0] / width - 1.0      # NDC; MV = P = identity so w = 1
    pos[..., 1] = 1.0 - 2.0 * xy[..., 1] / height
    pos[..., 2] = z
    pos = pos.reshape(-1, 3)
    nrm = np.tile(np.array([0.0, 0.0, -1.0]), (pos.shape[0], 1))
    uv = rng.random((pos.shape[0], 2))
    idx = np.arange(num_tris * 3, dtype=np.uint32).reshape(-1, 3)
    return Scene(name=name, width=width, height=height, vertices=_pack(pos, nrm, uv), indices=idx,
                 mv=cam.identity(), proj=cam.identity(), raster=cam.raster_matrix(width, height),
                 shader=SHADER_DEPTH_ONLY)


def stress_m1(width=1920, height=1080, num_tris=1_000_000, seed=0x111):
    """M1 - not a BASELINE config: a stress case for the routing of MID-SIZE triangles. One million
    triangles whose vertices scatter over boxes of 32..128 px, depth only: too large for the per-thread path, and with
    round 1's routing all of them landed on the tile path, where every 64x64 bin swept the whole list."""
    return config2(width, height, num_tris, max_px=128.0, seed=seed, min_px=32.0, name="M1_mid_tris")


# ---------------------------------------------------------------------------------------------
# C3: screen-covering triangles, fill / shading stress, perspective-correct attributes
# ---------------------------------------------------------------------------------------------
def config3(width=3840, height=2160, num_tris=2000, seed=0xC3):
    rng = np.random.default_rng(seed)
    near, far, fov = 0.5, 10.0, 65.0
    proj = cam.perspective_lh(fov, width / float(height), near, far)
    xs, ys = float(proj[0, 0]), float(proj[1, 1])
    i = np.arange(num_tris)
    level = ((i * 7919) % num_tris) / float(num_tris)
    zbase = 1.2 + 1.6 * level                           # view depth; z_ndc rises with it
    tilt = 1.0 + 0.08 * (rng.random((num_tris, 3)) - 0.5)
    zv = zbase[:, None] * tilt                          # per-vertex view depth == clip w in [1, 3]
    ndc = np.array([[-1.5, -1.5], [-1.5, 4.0], [4.0, -1.5]])[None] + 0.2 * (rng.random((num_tris, 3, 2)) - 0.5)
    pos = np.empty((num_tris, 3, 3))
    pos[..., 0] = ndc[..., 0] * zv / xs
    pos[..., 1] = ndc[..., 1] * zv / ys
    pos[..., 2] = zv
    nrm = np.array([0.0, 0.0, -1.0])[None, None] + 0.6 * (rng.random((num_tris, 3, 3)) - 0.5)
    uv = rng.random((num_tris, 3, 2))
    idx = np.arange(num_tris * 3, dtype=np.uint32).reshape(-1, 3)
    return Scene(name="C3_fill", width=width, height=height,
                 vertices=_pack(pos.reshape(-1, 3), nrm.reshape(-1, 3), uv.reshape(-1, 2)), indices=idx,
                 mv=cam.identity(), proj=proj, raster=cam.raster_matrix(width, height), shader=SHADER_BLINN_PHONG)


# ---------------------------------------------------------------------------------------------
# C4 / C5: displaced grid seen from just above its surface (clipping + setup bound)
# ---------------------------------------------------------------------------------------------
def _value_noise(x, z, seed, lattice=64):
    rng = np.random.default_rng(seed)
    tab = rng.random((lattice, lattice))
    out = np.zeros_like(x)
    amp, freq = 1.0, 1.0
    for _ in range(3):
        fx, fz = x * freq, z * freq
        ix, iz = np.floor(fx).astype(np.int64), np.floor(fz).astype(np.int64)
        tx, tz = fx - ix, fz - iz
        tx, tz = tx * tx * (3.0 - 2.0 * tx), tz * tz * (3.0 - 2.0 * tz)
        a = tab[ix % lattice, iz % lattice]
        b = tab[(ix + 1) % lattice, iz % lattice]
        c = tab[ix % lattice, (iz + 1) % lattice]
        d = tab[(ix + 1) % lattice, (iz + 1) % lattice]
        out += amp * ((a * (1 - tx) + b * tx) * (1 - tz) + (c * (1 - tx) + d * tx) * tz)
        amp *= 0.5
        freq *= 2.0
    return out / 1.75


def displaced_grid(quads_x=2500, quads_z=2000, cell=0.02, amplitude=0.12, seed=0xC4):
    nx, nz = quads_x + 1, quads_z + 1
    gx = (np.arange(nx) - quads_x * 0.5) * cell
    gz = (np.arange(nz) - quads_z * 0.5) * cell
    x, z = np.meshgrid(gx, gz, indexing="xy")            # shape (nz, nx), x fastest
    y = amplitude * _value_noise(x * 0.9, z * 0.9, seed)
    dydx = np.gradient(y, cell, axis=1)
    dydz = np.gradient(y, cell, axis=0)
    n = np.stack([-dydx, np.ones_like(y), -dydz], -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    pos = np.stack([x, y, z], -1).reshape(-1, 3)
    uv = np.stack([(x - gx[0]) / (gx[-1] - gx[0]), (z - gz[0]) / (gz[-1] - gz[0])], -1).reshape(-1, 2)
    j, i = np.meshgrid(np.arange(quads_z), np.arange(quads_x), indexing="ij")
    a = (j * nx + i).reshape(-1)
    b = a + 1
    c = a + nx
    d = c + 1
    idx = np.stack([np.stack([a, c, b], -1), np.stack([b, c, d], -1)], 1).reshape(-1, 3)
    return _pack(pos, n.reshape(-1, 3), uv), idx.astype(np.uint32), (gx, gz, y)


def grid_camera(height_field, width, height, yaw=0.0, ring=8.0, above=0.05, fov=65.0, near=0.1, far=100.0):
    gx, gz, y = height_field
    ex, ez = ring * math.sin(yaw + math.pi), ring * math.cos(yaw + math.pi)
    ix = int(np.clip(np.searchsorted(gx, ex), 1, len(gx) - 2))
    iz = int(np.clip(np.searchsorted(gz, ez), 1, len(gz) - 2))
    r = max(1, int(0.3 / (gx[1] - gx[0])))               # stay above the terrain within 30 cm
    ground = float(y[max(iz - r, 0):iz + r + 1, max(ix - r, 0):ix + r + 1].max())
    eye = (ex, ground + above, ez)
    target = (ex + math.sin(yaw) * 10.0, ground + above - 0.35, ez + math.cos(yaw) * 10.0)
    return cam.Camera(eye, target, (0.0, 1.0, 0.0), width, height, fov, near, far)


def config4(width=1920, height=1080, quads_x=2500, quads_z=2000, yaw=0.0, cell=None):
    if cell is None:
        cell = 50.0 / quads_x                            # keep the 50 x 40 unit footprint at any scale
    v, i, hf = displaced_grid(quads_x, quads_z, cell)
    ring = min(8.0, 0.16 * quads_x * cell)
    c = grid_camera(hf, width, height, yaw, ring=ring)
    s = Scene(name="C4_grid", width=width, height=height, vertices=v, indices=i,
              mv=c.view, proj=c.proj, raster=c.raster, shader=SHADER_BLINN_PHONG)
    s["height_field"] = hf
    s["ring"] = ring
    return s


def config5_views(scene4, num_views=256):
    """C5: `num_views` cameras on a circle over the C4 mesh (yaw_i = 2*pi*i/num_views)."""
    out = []
    for k in range(num_views):
        c = grid_camera(scene4["height_field"], scene4.width, scene4.height, 2.0 * math.pi * k / num_views,
                        ring=scene4["ring"])
        out.append((c.view, c.proj, c.raster))
    return out


def by_name(name, scale=1.0):
    """Config by short name at a linear `scale` of its triangle count (1.0 = BASELINE.json size)."""
    name = name.upper()
    if name == "C1":
        n = max(4, int(round(100 * math.sqrt(scale))))
        return config1(slices=n, stacks=n)
    if name == "C2":
        return config2(num_tris=max(1, int(1_000_000 * scale)))
    if name == "C3":
        return config3(num_tris=max(1, int(2000 * scale)))
    if name == "C4":
        s = math.sqrt(scale)
        return config4(quads_x=max(2, int(2500 * s)), quads_z=max(2, int(2000 * s)))
    if name == "M1":
        return stress_m1(num_tris=max(1, int(1_000_000 * scale)))
    raise ValueError(name)


# ---------------------------------------------------------------------------------------------
# Textured cases (LambertianAlbedoPixelShader + filter modes). Not BASELINE configs.
# ---------------------------------------------------------------------------------------------
def noise_texture(width, height, seed=7, cell=4):
    """RGBA8 image with structure at several scales: coloured checker cells + per-texel noise (H x W x 4)."""
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:height, 0:width]
    cells = rng.integers(40, 256, ((height + cell - 1) // cell + 1, (width + cell - 1) // cell + 1, 3))
    img = cells[ys // cell, xs // cell].astype(np.int32) + rng.integers(-30, 31, (height, width, 3))
    out = np.empty((height, width, 4), np.uint8)
    out[..., :3] = np.clip(img, 0, 255)
    out[..., 3] = 255
    return out


def textured_plane(width=640, height=360, quads=24, uv_scale=6.0, tex=(128, 64), tex_filter=2, seed=11):
    """A ground plane seen at a grazing angle: minification grows towards the horizon (every mip level is used,
    footprints are strongly anisotropic) and the near edge crosses the near plane (texcoords of clipped vertices)."""
    n = quads + 1
    g = np.linspace(-12.0, 12.0, n)
    xx, zz = np.meshgrid(g, g)
    pos = np.stack([xx.ravel(), np.zeros(n * n), zz.ravel()], axis=1)
    nrm = np.tile(np.array([0.0, 1.0, 0.0]), (n * n, 1))
    uv = np.stack([xx.ravel(), zz.ravel()], axis=1) * (uv_scale / 24.0)
    idx = []
    for j in range(quads):
        for i in range(quads):
            a = j * n + i
            idx += [[a, a + n, a + 1], [a + 1, a + n, a + n + 1]]
    c = cam.Camera((0.3, 0.7, -11.0), (0.0, 0.0, 4.0), (0.0, 1.0, 0.0), width, height, 60.0, 0.5, 100.0)
    return Scene(name="textured_plane", width=width, height=height, vertices=_pack(pos, nrm, uv),
                 indices=np.array(idx, np.uint32), mv=c.view, proj=c.proj, raster=c.raster, shader=SHADER_LAMBERT_ALBEDO,
                 textures=[("image", noise_texture(tex[0], tex[1], seed))], tex_ids=None, tex_filter=tex_filter)


def textured_sphere(width=640, height=360, slices=48, stacks=48, tex_filter=2):
    """C1's sphere with three slots - a constant colour, an odd-sized image and a tiny image - assigned per triangle."""
    sc = config1(width=width, height=height, slices=slices, stacks=stacks)
    sc["shader"] = SHADER_LAMBERT_ALBEDO
    sc["textures"] = [("constant", (0.9, 0.5, 0.2)), ("image", noise_texture(37, 21, 3, cell=3)), ("image", noise_texture(2, 2, 5, cell=1))]
    sc["tex_ids"] = (np.arange(sc.num_tris, dtype=np.uint32) // 7) % 3
    sc["tex_filter"] = tex_filter
    v = sc.vertices.copy()
    v[:, 6:8] *= np.array([3.0, 2.0], np.float32)           # repeat addressing
    sc["vertices"] = v
    return sc
