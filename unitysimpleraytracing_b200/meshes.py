"""Seeded synthetic meshes, cameras and ray sets (host side, numpy only).

These are the inputs of BASELINE.json's configs: uniform triangle soups,
tessellated spheres and height-field grids, all inside the reference's fixed +-125 world box
(Assets/_Scripts/MeshBufferContainer.cs:9-15). They are generated ONCE on the host and fed
identically to the CUDA path and to the test oracle, so nothing here needs to be bit-reproducible
on the GPU; a counter-based integer hash keeps them reproducible across numpy versions.

The output is the reference's 128-byte `Triangle` (MeshBufferContainer.cs:133-144): positions,
per-vertex UVs and normals.
"""
import numpy as np

from .scene_types import Triangle

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def hash_u64(x):
    """splitmix64 finaliser on a uint64 array."""
    x = np.asarray(x, np.uint64).copy()
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def uniform01(seed, stream, count):
    """count floats in [0,1) with 24 random bits each (exact in fp32)."""
    with np.errstate(over="ignore"):
        idx = np.arange(count, dtype=np.uint64) + (np.uint64(seed) << np.uint64(32)) * np.uint64(stream + 1)
        h = hash_u64(idx ^ hash_u64(np.uint64(seed) + np.uint64(stream) * np.uint64(0x632BE59BD9B4E019)))
    return ((h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def random_u32(seed, stream, count):
    with np.errstate(over="ignore"):
        idx = np.arange(count, dtype=np.uint64) + (np.uint64(seed) << np.uint64(32)) * np.uint64(stream + 1)
        h = hash_u64(idx ^ hash_u64(np.uint64(seed) + np.uint64(stream) * np.uint64(0x632BE59BD9B4E019)))
    return (h >> np.uint64(32)).astype(np.uint32)


def pack_triangles(a, b, c, uv=None, normals=None):
    """(n,3) float32 vertex arrays -> Triangle[n] (pads zero like C# default(struct))."""
    n = len(a)
    t = np.zeros(n, Triangle)
    t["a"], t["b"], t["c"] = a, b, c
    if uv is not None:
        t["a_uv"], t["b_uv"], t["c_uv"] = uv
    if normals is None:
        e1 = (b - a).astype(np.float64); e2 = (c - a).astype(np.float64)
        nn = np.cross(e1, e2)
        ln = np.linalg.norm(nn, axis=1, keepdims=True)
        nn = np.where(ln > 0, nn / np.maximum(ln, 1e-300), 0.0).astype(np.float32)
        normals = (nn, nn, nn)
    t["a_normal"], t["b_normal"], t["c_normal"] = normals
    return t


def uniform_soup(n, seed=0x5EED0001, extent=100.0, edge=None):
    """Triangle soup: centre ~ U[-extent,extent]^3, vertices = centre + U[-s,s]^3."""
    if edge is None:
        edge = 2.0 * extent / max(1.0, float(n) ** (1.0 / 3.0))
    s = np.float32(edge * 0.75)
    u = [uniform01(seed, k, n) for k in range(12)]
    centre = np.stack([(u[k] * np.float32(2) - np.float32(1)) * np.float32(extent) for k in range(3)], 1)
    verts = []
    for v in range(3):
        off = np.stack([(u[3 + 3 * v + k] * np.float32(2) - np.float32(1)) * s for k in range(3)], 1)
        verts.append((centre + off).astype(np.float32))
    uv = tuple(np.stack([u[3 + v], u[6 + v]], 1) for v in range(3))
    return pack_triangles(verts[0], verts[1], verts[2], uv=uv)


def _grid_triangles(P, N=None, UV=None):
    """P: (rows+1, cols+1, 3) vertex grid -> 2*rows*cols triangles, row-major quads, fixed diagonal."""
    p00, p01, p10, p11 = P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]
    a = np.stack([p00, p11], 2).reshape(-1, 3)
    b = np.stack([p01, p10], 2).reshape(-1, 3)   # tri0: p00,p01,p11 ; tri1: p11,p10,p00
    c = np.stack([p11, p00], 2).reshape(-1, 3)
    out = {}
    for name, G in (("n", N), ("uv", UV)):
        if G is None:
            out[name] = None
            continue
        g00, g01, g10, g11 = G[:-1, :-1], G[:-1, 1:], G[1:, :-1], G[1:, 1:]
        k = G.shape[-1]
        out[name] = (np.stack([g00, g11], 2).reshape(-1, k), np.stack([g01, g10], 2).reshape(-1, k),
                     np.stack([g11, g00], 2).reshape(-1, k))
    return pack_triangles(a.astype(np.float32), b.astype(np.float32), c.astype(np.float32), uv=out["uv"],
                          normals=out["n"])


def sphere(n_lat, n_lon, radius=35.0, centre=(0.0, 0.0, 35.0)):
    """Lat-long tessellated sphere: 2*n_lat*n_lon triangles (pole rows hold zero-area triangles,
    which the reference's |det| < 1e-8 test rejects -- Raytracing.compute:47)."""
    th = np.linspace(0.0, np.pi, n_lat + 1)[:, None]
    ph = np.linspace(0.0, 2.0 * np.pi, n_lon + 1)[None, :]
    nrm = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th) * np.ones_like(ph)], 2)
    nrm[:, -1] = nrm[:, 0]                      # close the seam exactly
    P = (nrm * radius + np.asarray(centre)[None, None, :]).astype(np.float32)
    UV = np.stack([np.broadcast_to(ph / (2 * np.pi), nrm.shape[:2]), np.broadcast_to(th / np.pi, nrm.shape[:2])], 2)
    return _grid_triangles(P, nrm.astype(np.float32), UV.astype(np.float32))


def _value_noise(x, y, seed, cells):
    xi = np.floor(x * cells).astype(np.int64); yi = np.floor(y * cells).astype(np.int64)
    fx = x * cells - xi; fy = y * cells - yi

    def lattice(ix, iy):
        h = hash_u64((ix.astype(np.uint64) << np.uint64(32)) ^ iy.astype(np.uint64) ^ (np.uint64(seed) << np.uint64(48)))
        return (h >> np.uint64(40)).astype(np.float64) * 2.0 ** -24

    sx = fx * fx * (3 - 2 * fx); sy = fy * fy * (3 - 2 * fy)
    v00, v10, v01, v11 = lattice(xi, yi), lattice(xi + 1, yi), lattice(xi, yi + 1), lattice(xi + 1, yi + 1)
    return (v00 * (1 - sx) + v10 * sx) * (1 - sy) + (v01 * (1 - sx) + v11 * sx) * sy


def height_field(m, half_extent=120.0, amplitude=12.0, base=-15.0, seed=0x5EED0002, octaves=4):
    """(m+1)^2 vertices on the x,y plane, z = hashed value-noise octaves; 2*m*m triangles. The
    reference's scene mesh ExampleObject3.obj is the m=80, amplitude=0 instance (x,y in [-4,4], z=0)."""
    g = np.linspace(0.0, 1.0, m + 1)
    X, Y = np.meshgrid(g, g, indexing="xy")
    Z = np.zeros_like(X)
    amp = 1.0
    for o in range(octaves):
        Z += amp * (_value_noise(X, Y, seed + o, 4 * 2 ** o) - 0.5)
        amp *= 0.5
    P = np.stack([(X * 2 - 1) * half_extent, (Y * 2 - 1) * half_extent, base + amplitude * Z], 2).astype(np.float32)
    UV = np.stack([X, Y], 2).astype(np.float32)
    return _grid_triangles(P, None, UV)


def reference_scene_grid():
    """The reference's shipped scene mesh restated: 80x80 quads, z=0, x,y in [-4,4] -> 12,800 tris
    (Assets/__Scenes/Scene.unity:364 -> ExampleObject3.obj). ~11 triangles per Morton cell, which is
    the duplicate-key case DistributeKeys exists for."""
    return height_field(80, half_extent=4.0, amplitude=0.0, base=0.0)


def scene_c1(n=65536):
    """BASELINE config 1: 65,536-triangle uniform soup."""
    return uniform_soup(n, seed=0x5EED0001)


def scene_c2(m_sphere=512, m_grid=512):
    """BASELINE config 2: tessellated sphere (2*512*512) + height field (2*512*512) = 1,048,576 tris."""
    return np.concatenate([sphere(m_sphere, m_sphere), height_field(m_grid)])


# ---------------------------------------------------------------------------------------------
# cameras and rays (Raytracing.compute:108-126, RaytracingMeshDrawer.cs:78-81, Scene.unity:315-343)
# ---------------------------------------------------------------------------------------------
DEG2RAD = np.float32(0.0174532924)   # UnityEngine.Mathf.Deg2Rad


def tan_half_fov(fov_deg=60.0):
    """RaytracingMeshDrawer.cs:80: Mathf.Tan(fieldOfView * Mathf.Deg2Rad / 2)."""
    return np.float32(np.tan(np.float64(np.float32(fov_deg) * DEG2RAD / np.float32(2))))


def camera_to_world(position=(0.0, 0.0, 15.7)):
    """Row-major 4x4 of the reference scene's camera: yaw 180 deg about Y (Scene.unity:342) and Unity's
    camera space looking down -Z, i.e. cameraToWorld = T * R_y(180) * diag(1,1,-1)."""
    m = np.array([[-1, 0, 0, position[0]], [0, 1, 0, position[1]], [0, 0, 1, position[2]], [0, 0, 0, 1]],
                 np.float32)
    return m


REFERENCE_CAMERA = dict(near=np.float32(0.3), tan_half_fov=tan_half_fov(60.0), cam_to_world=camera_to_world())
SCENE_C2_CAMERA = dict(near=np.float32(0.3), tan_half_fov=tan_half_fov(60.0),
                       cam_to_world=camera_to_world((0.0, 0.0, 120.0)))
SCENE_SOUP_CAMERA = dict(near=np.float32(0.3), tan_half_fov=tan_half_fov(60.0),
                         cam_to_world=camera_to_world((0.0, 0.0, 124.0)))


def incoherent_rays(count, seed=0x5EED0003, extent=100.0):
    """Origins uniform in the scene box, directions uniform on the sphere by hashed rejection
    sampling (trig-free), normalised in fp64 then rounded: rays are INPUTS, fed identically to both
    sides. Layout: (count, 8) float32 = origin.xyz, 0, dir.xyz, 0 (the ray-buffer ABI)."""
    rays = np.zeros((count, 8), np.float32)
    for k in range(3):
        rays[:, k] = (uniform01(seed, k, count) * np.float32(2) - np.float32(1)) * np.float32(extent)
    d = np.zeros((count, 3), np.float64)
    todo = np.arange(count)
    rnd = 0
    while len(todo):
        c = np.stack([uniform01(seed + 17 + rnd, 3 + k, count)[todo].astype(np.float64) * 2 - 1 for k in range(3)], 1)
        l2 = (c * c).sum(1)
        ok = (l2 <= 1.0) & (l2 > 1e-4)
        d[todo[ok]] = c[ok] / np.sqrt(l2[ok])[:, None]
        todo = todo[~ok]
        rnd += 1
    rays[:, 4:7] = d.astype(np.float32)
    return rays
