"""Developer tool (GPU, under compute-sanitizer): every kernel once on small inputs, checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import usrt_oracle as O
from unitysimpleraytracing_b200 import host, meshes, _lib
tris = meshes.uniform_soup(20000, seed=3); cam = meshes.SCENE_SOUP_CAMERA
ref = O.Scene(tris)
d = host.RaytracingMeshDrawer(tris).Awake(); c = d.container.ctx
assert c.download(_lib.BUF_BVH_DATA, len(tris) - 1).tobytes() == ref.bvhData[:len(tris) - 1].tobytes()
d.Rebuild()
hits = d.Update(96, 64, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
assert hits.tobytes() == ref.trace_primary(96, 64, cam["near"], cam["tan_half_fov"], cam["cam_to_world"]).tobytes()
rays = meshes.incoherent_rays(2000, seed=1)
assert c.trace_rays(rays).tobytes() == ref.trace_rays(rays).tobytes()
c.upload_texture(np.random.default_rng(0).random((8, 8, 4), dtype=np.float32)); c.shade()
k = np.random.default_rng(1).integers(0, 2**32, 50001, dtype=np.uint64).astype(np.uint32); v = np.arange(50001, dtype=np.uint32)
o = np.argsort(k, kind="stable"); kk, vv = k.copy(), v.copy(); c.sort_pairs_host(kk, vv)
assert np.array_equal(kk, k[o]) and np.array_equal(vv, v[o])
k = np.random.default_rng(2).integers(0, 2**32, 700001, dtype=np.uint64).astype(np.uint32); v = np.arange(700001, dtype=np.uint32)
o = np.argsort(k, kind="stable"); kk, vv = k.copy(), v.copy(); c.sort_pairs_host(kk, vv)
assert np.array_equal(kk, k[o]) and np.array_equal(vv, v[o])
# kernels added later in round 1: diffuse bounce rays, hit mirrors + zero-copy frame, scene-box reduction, per-digit-address partition
import torch
W, H = 96, 64
args = (W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
prim = ref.trace_primary(*args)
c.trace_primary(*args, download=False)
rays_d = torch.zeros(3 * W * H * 8, dtype=torch.float32, device="cuda")
c.diffuse_rays_device(*args, 5, 1, 3, rays_d.data_ptr()); c.sync()
want = O.diffuse_rays(prim, tris, *args, 5, 1, 3)
assert rays_d.cpu().numpy().tobytes() == want.tobytes()
assert c.trace_rays(want).tobytes() == ref.trace_rays(want).tobytes()
mir = [torch.zeros(W * H * 4, dtype=torch.float32, device="cuda") for _ in range(8)]
c.set_hit_mirrors([m_.data_ptr() for m_ in mir])
pin = torch.zeros(W * H * 16, dtype=torch.uint8).pin_memory()
c.trace_primary_async(*args, pin.numpy().view(prim.dtype)); c.sync()
assert pin.numpy().tobytes() == prim.tobytes() and all(m_.cpu().numpy().tobytes() == prim.tobytes() for m_ in mir)
c.set_hit_mirrors([])
lo, hi = c.fit_world_box(); rlo, rhi = O.scene_box(tris)
assert np.array_equal(lo, rlo) and np.array_equal(hi, rhi)
c.rebuild(); assert np.array_equal(c.download(_lib.BUF_KEYS), O.Scene(tris, rlo, rhi).sortedMortonCodes)
for n_ in (3001, 300001):                                  # small-tile and big-tile instantiations of the peer scatter
    k = np.random.default_rng(n_).integers(0, 2**32, n_, dtype=np.uint64).astype(np.uint32); v = np.arange(n_, dtype=np.uint32)
    kd, vd = torch.from_numpy(k.view(np.int32)).cuda(), torch.from_numpy(v.view(np.int32)).cuda()
    hist = torch.zeros(256, dtype=torch.int32, device="cuda")
    c.digit_histogram_device(kd.data_ptr(), n_, 24, hist.data_ptr()); c.sync()
    h_ = hist.cpu().numpy().astype(np.int64); assert np.array_equal(h_, np.bincount(k >> 24, minlength=256))
    ok_, ov_ = torch.zeros_like(kd), torch.zeros_like(vd)
    base = np.concatenate([[0], np.cumsum(h_)[:-1]]) * 4
    kp = torch.from_numpy((ok_.data_ptr() + base).astype(np.int64)).cuda(); vp = torch.from_numpy((ov_.data_ptr() + base).astype(np.int64)).cuda()
    c.partition_scatter_device(kd.data_ptr(), vd.data_ptr(), n_, 24, kp.data_ptr(), vp.data_ptr()); c.sync()
    o = np.argsort(k >> 24, kind="stable")
    assert np.array_equal(ok_.cpu().numpy().view(np.uint32), k[o]) and np.array_equal(ov_.cpu().numpy().view(np.uint32), v[o])
# round 2: persistent sort CTAs -- more tiles than resident CTAs (444 x 6144 / 296 x 4096 pairs), so every CTA loops
# (tile tickets, shared-memory reuse from one tile to the next), 32- and 64-bit keys
k = np.random.default_rng(4).integers(0, 2**32, 3000001, dtype=np.uint64).astype(np.uint32); v = np.arange(3000001, dtype=np.uint32)
o = np.argsort(k, kind="stable"); kk, vv = k.copy(), v.copy(); c.sort_pairs_host(kk, vv)
assert np.array_equal(kk, k[o]) and np.array_equal(vv, v[o])
k64 = np.random.default_rng(5).integers(0, 2**63, 1500001, dtype=np.uint64); v = np.arange(1500001, dtype=np.uint32)
o = np.argsort(k64, kind="stable"); kk, vv = k64.copy(), v.copy(); c.sort_pairs64_host(kk, vv)
assert np.array_equal(kk, k64[o]) and np.array_equal(vv, v[o])
print("sanitize workload ok")
