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
print("sanitize workload ok")
