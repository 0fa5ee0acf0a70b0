"""One-off evidence run (GPU): BASELINE config 4 at full size -- 16,777,216-triangle soup, every build buffer
compared with the oracle by SHA-256, plus incoherent rays and a primary frame compared record by record."""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import usrt_oracle as O
from unitysimpleraytracing_b200 import host, meshes, _lib
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
t0 = time.time(); tris = meshes.uniform_soup(n, seed=0x5EED0004); print("generated %d triangles in %.1fs" % (n, time.time() - t0), flush=True)
tm = {}; t0 = time.time(); ref = O.Scene(tris, timings=tm); print("oracle build %.1fs %s" % (time.time() - t0, {k: round(v, 2) for k, v in tm.items()}), flush=True)
ctx = host.Context(n); ctx.upload_triangles(tris); ctx.enable_stage_timing(True); ctx.rebuild(); ctx.rebuild()
print("gpu rebuild ms", {k: round(v, 3) for k, v in ctx.last_rebuild_ms().items()}, flush=True)
ok = True
for name, buf, want, cnt in (("distributed keys", _lib.BUF_KEYS, ref.sortedMortonCodes, n), ("sorted indices", _lib.BUF_TRIANGLE_INDEX, ref.sortedTriangleIndices, n),
                             ("triangle AABBs", _lib.BUF_TRIANGLE_AABB, ref.triangleAABB, n), ("internal nodes", _lib.BUF_INTERNAL_NODES, ref.internalNodes[:n - 1], n - 1),
                             ("leaf nodes", _lib.BUF_LEAF_NODES, ref.leafNodes, n), ("node AABBs", _lib.BUF_BVH_DATA, ref.bvhData[:n - 1], n - 1)):
    a, b = sha(ctx.download(buf, cnt)), sha(want)
    print("  %-18s gpu %s oracle %s %s" % (name, a, b, "OK" if a == b else "MISMATCH"), flush=True); ok &= a == b
print("  corrupted nodes:", ctx.count_corrupted_nodes())
rays = meshes.incoherent_rays(4000, seed=0x5EED0005)
t0 = time.time(); want = ref.trace_rays(rays, threads=os.cpu_count()); print("oracle 4000 incoherent rays %.1fs" % (time.time() - t0), flush=True)
got = ctx.trace_rays(rays); same = got.tobytes() == want.tobytes(); ok &= same
print("  incoherent rays: %s (hits %d)" % ("bit-exact" if same else "MISMATCH", int((want["distance"] != O.max_float()).sum())))
cam = meshes.SCENE_SOUP_CAMERA
want = ref.trace_primary(64, 36, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], threads=os.cpu_count())
got = ctx.trace_primary(64, 36, cam["near"], cam["tan_half_fov"], cam["cam_to_world"]); same = got.tobytes() == want.tobytes(); ok &= same
print("  64x36 primary frame: %s" % ("bit-exact" if same else "MISMATCH"))
print("C4 PARITY", "OK" if ok else "FAILED")
