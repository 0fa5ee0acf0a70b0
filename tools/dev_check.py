"""Developer diagnostic (GPU): stage-by-stage comparison against the oracle with detailed output."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import usrt_oracle as O
from unitysimpleraytracing_b200 import host, meshes, _lib

def cmp(name, a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    ok = a.tobytes() == b.tobytes()
    if not ok:
        au = a.view(np.uint32).reshape(len(a), -1); bu = b.view(np.uint32).reshape(len(b), -1)
        bad = np.nonzero((au != bu).any(1))[0]
        print("  MISMATCH %s: %d / %d rows differ; first %s" % (name, len(bad), len(a), bad[:5]))
        for i in bad[:3]:
            print("    row", i, "got", a[i], "want", b[i])
    else:
        print("  ok", name)
    return ok

def check_scene(name, tris, cam, W, H):
    print("== scene", name, "n =", len(tris))
    ref = O.Scene(tris)
    n = len(tris)
    ctx = host.Context(n)
    ctx.upload_triangles(tris)
    ctx.morton(); ctx.sync()
    ok = cmp("morton keys", ctx.download(_lib.BUF_KEYS), ref.mortonCodes)
    ok &= cmp("tri aabb", ctx.download(_lib.BUF_TRIANGLE_AABB), ref.triangleAABB)
    ctx.sort(); ctx.sync()
    ok &= cmp("sorted keys", ctx.download(_lib.BUF_KEYS), ref.sortedMortonRaw)
    ok &= cmp("sorted idx", ctx.download(_lib.BUF_TRIANGLE_INDEX), ref.sortedTriangleIndices)
    ctx.distribute_keys(); ctx.sync()
    ok &= cmp("distributed keys", ctx.download(_lib.BUF_KEYS), ref.sortedMortonCodes)
    ctx.construct_tree(); ctx.sync()
    ok &= cmp("internal", ctx.download(_lib.BUF_INTERNAL_NODES, n - 1), ref.internalNodes[:n - 1])
    ok &= cmp("leaf", ctx.download(_lib.BUF_LEAF_NODES), ref.leafNodes)
    ctx.construct_bvh(); ctx.sync()
    ok &= cmp("bvh", ctx.download(_lib.BUF_BVH_DATA, n - 1), ref.bvhData[:n - 1])
    print("  corrupted:", ctx.count_corrupted_nodes())
    hits = ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    t0 = time.perf_counter()
    want = ref.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], threads=os.cpu_count())
    print("  oracle trace %.2fs" % (time.perf_counter() - t0))
    ok &= cmp("hits", hits, want)
    print("  hit fraction", float((want["distance"] != O.max_float()).mean()))
    # rebuild (fused) + timing
    ctx.enable_stage_timing(True)
    for _ in range(3):
        ctx.rebuild()
    print("  rebuild ms:", {k: round(v, 4) for k, v in ctx.last_rebuild_ms().items()})
    ok &= cmp("bvh after rebuild", ctx.download(_lib.BUF_BVH_DATA, n - 1), ref.bvhData[:n - 1])
    ok &= cmp("internal after rebuild", ctx.download(_lib.BUF_INTERNAL_NODES, n - 1), ref.internalNodes[:n - 1])
    import torch
    for mode in (0, 1):
        ctx.set_trace_mode(mode)
        for Wb, Hb in ((W, H), (1920, 1080)):
            ctx.trace_primary(Wb, Hb, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=False); ctx.sync()
            t0 = time.perf_counter()
            for _ in range(3):
                ctx.trace_primary(Wb, Hb, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=False)
            ctx.sync()
            dt = (time.perf_counter() - t0) / 3
            print("  trace mode %d %dx%d: %.3f ms  %.1f Mrays/s" % (mode, Wb, Hb, dt * 1e3, Wb * Hb / dt / 1e6))
    ctx.set_trace_mode(1)
    hc = ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    print("  culled-vs-strict mismatches:", int((hc.view(np.uint32).reshape(-1, 4) != want.view(np.uint32).reshape(-1, 4)).any(1).sum()))
    ctx.close()
    return ok

def check_sort(n, kind):
    rng = np.random.default_rng(n + len(kind))
    if kind == "uniform": k = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    elif kind == "morton30": k = rng.integers(0, 2**30, n, dtype=np.uint64).astype(np.uint32)
    elif kind == "low": k = rng.integers(0, 7, n, dtype=np.uint64).astype(np.uint32) * np.uint32(0x01010101)
    elif kind == "equal": k = np.full(n, 0xDEADBEEF, np.uint32)
    v = np.arange(n, dtype=np.uint32)
    order = np.argsort(k, kind="stable")
    wk, wv = k[order], v[order]
    ctx = host.Context(2)
    gk, gv = k.copy(), v.copy()
    t0 = time.perf_counter()
    ctx.sort_pairs_host(gk, gv)
    dt = time.perf_counter() - t0
    ok = np.array_equal(gk, wk) and np.array_equal(gv, wv)
    print("sort n=%d %s: %s (%.1f ms e2e)" % (n, kind, "ok" if ok else "MISMATCH", dt * 1e3))
    if not ok:
        bad = np.nonzero((gk != wk) | (gv != wv))[0]
        print("   first bad", bad[:5], gk[bad[:3]], wk[bad[:3]], gv[bad[:3]], wv[bad[:3]])
    ctx.close()
    return ok

if __name__ == "__main__":
    ok = True
    for n in (1, 2, 31, 4096, 4097, 100003, 1 << 20):
        for kind in ("uniform", "morton30", "low", "equal"):
            ok &= check_sort(n, kind)
    ok &= check_scene("soup4k", meshes.uniform_soup(4096, seed=7), meshes.SCENE_SOUP_CAMERA, 128, 128)
    ok &= check_scene("refgrid", meshes.reference_scene_grid(), meshes.REFERENCE_CAMERA, 128, 128)
    ok &= check_scene("c1", meshes.scene_c1(), meshes.SCENE_SOUP_CAMERA, 256, 256)
    if "--big" in sys.argv:
        ok &= check_scene("c2", meshes.scene_c2(), meshes.SCENE_C2_CAMERA, 480, 270)
    print("ALL OK" if ok else "FAILURES")
    sys.exit(0 if ok else 1)
