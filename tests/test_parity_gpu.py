"""GPU parity tests proper: every stage of the CUDA path, called through the C ABI, against the
oracle on the same seeded inputs. Bit-exact for keys, indices, topology and boxes; hit records are
compared bit-for-bit as well (the north_star tolerance is 1e-5 relative on distance -- the arithmetic
is defined op-for-op on both sides, so exact equality is what we assert)."""
import numpy as np
import pytest

from unitysimpleraytracing_b200 import _lib, meshes

pytestmark = pytest.mark.gpu


def _mesh(name):
    if name == "soup2":
        return meshes.uniform_soup(2, seed=41)
    if name == "soup3":
        return meshes.uniform_soup(3, seed=42)
    if name == "soup4097":
        return meshes.uniform_soup(4097, seed=43)
    if name == "refgrid":
        return meshes.reference_scene_grid()
    if name == "sphere":
        return meshes.sphere(96, 192)
    if name == "c1":
        return meshes.scene_c1()
    if name == "outside":               # centroids beyond the +-125 world box clamp to 0 / 1023
        return meshes.uniform_soup(20000, seed=44, extent=190.0)
    if name == "identical":             # every key equal: DistributeKeys spreads them 0,1,2,...
        t = meshes.uniform_soup(1, seed=45)
        return np.repeat(t, 3001)
    if name == "degenerate":            # zero-area and point triangles mixed in
        t = meshes.uniform_soup(5000, seed=46)
        t["b"][::3] = t["a"][::3]
        t["c"][::7] = t["a"][::7]
        return t
    raise ValueError(name)


MESHES = ["soup2", "soup3", "soup4097", "refgrid", "sphere", "c1", "outside", "identical", "degenerate"]


def _same(a, b):
    return np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes()


@pytest.mark.parametrize("name", MESHES)
def test_build_stage_by_stage(usrt, oracle, name):
    """The reference's Awake() sequence (RaytracingMeshDrawer.cs:34-51), one entry point at a time."""
    tris = _mesh(name)
    n = len(tris)
    ref = oracle.Scene(tris)
    c = usrt.MeshBufferContainer(tris, capacity=n + 1000)        # capacity > n: padding slots stay 0xFFFFFFFF
    ctx = c.ctx
    assert c.TrianglesLength == n
    assert np.array_equal(c.Keys.GetData(n), ref.mortonCodes)
    assert np.array_equal(c.TriangleIndex.GetData(n), np.arange(n, dtype=np.uint32))
    assert _same(c.TriangleAABB.GetData(n), ref.triangleAABB)

    usrt.ComputeBufferSorter(c.TrianglesLength, c.Keys, c.TriangleIndex).Sort()
    assert np.array_equal(c.Keys.GetData(n), ref.sortedMortonRaw)
    assert np.array_equal(c.TriangleIndex.GetData(n), ref.sortedTriangleIndices)
    full_keys = c.Keys.GetData(n + 1000)
    assert (full_keys[n:] == 0xFFFFFFFF).all() and (c.TriangleIndex.GetData(n + 1000)[n:] == 0xFFFFFFFF).all()

    c.DistributeKeys()
    assert np.array_equal(c.Keys.GetData(n), ref.sortedMortonCodes)
    assert (c.Keys.GetData(n + 1000)[n:] == 0xFFFFFFFF).all()

    b = usrt.BVHConstructor(c.TrianglesLength, c.Keys, c.TriangleIndex, c.TriangleAABB, c.BvhInternalNode,
                            c.BvhLeafNode, c.BvhData)
    b.ConstructTree()
    assert _same(c.BvhInternalNode.GetData(n - 1), ref.internalNodes[:n - 1])
    assert _same(c.BvhLeafNode.GetData(n), ref.leafNodes)
    # untouched slots keep NullLeaf (MeshBufferContainer.cs:114-115)
    assert (c.BvhInternalNode.GetData(n + 10).view(np.uint32).reshape(-1, 6)[n - 1:] == 0xFFFFFFFF).all()
    assert (c.BvhLeafNode.GetData(n + 10).view(np.uint32).reshape(-1, 2)[n:] == 0xFFFFFFFF).all()
    b.ConstructBVH()
    assert _same(c.BvhData.GetData(n - 1), ref.bvhData[:n - 1])
    assert c.GetAllGpuData() == (0, 0)                          # MeshBufferContainer.cs:181-195
    assert ctx.count_corrupted_nodes() == (0, 0)
    b.ConstructBVH()                                             # re-runnable (self-resetting counters)
    assert _same(c.BvhData.GetData(n - 1), ref.bvhData[:n - 1])
    c.Dispose()


@pytest.mark.parametrize("name", ["soup4097", "refgrid", "c1", "identical"])
def test_fused_rebuild_equals_stages_and_is_repeatable(usrt, oracle, name):
    tris = _mesh(name)
    n = len(tris)
    ref = oracle.Scene(tris)
    ctx = usrt.Context(n)
    ctx.upload_triangles(tris)
    for _ in range(3):
        ctx.rebuild()
        assert np.array_equal(ctx.download(_lib.BUF_KEYS), ref.sortedMortonCodes)
        assert np.array_equal(ctx.download(_lib.BUF_TRIANGLE_INDEX), ref.sortedTriangleIndices)
        assert _same(ctx.download(_lib.BUF_INTERNAL_NODES, n - 1), ref.internalNodes[:n - 1])
        assert _same(ctx.download(_lib.BUF_LEAF_NODES), ref.leafNodes)
        assert _same(ctx.download(_lib.BUF_BVH_DATA, n - 1), ref.bvhData[:n - 1])
    # a smaller mesh in the same context: stale nodes must not leak
    small = meshes.uniform_soup(max(n // 3, 2), seed=77)
    ref2 = oracle.Scene(small)
    ctx.upload_triangles(small)
    ctx.rebuild()
    m = len(small)
    assert _same(ctx.download(_lib.BUF_INTERNAL_NODES, m - 1), ref2.internalNodes[:m - 1])
    assert _same(ctx.download(_lib.BUF_BVH_DATA, m - 1), ref2.bvhData[:m - 1])
    assert (ctx.download(_lib.BUF_INTERNAL_NODES, n).view(np.uint32).reshape(-1, 6)[m - 1:] == 0xFFFFFFFF).all()
    ctx.close()


def _report_ties(got, want):
    """IDs must be exact; where they differ but the distances are equal it is an edge/vertex tie,
    counted separately (north_star). Returns (id_mismatches_not_ties, ties)."""
    bad = got["triangleIndex"] != want["triangleIndex"]
    ties = bad & (got["distance"] == want["distance"])
    return int((bad & ~ties).sum()), int(ties.sum())


@pytest.mark.parametrize("name,cam,w,h", [
    ("soup4097", "SCENE_SOUP_CAMERA", 160, 90), ("refgrid", "REFERENCE_CAMERA", 160, 90),
    ("sphere", "SCENE_C2_CAMERA", 192, 108), ("c1", "SCENE_SOUP_CAMERA", 128, 128),
    ("degenerate", "SCENE_SOUP_CAMERA", 96, 96), ("identical", "SCENE_SOUP_CAMERA", 33, 17),
    ("soup2", "SCENE_SOUP_CAMERA", 31, 9)])
def test_primary_rays(usrt, oracle, name, cam, w, h):
    tris = _mesh(name)
    cam = getattr(meshes, cam)
    ref = oracle.Scene(tris)
    want = ref.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], threads=8)
    d = usrt.RaytracingMeshDrawer(tris).Awake()
    got = d.Update(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    assert _report_ties(got, want) == (0, 0)
    assert np.array_equal(got["triangleIndex"], want["triangleIndex"])
    hit = want["distance"] != oracle.max_float()
    rel = np.abs(got["distance"][hit] - want["distance"][hit]) / np.maximum(np.abs(want["distance"][hit]), 1e-30)
    assert (rel <= 1e-5).all()                                   # the north_star bar ...
    assert _same(got, want)                                      # ... and what actually holds: bit-exact
    assert np.array_equal(got["distance"][~hit].view(np.uint32), np.full((~hit).sum(), 0x4EFF0000, np.uint32))
    # row sharding hook: two half-frames == the full frame
    ctx = d.container.ctx
    out = np.zeros(w * h, got.dtype)
    ctx.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], 0, h // 2, out=out)
    ctx.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], h // 2, h, out=out)
    assert _same(out, want)
    d.OnDestroy()


@pytest.mark.parametrize("name,extent,count", [("c1", 100.0, 20000), ("refgrid", 4.0, 5000), ("sphere", 60.0, 5000)])
def test_incoherent_rays(usrt, oracle, name, extent, count):
    """Random-origin random-direction rays (config 4's stress shape): origins inside the geometry, hits
    behind the origin (negative t is accepted by the reference), axis-aligned directions (inv_dir = inf)."""
    tris = _mesh(name)
    ref = oracle.Scene(tris)
    rays = meshes.incoherent_rays(count, seed=9, extent=extent)
    rays[:50, 4:7] = [0, 0, -1]           # inv_dir = (inf, inf, -1): the 0*inf = NaN slab cases
    rays[50:100, 4:7] = [1, 0, 0]
    rays[100:120, 0:3] = tris["a"][:20]   # origins exactly on vertices
    want = ref.trace_rays(rays, threads=8)
    d = usrt.RaytracingMeshDrawer(tris).Awake()
    got = d.container.ctx.trace_rays(rays)
    assert _report_ties(got, want) == (0, 0)
    assert _same(got, want)
    assert (want["distance"] != oracle.max_float()).sum() > 10
    d.OnDestroy()


@pytest.mark.parametrize("name,cam", [("c1", "SCENE_SOUP_CAMERA"), ("sphere", "REFERENCE_CAMERA")])
def test_diffuse_bounce_rays(usrt, oracle, name, cam):
    """BASELINE configs[4] ("64 spp random diffuse rays"): the bounce rays generated on the device from the primary
    hit records are bit-identical to the oracle's, for any sample window, and so are their hit records."""
    import torch
    tris = _mesh(name) if name != "sphere" else meshes.sphere(48, 96)
    cam = getattr(meshes, cam)
    w, h, seed = 96, 54, 0x5EED0007
    args = (w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    ref = oracle.Scene(tris)
    hits = ref.trace_primary(*args, threads=8)
    assert 0 < (hits["distance"] != oracle.max_float()).sum()
    want_rays = oracle.diffuse_rays(hits, tris, *args, seed, 0, 5)
    d = usrt.RaytracingMeshDrawer(tris).Awake()
    ctx = d.container.ctx
    got_hits = d.Update(*args)
    assert _same(got_hits, hits)
    rays = torch.zeros(5 * w * h * 8, dtype=torch.float32, device="cuda")
    ctx.diffuse_rays_device(*args, seed, 0, 5, rays.data_ptr())          # primary hits = the context's last trace
    ctx.sync()
    assert rays.cpu().numpy().tobytes() == want_rays.tobytes()
    # a window of samples, from an explicit copy of the primary hit records
    hits_dev = torch.from_numpy(hits.view(np.float32).reshape(-1).copy()).cuda()
    win = torch.zeros(2 * w * h * 8, dtype=torch.float32, device="cuda")
    ctx.diffuse_rays_device(*args, seed, 3, 2, win.data_ptr(), primary_hits_ptr=hits_dev.data_ptr())
    ctx.sync()
    assert win.cpu().numpy().tobytes() == want_rays[3 * w * h:].tobytes()
    # unit directions in the hemisphere of the (ray-facing) normal, null rays exactly where the pixel missed
    wr = want_rays.reshape(5, w * h, 8)
    miss = hits["distance"] == oracle.max_float()
    assert not wr[:, miss].any() and (np.abs(np.linalg.norm(wr[:, ~miss, 4:7], axis=2) - 1) < 1e-6).all()
    assert len(np.unique(wr[:, ~miss, 4:7].reshape(-1, 3), axis=0)) > 0.99 * 5 * (~miss).sum()
    # the bounce itself
    want = ref.trace_rays(want_rays, threads=8)
    out = torch.zeros(5 * w * h * 4, dtype=torch.float32, device="cuda")
    ctx.trace_rays_device(rays.data_ptr(), 5 * w * h, out.data_ptr())
    ctx.sync()
    got = out.cpu().numpy().view(want.dtype)
    assert _report_ties(got, want) == (0, 0)
    assert _same(got, want)
    assert (want["distance"][np.tile(miss, 5)] == oracle.max_float()).all()      # null rays hit nothing
    d.OnDestroy()


@pytest.mark.parametrize("name", ["soup4097", "refgrid", "c1"])
def test_fitted_world_box_opt_in(usrt, oracle, name):
    """usrt_fit_world_box (the reference's runtime-scene-AABB TODO, opt-in): the device reduction equals the oracle's
    box (flat axes padded), and a rebuild with it is bit-identical to the oracle built with the same box."""
    tris = _mesh(name)
    lo, hi = oracle.scene_box(tris)
    ctx = usrt.Context(len(tris))
    ctx.upload_triangles(tris)
    got_lo, got_hi = ctx.fit_world_box()
    assert np.array_equal(got_lo, lo) and np.array_equal(got_hi, hi)
    ctx.rebuild()
    ref = oracle.Scene(tris, lo, hi)
    n = len(tris)
    assert np.array_equal(ctx.download(_lib.BUF_KEYS), ref.sortedMortonCodes)
    assert np.array_equal(ctx.download(_lib.BUF_TRIANGLE_INDEX), ref.sortedTriangleIndices)
    assert ctx.download(_lib.BUF_INTERNAL_NODES, n - 1).tobytes() == ref.internalNodes[:n - 1].tobytes()
    assert ctx.download(_lib.BUF_BVH_DATA, n - 1).tobytes() == ref.bvhData[:n - 1].tobytes()
    assert ctx.count_corrupted_nodes() == (0, 0)
    cam = meshes.SCENE_SOUP_CAMERA if name != "refgrid" else meshes.REFERENCE_CAMERA
    want = ref.trace_primary(64, 36, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], threads=8)
    assert _same(ctx.trace_primary(64, 36, cam["near"], cam["tan_half_fov"], cam["cam_to_world"]), want)
    # an explicit per-axis box, then back to the reference's cube: keys return to the default ones
    ctx.set_world_box(lo - 1, hi + 2)
    ctx.rebuild()
    assert np.array_equal(ctx.download(_lib.BUF_KEYS), oracle.Scene(tris, lo - 1, hi + 2).sortedMortonCodes)
    ctx.set_world_bounds(-125.0, 125.0)
    ctx.rebuild()
    assert np.array_equal(ctx.download(_lib.BUF_KEYS), oracle.Scene(tris).sortedMortonCodes)
    with pytest.raises(_lib.UsrtError):
        ctx.set_world_box([0, 0, 0], [1, 0, 1])
    ctx.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_culled_modes_are_reported_separately(usrt, oracle, mode):
    """Modes 1 (distance-culled) and 2 (culled, near child first) are NOT part of the parity contract;
    they must still find a hit wherever strict does, at the same distance up to ties. We only bound how far the
    triangle ids are from strict (equally distant triangles may resolve differently)."""
    tris = _mesh("c1"); cam = meshes.SCENE_SOUP_CAMERA
    d = usrt.RaytracingMeshDrawer(tris).Awake()
    strict = d.Update(128, 128, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    d.container.ctx.set_trace_mode(mode)
    culled = d.Update(128, 128, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    d.container.ctx.set_trace_mode(0)
    assert np.array_equal(strict["distance"] == oracle.max_float(), culled["distance"] == oracle.max_float())
    differing = int((strict["triangleIndex"] != culled["triangleIndex"]).sum())
    assert differing <= len(strict) // 1000
    assert int((strict["distance"] != culled["distance"]).sum()) <= len(strict) // 1000
    again = d.Update(128, 128, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    assert again.tobytes() == strict.tobytes()                      # back in strict mode
    with pytest.raises(_lib.UsrtError):
        d.container.ctx.set_trace_mode(3)
    d.OnDestroy()


@pytest.mark.parametrize("log2n", [22, 24])
def test_large_soup_build_and_incoherent_rays(usrt, oracle, log2n):
    """BASELINE config 4 shape at 2^22 triangles and at its stated size, 2^24 = 16,777,216: full build
    parity by digest of every buffer, plus incoherent random rays and a small primary frame."""
    import hashlib
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    n = 1 << log2n
    tris = meshes.uniform_soup(n, seed=0x5EED0004)
    ref = oracle.Scene(tris)
    ctx = usrt.Context(n)
    ctx.upload_triangles(tris)
    ctx.rebuild()
    assert sha(ctx.download(_lib.BUF_KEYS)) == sha(ref.sortedMortonCodes)
    assert sha(ctx.download(_lib.BUF_TRIANGLE_INDEX)) == sha(ref.sortedTriangleIndices)
    assert sha(ctx.download(_lib.BUF_INTERNAL_NODES, n - 1)) == sha(ref.internalNodes[:n - 1])
    assert sha(ctx.download(_lib.BUF_LEAF_NODES)) == sha(ref.leafNodes)
    assert sha(ctx.download(_lib.BUF_BVH_DATA, n - 1)) == sha(ref.bvhData[:n - 1])
    assert ctx.count_corrupted_nodes() == (0, 0)
    rays = meshes.incoherent_rays(3000, seed=0x5EED0005)
    assert _same(ctx.trace_rays(rays), ref.trace_rays(rays, threads=16))
    cam = meshes.SCENE_SOUP_CAMERA
    got = ctx.trace_primary(48, 27, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    assert _same(got, ref.trace_primary(48, 27, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], threads=16))
    ctx.close()


def _checker_texture(w=64, h=48, seed=3):
    rng = np.random.default_rng(seed)
    tex = rng.random((h, w, 4), dtype=np.float32)
    tex[::7, :, :3] *= 40.0            # some values beyond [0,1] so fp16 rounding/overflow paths are hit
    return tex


@pytest.mark.parametrize("name,cam", [("sphere", "SCENE_C2_CAMERA"), ("soup4097", "SCENE_SOUP_CAMERA"), ("refgrid", "REFERENCE_CAMERA")])
def test_shading_epilogue(usrt, oracle, name, cam):
    """SURVEY 8(f)-1, Raytracing.compute:178-184: RGBA16F image bit-exact against the oracle, misses included
    (alpha 0, colour of triangle 0 as the reference computes it)."""
    tris = _mesh(name); cam = getattr(meshes, cam)
    w, h = 160, 90
    tex = _checker_texture()
    d = usrt.RaytracingMeshDrawer(tris).Awake()
    ctx = d.container.ctx
    ctx.upload_texture(tex)
    hits = d.Update(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    img = ctx.shade()
    want = oracle.shade(hits, tris, tex)
    assert img.shape == (w * h, 4)
    assert img.view(np.uint16).tobytes() == want.view(np.uint16).tobytes()
    hit = hits["distance"] != oracle.max_float()
    assert np.array_equal(img[:, 3] == 1.0, hit) and np.array_equal(img[:, 3] == 0.0, ~hit)
    d.OnDestroy()


def test_bvh_dump_and_reload(usrt, oracle, tmp_path):
    """save the seven buffers, load them into a FRESH context (no rebuild) and trace; also
    install a BVH that was built elsewhere -- here by the oracle -- straight from host arrays."""
    from unitysimpleraytracing_b200 import bvh_io
    tris = _mesh("c1"); cam = meshes.SCENE_SOUP_CAMERA
    n = len(tris)
    a = usrt.Context(n); a.upload_triangles(tris); a.rebuild()
    want = a.trace_primary(96, 54, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    path = str(tmp_path / "scene.usrtbvh")
    saved = bvh_io.save_bvh(a, path)
    a.close()
    b = usrt.Context(n + 7)
    assert bvh_io.load_bvh(b, path) == n
    assert _same(b.trace_primary(96, 54, cam["near"], cam["tan_half_fov"], cam["cam_to_world"]), want)
    for name, buf in zip(("keys", "triangleIndex", "triangleAABB", "bvhData", "leafNodes", "internalNodes"),
                         (_lib.BUF_KEYS, _lib.BUF_TRIANGLE_INDEX, _lib.BUF_TRIANGLE_AABB, _lib.BUF_BVH_DATA,
                          _lib.BUF_LEAF_NODES, _lib.BUF_INTERNAL_NODES)):
        assert _same(b.download(buf, len(saved[name])), saved[name]), name
    rays = meshes.incoherent_rays(2000, seed=11)
    ref = oracle.Scene(tris)
    c = usrt.Context(n)
    bvh_io.upload_bvh(c, n, dict(keys=ref.sortedMortonCodes, triangleIndex=ref.sortedTriangleIndices, triangleData=tris,
                                 triangleAABB=ref.triangleAABB, bvhData=ref.bvhData, leafNodes=ref.leafNodes,
                                 internalNodes=ref.internalNodes))
    assert _same(c.trace_rays(rays), ref.trace_rays(rays, threads=8))
    assert _same(c.trace_primary(96, 54, cam["near"], cam["tan_half_fov"], cam["cam_to_world"]), want)
    b.close(); c.close()


def test_many_tiny_meshes_and_ragged_ray_counts(usrt, oracle):
    """n = 2 .. 70 triangles (single block, partial warps, every internal node spanning the one block) and ray
    batches whose size is not a multiple of the CTA / warp size, all in one re-used context."""
    ctx = usrt.Context(128)
    rng = np.random.default_rng(2024)
    for trial in range(40):
        n = int(rng.integers(2, 71))
        tris = meshes.uniform_soup(n, seed=1000 + trial, extent=20.0, edge=8.0)
        if trial % 5 == 0:
            tris["b"][::2] = tris["a"][::2]                      # degenerate ones
        if trial % 7 == 0:
            tris[n // 2:] = tris[:n - n // 2]                    # exact duplicates -> equal Morton keys
        ref = oracle.Scene(tris)
        ctx.upload_triangles(tris)
        ctx.rebuild()
        assert np.array_equal(ctx.download(_lib.BUF_KEYS), ref.sortedMortonCodes), (trial, n)
        assert np.array_equal(ctx.download(_lib.BUF_TRIANGLE_INDEX), ref.sortedTriangleIndices), (trial, n)
        assert _same(ctx.download(_lib.BUF_INTERNAL_NODES, n - 1), ref.internalNodes[:n - 1]), (trial, n)
        assert _same(ctx.download(_lib.BUF_LEAF_NODES), ref.leafNodes), (trial, n)
        assert _same(ctx.download(_lib.BUF_BVH_DATA, n - 1), ref.bvhData[:n - 1]), (trial, n)
        m = int(rng.integers(1, 300))
        rays = meshes.incoherent_rays(m, seed=50 + trial, extent=25.0)
        assert _same(ctx.trace_rays(rays), ref.trace_rays(rays)), (trial, n, m)
    assert len(ctx.trace_rays(np.zeros((0, 8), np.float32))) == 0
    ctx.close()


# ---- BASELINE configurations at their stated sizes ------------------------------------------------------------------
def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _ref_digests(name):
    """sha256s written by tests/golden/make_ref_golden.py from the reference's own code (oracle/_ref)."""
    import json, os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_digests.json")))[name]


def _check_config(usrt, oracle, tris, cam, w, h, golden):
    """All six build buffers + the full primary frame: CUDA == oracle (live, byte for byte) == reference-generated digest."""
    import os
    n = len(tris)
    assert _sha(tris) == golden["triangles"], "mesh generator output differs from the one the goldens were made with"
    ref = oracle.Scene(tris)
    ctx = usrt.Context(n)
    ctx.upload_triangles(tris)
    ctx.rebuild()
    got = dict(sortedMortonCodes=ctx.download(_lib.BUF_KEYS), sortedTriangleIndices=ctx.download(_lib.BUF_TRIANGLE_INDEX),
               triangleAABB=ctx.download(_lib.BUF_TRIANGLE_AABB), internalNodes=ctx.download(_lib.BUF_INTERNAL_NODES, n - 1),
               leafNodes=ctx.download(_lib.BUF_LEAF_NODES), bvhData=ctx.download(_lib.BUF_BVH_DATA, n - 1))
    want = dict(sortedMortonCodes=ref.sortedMortonCodes, sortedTriangleIndices=ref.sortedTriangleIndices,
                triangleAABB=ref.triangleAABB, internalNodes=ref.internalNodes[:n - 1], leafNodes=ref.leafNodes,
                bvhData=ref.bvhData[:n - 1])
    for k in got:
        assert _same(got[k], want[k]), k                         # vs the oracle, run here
        assert _sha(got[k]) == golden[k], k                      # vs the reference's own code (committed digest)
    assert ctx.count_corrupted_nodes() == (0, 0)
    frame = ctx.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    want_frame = ref.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], threads=os.cpu_count() or 1)
    assert _report_ties(frame, want_frame) == (0, 0)
    assert _same(frame, want_frame)
    assert _sha(frame) == golden["primary_%dx%d" % (w, h)]
    assert int((frame["distance"] != oracle.max_float()).sum()) == golden["primary_hit_count"]
    ctx.close()


def test_config0_as_stated_65536_triangles_512x512(usrt, oracle):
    """BASELINE configs[0]: 65,536-triangle soup, Morton -> sort -> LBVH -> refit -> 512x512 primary rays, every record."""
    _check_config(usrt, oracle, meshes.scene_c1(), meshes.SCENE_SOUP_CAMERA, 512, 512, _ref_digests("config0_soup_65536"))


def test_config1_as_stated_1048576_triangles_1080p(usrt, oracle):
    """BASELINE configs[1] -- the bench workload -- at full size: 1,048,576 triangles (sphere + height field), full
    rebuild, 1920x1080 primary rays; all six build buffers and all 2,073,600 hit records."""
    _check_config(usrt, oracle, meshes.scene_c2(), meshes.SCENE_C2_CAMERA, 1920, 1080, _ref_digests("config1_scene_1048576"))


def test_reference_scene_mesh_full_frame(usrt, oracle):
    _check_config(usrt, oracle, meshes.reference_scene_grid(), meshes.REFERENCE_CAMERA, 480, 270, _ref_digests("refgrid_12800"))
