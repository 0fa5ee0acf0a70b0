"""SURVEY.md 8(f)-4 key variants on the device, each bit-exact against its oracle twin (oracle.VariantScene):
index tie-break instead of DistributeKeys, 63-bit Morton keys with the 8-pass 64-bit sorter."""
import numpy as np
import pytest

from unitysimpleraytracing_b200 import _lib, meshes

pytestmark = pytest.mark.gpu


def _same(a, b):
    return np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes()


def _mesh(name):
    if name == "refgrid":                 # ~11 triangles per Morton cell: 11,644 duplicate codes
        return meshes.reference_scene_grid()
    if name == "identical":
        return np.repeat(meshes.uniform_soup(1, seed=45), 3001)
    if name == "soup4097":
        return meshes.uniform_soup(4097, seed=43)
    if name == "soup3":
        return meshes.uniform_soup(3, seed=42)
    if name == "c1":
        return meshes.scene_c1()
    raise ValueError(name)


@pytest.mark.parametrize("n,kind", [(1, "u"), (2, "u"), (4097, "u"), (100000, "u"), (100000, "few"), (1 << 20, "u"),
                                    ((1 << 21) + 13, "low32"), (300001, "high32")])
def test_sort_pairs64_matches_oracle(usrt, oracle, n, kind):
    """ComputeBufferSorter<ulong, uint>: 8 passes over 64-bit keys == the oracle's 8-pass LSD == std::stable_sort."""
    rng = np.random.default_rng(n + len(kind))
    k = rng.integers(0, 2 ** 64, n, dtype=np.uint64)
    if kind == "few":
        k = (k % 23) << np.uint64(40)
    elif kind == "low32":
        k &= np.uint64(0xFFFFFFFF)           # upper four digits all zero
    elif kind == "high32":
        k &= np.uint64(0xFFFFFFFF00000000)
    v = np.arange(n, dtype=np.uint32)
    want_k, want_v = oracle.sort64(k, v)
    sk, sv = oracle.stable_sort64(k, v)
    assert np.array_equal(want_k, sk) and np.array_equal(want_v, sv)
    ctx = usrt.Context(2)
    gk, gv = k.copy(), v.copy()
    ctx.sort_pairs64_host(gk, gv)
    assert np.array_equal(gk, want_k) and np.array_equal(gv, want_v)
    gk2 = k.copy()
    ctx.sort_pairs64_host(gk2)                # keys only
    assert np.array_equal(gk2, want_k)
    ctx.close()


@pytest.mark.parametrize("mode", [_lib.KEYS_INDEX_TIEBREAK, _lib.KEYS_MORTON64])
@pytest.mark.parametrize("name", ["soup3", "soup4097", "refgrid", "identical", "c1"])
def test_variant_build_and_trace_match_oracle_twin(usrt, oracle, mode, name):
    tris = _mesh(name)
    n = len(tris)
    ref = oracle.VariantScene(tris, mode)
    ctx = usrt.Context(n + 5)
    ctx.set_key_mode(mode)
    ctx.upload_triangles(tris)
    for _ in range(2):                        # stage by stage, then fused; both repeatable
        ctx.morton(); ctx.sort()
        if mode == _lib.KEYS_INDEX_TIEBREAK:
            with pytest.raises(_lib.UsrtError):
                ctx.distribute_keys()
        else:
            ctx.distribute_keys()
        ctx.construct_tree(); ctx.construct_bvh()
        keys = ctx.download(_lib.BUF_KEYS64 if mode == _lib.KEYS_MORTON64 else _lib.BUF_KEYS)
        assert np.array_equal(keys, ref.sortedMortonCodes)
        assert np.array_equal(ctx.download(_lib.BUF_TRIANGLE_INDEX), ref.sortedTriangleIndices)
        assert _same(ctx.download(_lib.BUF_INTERNAL_NODES, n - 1), ref.internalNodes[:n - 1])
        assert _same(ctx.download(_lib.BUF_LEAF_NODES), ref.leafNodes)
        assert _same(ctx.download(_lib.BUF_BVH_DATA, n - 1), ref.bvhData[:n - 1])
        assert ctx.count_corrupted_nodes() == (0, 0)
    ctx.rebuild()
    assert _same(ctx.download(_lib.BUF_INTERNAL_NODES, n - 1), ref.internalNodes[:n - 1])
    assert _same(ctx.download(_lib.BUF_BVH_DATA, n - 1), ref.bvhData[:n - 1])
    cam = meshes.REFERENCE_CAMERA if name == "refgrid" else meshes.SCENE_SOUP_CAMERA
    got = ctx.trace_primary(96, 54, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    assert _same(got, ref.trace_primary(96, 54, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], threads=4))
    rays = meshes.incoherent_rays(800, seed=9, extent=4.0 if name == "refgrid" else 100.0)
    assert _same(ctx.trace_rays(rays), ref.trace_rays(rays, threads=4))
    with pytest.raises(_lib.UsrtError):       # the other key buffer is void in this mode
        ctx.download(_lib.BUF_KEYS if mode == _lib.KEYS_MORTON64 else _lib.BUF_KEYS64)
    # back to the reference mode on the same context: the reference result, bit for bit
    ctx.set_key_mode(_lib.KEYS_REFERENCE)
    with pytest.raises(_lib.UsrtError):       # the variant's tree is void after the switch
        ctx.trace_primary(4, 4, 0.3, 0.5, np.eye(4, dtype=np.float32))
    ctx.rebuild()
    r0 = oracle.Scene(tris)
    assert np.array_equal(ctx.download(_lib.BUF_KEYS), r0.sortedMortonCodes)
    assert _same(ctx.download(_lib.BUF_BVH_DATA, n - 1), r0.bvhData[:n - 1])
    ctx.close()


def test_variants_find_the_same_closest_hits_as_the_reference_mode(usrt, oracle):
    """Different keys give a different tree, but the closest hit of a ray is a property of the triangles: distances agree
    with the reference mode everywhere; ids may differ only where two triangles tie on distance (visit order differs)."""
    tris = meshes.scene_c1(); cam = meshes.SCENE_SOUP_CAMERA
    ctx = usrt.Context(len(tris)); ctx.upload_triangles(tris); ctx.rebuild()
    base = ctx.trace_primary(160, 90, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    for mode in (_lib.KEYS_INDEX_TIEBREAK, _lib.KEYS_MORTON64):
        ctx.set_key_mode(mode); ctx.rebuild()
        got = ctx.trace_primary(160, 90, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
        assert np.array_equal(got["distance"], base["distance"])
        differ = got["triangleIndex"] != base["triangleIndex"]
        assert differ.sum() <= 4                                  # exact distance ties only
    ctx.close()
