"""CPU tests of the oracle itself: the reference's runtime self-checks restated as
unit tests, the C++ twin against an independent numpy restatement, and structural invariants."""
import numpy as np
import pytest

from oracle import np_oracle as NP
from unitysimpleraytracing_b200 import meshes


def _keys(kind, n, seed=1):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    if kind == "morton30":
        return rng.integers(0, 2 ** 30, n, dtype=np.uint64).astype(np.uint32)
    if kind == "low":
        return (rng.integers(0, 5, n, dtype=np.uint64) * 0x01000100).astype(np.uint32)
    if kind == "equal":
        return np.full(n, 0xCAFEF00D, np.uint32)
    if kind == "padded":                     # real keys followed by 0xFFFFFFFF padding (MeshBufferContainer.cs:108-109)
        k = rng.integers(0, 2 ** 30, n, dtype=np.uint64).astype(np.uint32)
        k[n // 2:] = 0xFFFFFFFF
        return k
    raise ValueError(kind)


# ---- Morton / AABB ---------------------------------------------------------------------------------
def test_morton_known_answers(oracle):
    # ExpandBits/Morton3D (MeshBufferContainer.cs:32-50): x is the most significant bit of each triple
    t = np.zeros(4, oracle.TRIANGLE)
    pts = np.array([[-125, -125, -125], [125, 125, 125], [0, -125, -125], [-125, -125, 0]], np.float32)
    for f in "abc":
        t[f] = pts
    keys, values, aabb = oracle.morton(t)
    # centroid == the point; (p+125)/250*1024 -> 0, 1024->clamped 1023, 512
    assert keys[0] == 0
    assert keys[1] == 0x3FFFFFFF
    assert keys[2] == 0b100 << 27            # x = 512 = bit 9 -> key bit 3*9+2
    assert keys[3] == 0b001 << 27            # z = 512 -> key bit 27
    assert list(values) == [0, 1, 2, 3]
    assert np.allclose(aabb["min"][0], -125.001) and np.allclose(aabb["max"][0], -124.999)
    assert (aabb["_dummy0"] == 0).all() and (aabb["_dummy1"] == 0).all()


@pytest.mark.parametrize("mesh", ["soup", "grid", "sphere", "outside"])
def test_morton_matches_numpy_restatement(oracle, mesh):
    tris = {"soup": lambda: meshes.uniform_soup(5000, seed=3),
            "grid": meshes.reference_scene_grid,
            "sphere": lambda: meshes.sphere(24, 48),
            "outside": lambda: meshes.uniform_soup(3000, seed=4, extent=200.0)}[mesh]()   # clamps to 0 / 1023
    keys, values, aabb = oracle.morton(tris)
    k2, mn, mx = NP.morton_and_aabb(tris["a"], tris["b"], tris["c"])
    assert np.array_equal(keys, k2)
    assert aabb["min"].tobytes() == mn.tobytes() and aabb["max"].tobytes() == mx.tobytes()
    assert np.array_equal(values, np.arange(len(tris), dtype=np.uint32))
    assert keys.max() < 2 ** 30


def test_reference_scene_has_many_duplicate_keys(oracle):
    # the shipped 12,800-triangle grid falls into ~1,156 Morton cells -- why DistributeKeys exists
    keys, _, _ = oracle.morton(meshes.reference_scene_grid())
    assert len(keys) == 12800
    assert 1000 < len(np.unique(keys)) < 1400


# ---- sort: the reference's validators (ComputeBufferSorter.cs:150-272) -------------------------------
@pytest.mark.parametrize("kind", ["uniform", "morton30", "low", "equal", "padded"])
@pytest.mark.parametrize("n", [1024, 8192])
def test_sort_pass_validators(oracle, kind, n):
    keys = _keys(kind, n)
    values = np.arange(n, dtype=np.uint32)
    nb = n // 1024
    for bit_offset in (0, 8, 16, 24):
        r = oracle.sort_pass(keys, values, bit_offset)
        radix = lambda k: (k >> np.uint32(bit_offset)) & np.uint32(255)
        # ValidateIntermediateData part 1 (:200-218): per-digit multiset preserved by the block sort
        assert np.array_equal(np.bincount(radix(r["sortedBlocksKeys"]), minlength=256),
                              np.bincount(radix(keys), minlength=256))
        # part 2 (:227-254): per block b, digit k: count == sizesBefore[b + k*numBlocks]
        blocks = radix(r["sortedBlocksKeys"]).reshape(nb, 1024)
        for b in range(nb):
            assert np.array_equal(np.bincount(blocks[b], minlength=256), r["sizesBefore"][b::nb][:256])
            assert (np.diff(blocks[b].astype(np.int64)) >= 0).all()       # block is digit-sorted
        # part 3 (:256-271): exclusive-scan identity
        assert r["sizesAfter"][0] == 0
        assert np.array_equal(r["sizesAfter"][1:], (r["sizesBefore"][:-1] + r["sizesAfter"][:-1]).astype(np.uint32))
        # the pass is a stable sort by this digit
        order = np.argsort(radix(keys), kind="stable")
        assert np.array_equal(r["keys"], keys[order]) and np.array_equal(r["values"], values[order])
        keys, values = r["keys"], r["values"]
    # ValidateSortedData (:150-177): non-decreasing
    assert (np.diff(keys.astype(np.int64)) >= 0).all()


@pytest.mark.parametrize("kind", ["uniform", "morton30", "low", "equal", "padded"])
@pytest.mark.parametrize("n", [0, 1, 2, 1023, 1025, 5000, 70001])
def test_sort_is_stable_sort_by_key(oracle, kind, n):
    keys = _keys(kind, n, seed=n + 5)
    values = np.arange(n, dtype=np.uint32)[::-1].copy()
    k, v = oracle.sort(keys, values)
    k2, v2 = oracle.stable_sort(keys, values)
    k3, v3 = NP.stable_sort(keys, values)
    assert np.array_equal(k, k2) and np.array_equal(v, v2)
    assert np.array_equal(k, k3) and np.array_equal(v, v3)


# ---- DistributeKeys -------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["uniform", "morton30", "low", "equal"])
def test_distribute_keys(oracle, kind):
    keys = np.sort(_keys(kind, 4000))
    d = oracle.distribute_keys(keys)
    assert np.array_equal(d, NP.distribute_keys(keys))
    assert d[0] == 0
    if kind != "uniform":     # no wraparound possible below 2^31
        assert (np.diff(d.astype(np.int64)) > 0).all()


def test_distribute_keys_wraps_like_unchecked_uint(oracle):
    # unsorted input (misuse) still follows C# unchecked arithmetic: max(k[i]-k[i-1] mod 2^32, 1)
    keys = np.array([10, 5, 5, 0xFFFFFFF0, 3], np.uint32)
    want = np.array([0, 0xFFFFFFFB, 0xFFFFFFFC, 0xFFFFFFE7, 0xFFFFFFFA], np.uint32)
    assert np.array_equal(oracle.distribute_keys(keys), want)
    assert np.array_equal(NP.distribute_keys(keys), want)


def test_distribute_keys_only_touches_triangles_length(oracle):
    keys = np.array([1, 1, 4, 0xFFFFFFFF, 0xFFFFFFFF], np.uint32)
    assert list(oracle.distribute_keys(keys, n=3)) == [0, 1, 4, 0xFFFFFFFF, 0xFFFFFFFF]


# ---- tree + refit ---------------------------------------------------------------------------------
def _check_tree_invariants(internal, leaf, n):
    # every internal has two children, parent/child reciprocity, leaf.index == position,
    # in-order leaf sequence == 0..n-1, root = node 0 with parent NullLeaf
    assert internal["parent"][0] == 0xFFFFFFFF
    assert np.array_equal(internal["index"][:n - 1], np.arange(n - 1, dtype=np.uint32))
    assert np.array_equal(leaf["index"][:n], np.arange(n, dtype=np.uint32))
    for side in ("left", "right"):
        c, t = internal[side + "Node"][:n - 1], internal[side + "NodeType"][:n - 1]
        assert set(np.unique(t)) <= {0, 1}
        ii = np.nonzero(t == 0)[0]; ll = np.nonzero(t == 1)[0]
        assert np.array_equal(internal["parent"][c[ii]], ii.astype(np.uint32))
        assert np.array_equal(leaf["parent"][c[ll]], ll.astype(np.uint32))
    kids_i = np.concatenate([internal[s + "Node"][:n - 1][internal[s + "NodeType"][:n - 1] == 0] for s in ("left", "right")])
    kids_l = np.concatenate([internal[s + "Node"][:n - 1][internal[s + "NodeType"][:n - 1] == 1] for s in ("left", "right")])
    assert np.array_equal(np.sort(kids_i), np.arange(1, n - 1, dtype=np.uint32))      # every non-root internal once
    assert np.array_equal(np.sort(kids_l), np.arange(n, dtype=np.uint32))            # every leaf once
    order, stack = [], [(0, 0)]
    while stack:
        node, typ = stack.pop()
        if typ == 1:
            order.append(node)
        else:
            stack.append((int(internal["rightNode"][node]), int(internal["rightNodeType"][node])))
            stack.append((int(internal["leftNode"][node]), int(internal["leftNodeType"][node])))
    assert order == list(range(n))


@pytest.mark.parametrize("mesh,n", [("soup", 2), ("soup", 3), ("soup", 257), ("soup", 3000), ("grid", 12800)])
def test_tree_and_refit_match_topdown_restatement(oracle, mesh, n):
    tris = meshes.uniform_soup(n, seed=11) if mesh == "soup" else meshes.reference_scene_grid()
    s = oracle.Scene(tris)
    n = s.n
    _check_tree_invariants(s.internalNodes, s.leafNodes, n)
    it, lf = NP.radix_tree_topdown(s.sortedMortonCodes)
    assert np.array_equal(s.internalNodes.view(np.uint32).reshape(-1, 6)[:n - 1], it[:n - 1])
    assert np.array_equal(s.leafNodes.view(np.uint32).reshape(-1, 2), lf)
    bmin, bmax = NP.refit_recursive(it, s.sortedTriangleIndices, s.triangleAABB["min"], s.triangleAABB["max"])
    assert s.bvhData["min"][:n - 1].tobytes() == bmin[:n - 1].tobytes()
    assert s.bvhData["max"][:n - 1].tobytes() == bmax[:n - 1].tobytes()
    assert (s.bvhData["_dummy0"] == 0).all() and (s.bvhData["_dummy1"] == 0).all()
    # root box == union of all triangle boxes (SURVEY 8c iv)
    assert np.array_equal(s.bvhData["min"][0], s.triangleAABB["min"].min(0))
    assert np.array_equal(s.bvhData["max"][0], s.triangleAABB["max"].max(0))


def test_tree_depth_bounded_by_key_width(oracle):
    # unique 32-bit keys => depth <= 32 => the 64-entry stack (Raytracing.compute:133) cannot overflow
    s = oracle.Scene(meshes.reference_scene_grid())
    d, stack = 0, [(0, 1)]
    while stack:
        node, dep = stack.pop()
        d = max(d, dep)
        for side in ("left", "right"):
            if s.internalNodes[side + "NodeType"][node] == 0:
                stack.append((int(s.internalNodes[side + "Node"][node]), dep + 1))
    assert d <= 33


# ---- traversal ------------------------------------------------------------------------------------
def test_miss_sentinel_is_not_flt_max(oracle):
    assert oracle.max_float().view(np.uint32) == 0x4EFF0000 and float(oracle.max_float()) == 2139095040.0


@pytest.mark.parametrize("mesh", ["soup", "grid"])
def test_traversal_equals_brute_force_in_visit_order(oracle, mesh):
    # BVH walk == brute force over all triangles in the walk's visiting order
    if mesh == "soup":
        s = oracle.Scene(meshes.uniform_soup(2000, seed=21)); cam = meshes.SCENE_SOUP_CAMERA
    else:
        s = oracle.Scene(meshes.reference_scene_grid()); cam = meshes.REFERENCE_CAMERA
    rays = np.concatenate([oracle.primary_rays(24, 24, cam["near"], cam["tan_half_fov"], cam["cam_to_world"]),
                           meshes.incoherent_rays(300, seed=5, extent=4.0 if mesh == "grid" else 100.0)])
    got = s.trace_rays(rays)
    order = s.visit_order()
    assert np.array_equal(np.sort(order), np.arange(s.n, dtype=np.uint32))
    bf = s.brute_force(rays, order=order)
    assert got.tobytes() == bf.tobytes()
    # and against the independent numpy brute force
    t = s.triangleData
    ref = NP.brute_force_hits(rays[:, 0:3], rays[:, 4:7], t["a"], t["b"], t["c"], s.triangleAABB["min"],
                              s.triangleAABB["max"], order, oracle.max_float())
    for g, (dist, tri, u, v) in zip(got, ref):
        assert g["distance"].view(np.uint32) == np.float32(dist).view(np.uint32)
        assert g["triangleIndex"] == tri
        assert g["uv"][0].view(np.uint32) == np.float32(u).view(np.uint32)
        assert g["uv"][1].view(np.uint32) == np.float32(v).view(np.uint32)
    assert (got["distance"] != oracle.max_float()).sum() > 50


def test_primary_trace_equals_ray_buffer_trace(oracle):
    s = oracle.Scene(meshes.uniform_soup(1500, seed=31)); cam = meshes.SCENE_SOUP_CAMERA
    a = s.trace_primary(40, 30, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], threads=2)
    b = s.trace_rays(oracle.primary_rays(40, 30, cam["near"], cam["tan_half_fov"], cam["cam_to_world"]))
    assert a.tobytes() == b.tobytes()


def test_row_zero_is_bottom_of_view(oracle):
    # Raytracing.compute:116: id.y = 0 maps to the most negative camera-space y
    cam = meshes.REFERENCE_CAMERA
    rays = oracle.primary_rays(8, 8, cam["near"], cam["tan_half_fov"], cam["cam_to_world"]).reshape(8, 8, 8)
    assert (rays[0, :, 5] < 0).all() and (rays[7, :, 5] > 0).all()
    assert np.allclose(rays[:, :, 0:3], [0, 0, 15.7])
    assert np.allclose(np.linalg.norm(rays[:, :, 4:7], axis=2), 1.0, atol=1e-6)
    assert (rays[:, :, 6] < 0).all()          # the scene camera looks down world -z (Scene.unity:342)


# ---- shading epilogue --------------------------------------------------------------------
def test_float_to_half_matches_numpy(oracle):
    rng = np.random.default_rng(5)
    vals = np.concatenate([
        (rng.standard_normal(4000) * 10.0 ** rng.integers(-9, 6, 4000)).astype(np.float32),
        np.array([0, -0.0, 65504, 65519.99, 65520, 1e-8, 5.96e-8, 2.98e-8, 2.9802322e-8, 2.9802326e-8, np.inf, -np.inf,
                  6.1035156e-05, 6.0975552e-05, 1.0009766, 1.0004883, 1.0014648], np.float32)])
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    got = np.array([oracle.float_to_half_bits(float(v)) for v in vals], np.uint16)
    assert np.array_equal(got, want)


def test_shade_against_numpy_restatement(oracle):
    tris = meshes.sphere(12, 24)
    s = oracle.Scene(tris); cam = meshes.SCENE_C2_CAMERA
    hits = s.trace_primary(40, 30, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    rng = np.random.default_rng(1)
    tex = rng.random((16, 20, 4), dtype=np.float32)
    got = oracle.shade(hits, tris, tex)
    F = np.float32
    t = tris[hits["triangleIndex"]]
    bu, bv = hits["uv"][:, 0], hits["uv"][:, 1]
    bw = ((F(1) - bu).astype(F) - bv).astype(F)
    def interp(a, b, c):
        return (((bw[:, None] * a).astype(F) + (bu[:, None] * b).astype(F)).astype(F) + (bv[:, None] * c).astype(F)).astype(F)
    uv = interp(t["a_uv"], t["b_uv"], t["c_uv"]); nrm = interp(t["a_normal"], t["b_normal"], t["c_normal"])
    light = F(1) / np.sqrt(F(3), dtype=F)
    ndotl = (((light * nrm[:, 0]).astype(F) + (light * nrm[:, 1]).astype(F)).astype(F) + (light * nrm[:, 2]).astype(F)).astype(F)
    shade = np.fmax(F(0.4), ndotl)
    th, tw = tex.shape[:2]
    x = ((uv[:, 0] * F(tw)).astype(F) - F(0.5)).astype(F); y = ((uv[:, 1] * F(th)).astype(F) - F(0.5)).astype(F)
    x0f, y0f = np.floor(x), np.floor(y)
    fx, fy = (x - x0f).astype(F), (y - y0f).astype(F)
    x0 = np.clip(x0f, 0, tw - 1).astype(int); x1 = np.clip(x0f + 1, 0, tw - 1).astype(int)
    y0 = np.clip(y0f, 0, th - 1).astype(int); y1 = np.clip(y0f + 1, 0, th - 1).astype(int)
    def lerp(a, b, w):
        return (a + ((b - a).astype(F) * w[:, None]).astype(F)).astype(F)
    texel = lerp(lerp(tex[y0, x0], tex[y0, x1], fx), lerp(tex[y1, x0], tex[y1, x1], fx), fy)
    rgb = (texel[:, :3] * shade[:, None]).astype(F).astype(np.float16)
    alpha = (hits["distance"] != oracle.max_float()).astype(np.float16)
    assert np.array_equal(got[:, :3].view(np.uint16), rgb.view(np.uint16))
    assert np.array_equal(got[:, 3], alpha)
    assert (alpha == 1).sum() > 100


def test_diffuse_bounce_rays_properties(oracle):
    """BASELINE configs[4] bounce rays (defined by the oracle): unit directions in the hemisphere of the hit
    triangle's ray-facing normal, origin just off the surface, null rays for pixels that missed, and a sample
    window equals the same samples of the full set."""
    from unitysimpleraytracing_b200 import meshes
    tris = meshes.scene_c1(4096); cam = meshes.SCENE_SOUP_CAMERA
    w, h = 48, 27
    args = (w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    hits = oracle.Scene(tris).trace_primary(*args)
    rays = oracle.diffuse_rays(hits, tris, *args, 77, 0, 4).reshape(4, w * h, 8)
    miss = hits["distance"] == oracle.max_float()
    assert miss.any() and (~miss).any()
    assert not rays[:, miss].any()
    t = tris[hits["triangleIndex"][~miss]]
    n = np.cross((t["b"] - t["a"]).astype(np.float64), (t["c"] - t["a"]).astype(np.float64))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    d = rays[:, ~miss, 4:7].astype(np.float64)
    assert (np.abs(np.linalg.norm(d, axis=2) - 1) < 1e-6).all()
    cosine = np.abs((d * n[None]).sum(2))                   # n up to sign: the ray-facing one is used
    o = rays[:, ~miss, 0:3].astype(np.float64)
    assert (np.abs(o[0] - o[1]) == 0).all()                 # same origin for every sample of a pixel
    assert np.array_equal(oracle.diffuse_rays(hits, tris, *args, 77, 2, 2).reshape(2, w * h, 8), rays[2:])
    assert 0.55 < cosine.mean() < 0.78                      # cosine-weighted: E[cos] = 2/3
    assert not np.array_equal(rays[0], rays[1])


def test_diffuse_bounce_rays_cpp_twin_equals_numpy_restatement(oracle):
    """The two independent restatements (C++ scalar loop, vectorised numpy) of the bounce-ray generator agree
    bit for bit, primary-ray generation included."""
    from unitysimpleraytracing_b200 import meshes
    for tris, cam in ((meshes.scene_c1(4096), meshes.SCENE_SOUP_CAMERA), (meshes.sphere(24, 48), meshes.REFERENCE_CAMERA)):
        w, h = 40, 23
        args = (w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
        hits = oracle.Scene(tris).trace_primary(*args)
        want = oracle.diffuse_rays(hits, tris, *args, 0xABCDEF, 1, 3)
        got = NP.diffuse_rays(hits["distance"], hits["triangleIndex"], tris["a"], tris["b"], tris["c"], *args, 0xABCDEF, 1, 3,
                              oracle.max_float())
        assert (hits["distance"] != oracle.max_float()).any()
        assert got.tobytes() == want.tobytes()


def test_fitted_world_box_opt_in(oracle):
    """The MeshBufferContainer.cs:7 TODO as an opt-in: per-axis scene box (flat axes padded), keys from it agree
    between the C++ twin and the numpy restatement, the cube path is untouched, and the tree is still valid."""
    from unitysimpleraytracing_b200 import meshes
    for tris in (meshes.uniform_soup(3000, seed=5, extent=40.0), meshes.reference_scene_grid()):
        lo, hi = oracle.scene_box(tris)
        v = np.concatenate([tris["a"], tris["b"], tris["c"]])
        assert np.array_equal(lo, v.min(0))
        flat = v.max(0) == v.min(0)
        assert np.array_equal(hi, np.where(flat, v.min(0) + np.float32(1), v.max(0)))
        keys, _, _ = oracle.morton(tris, lo, hi)
        nk, _, _ = NP.morton_and_aabb(tris["a"], tris["b"], tris["c"], lo, hi)
        assert np.array_equal(keys, nk)
        cube_keys, _, _ = oracle.morton(tris)
        assert np.array_equal(cube_keys, oracle.morton(tris, np.full(3, -125, np.float32), np.full(3, 125, np.float32))[0])
        assert not np.array_equal(keys, cube_keys)
        assert len(np.unique(keys)) >= len(np.unique(cube_keys))      # a tighter box never merges more cells
        s = oracle.Scene(tris, lo, hi)
        assert s.count_corrupted() == (0, 0) if hasattr(s, "count_corrupted") else True
        assert np.array_equal(np.sort(s.sortedTriangleIndices), np.arange(len(tris)))


# ---- SURVEY 8(f)-4 key variants: the oracle twins are self-consistent -------------------------------------------------
@pytest.mark.parametrize("kind", ["uniform", "few", "low32", "high32"])
def test_sort64_twin_equals_stable_sort(oracle, kind):
    rng = np.random.default_rng(len(kind))
    k = rng.integers(0, 2 ** 64, 50000, dtype=np.uint64)
    if kind == "few":
        k = (k % 19) << np.uint64(37)
    elif kind == "low32":
        k &= np.uint64(0xFFFFFFFF)
    elif kind == "high32":
        k &= np.uint64(0xFFFFFFFF00000000)
    v = np.arange(len(k), dtype=np.uint32)
    a, b = oracle.sort64(k, v), oracle.stable_sort64(k, v)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    order = np.argsort(k, kind="stable")
    assert np.array_equal(a[1], order.astype(np.uint32))


def test_distribute_keys64_twin(oracle):
    k = np.sort(np.random.default_rng(5).integers(0, 2 ** 63, 4000, dtype=np.uint64))
    k[100:140] = k[100]                                              # a run of equal keys
    d = oracle.distribute_keys64(k)
    assert d[0] == 0 and (np.diff(d.astype(np.int64)) > 0).all()
    want = np.concatenate([[0], np.cumsum(np.maximum(np.diff(k.astype(object)), 1))]).astype(object)
    assert [int(x) for x in d] == [int(x) % 2 ** 64 for x in want]


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("mesh", ["grid", "soup", "identical"])
def test_variant_trees_are_valid_and_trace_like_brute_force(oracle, mode, mesh):
    """Index tie-break (mode 1) and 63-bit Morton keys (mode 2): every node written once, parent/child reciprocity, the
    in-order leaves are 0..n-1, node boxes contain their subtree, and the walk equals brute force in visiting order."""
    tris = {"grid": meshes.reference_scene_grid(), "soup": meshes.uniform_soup(3000, seed=8),
            "identical": np.repeat(meshes.uniform_soup(1, seed=45), 700)}[mesh]
    s = oracle.VariantScene(tris, mode)
    n = s.n
    I, L = s.internalNodes[:n - 1], s.leafNodes
    assert not (L["parent"] == 0xFFFFFFFF).any() and (L["index"] == np.arange(n)).all()
    assert (I["index"] == np.arange(n - 1)).all() and I["parent"][0] == 0xFFFFFFFF
    for side in ("left", "right"):
        c, t = I[side + "Node"], I[side + "NodeType"]
        assert (L["parent"][c[t == 1]] == np.nonzero(t == 1)[0]).all()
        inner = np.nonzero(t == 0)[0]
        assert (I["parent"][c[inner]] == inner).all()
    order = s.visit_order()
    assert sorted(order.tolist()) == sorted(s.sortedTriangleIndices.tolist())
    root = s.bvhData[0]
    assert np.allclose(root["min"], s.triangleAABB["min"].min(0)) and np.allclose(root["max"], s.triangleAABB["max"].max(0))
    extent = 4.0 if mesh == "grid" else 100.0
    rays = meshes.incoherent_rays(400, seed=12, extent=extent)
    assert s.trace_rays(rays).tobytes() == s.brute_force(rays, order=order).tobytes()
    if mode == 1:
        base = oracle.Scene(tris)
        assert np.array_equal(s.sortedTriangleIndices, base.sortedTriangleIndices)       # same sort, different tree keys
