"""The pin of the oracle: oracle/usrt_oracle.cpp (hand-written restatement) must agree BIT FOR BIT with
oracle/_ref/libusrt_ref.so, which is the reference's own text -- BVH.compute, Raytracing.compute and the static
functions + DistributeKeys of MeshBufferContainer.cs -- compiled with g++ by oracle/build_ref.sh.

Runs wherever the _ref library exists: in the build container it is (re)built from /root/reference; on the GPU box the
prebuilt .so travels with the snapshot. The committed digests (tests/golden/ref_digests.json) were produced by the same
library, so the last test pins the oracle to them even when neither is present."""
import hashlib
import json
import os

import numpy as np
import pytest

from unitysimpleraytracing_b200 import meshes

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BUFS = ("mortonCodes", "triangleAABB", "sortedMortonRaw", "sortedTriangleIndices", "sortedMortonCodes", "leafNodes")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ref():
    from oracle import usrt_ref
    if not usrt_ref.available():
        pytest.skip("oracle/_ref/libusrt_ref.so not built and no reference checkout here")
    return usrt_ref


def _mesh(name):
    if name == "soup2":
        return meshes.uniform_soup(2, seed=41)
    if name == "soup3":
        return meshes.uniform_soup(3, seed=42)
    if name == "soup4097":
        return meshes.uniform_soup(4097, seed=43)
    if name == "refgrid":                # the reference's own scene mesh: ~11 triangles per Morton cell
        return meshes.reference_scene_grid()
    if name == "sphere":
        return meshes.sphere(96, 192)
    if name == "outside":                # centroids beyond the +-125 world box clamp to 0 / 1023
        return meshes.uniform_soup(20000, seed=44, extent=190.0)
    if name == "identical":              # every key equal: DistributeKeys spreads them 0,1,2,...
        return np.repeat(meshes.uniform_soup(1, seed=45), 3001)
    if name == "degenerate":
        t = meshes.uniform_soup(5000, seed=46)
        t["b"][::3] = t["a"][::3]
        t["c"][::7] = t["a"][::7]
        return t
    raise ValueError(name)


@pytest.mark.parametrize("name", ["soup2", "soup3", "soup4097", "refgrid", "sphere", "outside", "identical", "degenerate"])
def test_build_buffers_match_the_reference_text(oracle, ref, name):
    tris = _mesh(name)
    o, r = oracle.Scene(tris), ref.Scene(tris)
    n = o.n
    for f in BUFS:
        assert np.ascontiguousarray(getattr(o, f)).tobytes() == np.ascontiguousarray(getattr(r, f)).tobytes(), f
    assert o.internalNodes[:n - 1].tobytes() == r.internalNodes[:n - 1].tobytes()      # TreeConstructor, BVH.compute:94-149
    assert o.bvhData[:n - 1].tobytes() == r.bvhData[:n - 1].tobytes()                  # BVHConstructor, :172-220
    # the root keeps its NullLeaf parent and no NullLeaf entry is left (MeshBufferContainer.GetAllGpuData :181-195)
    assert r.internalNodes["parent"][0] == 0xFFFFFFFF
    assert not ((r.leafNodes["index"] == 0xFFFFFFFF) & (r.leafNodes["parent"] == 0xFFFFFFFF)).any()


@pytest.mark.parametrize("name,cam,w,h", [("refgrid", "REFERENCE_CAMERA", 160, 90), ("sphere", "SCENE_C2_CAMERA", 128, 72),
                                          ("soup4097", "SCENE_SOUP_CAMERA", 96, 96), ("degenerate", "SCENE_SOUP_CAMERA", 64, 64),
                                          ("identical", "SCENE_SOUP_CAMERA", 48, 48)])
def test_hit_records_match_the_raytracing_kernel(oracle, ref, name, cam, w, h):
    tris = _mesh(name)
    c = getattr(meshes, cam)
    o, r = oracle.Scene(tris), ref.Scene(tris)
    want, _ = r.trace_primary(w, h, c["near"], c["tan_half_fov"], c["cam_to_world"])
    got = o.trace_primary(w, h, c["near"], c["tan_half_fov"], c["cam_to_world"], threads=8)
    assert got.tobytes() == want.tobytes()
    assert (want["distance"] != oracle.max_float()).any() or name == "identical"


def test_config0_full_frame_matches_the_raytracing_kernel(oracle, ref):
    """BASELINE configs[0] as stated: 65,536-triangle soup, 512x512 primary rays -- every record of the frame."""
    tris = meshes.scene_c1(); c = meshes.SCENE_SOUP_CAMERA
    o, r = oracle.Scene(tris), ref.Scene(tris)
    want, _ = r.trace_primary(512, 512, c["near"], c["tan_half_fov"], c["cam_to_world"])
    got = o.trace_primary(512, 512, c["near"], c["tan_half_fov"], c["cam_to_world"], threads=8)
    assert got.tobytes() == want.tobytes()


def test_shading_epilogue_matches_the_raytracing_kernel(oracle, ref):
    """Raytracing.compute:178-184 incl. the scalar `lightDir` (:181) and the float4 -> float3 truncation (:183)."""
    tris = meshes.sphere(48, 96); c = meshes.SCENE_C2_CAMERA
    rng = np.random.default_rng(3)
    tex = rng.random((48, 64, 4), dtype=np.float32)
    r = ref.Scene(tris)
    hits, rgba = r.trace_primary(120, 68, c["near"], c["tan_half_fov"], c["cam_to_world"], texture=tex)
    want = rgba.astype(np.float16)                                   # the R16G16B16A16_SFloat target (RaytracingMeshDrawer.cs:56)
    got = oracle.shade(hits, tris, tex)
    assert got.view(np.uint16).tobytes() == want.view(np.uint16).tobytes()
    assert (want[:, 3] == 1).any() and (want[:, 3] == 0).any()


def test_oracle_reproduces_reference_generated_digests(oracle):
    """tests/golden/ref_digests.json was written by make_ref_golden.py from the reference's own code; the oracle must
    land on the same sha256 for every buffer and every full frame, up to configs[1] at full size (1,048,576 triangles,
    1920x1080 rays)."""
    d = json.load(open(os.path.join(HERE, "ref_digests.json")))
    for name, make, cam, (w, h) in (("refgrid_12800", meshes.reference_scene_grid, meshes.REFERENCE_CAMERA, (480, 270)),
                                    ("config0_soup_65536", meshes.scene_c1, meshes.SCENE_SOUP_CAMERA, (512, 512)),
                                    ("config1_scene_1048576", meshes.scene_c2, meshes.SCENE_C2_CAMERA, (1920, 1080))):
        tris = make()
        assert sha(tris) == d[name]["triangles"], "mesh generator output changed: regenerate the goldens in the build container"
        s = oracle.Scene(tris)
        n = s.n
        for f in BUFS:
            assert sha(getattr(s, f)) == d[name][f], (name, f)
        assert sha(s.internalNodes[:n - 1]) == d[name]["internalNodes"], name
        assert sha(s.bvhData[:n - 1]) == d[name]["bvhData"], name
        frame = s.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], threads=os.cpu_count() or 1)
        assert sha(frame) == d[name]["primary_%dx%d" % (w, h)], name


# ---- the reference's five sort kernels, run as written under the lock-step wave emulator -----------------------------
def _keys(kind, n, seed):
    rng = np.random.default_rng(seed)
    k = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    if kind == "morton30":
        k >>= np.uint32(2)
    elif kind == "few":
        k = ((k % 7) << np.uint32(21)).astype(np.uint32)
    elif kind == "equal":
        k[:] = 0x12345678
    elif kind == "padded":                    # real keys followed by the 0xFFFFFFFF padding of MeshBufferContainer.cs:108-109
        k[n - n // 3:] = 0xFFFFFFFF
    return k


@pytest.mark.parametrize("kind", ["uniform", "morton30", "few", "equal", "padded"])
@pytest.mark.parametrize("bit_offset", [0, 8, 16, 24])
def test_sort_pass_intermediates_match_the_sort_kernels(oracle, ref, kind, bit_offset):
    """LocalRadixSort -> PreScan / BlockSum / GlobalScan -> GlobalRadixSort: block-sorted keys / values, per-block digit
    offsets, the digit-major count table before and after the scan, and the scattered output, all byte for byte."""
    n = 8192
    k = _keys(kind, n, 100 + bit_offset); v = np.arange(n, dtype=np.uint32)
    r, o = ref.sort_pass(k, v, bit_offset), oracle.sort_pass(k, v, bit_offset)
    nb = n // 1024
    for f in ("sortedBlocksKeys", "sortedBlocksValues", "keys", "values"):
        assert np.array_equal(r[f], o[f]), f
    assert np.array_equal(r["offsets"][:nb * 256], o["offsets"])
    # the reference lays the count table out for its fixed 512 blocks (sizes[digit * 512 + block]), the oracle for nb blocks
    assert np.array_equal(r["sizesBefore"].reshape(256, 512)[:, :nb], o["sizesBefore"].reshape(256, nb))
    assert np.array_equal(r["sizesAfter"].reshape(256, 512)[:, :nb], o["sizesAfter"].reshape(256, nb))
    assert not r["sizesBefore"].reshape(256, 512)[:, nb:].any()


@pytest.mark.parametrize("kind,n", [("uniform", 65536), ("morton30", 20000), ("few", 5000), ("equal", 3000), ("uniform", 1)])
def test_full_sort_matches_the_sort_kernels(oracle, ref, kind, n):
    k = _keys(kind, n, n); v = np.arange(n, dtype=np.uint32)
    rk, rv = ref.sort(k, v)
    ok, ov = oracle.sort(k, v)
    assert np.array_equal(rk, ok) and np.array_equal(rv, ov)
    assert np.array_equal(rv, np.argsort(k, kind="stable").astype(np.uint32))      # the net contract: a stable sort


def test_one_pass_at_the_reference_capacity(oracle, ref):
    """All 512 blocks x 1024 elements, the only size the reference itself ever runs (Constants.cs:3-6)."""
    n = 512 * 1024
    k = _keys("morton30", n, 7); k[400000:] = 0xFFFFFFFF                            # a 400,000-triangle mesh, padded
    v = np.arange(n, dtype=np.uint32); v[400000:] = 0xFFFFFFFF
    r, o = ref.sort_pass(k, v, 8), oracle.sort_pass(k, v, 8)
    assert np.array_equal(r["keys"], o["keys"]) and np.array_equal(r["values"], o["values"])
    assert np.array_equal(r["sizesAfter"].reshape(256, 512), o["sizesAfter"].reshape(256, 512))
