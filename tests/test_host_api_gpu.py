"""Error behaviour and ownership rules of the boundary on a real device."""
import numpy as np
import pytest

from unitysimpleraytracing_b200 import _lib, meshes

pytestmark = pytest.mark.gpu


def test_stage_order_is_enforced(usrt):
    ctx = usrt.Context(100)
    with pytest.raises(_lib.UsrtError) as e:
        ctx.morton()
    assert e.value.code == -3
    ctx.upload_triangles(meshes.uniform_soup(50, seed=1))
    for call in (ctx.sort, ctx.distribute_keys, ctx.construct_tree, ctx.construct_bvh):
        with pytest.raises(_lib.UsrtError):
            call()
    with pytest.raises(_lib.UsrtError):
        ctx.trace_primary(4, 4, 0.3, 0.5, np.eye(4, dtype=np.float32))
    ctx.morton(); ctx.sort(); ctx.distribute_keys()
    with pytest.raises(_lib.UsrtError):
        ctx.distribute_keys()            # applying it twice would change the keys the tree is built on
    ctx.construct_tree(); ctx.construct_bvh()
    ctx.trace_primary(4, 4, 0.3, 0.5, np.eye(4, dtype=np.float32))
    ctx.close()


def test_fewer_than_two_triangles_is_rejected(usrt):
    # BVH.compute:101 with n = 1 builds nothing and traversal would read an unwritten root box
    ctx = usrt.Context(10)
    ctx.upload_triangles(meshes.uniform_soup(1, seed=2))
    with pytest.raises(_lib.UsrtError) as e:
        ctx.rebuild()
    assert e.value.code == -1
    ctx.close()


def test_capacity_is_enforced(usrt):
    ctx = usrt.Context(10)
    with pytest.raises(_lib.UsrtError):
        ctx.upload_triangles(meshes.uniform_soup(11, seed=3))
    ctx.close()


def test_two_contexts_are_independent(usrt, oracle):
    a, b = meshes.uniform_soup(3000, seed=4), meshes.uniform_soup(2000, seed=5)
    ca, cb = usrt.Context(3000), usrt.Context(2000)
    ca.upload_triangles(a); cb.upload_triangles(b)
    ca.rebuild(); cb.rebuild()
    ra, rb = oracle.Scene(a), oracle.Scene(b)
    assert ca.download(_lib.BUF_BVH_DATA, 2999).tobytes() == ra.bvhData[:2999].tobytes()
    assert cb.download(_lib.BUF_BVH_DATA, 1999).tobytes() == rb.bvhData[:1999].tobytes()
    ca.close(); cb.close()


def test_world_bounds_override(usrt, oracle):
    tris = meshes.uniform_soup(5000, seed=6, extent=10.0)
    ctx = usrt.Context(5000)
    ctx.set_world_bounds(-16.0, 16.0)
    ctx.upload_triangles(tris); ctx.morton()
    want, _, _ = oracle.morton(tris, -16.0, 16.0)
    assert np.array_equal(ctx.download(_lib.BUF_KEYS), want)
    ctx.close()


def test_stage_timing_and_launch_counter(usrt):
    ctx = usrt.Context(70000)
    ctx.upload_triangles(meshes.scene_c1())
    ctx.enable_stage_timing(True)
    before = ctx.kernel_launches
    ctx.rebuild()
    ms = ctx.last_rebuild_ms()
    assert ctx.kernel_launches - before >= 9
    assert 0 < ms["total"] < 50 and abs(sum(ms[k] for k in ("morton", "sort", "distribute", "tree", "bvh")) - ms["total"]) < 0.05
    ctx.close()


def test_runs_on_a_torch_stream_with_device_buffers(usrt, oracle):
    import torch
    tris = meshes.uniform_soup(4000, seed=8)
    ref = oracle.Scene(tris)
    dev = torch.device("cuda:0")
    ctx = usrt.Context(4000)
    s = torch.cuda.Stream()
    ctx.set_stream(s.cuda_stream)
    with torch.cuda.stream(s):
        t = torch.from_numpy(tris.view(np.uint8).reshape(-1)).to(dev)
        ctx.set_triangles_device(t.data_ptr(), len(tris))
        ctx.rebuild()
        rays = meshes.incoherent_rays(1000, seed=3)
        tr = torch.from_numpy(rays).to(dev)
        out = torch.empty(1000 * 4, dtype=torch.float32, device=dev)
        ctx.trace_rays_device(tr.data_ptr(), 1000, out.data_ptr())
    s.synchronize()
    got = out.cpu().numpy().view(oracle.RAYCAST_RESULT)
    assert got.tobytes() == ref.trace_rays(rays).tobytes()
    ctx.close()


def test_pinned_host_frame_is_written_directly_and_matches_staged_copy(usrt):
    """usrt_trace_primary writes hit records straight into page-locked host memory (zero-copy) and into
    pageable memory through a staged copy; both must hold the same frame, including partial row ranges."""
    import torch
    tris = meshes.scene_c1(); cam = meshes.SCENE_SOUP_CAMERA
    ctx = usrt.Context(len(tris)); ctx.upload_triangles(tris); ctx.rebuild()
    w, h = 321, 123
    pageable = ctx.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    pinned_t = torch.zeros(w * h * 16, dtype=torch.uint8).pin_memory()
    pinned = pinned_t.numpy().view(pageable.dtype)
    ctx.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], out=pinned)
    assert pinned.tobytes() == pageable.tobytes()
    pinned[:] = 0
    ctx.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], 40, 77, out=pinned)
    assert pinned[40 * w:77 * w].tobytes() == pageable[40 * w:77 * w].tobytes()
    assert not pinned[:40 * w].view(np.uint8).any() and not pinned[77 * w:].view(np.uint8).any()
    ctx.close()


def test_async_upload_and_trace_pipeline_over_two_contexts(usrt):
    """usrt_upload_triangles_async / usrt_trace_primary_async: frames alternating over two contexts (the e2e
    pipeline of bench.py) give the frames the blocking calls give; pageable memory is refused, not copied."""
    import torch
    from unitysimpleraytracing_b200.scene_types import RaycastResult
    cam = meshes.SCENE_SOUP_CAMERA
    scenes = [meshes.uniform_soup(20000, seed=s) for s in (1, 2, 3, 4, 5)]
    w, h = 256, 144
    want = []
    with usrt.Context(20000) as c:
        for t in scenes:
            c.upload_triangles(t); c.rebuild()
            want.append(c.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"]).tobytes())
    pin_t = [torch.from_numpy(t.view(np.uint8).reshape(-1).copy()).pin_memory() for t in scenes]
    pin_h = [torch.zeros(w * h * 16, dtype=torch.uint8).pin_memory() for _ in range(2)]
    ctxs = [usrt.Context(20000) for _ in range(2)]
    got = {}
    for i, t in enumerate(scenes):
        c = ctxs[i % 2]
        c.sync()
        if i >= 2:
            got[i - 2] = pin_h[i % 2].numpy().tobytes()
        c.upload_triangles_async(pin_t[i].numpy().view(scenes[i].dtype)); c.rebuild()
        c.trace_primary_async(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], pin_h[i % 2].numpy().view(RaycastResult))
    for i in (len(scenes) - 2, len(scenes) - 1):
        ctxs[i % 2].sync()
        got[i] = pin_h[i % 2].numpy().tobytes()
    assert [got[i] for i in range(len(scenes))] == want
    with pytest.raises(_lib.UsrtError, match="page-locked"):
        ctxs[0].upload_triangles_async(scenes[0].copy())
    with pytest.raises(_lib.UsrtError, match="page-locked"):
        ctxs[0].trace_primary_async(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], np.zeros(w * h, RaycastResult))
    for c in ctxs:
        c.close()


def test_hit_mirrors_receive_every_record(usrt):
    """usrt_set_hit_mirrors: the trace kernel stores each record to up to 8 more device frames (peer GPUs' slots
    in the multi-GPU drawer; plain device buffers here), together with the zero-copy pinned host frame."""
    import torch
    tris = meshes.scene_c1(); cam = meshes.SCENE_SOUP_CAMERA
    ctx = usrt.Context(len(tris)); ctx.upload_triangles(tris); ctx.rebuild()
    w, h = 200, 77
    want = ctx.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    mirrors = [torch.zeros(w * h * 4, dtype=torch.float32, device="cuda") for _ in range(8)]
    ctx.set_hit_mirrors([m.data_ptr() for m in mirrors])
    pinned_t = torch.zeros(w * h * 16, dtype=torch.uint8).pin_memory()
    ctx.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], out=pinned_t.numpy().view(want.dtype))
    assert pinned_t.numpy().tobytes() == want.tobytes()
    for m in mirrors:
        assert m.cpu().numpy().tobytes() == want.tobytes()
    # sharded call: same compact layout in the mirrors as in the primary output
    for m in mirrors:
        m.fill_(7.0)
    ctx.set_hit_mirrors([m.data_ptr() for m in mirrors[:3]])
    ctx.trace_primary_sharded(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], 8, 1, 3)
    ptr, cnt = ctx.hits_device()
    ctx.sync()
    own = torch.zeros(cnt * 4, dtype=torch.float32, device="cuda")
    ctx.trace_primary_sharded(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], 8, 1, 3, dev_out=own.data_ptr())
    ctx.sync()
    for m in mirrors[:3]:
        assert m[:cnt * 4].cpu().numpy().tobytes() == own.cpu().numpy().tobytes()
    assert float(mirrors[3][0]) == 7.0                      # not a mirror any more
    with pytest.raises(_lib.UsrtError):
        ctx.set_hit_mirrors([1] * 9)
    ctx.set_hit_mirrors([])
    ctx.close()


def test_rebuild_graph_replay_and_legacy_stream_fallback(usrt, oracle):
    """usrt_rebuild replays a CUDA graph; it must re-capture when n changes, survive a world-box change, and fall
    back to plain launches on a stream that cannot be captured (torch's default stream = the legacy stream)."""
    import torch
    a, b = meshes.uniform_soup(5000, seed=21), meshes.uniform_soup(3000, seed=22)
    ra, rb = oracle.Scene(a), oracle.Scene(b)
    ctx = usrt.Context(6000)
    ctx.upload_triangles(a)
    for _ in range(3):
        ctx.rebuild()
    assert ctx.download(_lib.BUF_BVH_DATA, 4999).tobytes() == ra.bvhData[:4999].tobytes()
    ctx.upload_triangles(b)                      # n changed -> re-capture
    ctx.rebuild(); ctx.rebuild()
    assert ctx.download(_lib.BUF_BVH_DATA, 2999).tobytes() == rb.bvhData[:2999].tobytes()
    assert np.array_equal(ctx.download(_lib.BUF_KEYS), rb.sortedMortonCodes)
    ctx.set_world_bounds(-64.0, 64.0)            # kernel argument changed -> graph dropped
    ctx.rebuild()
    want = oracle.Scene(b, -64.0, 64.0)
    assert np.array_equal(ctx.download(_lib.BUF_KEYS), want.sortedMortonCodes)
    ctx.close()
    ctx = usrt.Context(6000)
    ctx.use_torch_stream(torch.cuda.default_stream())     # handle 0 -> cudaStreamLegacy: not capturable
    ctx.upload_triangles(a)
    ctx.rebuild(); ctx.rebuild()
    assert ctx.download(_lib.BUF_BVH_DATA, 4999).tobytes() == ra.bvhData[:4999].tobytes()
    ctx.close()


def test_rebuild_graph_survives_a_larger_standalone_sort(usrt, oracle):
    """The captured rebuild graph points into the sort scratch; a standalone sort of more pairs than the context was
    created for re-allocates that scratch, so the next rebuild must re-capture (not replay into freed memory)."""
    tris = meshes.uniform_soup(6000, seed=21)
    ref = oracle.Scene(tris)
    ctx = usrt.Context(6000)
    ctx.upload_triangles(tris)
    ctx.rebuild(); ctx.rebuild()                                  # second call replays the graph
    rng = np.random.default_rng(5)
    k = rng.integers(0, 2 ** 32, 3_000_000, dtype=np.uint64).astype(np.uint32); v = np.arange(len(k), dtype=np.uint32)
    order = np.argsort(k, kind="stable")
    ctx.sort_pairs_host(k, v)                                     # needs 500x the status words: scratch is re-allocated
    assert np.array_equal(v, order.astype(np.uint32))
    for _ in range(3):
        ctx.rebuild()
        ctx.sync()
    assert ctx.download(_lib.BUF_BVH_DATA, 5999).tobytes() == ref.bvhData[:5999].tobytes()
    assert np.array_equal(ctx.download(_lib.BUF_KEYS), ref.sortedMortonCodes)
    assert ctx.count_corrupted_nodes() == (0, 0)
    ctx.close()


def _oracle_bvh(ref, tris):
    return dict(keys=ref.sortedMortonCodes, triangleIndex=ref.sortedTriangleIndices, triangleData=tris,
                triangleAABB=ref.triangleAABB, bvhData=ref.bvhData, leafNodes=ref.leafNodes, internalNodes=ref.internalNodes)


def test_refit_after_import_uses_rebuilt_links(usrt, oracle):
    """usrt_upload_bvh derives the K4 -> K5 parent links itself, so ConstructBVH on an imported tree (here with every
    node box wiped first) reproduces the oracle's boxes and traces identically."""
    from unitysimpleraytracing_b200 import bvh_io
    tris = meshes.uniform_soup(5000, seed=22)
    ref = oracle.Scene(tris)
    n = len(tris)
    bufs = _oracle_bvh(ref, tris)
    bufs["bvhData"] = np.zeros_like(ref.bvhData)                  # boxes are what the refit must recompute
    ctx = usrt.Context(n)
    bvh_io.upload_bvh(ctx, n, bufs)
    for _ in range(2):                                            # re-runnable
        ctx.construct_bvh()
        assert ctx.download(_lib.BUF_BVH_DATA, n - 1).tobytes() == ref.bvhData[:n - 1].tobytes()
    rays = meshes.incoherent_rays(1500, seed=23)
    assert ctx.trace_rays(rays).tobytes() == ref.trace_rays(rays).tobytes()
    ctx.close()


@pytest.mark.parametrize("damage", ["child_out_of_range", "two_parents", "cycle", "triangle_index", "bad_type", "root_as_child"])
def test_corrupt_import_is_rejected(usrt, oracle, damage):
    from unitysimpleraytracing_b200 import bvh_io
    tris = meshes.uniform_soup(300, seed=24)
    ref = oracle.Scene(tris)
    n = len(tris)
    bufs = {k: np.array(v, copy=True) for k, v in _oracle_bvh(ref, tris).items()}
    nodes = bufs["internalNodes"]
    inner = [i for i in range(1, n - 1) if nodes["leftNodeType"][i] == 0]       # nodes with an internal left child
    if damage == "child_out_of_range":
        nodes["rightNode"][5] = n + 3
    elif damage == "two_parents":
        nodes["leftNode"][inner[0]] = nodes["leftNode"][inner[1]]
    elif damage == "cycle":                                        # a node's left child becomes one of its ancestors
        i = [j for j in inner if nodes["parent"][j] != 0][-1]
        a, b = int(nodes["parent"][i]), int(nodes["leftNode"][i])
        nodes["leftNode"][i] = a
        side = "leftNode" if nodes["leftNode"][int(nodes["parent"][a])] == a and nodes["leftNodeType"][int(nodes["parent"][a])] == 0 else "rightNode"
        nodes[side][int(nodes["parent"][a])] = b                   # keeps "one parent each": only the depth walk can see it
    elif damage == "triangle_index":
        bufs["triangleIndex"][7] = 0xFFFFFFF0
    elif damage == "bad_type":
        nodes["leftNodeType"][3] = 7
    elif damage == "root_as_child":
        nodes["rightNode"][inner[0]] = 0; nodes["rightNodeType"][inner[0]] = 0
    ctx = usrt.Context(n)
    with pytest.raises(_lib.UsrtError) as e:
        bvh_io.upload_bvh(ctx, n, bufs)
    assert e.value.code == -1
    with pytest.raises(_lib.UsrtError):                            # nothing usable was installed
        ctx.trace_primary(4, 4, 0.3, 0.5, np.eye(4, dtype=np.float32))
    ctx.upload_triangles(tris); ctx.rebuild()                      # the context itself is still fine
    assert ctx.download(_lib.BUF_BVH_DATA, n - 1).tobytes() == ref.bvhData[:n - 1].tobytes()
    ctx.close()


def test_partial_trace_then_shade_sees_misses_not_garbage(usrt, oracle):
    """A freshly allocated hit buffer is filled with miss records, so shading after a row-range trace never indexes
    the triangle buffer with uninitialised triangleIndex values."""
    tris = meshes.scene_c1(); cam = meshes.SCENE_SOUP_CAMERA
    ctx = usrt.Context(len(tris)); ctx.upload_triangles(tris); ctx.rebuild()
    w, h = 80, 60
    part = ctx.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], y0=20, y1=30)
    ctx.upload_texture(np.ones((4, 4, 4), np.float32))
    img = ctx.shade()
    rows = img.reshape(h, w, 4)
    assert (rows[:20, :, 3] == 0).all() and (rows[30:, :, 3] == 0).all()          # untraced rows are misses
    full = oracle.Scene(tris).trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    assert part[20 * w:30 * w].tobytes() == full[20 * w:30 * w].tobytes()
    ctx.close()


def _positions_of(tris):
    """First 48 bytes of every Triangle as (n, 12) float32: a.xyz, pad, b.xyz, pad, c.xyz, pad."""
    return np.ascontiguousarray(tris.view(np.float32).reshape(len(tris), 32)[:, :12])


def test_positions_only_upload_is_bit_identical_to_a_full_upload(usrt, oracle):
    """usrt_upload_positions (48 B/triangle) feeds the build and the traversal exactly what they read of a Triangle;
    every buffer, hit record and bounce ray equals the full-struct path, also when the two are interleaved."""
    import torch
    a, b = meshes.uniform_soup(30000, seed=31), meshes.uniform_soup(30000, seed=32)
    ra, rb = oracle.Scene(a), oracle.Scene(b)
    cam = meshes.SCENE_SOUP_CAMERA
    ctx = usrt.Context(30000)
    ctx.upload_triangles(a); ctx.rebuild(); ctx.rebuild()                       # graph captured on the Triangle source
    for tris, ref, pinned in ((b, rb, False), (a, ra, True), (b, rb, False)):
        pos = _positions_of(tris)
        if pinned:
            keep = torch.from_numpy(pos).pin_memory(); pos = keep.numpy()
        ctx.upload_positions(pos, pinned=pinned)
        ctx.rebuild(); ctx.rebuild()
        n = len(tris)
        assert np.array_equal(ctx.download(_lib.BUF_KEYS), ref.sortedMortonCodes)
        assert ctx.download(_lib.BUF_TRIANGLE_AABB).tobytes() == ref.triangleAABB.tobytes()
        assert ctx.download(_lib.BUF_INTERNAL_NODES, n - 1).tobytes() == ref.internalNodes[:n - 1].tobytes()
        assert ctx.download(_lib.BUF_BVH_DATA, n - 1).tobytes() == ref.bvhData[:n - 1].tobytes()
        hits = ctx.trace_primary(96, 54, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
        assert hits.tobytes() == ref.trace_primary(96, 54, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], threads=4).tobytes()
        rays = torch.empty(96 * 54 * 8, dtype=torch.float32, device="cuda:0")
        ctx.diffuse_rays_device(96, 54, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], 77, 0, 1, rays.data_ptr())
        ctx.sync()
        want = oracle.diffuse_rays(hits, tris, 96, 54, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], 77, 0, 1)
        assert rays.cpu().numpy().tobytes() == want.tobytes()
    ctx.upload_triangles(a); ctx.rebuild()                                        # and back to full structs
    assert ctx.download(_lib.BUF_BVH_DATA, 29999).tobytes() == ra.bvhData[:29999].tobytes()
    with pytest.raises(_lib.UsrtError):
        ctx.upload_positions(np.zeros((30001, 12), np.float32))
    ctx.close()


def test_library_pinned_buffers_feed_the_async_entry_points(usrt, oracle):
    """usrt_host_alloc / usrt_host_free: page-locked memory behind the ABI (for hosts without a CUDA binding). The
    async upload and the zero-copy frame accept it (they refuse pageable memory), results are the usual ones."""
    tris = meshes.uniform_soup(3000, seed=21); cam = meshes.SCENE_SOUP_CAMERA
    ctx = usrt.Context(len(tris))
    buf = ctx.host_alloc(tris.nbytes)
    buf[:] = tris.view(np.uint8).reshape(-1)
    frame = ctx.host_alloc(64 * 48 * 16)
    with pytest.raises(_lib.UsrtError):
        ctx.upload_triangles_async(tris)                                   # pageable: refused, not staged
    ctx.upload_triangles_async(buf.view(tris.dtype))
    ctx.rebuild()
    hits = frame.view(usrt.RaycastResult)
    ctx.trace_primary_async(64, 48, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], hits)
    ctx.sync()
    want = oracle.Scene(tris).trace_primary(64, 48, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    assert hits.tobytes() == want.tobytes()
    del hits
    ctx.host_free(frame)
    with pytest.raises(ValueError):
        ctx.host_free(frame)                                               # already given back
    with pytest.raises(_lib.UsrtError):
        ctx.host_alloc(0)
    ctx.close()                                                            # frees `buf`, which nobody gave back
