"""Host-side logic of the multi-GPU paths under gloo, world_size 2 and 3, on CPU. The per-rank GPU
work is replaced by numpy stand-ins defined HERE (tests only); the product's defaults call CUDA."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unitysimpleraytracing_b200 import dist as udist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _np_partition(k, v):
    ku = k.numpy().view(np.uint32); vu = v.numpy().view(np.uint32)
    order = np.argsort(ku >> 24, kind="stable")
    hist = np.bincount(ku >> 24, minlength=256).astype(np.int64)
    return (torch.from_numpy(ku[order].view(np.int32).copy()), torch.from_numpy(vu[order].view(np.int32).copy()),
            torch.from_numpy(hist))


def _np_sort(k, v):
    ku = k.numpy().view(np.uint32); vu = v.numpy().view(np.uint32)
    order = np.argsort(ku, kind="stable")
    k.copy_(torch.from_numpy(ku[order].view(np.int32).copy())); v.copy_(torch.from_numpy(vu[order].view(np.int32).copy()))


def _sort_worker(rank, world, port, kind, n_per_rank, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    n = n_per_rank + 37 * rank                      # ragged
    if kind == "uniform":
        keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    elif kind == "morton30":
        keys = rng.integers(0, 2 ** 30, n, dtype=np.uint64).astype(np.uint32)
    elif kind == "skew":                            # almost everything in one top-byte bucket
        keys = (rng.integers(0, 2 ** 20, n, dtype=np.uint64) | (7 << 24)).astype(np.uint32)
        keys[:10] = rng.integers(0, 2 ** 32, 10, dtype=np.uint64).astype(np.uint32)
    else:                                           # duplicates: cross-rank stability matters
        keys = (rng.integers(0, 50, n, dtype=np.uint64) << 22).astype(np.uint32)
    base = sum(n_per_rank + 37 * r for r in range(rank))
    vals = (np.arange(n) + base).astype(np.uint32)  # global original index
    k, v = udist.dist_sort_pairs(torch.from_numpy(keys.view(np.int32).copy()), torch.from_numpy(vals.view(np.int32).copy()),
                                 local_partition=_np_partition, local_sort=_np_sort)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), k=k.numpy().view(np.uint32), v=v.numpy().view(np.uint32),
             ik=keys, iv=vals)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("kind", ["uniform", "morton30", "skew", "dups"])
def test_dist_sort_equals_global_stable_sort(tmp_path, world, kind):
    port = _free_port()
    mp.spawn(_sort_worker, args=(world, port, kind, 5000, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / ("r%d.npz" % r)) for r in range(world)]
    all_k = np.concatenate([p["ik"] for p in parts]); all_v = np.concatenate([p["iv"] for p in parts])
    order = np.argsort(all_k, kind="stable")
    got_k = np.concatenate([p["k"] for p in parts]); got_v = np.concatenate([p["v"] for p in parts])
    assert np.array_equal(got_k, all_k[order])
    assert np.array_equal(got_v, all_v[order])      # globally stable (values = global original index)
    if kind in ("uniform", "morton30"):             # balanced to within a bucket
        sizes = [len(p["k"]) for p in parts]
        assert max(sizes) < 1.3 * len(all_k) / world


def test_choose_bucket_ranges():
    h = np.zeros(256, np.int64); h[:64] = 100
    b = udist.choose_bucket_ranges(h, 4)
    assert b == [0, 16, 32, 48, 256]
    h = np.zeros(256, np.int64); h[7] = 1000
    b = udist.choose_bucket_ranges(h, 3)
    assert b[0] == 0 and b[-1] == 256 and all(x <= y for x, y in zip(b, b[1:]))
    assert udist.choose_bucket_ranges(np.zeros(256), 2) == [0, 0, 256]


@pytest.mark.parametrize("world", [1, 2, 5])
@pytest.mark.parametrize("kind", ["uniform", "morton30", "onebucket"])
def test_peer_scatter_plan_lands_a_stable_exchange(world, kind):
    """Replay the plan of PeerSortExchange with numpy writes: every run lands in a disjoint slot, the
    owner's buffer is source-rank-major, and a stable local sort of it gives the global stable sort."""
    rng = np.random.default_rng(world)
    n = 3000
    keys = [rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32) for _ in range(world)]
    if kind == "morton30":
        keys = [k >> np.uint32(2) for k in keys]
    if kind == "onebucket":
        keys = [(k & np.uint32(0x00FFFFFF)) | np.uint32(0x21000000) for k in keys]
    vals = [np.arange(n, dtype=np.uint32) + r * n for r in range(world)]
    all_hist = np.stack([np.bincount(k >> 24, minlength=256) for k in keys])
    bounds = udist.choose_bucket_ranges(all_hist.sum(0), world)
    owner, offset, recv_total = udist.peer_scatter_plan(all_hist, bounds)
    assert recv_total.sum() == world * n
    bufk = [np.full(int(t), 0xDEADBEEF, np.uint32) for t in recv_total]
    bufv = [np.zeros(int(t), np.uint32) for t in recv_total]
    written = [np.zeros(int(t), bool) for t in recv_total]
    for s in range(world):
        order = np.argsort(keys[s] >> 24, kind="stable")          # the stable partition pass
        pk, pv = keys[s][order], vals[s][order]
        pos = 0
        for d in range(256):
            c = int(all_hist[s, d])
            o, at = int(owner[d]), int(offset[s, d])
            assert not written[o][at:at + c].any()
            bufk[o][at:at + c], bufv[o][at:at + c] = pk[pos:pos + c], pv[pos:pos + c]
            written[o][at:at + c] = True
            pos += c
    assert all(w.all() for w in written)
    got_k, got_v = [], []
    for o in range(world):
        order = np.argsort(bufk[o], kind="stable")
        got_k.append(bufk[o][order]); got_v.append(bufv[o][order])
    all_k, all_v = np.concatenate(keys), np.concatenate(vals)
    order = np.argsort(all_k, kind="stable")
    assert np.array_equal(np.concatenate(got_k), all_k[order])
    assert np.array_equal(np.concatenate(got_v), all_v[order])


def _frame_worker(rank, world, port, W, H, block_rows, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # stand-in for the per-rank trace: record (y, x) so placement is checkable
    rows = udist.frame_rows_of_shard(H, rank, world, block_rows)
    local = np.zeros((len(rows), W, 4), np.float32)
    for lr, y in enumerate(rows):
        if y >= 0:
            local[lr, :, 0] = y; local[lr, :, 1] = np.arange(W); local[lr, :, 2] = rank
    t = torch.from_numpy(local.reshape(-1))
    g = torch.empty(world * t.numel(), dtype=torch.float32)
    dist.all_gather_into_tensor(g, t)
    frame = udist.assemble_frame(g.numpy().reshape(world, -1, 4), W, H, world, block_rows)
    np.save(os.path.join(out_dir, "f%d.npy" % rank), frame)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,H,block_rows", [(2, 1080, 8), (3, 50, 4), (2, 17, 8)])
def test_ray_shard_gather_and_assembly(tmp_path, world, H, block_rows):
    W = 12
    port = _free_port()
    mp.spawn(_frame_worker, args=(world, port, W, H, block_rows, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        f = np.load(tmp_path / ("f%d.npy" % r)).reshape(H, W, 4)
        assert np.array_equal(f[:, :, 0], np.repeat(np.arange(H, dtype=np.float32)[:, None], W, 1))
        assert np.array_equal(f[:, :, 1], np.repeat(np.arange(W, dtype=np.float32)[None, :], H, 0))
        owner = (np.arange(H) // block_rows) % world
        assert np.array_equal(f[:, 0, 2], owner.astype(np.float32))


def test_shard_rows_cover_the_frame_exactly_once():
    for H, world, br in [(1080, 8, 8), (2160, 4, 8), (50, 3, 4), (7, 2, 8)]:
        seen = np.concatenate([udist.frame_rows_of_shard(H, s, world, br) for s in range(world)])
        seen = np.sort(seen[seen >= 0])
        assert np.array_equal(seen, np.arange(H))
        assert len({udist.shard_layout(H, world, br) for _ in range(2)}) == 1


class _FakePeerCtx:
    """Stands in for the CUDA context in map_peer_buffers: handles are the rank number, mapping can be made to fail."""

    def __init__(self, rank, fail_create_on=None, fail_open_on=None):
        self.rank, self.fail_create_on, self.fail_open_on = rank, fail_create_on, fail_open_on
        self.live = set()

    def peer_buffer_create(self, nbytes):
        if self.rank == self.fail_create_on:
            raise RuntimeError("out of memory")
        self.live.add(("own", self.rank))
        return 1000 + self.rank, bytes([self.rank]) * 64

    def peer_buffer_open(self, handle):
        if self.rank == self.fail_open_on:
            raise RuntimeError("no peer access")
        self.live.add(("peer", handle[0]))
        return 2000 + handle[0]

    def peer_buffer_close(self, ptr, opened):
        self.live.discard(("peer", ptr - 2000) if opened else ("own", ptr - 1000))


def _peer_map_worker(rank, world, port, mode, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _FakePeerCtx(rank, fail_create_on=1 if mode == "create" else None, fail_open_on=2 if mode == "open" else None)
    try:
        own, peers = udist.map_peer_buffers(ctx, 4096, device="cpu")
        res = ("ok", own, peers)
    except RuntimeError as e:
        res = ("raised", str(e), sorted(ctx.live))
    np.save(os.path.join(out_dir, "m%d.npy" % rank), np.array([repr(res)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["fine", "create", "open"])
def test_peer_buffer_mapping_is_collectively_consistent(tmp_path, mode):
    """map_peer_buffers: every rank gets every rank's buffer, or -- when ANY rank cannot allocate or map -- every
    rank raises, nothing stays allocated or mapped, and no rank is left waiting in a collective."""
    world = 3
    mp.spawn(_peer_map_worker, args=(world, _free_port(), mode, str(tmp_path)), nprocs=world, join=True)
    res = [eval(str(np.load(tmp_path / ("m%d.npy" % r))[0])) for r in range(world)]
    if mode == "fine":
        for r, (tag, own, peers) in enumerate(res):
            assert tag == "ok" and own == 1000 + r
            assert peers == [1000 + q if q == r else 2000 + q for q in range(world)]
    else:
        assert all(t[0] == "raised" and t[2] == [] for t in res), res
