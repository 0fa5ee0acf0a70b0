"""The two sharded paths on real GPUs. With >= 2 visible devices: one rank per GPU over NCCL (gpurun --gpus 2). With ONE
device the peer-memory forms still run for real -- two processes share the GPU, map each other's buffers through CUDA IPC
and the kernels store into the other process's memory exactly as they do over NVLink; only the control messages go over
gloo instead of NCCL (which refuses two ranks on one device). The NCCL all-gather / all-to-all baselines need 2 GPUs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _need_gpus(k):
    import torch
    if torch.cuda.device_count() < k:
        pytest.skip("needs %d GPUs, %d visible" % (k, torch.cuda.device_count()))


def _shared_device():
    """True: fewer GPUs than ranks -- every rank uses device 0 and gloo carries the control messages."""
    import torch
    return torch.cuda.device_count() < 2


def _init(rank, world, port, shared):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    device = 0 if shared else rank
    torch.cuda.set_device(device)
    if shared:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    return device


def _ray_worker(rank, world, port, out_dir, exchange="peer", shared=False):
    import torch.distributed as dist
    from unitysimpleraytracing_b200 import dist as udist, meshes
    device = _init(rank, world, port, shared)
    tris = meshes.scene_c1()
    cam = meshes.SCENE_SOUP_CAMERA
    d = udist.RayShardedDrawer(tris, rank, world, device=device, exchange=exchange).Awake()
    d.Update(160, 90, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])       # buffers are re-made / reused
    frame = d.Update(320, 180, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    frame = d.Update(320, 180, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    np.save(os.path.join(out_dir, "frame%d.npy" % rank), frame)
    d.OnDestroy()
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_ray_sharded_frame_matches_oracle(tmp_path, oracle, exchange):
    """peer: the trace kernel stores every record into both ranks' frames (peer memory: NVLink between two GPUs, CUDA IPC
    between two processes on one GPU); nccl: all-gather (two GPUs only)."""
    shared = _shared_device()
    if exchange == "nccl":
        _need_gpus(2)
    import torch.multiprocessing as mp
    from unitysimpleraytracing_b200 import meshes
    world = 2
    mp.spawn(_ray_worker, args=(world, _free_port(), str(tmp_path), exchange, shared), nprocs=world, join=True)
    cam = meshes.SCENE_SOUP_CAMERA
    want = oracle.Scene(meshes.scene_c1()).trace_primary(320, 180, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], threads=8)
    for r in range(world):
        got = np.load(tmp_path / ("frame%d.npy" % r))
        assert got.tobytes() == want.tobytes()


def _sort_worker(rank, world, port, n, out_dir):
    import torch
    import torch.distributed as dist
    from unitysimpleraytracing_b200 import dist as udist, host
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    rng = np.random.default_rng(7 + rank)
    keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    keys[::5] = keys[0]                                   # duplicates across ranks: stability matters
    vals = (np.arange(n) + rank * n).astype(np.uint32)
    ctx = host.Context(2, device=rank)
    k, v = udist.dist_sort_pairs(torch.from_numpy(keys.view(np.int32)).to(dev), torch.from_numpy(vals.view(np.int32)).to(dev), ctx=ctx)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, "s%d.npz" % rank), k=k.cpu().numpy().view(np.uint32), v=v.cpu().numpy().view(np.uint32), ik=keys, iv=vals)
    ctx.close()
    dist.destroy_process_group()


def test_dist_sort_matches_global_stable_sort(tmp_path):
    _need_gpus(2)
    import torch.multiprocessing as mp
    world, n = 2, 1 << 20
    mp.spawn(_sort_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / ("s%d.npz" % r)) for r in range(world)]
    all_k = np.concatenate([p["ik"] for p in parts]); all_v = np.concatenate([p["iv"] for p in parts])
    order = np.argsort(all_k, kind="stable")
    assert np.array_equal(np.concatenate([p["k"] for p in parts]), all_k[order])
    assert np.array_equal(np.concatenate([p["v"] for p in parts]), all_v[order])


def _peer_sort_worker(rank, world, port, n, rounds, out_dir, shared=False):
    import torch
    import torch.distributed as dist
    from unitysimpleraytracing_b200 import dist as udist, host
    device = _init(rank, world, port, shared)
    dev = torch.device("cuda", device)
    ctx = host.Context(2, device=device)
    ex = udist.PeerSortExchange(ctx, int(n * 1.5) + 4096)
    for it in range(rounds):                              # buffers are reused: the second round checks the fences
        rng = np.random.default_rng(100 * it + 7 + rank)
        keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        keys[::5] = keys[0]
        if it == 1:
            keys &= np.uint32(0x3FFFFFFF)                  # Morton-like: top byte < 64
        vals = (np.arange(n) + rank * n).astype(np.uint32)
        k, v = ex.sort(torch.from_numpy(keys.view(np.int32)).to(dev), torch.from_numpy(vals.view(np.int32)).to(dev))
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, "p%d_%d.npz" % (it, rank)), k=k.cpu().numpy().view(np.uint32),
                 v=v.cpu().numpy().view(np.uint32), ik=keys, iv=vals)
    ex.close()
    ctx.close()
    dist.destroy_process_group()


def _check_peer_sort(tmp_path, world, rounds):
    for it in range(rounds):
        parts = [np.load(tmp_path / ("p%d_%d.npz" % (it, r))) for r in range(world)]
        all_k = np.concatenate([p["ik"] for p in parts]); all_v = np.concatenate([p["iv"] for p in parts])
        order = np.argsort(all_k, kind="stable")
        assert np.array_equal(np.concatenate([p["k"] for p in parts]), all_k[order])
        assert np.array_equal(np.concatenate([p["v"] for p in parts]), all_v[order])


@pytest.mark.parametrize("n", [1000, (1 << 20) + 77])
def test_peer_scatter_sort_single_rank(tmp_path, n):
    """world = 1: the per-digit-address partition kernel and the landing plan, on one GPU."""
    import torch.multiprocessing as mp
    mp.spawn(_peer_sort_worker, args=(1, _free_port(), n, 2, str(tmp_path)), nprocs=1, join=True)
    _check_peer_sort(tmp_path, 1, 2)


def test_peer_scatter_sort_matches_global_stable_sort(tmp_path):
    """The partition pass writing into the other rank's receive buffer (CUDA IPC: over NVLink between two GPUs, or into
    the other process's memory on one GPU)."""
    shared = _shared_device()
    import torch.multiprocessing as mp
    world, n = 2, (1 << 20) + 13
    mp.spawn(_peer_sort_worker, args=(world, _free_port(), n, 2, str(tmp_path), shared), nprocs=world, join=True)
    _check_peer_sort(tmp_path, world, 2)


def test_peer_scatter_sort_three_ranks_on_whatever_is_there(tmp_path):
    """Three ranks (an odd world: uneven bucket ranges): three GPUs if visible, otherwise three processes on one."""
    import torch
    import torch.multiprocessing as mp
    shared = torch.cuda.device_count() < 3
    world, n = 3, 300007
    mp.spawn(_peer_sort_worker, args=(world, _free_port(), n, 2, str(tmp_path), shared), nprocs=world, join=True)
    _check_peer_sort(tmp_path, world, 2)
