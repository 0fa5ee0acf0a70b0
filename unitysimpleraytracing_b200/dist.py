"""Multi-GPU use of the path, one process per GPU over torch.distributed (NCCL on NVLink/NVSwitch;
gloo for the CPU tests of the host logic). Only the two places where the path shards naturally
(BASELINE.json north_star):

 (a) ray sharding  -- every rank holds a replica of the BVH (the build is deterministic, so each rank
     simply rebuilds it; no broadcast needed), traces an interleaved set of row blocks in one launch
     and the hit records are all-gathered: the ONLY collective on the path.
 (b) sorts of >= 2^28 pairs -- keys are split into MSD buckets (top byte), bucket ranges of roughly
     equal mass are assigned to ranks, pairs are exchanged all-to-all and sorted locally. Global order
     = rank order; stable across ranks because sources are concatenated in rank order and both the
     bucket split and the local sort are stable.

Morton / DistributeKeys / tree / refit do not shard without a global exchange: replicas only.

The local GPU work is injected (`local_*` callables) so the host logic can be exercised on CPU under
gloo; the defaults call the CUDA library. Nothing here falls back to a CPU path by itself.
"""
import numpy as np

from .scene_types import RaycastResult


# ==================================================================================================
# (a) ray sharding
# ==================================================================================================
def shard_layout(height, num_shards, block_rows=8):
    """(blocks_per_shard, local_rows) of usrt_trace_primary_sharded."""
    blocks = -(-height // block_rows)
    per = -(-blocks // num_shards)
    return per, per * block_rows


def frame_rows_of_shard(height, shard, num_shards, block_rows=8):
    """Frame row of every local row of `shard` (-1 for padding rows past the frame)."""
    per, local_rows = shard_layout(height, num_shards, block_rows)
    lr = np.arange(local_rows)
    y = ((lr // block_rows) * num_shards + shard) * block_rows + lr % block_rows
    return np.where(y < height, y, -1)


def assemble_frame(gathered, width, height, num_shards, block_rows=8):
    """gathered: (num_shards, local_rows * width) hit records as all-gathered -> (height * width) in
    frame order (record index y * width + x). Works on numpy structured arrays or (.., 4) float views."""
    per, local_rows = shard_layout(height, num_shards, block_rows)
    g = gathered.reshape(num_shards, local_rows, width, *gathered.shape[2:]) if gathered.ndim > 2 else \
        gathered.reshape(num_shards, local_rows, width)
    out = np.zeros((height, width) + g.shape[3:], g.dtype)
    for s in range(num_shards):
        rows = frame_rows_of_shard(height, s, num_shards, block_rows)
        ok = rows >= 0
        out[rows[ok]] = g[s][ok]
    return out.reshape((height * width,) + g.shape[3:])


class RayShardedDrawer:
    """RaytracingMeshDrawer across GPUs: Awake() builds the replica on this rank's GPU, Update() traces
    this rank's row blocks and all-gathers the records so every rank ends with the whole frame."""

    def __init__(self, mesh, rank, world, device=None, block_rows=8, group=None):
        from . import host
        self.rank, self.world, self.block_rows, self.group = rank, world, block_rows, group
        self.device = rank if device is None else device
        self.drawer = host.RaytracingMeshDrawer(mesh, device=self.device)

    def Awake(self):
        self.drawer.Awake()
        return self

    @property
    def ctx(self):
        return self.drawer.container.ctx

    def Update(self, width, height, near, cameraFov, cameraToWorldMatrix):
        """Returns (height*width) RaycastResult in frame order on every rank (numpy, host)."""
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", self.device)
        per, local_rows = shard_layout(height, self.world, self.block_rows)
        local = torch.empty(local_rows * width * 4, dtype=torch.float32, device=dev)
        self.ctx.use_torch_stream()
        self.ctx.trace_primary_sharded(width, height, near, cameraFov, cameraToWorldMatrix, self.block_rows,
                                       self.rank, self.world, dev_out=local.data_ptr())
        self.ctx.sync()
        if self.world > 1:
            gathered = torch.empty(self.world * local.numel(), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(gathered, local, group=self.group)
        else:
            gathered = local
        g = gathered.cpu().numpy().view(RaycastResult).reshape(self.world, local_rows * width)
        return assemble_frame(g, width, height, self.world, self.block_rows)

    def OnDestroy(self):
        self.drawer.OnDestroy()


# ==================================================================================================
# (b) distributed key/value sort by MSD buckets
# ==================================================================================================
def choose_bucket_ranges(global_hist, world):
    """Split the 256 top-byte buckets into `world` contiguous ranges of ~equal mass.
    Returns `bounds` (world+1 ints): rank r owns buckets [bounds[r], bounds[r+1])."""
    h = np.asarray(global_hist, np.int64)
    total = int(h.sum())
    csum = np.concatenate([[0], np.cumsum(h)])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(csum, target, side="left"))
        # csum[b] >= target; pick the closer of b-1 / b, but keep bounds non-decreasing
        if b > 0 and abs(csum[b - 1] - target) <= abs(csum[min(b, 256)] - target):
            b -= 1
        bounds.append(min(max(b, bounds[-1]), 256))
    bounds.append(256)
    return bounds


def dist_sort_pairs(keys_t, vals_t, group=None, local_partition=None, local_sort=None, ctx=None):
    """Sort the concatenation (in rank order) of every rank's (keys_t, vals_t) -- int32/uint32 torch
    tensors holding uint32 bit patterns, on this rank's device. Returns (keys, vals) of this rank's
    slice of the global result: rank r holds the r-th contiguous chunk of the sorted sequence.

    local_partition(keys, vals) -> (keys_by_bucket, vals_by_bucket, hist256): stable split by top byte.
    local_sort(keys, vals) -> None: stable in-place sort. Defaults call the CUDA library via `ctx`."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = keys_t.numel()
    dev = keys_t.device

    if local_partition is None or local_sort is None:
        if ctx is None:
            raise ValueError("dist_sort_pairs needs a usrt Context (ctx=) for the CUDA local steps")

        def local_partition(k, v):                                        # noqa: F811
            ok, ov = torch.empty_like(k), torch.empty_like(v)
            hist = torch.zeros(256, dtype=torch.int32, device=k.device)
            ctx.use_torch_stream()
            ctx.partition_pass_device(k.data_ptr(), v.data_ptr(), ok.data_ptr(), ov.data_ptr(), k.numel(), 24,
                                      hist.data_ptr())
            return ok, ov, hist.to(torch.int64)

        def local_sort(k, v):                                             # noqa: F811
            ctx.use_torch_stream()
            ctx.sort_pairs_device(k.data_ptr(), v.data_ptr(), k.numel())

    # 1. stable local split by the top byte + its histogram
    pk, pv, hist = local_partition(keys_t, vals_t)
    hist = hist.to(torch.int64)
    # 2. every rank learns every rank's histogram (256 x world int64: tiny)
    all_hist = torch.empty(world * 256, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_hist, hist.contiguous(), group=group)
    all_hist = all_hist.view(world, 256).cpu().numpy()
    # 3. contiguous bucket ranges of ~equal mass, identical on every rank
    bounds = choose_bucket_ranges(all_hist.sum(0), world)
    send_counts = [int(all_hist[rank, bounds[r]:bounds[r + 1]].sum()) for r in range(world)]
    recv_counts = [int(all_hist[s, bounds[rank]:bounds[rank + 1]].sum()) for s in range(world)]
    assert sum(send_counts) == n
    # 4. all-to-all of the pairs (the one data-path collective)
    rk = torch.empty(sum(recv_counts), dtype=keys_t.dtype, device=dev)
    rv = torch.empty(sum(recv_counts), dtype=vals_t.dtype, device=dev)
    dist.all_to_all_single(rk, pk, recv_counts, send_counts, group=group)
    dist.all_to_all_single(rv, pv, recv_counts, send_counts, group=group)
    # 5. stable local sort of what arrived (sources are concatenated in rank order)
    local_sort(rk, rv)
    return rk, rv
