"""Multi-GPU use of the path, one process per GPU over torch.distributed (NCCL on NVLink/NVSwitch;
gloo for the CPU tests of the host logic). Only the two places where the path shards naturally
(BASELINE.json north_star):

 (a) ray sharding  -- every rank holds a replica of the BVH (the build is deterministic, so each rank
     simply rebuilds it; no broadcast needed), traces an interleaved set of row blocks in one launch
     and the hit records are all-gathered: the ONLY collective on the path.
 (b) sorts of >= 2^28 pairs -- keys are split into MSD buckets (top byte), bucket ranges of roughly
     equal mass are assigned to ranks, pairs are exchanged all-to-all and sorted locally. Global order
     = rank order; stable across ranks because sources are concatenated in rank order and both the
     bucket split and the local sort are stable. Two exchanges: `dist_sort_pairs` (local split, then
     NCCL all-to-all) and `PeerSortExchange` (the split pass itself scatters into the owners' receive
     buffers through NVLink peer memory -- one kernel is both the split and the all-to-all).

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


def _all_gather_into(out, mine, group=None):
    """dist.all_gather_into_tensor -- or, on a backend without a CUDA all-gather (gloo: the single-GPU form of the GPU
    tests, several processes sharing one device and talking through CUDA IPC), the same result from an all-reduce of a
    buffer in which every rank fills its own slot. Only used for the small control messages (handles, histograms)."""
    import torch.distributed as dist
    if not (mine.is_cuda and dist.get_backend(group) == "gloo"):
        dist.all_gather_into_tensor(out, mine, group=group)
        return
    r, m = dist.get_rank(group), mine.numel()
    flat = out.view(-1)
    flat.zero_()
    flat[r * m:(r + 1) * m] = mine.reshape(-1)
    dist.all_reduce(out, group=group)


def map_peer_buffers(ctx, nbytes, group=None, device=None):
    """Allocate `nbytes` on this rank's GPU and map every other rank's buffer of the same call into this process
    (CUDA IPC over torch.distributed). Returns (own_ptr, [ptr of rank 0's buffer, ...]) with own_ptr at index
    rank. Collective, and collectively consistent: if any rank cannot allocate or map (no peer access, separate
    IPC namespaces), every rank releases what it holds and raises RuntimeError -- no rank is left in a collective."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    on_gpu = torch.device(device).type == "cuda"

    def everyone(ok):
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        return bool(flag.item())

    own_ptr, handle, err = 0, bytes(64), None
    try:
        own_ptr, handle = ctx.peer_buffer_create(nbytes)
    except Exception as e:                                             # noqa: BLE001
        err = e
    if not everyone(err is None):
        if own_ptr:
            ctx.peer_buffer_close(own_ptr, False)
        raise RuntimeError("map_peer_buffers: allocation failed on some rank (this rank: %s)" % err)
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
    allh = torch.empty(world * 64, dtype=torch.uint8, device=device)
    _all_gather_into(allh, mine, group=group)
    allh = allh.cpu().numpy().reshape(world, 64)
    peer_ptr = [0] * world
    peer_ptr[rank] = own_ptr
    try:
        for r in range(world):
            if r != rank:
                peer_ptr[r] = ctx.peer_buffer_open(allh[r].tobytes())
    except Exception as e:                                             # noqa: BLE001
        err = e
    if not everyone(err is None):
        for r in range(world):
            if r != rank and peer_ptr[r]:
                ctx.peer_buffer_close(peer_ptr[r], True)
        if on_gpu:
            torch.cuda.synchronize()
        dist.barrier(group=group)                                      # nobody frees a buffer a peer still maps
        ctx.peer_buffer_close(own_ptr, False)
        raise RuntimeError("map_peer_buffers: CUDA IPC mapping failed on some rank (this rank: %s)" % err)
    return own_ptr, peer_ptr


class PeerFrameExchange:
    """The hit-record all-gather done by the trace kernel's own stores (usrt.h: usrt_set_hit_mirrors). Every
    rank owns `buffers` frames of world x slot_records records, mapped into every other rank through CUDA IPC;
    select(b) points this rank's trace calls at slot `rank` of frame b on EVERY rank (own copy included), so when
    all ranks' kernels have finished -- fence() -- every rank holds the whole frame. No data-path collective."""

    def __init__(self, ctx, slot_records, buffers=2, group=None):
        import torch
        import torch.distributed as dist
        self.ctx, self.group, self.slot, self.buffers = ctx, group, int(slot_records), buffers
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("PeerFrameExchange: at most 8 ranks (one NVSwitch domain)")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.frame_bytes = self.world * self.slot * 16
        self.own_ptr, self.peer_ptr = map_peer_buffers(ctx, self.frame_bytes * buffers, group)
        self._fence = torch.zeros(1, dtype=torch.int32, device=self.device)

    def slot_ptr(self, b, on_rank=None):
        """Address (in this process) of this rank's slot of frame b in the buffer owned by `on_rank`."""
        base = self.peer_ptr[self.rank if on_rank is None else on_rank]
        return base + b * self.frame_bytes + self.rank * self.slot * 16

    def select(self, b, include_own=True):
        self.ctx.set_hit_mirrors([self.slot_ptr(b, r) for r in range(self.world) if include_own or r != self.rank])

    def fence(self):
        """Enqueue (current torch stream) the point after which every rank's stores have landed everywhere."""
        import torch.distributed as dist
        dist.all_reduce(self._fence, group=self.group)

    def frame(self, b):
        """Frame b of this rank as a float32 torch tensor view [world * slot_records * 4]."""
        import torch
        return torch.as_tensor(_DeviceView(self.own_ptr + b * self.frame_bytes, self.world * self.slot * 4, "<f4"),
                               device=self.device)

    def close(self):
        import torch
        torch.cuda.synchronize()
        self.ctx.set_hit_mirrors([])
        for r, p in enumerate(self.peer_ptr):
            if r != self.rank and p:
                self.ctx.peer_buffer_close(p, True)
        if self.own_ptr:
            import torch.distributed as dist
            dist.barrier(group=self.group)             # collective: every peer has unmapped this rank's buffer
            self.ctx.peer_buffer_close(self.own_ptr, False)
        self.peer_ptr, self.own_ptr = [], 0


class RayShardedDrawer:
    """RaytracingMeshDrawer across GPUs: Awake() builds the replica on this rank's GPU, Update() traces
    this rank's row blocks and every rank ends with the whole frame. exchange="peer" (default): the trace
    kernel stores each record into every rank's frame over NVLink (PeerFrameExchange); "nccl": trace, then
    all-gather."""

    def __init__(self, mesh, rank, world, device=None, block_rows=8, group=None, exchange="peer"):
        from . import host
        self.rank, self.world, self.block_rows, self.group = rank, world, block_rows, group
        self.device = rank if device is None else device
        self.exchange, self._peer = exchange, None
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
        self.ctx.use_torch_stream()
        if self.world > 1 and self.exchange == "peer":
            if self._peer is None or self._peer.slot != local_rows * width:
                if self._peer is not None:
                    self._peer.close()
                self._peer = PeerFrameExchange(self.ctx, local_rows * width, buffers=1, group=self.group)
            px = self._peer
            dist.barrier(group=self.group)                     # nobody still reads the previous frame
            px.select(0, include_own=False)
            self.ctx.trace_primary_sharded(width, height, near, cameraFov, cameraToWorldMatrix, self.block_rows,
                                           self.rank, self.world, dev_out=px.slot_ptr(0))
            px.fence()
            g = px.frame(0).cpu().numpy().view(RaycastResult).reshape(self.world, local_rows * width)
            return assemble_frame(g, width, height, self.world, self.block_rows)
        local = torch.empty(local_rows * width * 4, dtype=torch.float32, device=dev)
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
        if self._peer is not None:
            self._peer.close()
            self._peer = None
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
    _all_gather_into(all_hist, hist.contiguous(), group=group)
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


def peer_scatter_plan(all_hist, bounds):
    """Where every (source rank, top-byte value) run lands. all_hist: (world, 256) counts; bounds from
    choose_bucket_ranges. Returns (owner[256], offset[world][256], recv_total[world]): source s writes its
    run of digit d at element `offset[s][d]` of rank `owner[d]`'s receive buffer. The layout inside an owner
    is source-rank-major, digits ascending within a source -- the order the all-to-all of dist_sort_pairs
    produces -- so a stable local sort of the buffer gives the globally stable result."""
    h = np.asarray(all_hist, np.int64)
    world = h.shape[0]
    b = np.asarray(bounds, np.int64)
    owner = np.repeat(np.arange(world), np.diff(b))                       # [256]
    C = np.zeros((world, 257), np.int64)
    np.cumsum(h, axis=1, out=C[:, 1:])                                    # C[s][j] = sum of h[s][:j]
    T = C[:, b[1:]] - C[:, b[:-1]]                                        # T[s][o]: pairs source s sends to owner o
    before = np.cumsum(T, axis=0) - T                                     # from the lower-ranked sources
    offset = before[:, owner] + C[:, :256] - C[:, b[owner]]
    recv_total = T.sum(axis=0)
    return owner, offset, recv_total


class _DeviceView:
    """A library-owned device range as a __cuda_array_interface__ object (for torch.as_tensor)."""

    def __init__(self, ptr, count, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerSortExchange:
    """Multi-GPU sort whose bucket exchange is fused into the partition kernel (usrt.h:
    usrt_partition_scatter_device). Every rank owns a receive buffer of `capacity` pairs, mapped into every
    other rank through CUDA IPC at construction. sort():
      1. top-byte counts of the local keys (one kernel), all-gathered (world x 256 counters);
      2. identical bucket ranges + landing addresses on every rank, computed ON THE DEVICE from the gathered
         counters (usrt_peer_scatter_plan_device; peer_scatter_plan below is its numpy twin for the CPU tests);
      3. ONE partition pass that writes each pair straight into its owner's receive buffer over NVLink;
      4. a one-element all-reduce as the "all scatters landed" fence; 5. stable local 4-pass sort.
    Collectives carry only counters; the pairs never pass through NCCL, and the host reads back nothing but the
    per-rank receive counts (world x 8 bytes, copied while the scatter runs) -- it needs its own for step 5."""

    def __init__(self, ctx, capacity, group=None):
        import torch
        import torch.distributed as dist
        self.ctx, self.group, self.capacity = ctx, group, int(capacity)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.own_ptr, self.peer_ptr = map_peer_buffers(ctx, self.capacity * 8, group)       # keys | values
        self._fence = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._hist = torch.zeros(256, dtype=torch.int32, device=self.device)
        self._all_hist = torch.empty(self.world * 256, dtype=torch.int32, device=self.device)
        self._peer_base = torch.tensor([int(p) for p in self.peer_ptr], dtype=torch.int64, device=self.device)
        self._ptrs = torch.empty(512, dtype=torch.int64, device=self.device)               # 256 key + 256 value addresses
        self._recv = torch.empty(self.world, dtype=torch.int64, device=self.device)
        self._recv_host = torch.empty(self.world, dtype=torch.int64).pin_memory()
        self._recv_ready = torch.cuda.Event()
        self.last_recv_total = None

    def sort(self, keys_t, vals_t):
        """-> (keys, vals): this rank's contiguous chunk of the global stable sort, as views of the receive
        buffer (valid until the next sort())."""
        import torch
        import torch.distributed as dist
        n = keys_t.numel()
        ctx = self.ctx
        ctx.use_torch_stream()
        ctx.digit_histogram_device(keys_t.data_ptr(), n, 24, self._hist.data_ptr())
        # also the "receive buffers are free again" barrier: a rank gets here only after its previous local sort
        _all_gather_into(self._all_hist, self._hist, group=self.group)
        ctx.peer_scatter_plan_device(self._all_hist.data_ptr(), self.world, self.rank, self._peer_base.data_ptr(),
                                     self.capacity, self._ptrs.data_ptr(), self._ptrs.data_ptr() + 256 * 8,
                                     self._recv.data_ptr())
        self._recv_host.copy_(self._recv, non_blocking=True)      # 8 bytes per rank, in flight beside the scatter
        self._recv_ready.record()
        ctx.partition_scatter_device(keys_t.data_ptr(), vals_t.data_ptr(), n, 24, self._ptrs.data_ptr(),
                                     self._ptrs.data_ptr() + 256 * 8)
        dist.all_reduce(self._fence, group=self.group)           # every rank's scatter has completed
        self._recv_ready.synchronize()
        recv_total = self._recv_host.numpy().copy()
        self.last_recv_total = recv_total
        if int(recv_total.max()) > self.capacity:
            # every rank sees the same counts and raises together; what the scatter wrote past the buffers' value halves
            # stayed inside the 2 x capacity allocation only if max <= capacity, hence the headroom callers leave
            raise ValueError(f"receive buffer of {self.capacity} pairs too small for {int(recv_total.max())}")
        m = int(recv_total[self.rank])
        kptr, vptr = self.own_ptr, self.own_ptr + 4 * self.capacity
        if m:
            ctx.sort_pairs_device(kptr, vptr, m)
        if m == 0:
            return keys_t.new_empty(0), vals_t.new_empty(0)
        rk = torch.as_tensor(_DeviceView(kptr, m), device=self.device)
        rv = torch.as_tensor(_DeviceView(vptr, m), device=self.device)
        return rk, rv

    def close(self):
        import torch
        torch.cuda.synchronize()
        for r, p in enumerate(self.peer_ptr):
            if r != self.rank and p:
                self.ctx.peer_buffer_close(p, True)
        if self.own_ptr:
            import torch.distributed as dist
            dist.barrier(group=self.group)             # collective: every peer has unmapped this rank's buffer
            self.ctx.peer_buffer_close(self.own_ptr, False)
        self.peer_ptr, self.own_ptr = [], 0
