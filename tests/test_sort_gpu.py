"""Standalone key/value sorter (ComputeBufferSorter<uint,uint>) on the GPU: exact against the oracle
at small sizes, and size-independent properties at BASELINE config-3 sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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
    if kind == "allmax":
        return np.full(n, 0xFFFFFFFF, np.uint32)
    if kind == "sorted":
        return np.sort(rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32))
    if kind == "reverse":
        return np.sort(rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32))[::-1].copy()
    raise ValueError(kind)


KINDS = ["uniform", "morton30", "low", "equal", "allmax", "sorted", "reverse"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n", [0, 1, 2, 33, 1024, 4095, 4096, 4097, 70001, 1 << 20])
def test_sort_pairs_matches_oracle(usrt, oracle, kind, n):
    keys = _keys(kind, n, seed=n + 3)
    values = np.arange(n, dtype=np.uint32)[::-1].copy()
    want_k, want_v = oracle.sort(keys, values) if n <= 70001 else oracle.stable_sort(keys, values)
    k, v = keys.copy(), values.copy()
    usrt.ComputeBufferSorter(n, k, v).Sort()
    assert np.array_equal(k, want_k)
    assert np.array_equal(v, want_v)


def test_sort_keys_only(usrt):
    keys = _keys("uniform", 300000)
    k = keys.copy()
    ctx = usrt.Context(2)
    ctx.sort_pairs_host(k, None)
    assert np.array_equal(k, np.sort(keys))
    ctx.close()


@pytest.mark.parametrize("log2n,kind", [(24, "uniform"), (26, "uniform"), (26, "morton30"), (24, "low")])
def test_sort_properties_at_scale(usrt, log2n, kind):
    """2^24..2^26 pairs: sortedness, stability, permutation (checksum of checksums) -- no O(n log n) CPU."""
    n = 1 << log2n
    keys = _keys(kind, n, seed=log2n)
    values = np.arange(n, dtype=np.uint32)
    k, v = keys.copy(), values.copy()
    ctx = usrt.Context(2)
    ctx.sort_pairs_host(k, v)
    ctx.close()
    dk = np.diff(k.astype(np.int64))
    assert (dk >= 0).all()                                            # ValidateSortedData (:150-177)
    assert (np.diff(v.astype(np.int64))[dk == 0] > 0).all()           # stable: equal keys keep input order
    assert np.array_equal(keys[v], k)                                 # each pair travelled together
    assert int(v.astype(np.uint64).sum()) == n * (n - 1) // 2         # values are a permutation ...
    assert int((v.astype(np.uint64) ** 2 % 1000003).sum()) == int((values.astype(np.uint64) ** 2 % 1000003).sum())
    assert np.array_equal(np.bincount(k >> 24, minlength=256), np.bincount(keys >> 24, minlength=256))


@pytest.mark.parametrize("n", [444 * 6144 - 1, 444 * 6144, 444 * 6144 + 1, 2 * 444 * 6144 + 6143, (1 << 22) - 1, (1 << 22) + 6145])
@pytest.mark.parametrize("kind", ["uniform", "low"])
def test_persistent_tile_loop_boundaries(usrt, n, kind):
    """The pass kernel's CTAs are persistent (148 SMs x 3 CTAs = 444 tiles of 6,144 pairs at a time): counts right at and
    around whole multiples of the resident tile set, ragged last tiles, and both sides of the 2^22-pair switch to
    interleaved records -- exact against a stable argsort."""
    keys = _keys(kind, n, seed=n & 0xFFFF)
    values = np.arange(n, dtype=np.uint32)
    order = np.argsort(keys, kind="stable")
    k, v = keys.copy(), values.copy()
    ctx = usrt.Context(2)
    ctx.sort_pairs_host(k, v)
    kk = keys.copy()
    ctx.sort_pairs_host(kk)                                            # keys only
    ctx.close()
    assert np.array_equal(k, keys[order]) and np.array_equal(v, order.astype(np.uint32))
    assert np.array_equal(kk, keys[order])


@pytest.mark.parametrize("log2n", [28, 30])
def test_device_sort_at_config3_sizes(usrt, log2n):
    """BASELINE config 3 beyond 2^26: 2^28 and 2^30 pairs sorted in place on the device (2^30 is the first size that
    uses the 64-bit look-back words), checked on the device: ascending keys, equal keys keep their input order (the
    values are the input positions), every pair travelled together, the values are a permutation of 0..n-1."""
    import torch
    n = 1 << log2n
    g = torch.Generator(device="cuda"); g.manual_seed(log2n)
    keys = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g)
    keys[::5] >>= 20                                                   # plenty of duplicates: stability is exercised
    orig = keys.clone()
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    ctx = usrt.Context(2)
    ctx.sort_pairs_device(keys.data_ptr(), vals.data_ptr(), n)
    ctx.sync()
    chunk = 1 << 26
    total = 0
    for a in range(0, n, chunk):                                       # chunked so that temporaries stay small
        b = min(n, a + chunk + 1)
        k = keys[a:b].to(torch.int64) & 0xFFFFFFFF                     # the sorter orders keys as unsigned
        v = vals[a:b].to(torch.int64) & 0xFFFFFFFF
        assert bool((k[1:] >= k[:-1]).all())
        assert bool(((v[1:] > v[:-1]) | (k[1:] != k[:-1])).all())
        assert bool((orig[v[:-1] if b < n else v] == keys[a:b - 1 if b < n else b]).all())
        total += int((v[:-1] if b < n else v).sum())
    assert total == n * (n - 1) // 2
    seen = torch.zeros(n, dtype=torch.bool, device="cuda")
    seen[vals.to(torch.int64) & 0xFFFFFFFF] = True
    assert bool(seen.all())
    ctx.close()


def test_partition_pass_is_stable_split_by_top_byte(usrt):
    import torch
    n = 200003
    keys = _keys("uniform", n)
    values = np.arange(n, dtype=np.uint32)
    dev = torch.device("cuda:0")
    tk = torch.from_numpy(keys.view(np.int32)).to(dev); tv = torch.from_numpy(values.view(np.int32)).to(dev)
    ok = torch.empty_like(tk); ov = torch.empty_like(tv); hist = torch.zeros(256, dtype=torch.int32, device=dev)
    ctx = usrt.Context(2)
    torch.cuda.synchronize()
    ctx.partition_pass_device(tk.data_ptr(), tv.data_ptr(), ok.data_ptr(), ov.data_ptr(), n, 24, hist.data_ptr())
    ctx.sync()
    order = np.argsort(keys >> 24, kind="stable")
    assert np.array_equal(ok.cpu().numpy().view(np.uint32), keys[order])
    assert np.array_equal(ov.cpu().numpy().view(np.uint32), values[order])
    assert np.array_equal(hist.cpu().numpy(), np.bincount(keys >> 24, minlength=256))
    ctx.close()


def test_sort_with_unaligned_device_pointers(usrt):
    """Caller buffers need not be 16-byte aligned (the histogram has a scalar head/tail around its 128-bit body)."""
    import torch
    dev = torch.device("cuda:0")
    ctx = usrt.Context(2)
    for off in (1, 2, 3):
        n = 100003
        keys = _keys("uniform", n, seed=off)
        tk = torch.zeros(n + 8, dtype=torch.int32, device=dev); tv = torch.zeros(n + 8, dtype=torch.int32, device=dev)
        tk[off:off + n] = torch.from_numpy(keys.view(np.int32)).to(dev)
        tv[off:off + n] = torch.arange(n, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        ctx.sort_pairs_device(tk.data_ptr() + 4 * off, tv.data_ptr() + 4 * off, n)
        ctx.sync()
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(tk[off:off + n].cpu().numpy().view(np.uint32), keys[order])
        assert np.array_equal(tv[off:off + n].cpu().numpy().view(np.uint32), order.astype(np.uint32))
        assert int(tk[:off].abs().sum()) == 0 and int(tk[off + n:].abs().sum()) == 0      # nothing outside the range touched
    ctx.close()


def test_wide_look_back_words_path():
    """Sorts of >= 2^30 pairs use 64-bit look-back words; exercise that kernel variant at a small size through
    the USRT_FORCE_WIDE_STATUS test hook (fresh process: the hook is read once)."""
    import subprocess, sys, os
    code = (
        "import numpy as np, sys; sys.path.insert(0, %r)\n"
        "from unitysimpleraytracing_b200 import host\n"
        "rng = np.random.default_rng(3); n = 300001\n"
        "k = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32); k[::3] = k[0]; v = np.arange(n, dtype=np.uint32)\n"
        "o = np.argsort(k, kind='stable'); kk, vv = k.copy(), v.copy()\n"
        "c = host.Context(2); c.sort_pairs_host(kk, vv); c.close()\n"
        "assert np.array_equal(kk, k[o]) and np.array_equal(vv, v[o]); print('wide ok')\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, USRT_FORCE_WIDE_STATUS="1"), capture_output=True, text=True)
    assert out.returncode == 0 and "wide ok" in out.stdout, out.stderr[-2000:]


def test_fast_ranking_path_is_the_one_that_runs():
    """The pass kernel ranks with one returning shared-memory atomic per key and falls back to an order-independent
    ballot ranking when its per-CTA canary fails. A canary that fails for the wrong reason (e.g. a compiler that
    splits the warp before it) is silent: results stay correct, passes get ~1.8x slower. So time a 2^24-pair sort
    in a fresh process with and without USRT_FORCE_SLOW_RANK (the hook that forces the fallback): the default must
    be clearly faster, i.e. it is NOT running the fallback."""
    import subprocess, sys, os
    code = (
        "import sys, torch; sys.path.insert(0, %r)\n"
        "from unitysimpleraytracing_b200 import host\n"
        "n = 1 << 24; g = torch.Generator(device='cuda'); g.manual_seed(5)\n"
        "k0 = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device='cuda', generator=g)\n"
        "c = host.Context(2); s = torch.cuda.Stream(); c.set_stream(s.cuda_stream); ts = []\n"
        "with torch.cuda.stream(s):\n"
        "    for it in range(8):\n"
        "        k = k0.clone(); v = torch.arange(n, dtype=torch.int32, device='cuda')\n"
        "        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)\n"
        "        a.record(s); c.sort_pairs_device(k.data_ptr(), v.data_ptr(), n); b.record(s); torch.cuda.synchronize()\n"
        "        ts.append(a.elapsed_time(b))\n"
        "assert bool((k[1:] >= k[:-1]).all())\n"
        "print('ms', sorted(ts[3:])[2]); c.close()\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def run(extra):
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **extra), capture_output=True, text=True)
        assert out.returncode == 0 and "ms " in out.stdout, out.stderr[-2000:]
        return float(out.stdout.split("ms ")[1].split()[0])

    fast, slow = run({}), run({"USRT_FORCE_SLOW_RANK": "1"})
    assert slow > 1.25 * fast, "default sort %.3f ms vs forced fallback %.3f ms: the fast ranking path is not being taken" % (fast, slow)


def _exact_bucket_ranges(global_hist, world):
    """choose_bucket_ranges in exact integer arithmetic (the rule k_peer_scatter_plan implements)."""
    csum = [0]
    for v in global_hist:
        csum.append(csum[-1] + int(v))
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r
        b = 0
        while b < 256 and csum[b] * world < target:
            b += 1
        if b > 0 and target - csum[b - 1] * world <= abs(csum[b] * world - target):
            b -= 1
        bounds.append(min(max(b, bounds[-1]), 256))
    bounds.append(256)
    return bounds


@pytest.mark.parametrize("world,kind", [(2, "uniform"), (4, "skewed"), (8, "uniform"), (8, "sparse"), (3, "empty_rank")])
def test_peer_scatter_plan_kernel_matches_the_host_plan(usrt, world, kind):
    """The landing plan of the multi-GPU bucket exchange, computed on ONE GPU for every rank of a pretend world and
    compared with dist.peer_scatter_plan (numpy): same bucket ranges, same landing addresses, same receive counts."""
    import torch
    from unitysimpleraytracing_b200 import dist as udist
    rng = np.random.default_rng(world * 7 + len(kind))
    if kind == "uniform":
        h = rng.integers(0, 5000, (world, 256))
    elif kind == "skewed":
        h = (rng.random((world, 256)) ** 6 * 90000).astype(np.int64)
    elif kind == "sparse":
        h = np.zeros((world, 256), np.int64); h[:, rng.integers(0, 256, 5)] = rng.integers(1, 10 ** 6, (world, 5))
    else:
        h = rng.integers(0, 3000, (world, 256)); h[1] = 0
    h = h.astype(np.uint32)
    bounds = _exact_bucket_ranges(h.sum(0, dtype=np.int64), world)
    owner, offset, recv_total = udist.peer_scatter_plan(h, bounds)
    capacity = int(recv_total.max()) + 5
    base = np.array([(r + 1) << 40 for r in range(world)], np.int64)           # pretend mappings, never dereferenced
    ctx = usrt.Context(2)
    dev = torch.device("cuda:0")
    ctx.use_torch_stream()
    d_h = torch.from_numpy(h.view(np.int32).reshape(-1)).to(dev)
    d_base = torch.from_numpy(base).to(dev)
    for rank in range(world):
        ptrs = torch.zeros(512, dtype=torch.int64, device=dev); recv = torch.zeros(world, dtype=torch.int64, device=dev)
        bnd = torch.zeros(world + 1, dtype=torch.int32, device=dev)
        ctx.peer_scatter_plan_device(d_h.data_ptr(), world, rank, d_base.data_ptr(), capacity, ptrs.data_ptr(),
                                     ptrs.data_ptr() + 2048, recv.data_ptr(), bnd.data_ptr())
        torch.cuda.synchronize()
        assert bnd.cpu().numpy().tolist() == bounds
        assert np.array_equal(recv.cpu().numpy(), recv_total)
        p = ptrs.cpu().numpy()
        assert np.array_equal(p[:256], base[owner] + 4 * offset[rank])
        assert np.array_equal(p[256:], base[owner] + 4 * (capacity + offset[rank]))
    # a receive buffer one pair too small: the plan is all null addresses (the scatter then writes nothing)
    ptrs = torch.ones(512, dtype=torch.int64, device=dev); recv = torch.zeros(world, dtype=torch.int64, device=dev)
    ctx.peer_scatter_plan_device(d_h.data_ptr(), world, 0, d_base.data_ptr(), int(recv_total.max()) - 1, ptrs.data_ptr(),
                                 ptrs.data_ptr() + 2048, recv.data_ptr())
    torch.cuda.synchronize()
    assert (ptrs.cpu().numpy() == 0).all() and np.array_equal(recv.cpu().numpy(), recv_total)
    ctx.close()
