"""torchrun tool: BASELINE config 3 beyond one GPU -- 2^k (key,value) pairs in total, split evenly over the
ranks, sorted by MSD buckets + all-to-all (unitysimpleraytracing_b200/dist.py). Prints Mkeys/s (max over ranks).

    python -m torch.distributed.run --nproc-per-node N tools/dist_sort_bench.py --log2 28 30
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from unitysimpleraytracing_b200 import host, dist as udist

ap = argparse.ArgumentParser(); ap.add_argument("--log2", type=int, nargs="*", default=[28]); ap.add_argument("--iters", type=int, default=3); ap.add_argument("--exchange", choices=["nccl", "peer"], default="nccl")
a = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1: dist.init_process_group("nccl", device_id=dev)
else: dist.init_process_group("gloo", rank=0, world_size=1, init_method="tcp://127.0.0.1:29533")
ctx = host.Context(2, device=lr)
for lg in a.log2:
    n = (1 << lg) // world
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    k0 = torch.randint(-2 ** 31, 2 ** 31 - 1, (n,), dtype=torch.int32, device=dev, generator=g)
    v0 = torch.arange(n, dtype=torch.int32, device=dev) + rank * n
    times = []
    ex = udist.PeerSortExchange(ctx, int(n * 1.25) + 4096) if (world > 1 and a.exchange == "peer") else None
    for it in range(a.iters + 1):
        k = k0.clone(); v = v0.clone()
        torch.cuda.synchronize(); 
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if world > 1:
            rk, rv = ex.sort(k, v) if ex else udist.dist_sort_pairs(k, v, ctx=ctx)
        else:
            ctx.use_torch_stream(); ctx.sort_pairs_device(k.data_ptr(), v.data_ptr(), n); rk, rv = k, v
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0: times.append(float(t.item()))
        kk = rk.to(torch.int64) & 0xFFFFFFFF
        ok = bool((kk[1:] >= kk[:-1]).all().item()) if kk.numel() > 1 else True
        lo = int(kk[0].item()) if kk.numel() else -1; hi = int(kk[-1].item()) if kk.numel() else -1
        del k, v, kk
    # global order across ranks: my first key >= previous rank's last key
    edges = torch.tensor([lo, hi], dtype=torch.int64, device=dev)
    if world > 1:
        allv = [torch.empty_like(edges) for _ in range(world)]; dist.all_gather(allv, edges)
        glob = all(int(allv[r][0]) >= int(allv[r - 1][1]) for r in range(1, world))
        cnt = torch.tensor([rk.numel()], dtype=torch.int64, device=dev); dist.all_reduce(cnt)
    else:
        glob, cnt = True, torch.tensor([rk.numel()])
    if rank == 0:
        best = min(times)
        print("dist sort [" + a.exchange + "] 2^%d pairs on %d GPU(s): %.3f ms  %.0f Mkeys/s  locally sorted=%s globally ordered=%s total=%d"
              % (lg, world, best * 1e3, (1 << lg) / best / 1e6, ok, glob, int(cnt.item())), flush=True)
    del k0, v0, rk, rv
    if ex: ex.close()
    torch.cuda.empty_cache()
ctx.close()
dist.destroy_process_group()
