"""Developer tool (GPU): per-kernel device times of the key/value sort at several sizes."""
import argparse, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unitysimpleraytracing_b200 import host

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="*", default=[20, 22, 24, 26, 28])
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--kind", default="uniform")
a = ap.parse_args()
dev = torch.device("cuda:0")
ctx = host.Context(2)
ctx.enable_stage_timing(True)
peak = 6550.1
for lg in a.n:
    n = 1 << lg
    g = torch.Generator(device=dev); g.manual_seed(lg)
    hi = 2 ** 31 - 1 if a.kind == "uniform" else 2 ** 30 - 1
    lo = -2 ** 31 if a.kind == "uniform" else 0
    k0 = torch.randint(lo, hi, (n,), dtype=torch.int32, device=dev, generator=g)
    v0 = torch.arange(n, dtype=torch.int32, device=dev)
    k = torch.empty_like(k0); v = torch.empty_like(v0)
    rec = []
    for i in range(a.warm + a.iters):
        k.copy_(k0); v.copy_(v0)
        torch.cuda.synchronize()
        ctx.sort_pairs_device(k.data_ptr(), v.data_ptr(), n)
        t = ctx.last_sort_ms()
        if i >= a.warm:
            rec.append(t)
    m = {key: statistics.mean(r[key] for r in rec) for key in rec[0]}
    ku = k.view(torch.uint8)  # cheap sortedness check in signed-int32 space is wrong for uniform; use int64 view trick
    kk = k.to(torch.int64) & 0xFFFFFFFF
    ok = bool((kk[1:] >= kk[:-1]).all().item())
    gbs = 68.0 * n / (m["total"] * 1e-3) / 1e9
    print("2^%d %s: total %.4f ms  hist %.4f  passes %.4f %.4f %.4f %.4f | %.1f Mkeys/s  %.0f GB/s  %.1f%% of %.0f  sorted=%s"
          % (lg, a.kind, m["total"], m["histogram"], m["pass0"], m["pass8"], m["pass16"], m["pass24"],
             n / (m["total"] * 1e-3) / 1e6, gbs, 100 * gbs / peak, peak, ok), flush=True)
    del k0, v0, k, v, kk
ctx.close()
