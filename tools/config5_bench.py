"""torchrun tool: BASELINE config 5 shape -- a large soup (default 2^26 triangles, generated ON the GPU so no
64 GB of host memory is needed) with the BVH replicated per rank (each rank rebuilds it: deterministic), and
`rays_total` random incoherent rays sharded evenly across ranks; hit records all-gathered in chunks.
Prints aggregate Mrays/s (max time over ranks).

    python -m torch.distributed.run --nproc-per-node N tools/config5_bench.py --log2tris 26 --rays 33177600
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from unitysimpleraytracing_b200 import host

ap = argparse.ArgumentParser()
ap.add_argument("--log2tris", type=int, default=26); ap.add_argument("--rays", type=int, default=3840 * 2160 * 4)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--diffuse", type=int, default=0, help="BASELINE configs[4] proper: 3840x2160 primary frame, then this many "
                "diffuse bounce rays per pixel (usrt_diffuse_rays_device), samples sharded across the ranks")
a = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1: dist.init_process_group("nccl", device_id=dev)
n = 1 << a.log2tris
g = torch.Generator(device=dev); g.manual_seed(0x5EED0006)            # same scene on every rank
edge = 200.0 / n ** (1.0 / 3.0) * 0.75
centre = (torch.rand((n, 1, 3), device=dev, generator=g) * 2 - 1) * 100.0
verts = centre + (torch.rand((n, 3, 3), device=dev, generator=g) * 2 - 1) * edge
tri = torch.zeros((n, 32), dtype=torch.float32, device=dev)            # 128-byte Triangle: a,b,c in float4 slots 0..2
tri[:, 0:3] = verts[:, 0]; tri[:, 4:7] = verts[:, 1]; tri[:, 8:11] = verts[:, 2]
del centre, verts
ctx = host.Context(n, device=lr); ctx.use_torch_stream()
ctx.set_triangles_device(tri.data_ptr(), n); ctx.enable_stage_timing(True)
ctx.rebuild(); ctx.rebuild(); st = ctx.last_rebuild_ms()
leaf_bad, int_bad = ctx.count_corrupted_nodes()
if a.diffuse:
    import numpy as np
    from unitysimpleraytracing_b200 import meshes
    Wd, Hd = 3840, 2160
    cam = meshes.SCENE_SOUP_CAMERA
    cargs = (Wd, Hd, cam["near"], cam["tan_half_fov"], np.asarray(cam["cam_to_world"], np.float32))
    frame = Wd * Hd
    spp_rank = a.diffuse // world                      # rank r casts samples [r * spp_rank, (r + 1) * spp_rank)
    ctx.set_trace_mode(a.mode)
    prim = torch.empty(frame * 4, dtype=torch.float32, device=dev)
    rays = torch.empty(frame * 8, dtype=torch.float32, device=dev)
    out = torch.empty(spp_rank * frame * 4, dtype=torch.float32, device=dev)
    gathered = torch.empty(world * out.numel(), dtype=torch.float32, device=dev) if world > 1 else None   # [sample][rank][pixel]
    side = torch.cuda.Stream() if world > 1 else None

    def run():
        ctx.trace_primary(*cargs, download=False)       # every rank: the primary frame against its BVH replica
        p_ptr, cnt = ctx.hits_device()
        prim.copy_(torch.as_tensor(type("V", (), {"__cuda_array_interface__": {"shape": (cnt * 4,), "typestr": "<f4",
                   "data": (p_ptr, False), "version": 2}})(), device=dev))
        for s_ in range(spp_rank):                      # one sample pass = one frame of bounce rays
            ctx.diffuse_rays_device(*cargs, 0x5EED0008, rank * spp_rank + s_, 1, rays.data_ptr(), primary_hits_ptr=prim.data_ptr())
            ctx.trace_rays_device(rays.data_ptr(), frame, out.data_ptr() + s_ * frame * 16)
            if world > 1:                               # the output gather (16 B per ray), pass by pass beside the next pass
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    dist.all_gather_into_tensor(gathered[s_ * world * frame * 4:(s_ + 1) * world * frame * 4],
                                                out[s_ * frame * 4:(s_ + 1) * frame * 4])
        if world > 1:
            torch.cuda.current_stream().wait_stream(side)

    spp_save, spp_rank = spp_rank, 1
    run(); torch.cuda.synchronize(); spp_rank = spp_save
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    run()
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    t = torch.tensor([t_all], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total = frame * (1 + spp_rank * world)             # the primary frame counted once (every rank traces it)
    hit_p = float((prim.view(-1, 4)[:, 0] != 2139095040.0).float().mean().item())
    hit_d = float((out.view(-1, 4)[:, 0] != 2139095040.0).float().mean().item())
    if rank == 0:
        print("config5 diffuse: %d tris, build %.2f ms, corrupted=%s | 3840x2160 primary + %d spp diffuse over %d GPU(s): "
              "%d rays in %.1f ms incl. gather -> %.1f Mrays/s aggregate (mode %d); primary hit fraction %.3f, bounce hit fraction %.3f"
              % (n, st["total"], (leaf_bad, int_bad), spp_rank * world, world, total, float(t[0]) * 1e3, total / float(t[0]) / 1e6,
                 a.mode, hit_p, hit_d), flush=True)
    ctx.close()
    if world > 1: dist.destroy_process_group()
    sys.exit(0)
per = a.rays // world
g2 = torch.Generator(device=dev); g2.manual_seed(77 + rank)
rays = torch.zeros((per, 8), dtype=torch.float32, device=dev)
rays[:, 0:3] = (torch.rand((per, 3), device=dev, generator=g2) * 2 - 1) * 100.0
d = torch.randn((per, 3), device=dev, generator=g2); rays[:, 4:7] = d / d.norm(dim=1, keepdim=True)
out = torch.empty(per * 4, dtype=torch.float32, device=dev)
ctx.set_trace_mode(a.mode)
ctx.trace_rays_device(rays.data_ptr(), min(per, 1 << 16), out.data_ptr()); torch.cuda.synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
ctx.trace_rays_device(rays.data_ptr(), per, out.data_ptr())
torch.cuda.synchronize(); t_trace = time.perf_counter() - t0
gathered = None
if world > 1:
    gathered = torch.empty(world * out.numel(), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(gathered, out)
torch.cuda.synchronize(); t_all = time.perf_counter() - t0
t = torch.tensor([t_trace, t_all], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
hits = int((out.view(-1, 4)[:, 0] != 2139095040.0).sum().item())
if rank == 0:
    print("config5: %d tris, build %.2f ms (%s), corrupted=%s | %d rays over %d GPU(s): trace %.1f ms, +gather %.1f ms -> %.1f Mrays/s aggregate (mode %d), hit fraction rank0 %.3f"
          % (n, st["total"], ", ".join("%s %.2f" % (k, v) for k, v in st.items() if k != "total"), (leaf_bad, int_bad), per * world, world,
             float(t[0]) * 1e3, float(t[1]) * 1e3, per * world / float(t[1]) / 1e6, a.mode, hits / per), flush=True)
ctx.close()
if world > 1: dist.destroy_process_group()
