"""Developer tool (GPU): BASELINE.json configs 1-4 on one B200 -> markdown rows (profiles/<round>_configs.md)."""
import os, sys, time, statistics, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unitysimpleraytracing_b200 import host, meshes

def timed(stream, fn, iters=5, warm=2):
    for _ in range(warm): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(stream); fn(); b.record(stream)
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)

def scene(name, tris, cam, frames, rays=None):
    n = len(tris)
    ctx = host.Context(n); s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
    ctx.upload_triangles(tris); ctx.enable_stage_timing(True)
    for _ in range(3): ctx.rebuild()
    st = ctx.last_rebuild_ms()
    out = ["| %s | %d | build %.3f ms (morton %.3f, sort %.3f, distribute %.3f, tree %.3f, refit %.3f) |" % (name, n, st["total"], st["morton"], st["sort"], st["distribute"], st["tree"], st["bvh"])]
    for (w, h) in frames:
        for mode in (0, 1):
            ctx.set_trace_mode(mode)
            ms = timed(s, lambda: ctx.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=False))
            out.append("| %s | %d | primary %dx%d %s: %.3f ms = %.0f Mrays/s |" % (name, n, w, h, "strict" if mode == 0 else "culled(non-parity)", ms, w * h / ms / 1e3))
    if rays is not None:
        tr = torch.from_numpy(rays).cuda()
        for mode in (0, 1):
            ctx.set_trace_mode(mode)
            ms = timed(s, lambda: ctx.trace_rays_device(tr.data_ptr(), len(rays)), iters=3, warm=1)
            out.append("| %s | %d | %d incoherent rays %s: %.3f ms = %.0f Mrays/s |" % (name, n, len(rays), "strict" if mode == 0 else "culled(non-parity)", ms, len(rays) / ms / 1e3))
    ctx.close()
    print("\n".join(out), flush=True)

which = sys.argv[1:] or ["c1", "c2", "c4"]
if "c1" in which:
    scene("C1 soup 65,536", meshes.scene_c1(), meshes.SCENE_SOUP_CAMERA, [(512, 512), (1920, 1080)])
if "c2" in which:
    scene("C2 sphere+heightfield", meshes.scene_c2(), meshes.SCENE_C2_CAMERA, [(1920, 1080), (3840, 2160)])
if "c4" in which:
    t0 = time.time(); tris = meshes.uniform_soup(1 << 24, seed=0x5EED0004); print("gen 16M %.1fs" % (time.time() - t0), flush=True)
    scene("C4 soup 16,777,216", tris, meshes.SCENE_SOUP_CAMERA, [], rays=meshes.incoherent_rays(3840 * 2160, seed=0x5EED0005))
