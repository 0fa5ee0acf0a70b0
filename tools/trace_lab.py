"""Developer tool (GPU): time variant builds of the library (tools/micro/build_trace_lab.sh) on the C2 / C1 scenes.
    python tools/trace_lab.py tools/micro/libusrt_s24.so [...]      (the default build is always timed first)
Each variant's strict 1080p frame of C2 is bit-compared with the default build's."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np, torch, hashlib
    from unitysimpleraytracing_b200 import _lib
    if sys.argv[2] != "default":
        _lib.LIB_PATH = os.path.abspath(sys.argv[2])
    from unitysimpleraytracing_b200 import host, meshes
    out = {}
    for name, tris, cam, (W, H) in (("c2", meshes.scene_c2(), meshes.SCENE_C2_CAMERA, (1920, 1080)),
                                   ("c1", meshes.scene_c1(), meshes.SCENE_SOUP_CAMERA, (512, 512))):
        ctx = host.Context(len(tris)); ctx.upload_triangles(tris); ctx.rebuild(); ctx.sync()
        s = torch.cuda.Stream(); ctx.set_stream(s.cuda_stream)
        for mode in (0, 1):
            ctx.set_trace_mode(mode)
            with torch.cuda.stream(s):
                for _ in range(5): ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=False)
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
                for a, b in ev:
                    a.record(s); ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=False); b.record(s)
            torch.cuda.synchronize()
            out["%s_mode%d_ms" % (name, mode)] = float(np.median([a.elapsed_time(b) for a, b in ev]))
        ctx.set_trace_mode(0)
        out[name + "_sha"] = hashlib.sha256(ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"]).tobytes()).hexdigest()[:16]
        ctx.close()
    # incoherent rays over a 2^22-triangle soup (268 MB of packed nodes: beyond L2), strict and culled
    tris = meshes.uniform_soup(1 << 22, seed=0x5EED0004)
    rays = torch.from_numpy(meshes.incoherent_rays(1 << 20, seed=9)).cuda()
    outb = torch.empty((1 << 20) * 4, dtype=torch.float32, device="cuda")
    ctx = host.Context(len(tris)); ctx.upload_triangles(tris); ctx.rebuild(); ctx.sync()
    s = torch.cuda.Stream(); ctx.set_stream(s.cuda_stream)
    for mode in (0, 1):
        ctx.set_trace_mode(mode)
        with torch.cuda.stream(s):
            ctx.trace_rays_device(rays.data_ptr(), 1 << 20, outb.data_ptr())
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
            for a, b in ev:
                a.record(s); ctx.trace_rays_device(rays.data_ptr(), 1 << 20, outb.data_ptr()); b.record(s)
        torch.cuda.synchronize()
        out["inc_mode%d_ms" % mode] = float(np.median([a.elapsed_time(b) for a, b in ev]))
        if mode == 0: out["inc_sha"] = hashlib.sha256(outb.cpu().numpy().tobytes()).hexdigest()[:16]
    ctx.close()
    print(json.dumps(out))
    sys.exit(0)
base = None
for lib in ["default"] + sys.argv[1:]:
    r = subprocess.run([sys.executable, __file__, "--one", lib], capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(lib, "FAILED", r.stderr[-500:]); continue
    d = json.loads(line[0])
    if base is None: base = d
    same = d["c2_sha"] == base["c2_sha"] and d["c1_sha"] == base["c1_sha"] and d["inc_sha"] == base["inc_sha"]
    print("%-34s c2 strict %.4f ms (%.0f Mrays/s) culled %.4f | c1 strict %.4f culled %.4f | 2^22 soup, 2^20 incoherent rays strict %.2f culled %.2f | results %s" % (
        os.path.basename(lib), d["c2_mode0_ms"], 1920 * 1080 / d["c2_mode0_ms"] / 1e3, d["c2_mode1_ms"], d["c1_mode0_ms"], d["c1_mode1_ms"],
        d["inc_mode0_ms"], d["inc_mode1_ms"], "identical" if same else "DIFFER"), flush=True)
