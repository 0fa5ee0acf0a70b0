"""Developer tool (GPU): trace throughput on the C2 scene (and C1 soup), strict and culled, with parity check."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unitysimpleraytracing_b200 import host, meshes
from oracle import usrt_oracle as O
def run(name, tris, cam, check_rows=24):
    ctx = host.Context(len(tris)); ctx.upload_triangles(tris); ctx.rebuild(); ctx.sync()
    s = torch.cuda.Stream(); ctx.set_stream(s.cuda_stream)
    W, H = 1920, 1080
    for mode in (0, 1):
        ctx.set_trace_mode(mode)
        with torch.cuda.stream(s):
            for _ in range(3): ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=False)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
            for a, b in ev:
                a.record(s); ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=False); b.record(s)
        torch.cuda.synchronize()
        ms = np.median([a.elapsed_time(b) for a, b in ev])
        print("%s mode %d: %.3f ms  %.1f Mrays/s" % (name, mode, ms, W * H / ms / 1e3), flush=True)
    ctx.set_trace_mode(0)
    if check_rows:
        ref = O.Scene(tris)
        y0 = H // 2 - check_rows // 2
        want = ref.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], y0=y0, y1=y0 + check_rows, threads=os.cpu_count())
        got = ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
        a = got[y0 * W:(y0 + check_rows) * W]; b = want[y0 * W:(y0 + check_rows) * W]
        print("   parity rows: %s" % ("bit-exact" if a.tobytes() == b.tobytes() else "MISMATCH %d" % int((a.view(np.uint32).reshape(-1,4) != b.view(np.uint32).reshape(-1,4)).any(1).sum())))
    ctx.close()
run("c2", meshes.scene_c2(), meshes.SCENE_C2_CAMERA)
run("c1", meshes.scene_c1(), meshes.SCENE_SOUP_CAMERA, check_rows=8)
