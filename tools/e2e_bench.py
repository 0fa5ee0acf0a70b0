"""Developer tool (GPU): host-buffer (e2e) step components with pinned memory, for several copy-band counts."""
import os, sys, time, statistics, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for b in ("1", "2", "4", "8"):
        subprocess.run([sys.executable, __file__, b], env=dict(os.environ, USRT_COPY_BANDS=b))
    sys.exit(0)
import numpy as np, torch
from unitysimpleraytracing_b200 import host, meshes
tris = meshes.scene_c2(); cam = meshes.SCENE_C2_CAMERA; W, H = 1920, 1080
pt = torch.from_numpy(tris.view(np.uint8).reshape(-1)).pin_memory(); ph = torch.empty(W * H * 16, dtype=torch.uint8).pin_memory()
th = pt.numpy().view(tris.dtype); hh = ph.numpy().view(np.dtype([("distance", "<f4"), ("triangleIndex", "<u4"), ("uv", "<f4", 2)]))
ctx = host.Context(len(tris))
def t(fn, n=8):
    fn(); fn()
    xs = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ctx.sync(); xs.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(xs)
up = t(lambda: ctx.upload_triangles(th))
rb = t(lambda: ctx.rebuild())
tr = t(lambda: ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=False))
td = t(lambda: ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=True, out=hh))
def step():
    ctx.upload_triangles(th); ctx.rebuild(); ctx.trace_primary(W, H, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=True, out=hh)
st = t(step)
print("bands=%s upload %.3f ms (%.1f GB/s)  rebuild %.3f  trace %.3f  trace+download %.3f  step %.3f ms -> %.0f Mrays/s" % (
    os.environ.get("USRT_COPY_BANDS"), up, len(tris) * 128 / up / 1e6, rb, tr, td, st, W * H / st / 1e3), flush=True)
ctx.close()
