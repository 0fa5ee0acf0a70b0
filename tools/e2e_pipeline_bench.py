"""Tool: end-to-end frames (host triangles in, host hit records out) through the C ABI, sequential vs pipelined
over two contexts on one GPU (frame i+1's upload overlaps frame i's kernels), and the same for device-resident
steps (frame i+1's rebuild beside frame i's trace).   python tools/e2e_pipeline_bench.py [--steps 40]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unitysimpleraytracing_b200 import host, meshes
from unitysimpleraytracing_b200.scene_types import RaycastResult

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=40); ap.add_argument("--depth", type=int, default=2)
a = ap.parse_args()
W, H = 1920, 1080
tris = meshes.scene_c2(); cam = meshes.SCENE_C2_CAMERA; n = len(tris); rays = W * H
m = np.array(cam["cam_to_world"], np.float32)
pin_t = torch.from_numpy(tris.view(np.uint8).reshape(-1).copy()).pin_memory()
tris_h = pin_t.numpy().view(tris.dtype)
D = a.depth
pin_h = [torch.zeros(rays * 16, dtype=torch.uint8).pin_memory() for _ in range(D)]
hits_h = [p.numpy().view(RaycastResult) for p in pin_h]
ctxs = [host.Context(n) for _ in range(D)]

def seq(c, out):
    c.upload_triangles(tris_h); c.rebuild()
    c.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m, out=out)

for _ in range(3): seq(ctxs[0], hits_h[0])
ref = hits_h[0].copy()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(a.steps): seq(ctxs[0], hits_h[0])
torch.cuda.synchronize(); t_seq = (time.perf_counter() - t0) / a.steps * 1e3
print("sequential e2e       : %.3f ms/frame  %.0f Mrays/s" % (t_seq, rays / t_seq / 1e3))

def pipe(steps):
    for i in range(steps):
        c = ctxs[i % D]
        c.sync()                                   # frame i-D is complete: hits_h[i % D] has been read out
        c.upload_triangles_async(tris_h); c.rebuild()
        c.trace_primary_async(W, H, cam["near"], cam["tan_half_fov"], m, hits_h[i % D])
    for c in ctxs: c.sync()

for h in hits_h: h["distance"] = -1
pipe(2 * D)
assert all(h.tobytes() == ref.tobytes() for h in hits_h), "pipelined frame differs"
torch.cuda.synchronize(); t0 = time.perf_counter()
pipe(a.steps)
t_pipe = (time.perf_counter() - t0) / a.steps * 1e3
print("pipelined e2e (x%d)   : %.3f ms/frame  %.0f Mrays/s" % (D, t_pipe, rays / t_pipe / 1e3))

# device-resident steps
for c in ctxs: c.upload_triangles(tris_h)
def dev(steps, k):
    for i in range(steps):
        c = ctxs[i % k]
        c.rebuild(); c.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m, download=False)
    for c in ctxs: c.sync()
for k in range(1, D + 1):
    dev(6, k); t0 = time.perf_counter(); dev(a.steps * 4, k)
    t = (time.perf_counter() - t0) / (a.steps * 4) * 1e3
    print("device-resident, %d context(s): %.3f ms/step  %.0f Mrays/s" % (k, t, rays / t / 1e3))
