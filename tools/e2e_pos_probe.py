"""Developer tool (GPU): where does the positions-only end-to-end frame time go? Variants of the pipelined loop."""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unitysimpleraytracing_b200 import host, meshes
tris = meshes.scene_c2(); cam = meshes.SCENE_C2_CAMERA; n = len(tris)
W, H = 1920, 1080
m = np.array(cam["cam_to_world"], np.float32)
pos = torch.from_numpy(np.ascontiguousarray(tris.view(np.float32).reshape(n, 32)[:, :12])).pin_memory(); pos_h = pos.numpy()
hit_dtype = np.dtype([("distance", "<f4"), ("triangleIndex", "<u4"), ("uv", "<f4", 2)])
TORCH_STREAMS = len(sys.argv) > 1 and sys.argv[1] == "torch"
def run(ncx, download, steps=40, upload=True):
    C = [host.Context(n) for _ in range(ncx)]
    if TORCH_STREAMS:
        S = [torch.cuda.Stream() for _ in C]
        for c, st in zip(C, S): c.set_stream(st.cuda_stream)
    hits = [torch.empty(W * H * 16, dtype=torch.uint8).pin_memory() for _ in C]
    hh = [h.numpy().view(hit_dtype) for h in hits]
    for c in C:
        c.upload_triangles(tris); c.upload_positions(pos_h, pinned=True); c.rebuild(); c.sync()
    def loop(k):
        for i in range(k):
            c = C[i % ncx]
            c.sync()
            if upload: c.upload_positions(pos_h, pinned=True)
            c.rebuild()
            if download: c.trace_primary_async(W, H, cam["near"], cam["tan_half_fov"], m, hh[i % ncx])
            else: c.trace_primary(W, H, cam["near"], cam["tan_half_fov"], m, download=False)
        for c in C: c.sync()
    loop(6); torch.cuda.synchronize()
    t0 = time.perf_counter(); loop(steps); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    for c in C: c.close()
    return ms
for ncx in (1, 2, 3, 4):
    print("contexts %d: upload+build+trace->host %.3f ms | no host frame %.3f | no upload %.3f | neither %.3f" % (
        ncx, run(ncx, True), run(ncx, False), run(ncx, True, upload=False), run(ncx, False, upload=False)), flush=True)
