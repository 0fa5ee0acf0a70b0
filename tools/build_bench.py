"""Developer tool (GPU): rebuild stage times on the C2 scene (or a soup of 2^k triangles), L2 flushed between
rebuilds, plus the graph-replayed rebuild and the 1080p trace."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from unitysimpleraytracing_b200 import host, meshes
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
tris = meshes.scene_c2() if which == "c2" else meshes.uniform_soup(1 << int(which), seed=5)
cam = meshes.SCENE_C2_CAMERA if which == "c2" else meshes.SCENE_SOUP_CAMERA
dev = torch.device("cuda:0")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = host.Context(len(tris)); ctx.set_stream(st.cuda_stream); ctx.upload_triangles(tris)
if len(sys.argv) > 3 and sys.argv[3] == "pos":          # positions-only source (48 B/triangle)
    ctx.upload_positions(np.ascontiguousarray(tris.view(np.float32).reshape(len(tris), 32)[:, :12]))
    print("positions-only source")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ctx.enable_stage_timing(True)
rec = []
for i in range(iters + 3):
    flush.fill_(i & 255)
    ctx.rebuild(); t = ctx.last_rebuild_ms(); s = ctx.last_sort_ms()
    if i >= 3: rec.append({**t, **{"sort_" + k: v for k, v in s.items()}})
ctx.enable_stage_timing(False)
print(which, len(tris), {k: round(statistics.median(r[k] for r in rec), 4) for k in rec[0]})
def timed(fn, n=30):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i, (a, b) in enumerate(ev):
        flush.fill_(i & 255); a.record(st); fn(); b.record(st)
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
ctx.rebuild(); ctx.rebuild()
print("rebuild (graph replay) %.4f ms" % timed(ctx.rebuild))
print("trace 1080p %.4f ms" % timed(lambda: ctx.trace_primary(1920, 1080, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=False)))
ctx.close()
