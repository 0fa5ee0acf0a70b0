"""Developer tool (GPU): rebuild stage times on the C2 scene (or a soup of 2^k triangles)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from unitysimpleraytracing_b200 import host, meshes
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
tris = meshes.scene_c2() if which == "c2" else meshes.uniform_soup(1 << int(which), seed=5)
ctx = host.Context(len(tris)); ctx.upload_triangles(tris); ctx.enable_stage_timing(True)
rec = []
for i in range(iters + 2):
    ctx.rebuild(); t = ctx.last_rebuild_ms(); s = ctx.last_sort_ms()
    if i >= 2: rec.append({**t, **{"sort_" + k: v for k, v in s.items()}})
print(which, len(tris), {k: round(statistics.median(r[k] for r in rec), 4) for k in rec[0]})
ctx.close()
