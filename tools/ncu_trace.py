"""Workload for the `ncu --set full` captures of the traversal kernels that round 1 never profiled:
  k_trace_primary on BASELINE configs[0] (65,536-triangle soup, 512x512 primary rays, strict)
  k_trace_rays    on BASELINE configs[3] (16,777,216-triangle soup, incoherent random-direction rays; a 2,073,600-ray
                  sample of the 3840x2160 frame keeps the ~40 profiler replays short)
    ncu --set full --clock-control none --import-source on -k regex:"k_trace_primary|k_trace_rays" \\
        --launch-skip 2 --launch-count 1 ... python tools/ncu_trace.py c1      (or: c4)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from unitysimpleraytracing_b200 import host, meshes
which = sys.argv[1] if len(sys.argv) > 1 else "c1"
if which == "c1":
    tris = meshes.scene_c1(); cam = meshes.SCENE_SOUP_CAMERA
    ctx = host.Context(len(tris)); ctx.upload_triangles(tris); ctx.rebuild()
    for _ in range(4):
        ctx.trace_primary(512, 512, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=False)
else:
    import torch
    tris = meshes.uniform_soup(1 << 24, seed=0x5EED0004)
    ctx = host.Context(len(tris)); ctx.upload_triangles(tris); ctx.rebuild()
    rays = torch.from_numpy(meshes.incoherent_rays(1920 * 1080, seed=0x5EED0005)).cuda()
    out = torch.empty(1920 * 1080 * 4, dtype=torch.float32, device="cuda")
    ctx.use_torch_stream()
    for _ in range(4):
        ctx.trace_rays_device(rays.data_ptr(), 1920 * 1080, out.data_ptr())
ctx.sync(); ctx.close()
