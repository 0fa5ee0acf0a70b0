"""Workload for the `ncu --set full` capture of the bench step's kernels (1,048,576 triangles, 1920x1080):
    USRT_NO_GRAPH=1 ncu --set full --clock-control none --import-source on \\
        -k regex:"k_trace_primary|k_construct_bvh|k_morton|k_histogram|k_distribute_keys|k_construct_tree" \\
        --launch-skip 12 --launch-count 6 -o gpurun_out/r01_prof_build_trace python tools/ncu_step.py
(plain launches instead of the CUDA-graph replay so that every kernel is a separate launch for the profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unitysimpleraytracing_b200 import host, meshes
tris = meshes.scene_c2(); cam = meshes.SCENE_C2_CAMERA
ctx = host.Context(len(tris)); ctx.upload_triangles(tris)
for _ in range(4):
    ctx.rebuild()
    ctx.trace_primary(1920, 1080, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], download=False)
ctx.sync(); ctx.close()
