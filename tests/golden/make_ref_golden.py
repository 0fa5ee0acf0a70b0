"""Generates tests/golden/ref_digests.json FROM THE REFERENCE'S OWN CODE (oracle/_ref/libusrt_ref.so, i.e.
BVH.compute, Raytracing.compute and MeshBufferContainer.cs compiled by oracle/build_ref.sh):

    python tests/golden/make_ref_golden.py          # needs /root/reference (build container only)

sha256 of every buffer the build produces and of full primary frames, for the BASELINE configurations the oracle can
be checked on in seconds: the reference's own scene mesh (12,800 triangles), configs[0] (65,536-triangle soup, 512x512
rays) and configs[1] at FULL size (1,048,576 triangles, 1920x1080 rays). The GPU parity tests and bench.py compare
the CUDA path with these digests on the GPU box, where /root/reference does not exist. The first two cases are sorted by the reference's own
Sorting/*.compute kernels (under the wave emulator); configs[1] exceeds their hard-wired 524,288-element capacity
(Scan.compute:50,64), so its sorted arrays are the stable sort of the reference-computed keys -- the contract
ComputeBufferSorter.ValidateSortedData checks (ComputeBufferSorter.cs:150-177). `sorted_by` records which.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import usrt_ref as R                         # noqa: E402
from unitysimpleraytracing_b200 import meshes            # noqa: E402

CASES = (("refgrid_12800", meshes.reference_scene_grid, "REFERENCE_CAMERA", (480, 270)),
         ("config0_soup_65536", meshes.scene_c1, "SCENE_SOUP_CAMERA", (512, 512)),
         ("config1_scene_1048576", meshes.scene_c2, "SCENE_C2_CAMERA", (1920, 1080)))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def scene_digests(s):
    n = s.n
    return dict(mortonCodes=sha(s.mortonCodes), triangleAABB=sha(s.triangleAABB), sortedMortonRaw=sha(s.sortedMortonRaw),
                sortedTriangleIndices=sha(s.sortedTriangleIndices), sortedMortonCodes=sha(s.sortedMortonCodes),
                internalNodes=sha(s.internalNodes[:n - 1]), leafNodes=sha(s.leafNodes), bvhData=sha(s.bvhData[:n - 1]))


def main():
    if not R.reference_present():
        raise SystemExit("needs the reference checkout (REF=/root/reference)")
    R.build(force=True)
    out = {"_generator": "tests/golden/make_ref_golden.py over oracle/_ref/libusrt_ref.so (reference text compiled by oracle/build_ref.sh)"}
    for name, make, cam_name, (w, h) in CASES:
        tris = make()
        cam = getattr(meshes, cam_name)
        t0 = time.time()
        s = R.Scene(tris)
        d = scene_digests(s)
        d["triangles"] = sha(tris)
        hits, _ = s.trace_primary(w, h, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
        d["primary_%dx%d" % (w, h)] = sha(hits)
        d["primary_hit_count"] = int((hits["distance"] != np.float32(2139095040.0)).sum())
        d["n"] = int(s.n)
        d["sorted_by"] = s.sorted_by
        out[name] = d
        print("%s: n=%d frame %dx%d hits=%d (%.1f s)" % (name, s.n, w, h, d["primary_hit_count"], time.time() - t0))
    with open(os.path.join(HERE, "ref_digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
