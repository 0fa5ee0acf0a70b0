"""Generates the committed golden fixtures from the oracle (oracle/usrt_oracle.cpp):

    python tests/golden/make_golden.py

The reference has no golden vectors of its own (PARITY UNPINNED); these are regression
pins of the oracle -- cross-checked against the independent numpy restatement in tests/test_oracle.py --
so that neither the oracle nor the CUDA path can drift silently. Inputs are stored with the outputs,
so the fixtures do not depend on the mesh generators staying bit-stable.
  small_soup.npz : 96-triangle soup, every buffer of the build + a 16x12 primary frame + 64 random rays
                   + 2 samples of diffuse bounce rays off that frame and their hit records
  digests.json   : sha256 of every buffer for larger seeded scenes (grid 12,800; soup 65,536)
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import usrt_oracle as O                      # noqa: E402
from unitysimpleraytracing_b200 import meshes            # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def scene_buffers(s):
    n = s.n
    return dict(mortonCodes=s.mortonCodes, triangleAABB=s.triangleAABB, sortedMortonRaw=s.sortedMortonRaw,
                sortedTriangleIndices=s.sortedTriangleIndices, sortedMortonCodes=s.sortedMortonCodes,
                internalNodes=s.internalNodes[:n - 1], leafNodes=s.leafNodes, bvhData=s.bvhData[:n - 1])


def main():
    tris = meshes.uniform_soup(96, seed=0x601D)
    s = O.Scene(tris)
    cam = meshes.SCENE_SOUP_CAMERA
    rays = meshes.incoherent_rays(64, seed=0x601E)
    out = {k: np.ascontiguousarray(v).view(np.uint8) for k, v in scene_buffers(s).items()}
    out["triangles"] = tris.view(np.uint8)
    out["rays"] = rays
    out["near"] = np.float32(cam["near"]); out["tan_half_fov"] = np.float32(cam["tan_half_fov"])
    out["cam_to_world"] = np.asarray(cam["cam_to_world"], np.float32)
    out["primary_16x12"] = s.trace_primary(16, 12, cam["near"], cam["tan_half_fov"], cam["cam_to_world"]).view(np.uint8)
    out["ray_hits"] = s.trace_rays(rays).view(np.uint8)
    prim = s.trace_primary(16, 12, cam["near"], cam["tan_half_fov"], cam["cam_to_world"])
    bounce = O.diffuse_rays(prim, tris, 16, 12, cam["near"], cam["tan_half_fov"], cam["cam_to_world"], 0x601F, 0, 2)
    out["diffuse_rays_16x12x2"] = bounce                       # BASELINE configs[4] bounce rays, seed 0x601F, samples 0-1
    out["diffuse_hits"] = s.trace_rays(bounce).view(np.uint8)
    np.savez_compressed(os.path.join(HERE, "small_soup.npz"), **out)

    digests = {}
    for name, t, c, (w, h) in (("refgrid_12800", meshes.reference_scene_grid(), meshes.REFERENCE_CAMERA, (96, 54)),
                               ("soup_65536", meshes.scene_c1(), meshes.SCENE_SOUP_CAMERA, (64, 64))):
        sc = O.Scene(t)
        d = {k: sha(v) for k, v in scene_buffers(sc).items()}
        d["triangles"] = sha(t)
        d["primary_%dx%d" % (w, h)] = sha(sc.trace_primary(w, h, c["near"], c["tan_half_fov"], c["cam_to_world"], threads=8))
        digests[name] = d
    json.dump(digests, open(os.path.join(HERE, "digests.json"), "w"), indent=1, sort_keys=True)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
