"""SURVEY 8(f)-2 mesh ingest (OBJ -> Triangle[128 B]) and 8(f)-3 file format, host-side parts (CPU)."""
import os

import numpy as np
import pytest

from unitysimpleraytracing_b200 import obj_ingest, scene_types as T

OBJ = """# unit quad + a triangle, with uvs and normals, mixed index styles
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0 0 1
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 1
vn 0 1 0
f 1/1/1 2/2/1 3/3/1 4/4/1
f 1//2 2//2 5//2
f -5 -4 -1
""".splitlines()


def test_obj_fan_triangulation_and_packing():
    t = obj_ingest.load_obj(OBJ)
    assert t.dtype == T.Triangle and len(t) == 4                     # quad -> 2, plus 2 triangles
    assert np.array_equal(t["a"][0], [0, 0, 0]) and np.array_equal(t["b"][0], [1, 0, 0]) and np.array_equal(t["c"][0], [1, 1, 0])
    assert np.array_equal(t["a"][1], [0, 0, 0]) and np.array_equal(t["b"][1], [1, 1, 0]) and np.array_equal(t["c"][1], [0, 1, 0])
    assert np.array_equal(t["c_uv"][0], [1, 1]) and np.array_equal(t["c_uv"][1], [0, 1])
    assert np.array_equal(t["a_normal"][0], [0, 0, 1]) and np.array_equal(t["b_normal"][2], [0, 1, 0])
    assert np.array_equal(t["a_uv"][2], [0, 0])                      # v//vn: no uv -> 0
    assert np.array_equal(t["a"][3], [0, 0, 0]) and np.array_equal(t["c"][3], [0, 0, 1])   # negative indices
    assert t.tobytes()[12:16] == bytes(4)                        # pads stay zero like C# default(struct)


def test_obj_flip_x_keeps_orientation():
    a = obj_ingest.load_obj(OBJ); b = obj_ingest.load_obj(OBJ, flip_x=True)
    na = np.cross(a["b"] - a["a"], a["c"] - a["a"]); nb = np.cross(b["b"] - b["a"], b["c"] - b["a"])
    assert np.allclose(nb, na * [-1, 1, 1])                          # mirrored geometry, same facing
    assert np.array_equal(b["a_normal"][0], [0, 0, 1])


def test_reference_scene_mesh_if_present():
    p = "/root/reference/Assets/_Assets/ExampleObject3.obj"
    if not os.path.exists(p):
        pytest.skip("reference checkout not present (GPU box)")
    t = obj_ingest.load_obj(p, flip_x=True)
    assert len(t) == 12800                                           # 80 x 80 quads
    assert np.abs(t["a"][:, 2]).max() < 1e-6 and np.abs(t["a"][:, :2]).max() <= 4.0 + 1e-6


def test_bvh_file_round_trip_on_host(tmp_path, oracle):
    from unitysimpleraytracing_b200 import bvh_io, meshes
    import struct
    tris = meshes.uniform_soup(300, seed=9)
    s = oracle.Scene(tris); n = s.n
    bufs = dict(keys=s.sortedMortonCodes, triangleIndex=s.sortedTriangleIndices, triangleData=tris, triangleAABB=s.triangleAABB,
                bvhData=s.bvhData[:n - 1], leafNodes=s.leafNodes, internalNodes=s.internalNodes[:n - 1])
    path = tmp_path / "x.usrtbvh"
    with open(path, "wb") as f:
        sizes = [np.ascontiguousarray(bufs[k]).nbytes for k, _, _ in bvh_io._SECTIONS] + [0]
        f.write(bvh_io.MAGIC + struct.pack("<II8I", n, 0, *sizes))
        for k, _, _ in bvh_io._SECTIONS:
            f.write(np.ascontiguousarray(bufs[k]).tobytes())
    m, back = bvh_io.read_bvh(str(path))
    assert m == n
    for k in bufs:
        assert back[k].tobytes() == np.ascontiguousarray(bufs[k]).tobytes()
    assert os.path.getsize(path) == 48 + n * (4 + 4 + 128 + 32 + 8) + (n - 1) * (32 + 24)


def test_multiple_meshes_merge_and_map_back():
    """MeshBufferContainer.cs:96 "TODO multiple meshes": merged scene + the map from a hit's triangleIndex back to
    (mesh, local triangle). The container's constructor is exercised with a stand-in context (no GPU here)."""
    from unitysimpleraytracing_b200 import host, meshes
    from unitysimpleraytracing_b200.scene_types import Triangle
    a, b, c = meshes.uniform_soup(5, seed=1), meshes.sphere(4, 8), meshes.uniform_soup(1, seed=2)
    merged, off = host.merge_meshes([a, b, c])
    assert merged.dtype == Triangle and len(merged) == len(a) + len(b) + len(c)
    assert off.tolist() == [0, 5, 5 + len(b), 6 + len(b)]
    assert merged[:5].tobytes() == a.tobytes() and merged[5:5 + len(b)].tobytes() == b.tobytes() and merged[-1:].tobytes() == c.tobytes()
    mesh, local = host.mesh_of_triangle(off, [0, 4, 5, 5 + len(b) - 1, 5 + len(b)])
    assert mesh.tolist() == [0, 0, 1, 1, 2] and local.tolist() == [0, 4, 0, len(b) - 1, 0]

    class FakeCtx:
        def __init__(self): self.calls = []
        def upload_triangles(self, t): self.calls.append(("upload", len(t), t.tobytes()))
        def morton(self): self.calls.append(("morton",))
        def fit_world_box(self): self.calls.append(("fit",)); return (np.zeros(3, np.float32), np.ones(3, np.float32))

    ctx = FakeCtx()
    cont = host.MeshBufferContainer([a, b, c], ctx=ctx, fitWorldBox=True)
    assert [x[0] for x in ctx.calls] == ["upload", "fit", "morton"] and ctx.calls[0][2] == merged.tobytes()
    assert cont.MeshOffsets.tolist() == off.tolist()
    assert [x.tolist() for x in cont.MeshOfTriangle([6])] == [[1], [1]]
    single = host.MeshBufferContainer(a, ctx=FakeCtx())
    assert single.MeshOffsets.tolist() == [0, 5] and [x.tolist() for x in single.MeshOfTriangle([3])] == [[0], [3]]
