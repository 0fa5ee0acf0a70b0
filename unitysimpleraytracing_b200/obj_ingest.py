"""mesh ingest, the step BEFORE the path -- Wavefront OBJ -> Triangle[128 B] packed as
Assets/_Scripts/MeshBufferContainer.cs:117-146 does from a UnityEngine.Mesh (positions, UVs and normals
by corner index; missing UVs/normals are zero like Unity's empty arrays would throw -- here they default).

Unity's own OBJ importer welds vertices, splits quads on a fixed diagonal and flips the X axis to
convert handedness (ExampleObject3.obj.meta); `flip_x=True` mirrors that flip (and reverses winding so
faces keep their orientation). Triangulation is a fan from the first corner. Host-side numpy only.
"""
import numpy as np

from .scene_types import Triangle


def load_obj(path_or_lines, flip_x=False):
    lines = open(path_or_lines).read().splitlines() if isinstance(path_or_lines, str) else list(path_or_lines)
    v, vt, vn, corners = [], [], [], []

    def idx(tok, count):
        if tok == "":
            return -1
        i = int(tok)
        return i - 1 if i > 0 else count + i          # negative = relative to the end

    for line in lines:
        p = line.split()
        if not p or p[0].startswith("#"):
            continue
        if p[0] == "v":
            v.append([float(x) for x in p[1:4]])
        elif p[0] == "vt":
            vt.append([float(x) for x in p[1:3]])
        elif p[0] == "vn":
            vn.append([float(x) for x in p[1:4]])
        elif p[0] == "f":
            face = []
            for tok in p[1:]:
                parts = (tok.split("/") + ["", ""])[:3]
                face.append((idx(parts[0], len(v)), idx(parts[1], len(vt)), idx(parts[2], len(vn))))
            for k in range(1, len(face) - 1):          # fan triangulation
                tri = (face[0], face[k], face[k + 1])
                corners.append(tri[::-1] if flip_x else tri)
    V = np.asarray(v, np.float32).reshape(-1, 3)
    VT = np.asarray(vt, np.float32).reshape(-1, 2)
    VN = np.asarray(vn, np.float32).reshape(-1, 3)
    if flip_x:
        V = V * np.array([-1, 1, 1], np.float32)
        VN = VN * np.array([-1, 1, 1], np.float32)
    c = np.asarray(corners, np.int64).reshape(-1, 3, 3)  # (tri, corner, {v, vt, vn})
    t = np.zeros(len(c), Triangle)
    for k, name in enumerate("abc"):
        t[name] = V[c[:, k, 0]]
        if len(VT):
            t[name + "_uv"] = np.where((c[:, k, 1] >= 0)[:, None], VT[np.maximum(c[:, k, 1], 0)], 0)
        if len(VN):
            t[name + "_normal"] = np.where((c[:, k, 2] >= 0)[:, None], VN[np.maximum(c[:, k, 2], 0)], 0)
    if not len(VN):                                    # Unity would recalculate normals on import
        n = np.cross((t["b"] - t["a"]).astype(np.float64), (t["c"] - t["a"]).astype(np.float64))
        ln = np.linalg.norm(n, axis=1, keepdims=True)
        n = np.where(ln > 0, n / np.maximum(ln, 1e-300), 0).astype(np.float32)
        t["a_normal"] = t["b_normal"] = t["c_normal"] = n
    return t
