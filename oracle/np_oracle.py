"""Independent numpy / pure-Python restatement of the reference path -- TEST INFRASTRUCTURE ONLY.

A second opinion on oracle/usrt_oracle.cpp, written a different way on purpose (bit interleave by
loops instead of magic multiplies, argsort instead of the block radix structure, a top-down
recursive radix tree instead of Karras's per-node searches, recursive box unions instead of the
atomic climb, vectorised brute force instead of the stack walk). Agreement between the two is the
pin this path has, because the reference ships no golden vectors (PARITY UNPINNED).
Pure-Python parts are for small cases only.

Reference: Assets/_Scripts/MeshBufferContainer.cs:32-83,123-169; Assets/_Shaders/BVH/BVH.compute;
Assets/_Shaders/Raytracing/Raytracing.compute:37-103.
"""
import numpy as np

F = np.float32


def morton_and_aabb(a, b, c, whole_min=-125.0, whole_max=125.0):
    """MeshBufferContainer.cs:52-83,41-50 with numpy fp32 ops (IEEE per-op)."""
    a, b, c = (np.asarray(x, F) for x in (a, b, c))
    mn = (np.minimum(np.minimum(a, b), c) - F(0.001)).astype(F)
    mx = (np.maximum(np.maximum(a, b), c) + F(0.001)).astype(F)
    cen = ((mn + mx).astype(F) * F(0.5)).astype(F)
    lo, hi = np.asarray(whole_min, F), np.asarray(whole_max, F)          # scalars (the cube) or per-axis 3-vectors
    cen = ((cen - lo).astype(F) / (hi - lo).astype(F)).astype(F)
    q = np.minimum(np.maximum((cen * F(1024.0)).astype(F), F(0.0)), F(1023.0)).astype(np.uint32)  # truncation
    key = np.zeros(len(a), np.uint32)
    for bit in range(10):                      # x is the most significant of each triple
        for axis, sh in ((0, 2), (1, 1), (2, 0)):
            key |= ((q[:, axis] >> np.uint32(bit)) & np.uint32(1)) << np.uint32(3 * bit + sh)
    return key, mn, mx


def stable_sort(keys, values):
    order = np.argsort(np.asarray(keys, np.uint32), kind="stable")
    return np.asarray(keys)[order], np.asarray(values)[order]


def distribute_keys(keys):
    """MeshBufferContainer.cs:154-169 as a wrapped cumulative sum."""
    k = np.asarray(keys, np.uint32)
    if len(k) == 0:
        return k.copy()
    d = np.maximum((k[1:] - k[:-1]).astype(np.uint32), np.uint32(1)).astype(np.uint64)
    out = np.zeros(len(k), np.uint64)
    out[1:] = np.cumsum(d)
    return (out & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def radix_tree_topdown(keys):
    """Binary radix tree over strictly increasing uint32 keys, built top-down. With Karras's numbering
    the internal node of a range [first,last] split at s has children `s` and `s+1` (leaf iff the
    child range has one key) and the root is node 0 -- BVH.compute:94-149."""
    k = [int(x) for x in keys]
    n = len(k)
    NULL = 0xFFFFFFFF
    internal = np.full((max(n - 1, 1), 6), NULL, np.uint32)   # left, ltype, right, rtype, parent, index
    leaf = np.full((n, 2), NULL, np.uint32)                   # parent, index
    if n < 2:
        return internal, leaf
    work = [(0, n - 1, 0, NULL)]                              # first, last, node id, parent
    while work:
        first, last, node, parent = work.pop()
        diff = k[first] ^ k[last]
        top = diff.bit_length() - 1                           # highest differing bit
        # split = last index whose bit `top` equals that of k[first]
        lo, hi = first, last
        while lo < hi:
            mid = (lo + hi + 1) // 2
            if (k[mid] >> top) & 1 == (k[first] >> top) & 1:
                lo = mid
            else:
                hi = mid - 1
        s = lo
        internal[node, 5] = node
        if parent != NULL:
            internal[node, 4] = parent
        for slot, (cf, cl, cid) in enumerate(((first, s, s), (s + 1, last, s + 1))):
            if cf == cl:
                internal[node, 2 * slot] = cid; internal[node, 2 * slot + 1] = 1
                leaf[cid] = (node, cid)
            else:
                internal[node, 2 * slot] = cid; internal[node, 2 * slot + 1] = 0
                work.append((cf, cl, cid, node))
    return internal, leaf


def refit_recursive(internal, sorted_indices, tri_min, tri_max):
    """Node boxes as unions over the subtree (BVH.compute:152-220), iterative post-order."""
    m = len(internal)
    bmin = np.zeros((m, 3), F); bmax = np.zeros((m, 3), F)
    done = np.zeros(m, bool)
    stack = [0]
    while stack:
        node = stack[-1]
        l, lt, r, rt = (int(x) for x in internal[node, :4])
        pending = [c for c, t in ((l, lt), (r, rt)) if t == 0 and not done[c]]
        if pending:
            stack.extend(pending)
            continue
        boxes = []
        for c, t in ((l, lt), (r, rt)):
            if t == 0:
                boxes.append((bmin[c], bmax[c]))
            else:
                tri = int(sorted_indices[c])
                boxes.append((tri_min[tri], tri_max[tri]))
        bmin[node] = np.minimum(boxes[0][0], boxes[1][0]); bmax[node] = np.maximum(boxes[0][1], boxes[1][1])
        done[node] = True
        stack.pop()
    return bmin, bmax


def _dot(a, b):
    return ((a[..., 0] * b[..., 0]).astype(F) + (a[..., 1] * b[..., 1]).astype(F)).astype(F) + (a[..., 2] * b[..., 2]).astype(F)


def _cross(a, b):
    return np.stack([((a[..., 1] * b[..., 2]).astype(F) - (a[..., 2] * b[..., 1]).astype(F)).astype(F),
                     ((a[..., 2] * b[..., 0]).astype(F) - (a[..., 0] * b[..., 2]).astype(F)).astype(F),
                     ((a[..., 0] * b[..., 1]).astype(F) - (a[..., 1] * b[..., 0]).astype(F)).astype(F)], -1)


def brute_force_hits(origin, direction, v0, v1, v2, tri_min, tri_max, order, max_float):
    """For each ray: closest candidate over all triangles visited in `order`, a candidate being a
    triangle whose padded AABB passes the slab test (Raytracing.compute:75-87,91) and that passes
    Moller-Trumbore (:37-73); strict '<' => the earliest visited wins ties (:95). Vectorised over
    triangles, looped over rays (small cases only). Returns (distance, triangleIndex, u, v)."""
    v0, v1, v2 = (np.asarray(x, F)[order] for x in (v0, v1, v2))
    bmn, bmx = np.asarray(tri_min, F)[order], np.asarray(tri_max, F)[order]
    out = []
    with np.errstate(all="ignore"):
        for o, d in zip(np.asarray(origin, F), np.asarray(direction, F)):
            inv = (F(1.0) / d).astype(F)
            t1 = ((bmn - o).astype(F) * inv).astype(F); t2 = ((bmx - o).astype(F) * inv).astype(F)
            lo = np.fmin(t1, t2); hi = np.fmax(t1, t2)
            tmin = np.fmax(lo[:, 0], np.fmax(lo[:, 1], lo[:, 2])); tmax = np.fmin(hi[:, 0], np.fmin(hi[:, 1], hi[:, 2]))
            box_ok = (tmax > tmin) & (tmax > 0)
            e1 = (v1 - v0).astype(F); e2 = (v2 - v0).astype(F)
            dd = np.broadcast_to(d, e2.shape)
            pvec = _cross(dd, e2)
            det = _dot(e1, pvec)
            ok = box_ok & ~((det < F(1e-8)) & (det > F(-1e-8)))
            inv_det = (F(1.0) / det).astype(F)
            tvec = (o - v0).astype(F)
            u = (_dot(tvec, pvec) * inv_det).astype(F)
            ok &= ~((u < 0) | (u > 1))
            qvec = _cross(tvec, e1)
            v = (_dot(dd, qvec) * inv_det).astype(F)
            ok &= ~((v < 0) | ((u + v).astype(F) > 1))
            t = (_dot(e2, qvec) * inv_det).astype(F)
            t = np.where(ok, t, F(np.inf))
            if not ok.any() or not (t.min() < max_float):
                out.append((max_float, 0, F(0), F(0)))
                continue
            j = int(np.argmin(t))             # first minimum in visiting order
            out.append((t[j], int(order[j]), u[j], v[j]))
    return out


# ---- diffuse bounce rays (BASELINE configs[4]; defined by oracle/usrt_oracle.cpp DiffuseRay) --------------------
def primary_rays(width, height, near, tan_half_fov, cam_to_world):
    """Raytracing.compute:108-126 + RaytracingMeshDrawer.cs:78-81, vectorised over the frame: (H*W, 3) origins
    and unit directions, pixel index y*W + x."""
    m = np.asarray(cam_to_world, F).reshape(4, 4)
    near, fov = F(near), F(tan_half_fov)
    h = F(F(F(2) * near) * fov)
    w = F(F(F(width) * h) / F(height))
    xs = (np.arange(width, dtype=np.uint32).astype(F) + F(0.5)).astype(F)
    ys = (np.arange(height, dtype=np.uint32).astype(F) + F(0.5)).astype(F)
    dx = (F(F(-w) / F(2)) + (F(w / F(width)) * xs).astype(F)).astype(F)
    dy = (F(F(-h) / F(2)) + (F(h / F(height)) * ys).astype(F)).astype(F)
    DX, DY = np.meshgrid(dx, dy)                        # [y][x]
    DZ = np.full_like(DX, -near)
    o = np.empty(3, F); d = np.empty((height, width, 3), F)
    for r in range(3):
        o[r] = F(F(F(m[r, 0] * F(0)) + F(m[r, 1] * F(0))) + F(m[r, 2] * F(0))) + F(m[r, 3] * F(1))
        d[..., r] = ((((m[r, 0] * DX).astype(F) + (m[r, 1] * DY).astype(F)).astype(F) + (m[r, 2] * DZ).astype(F)).astype(F)
                     + F(m[r, 3] * F(0))).astype(F)
    ln = np.sqrt(_dot(d, d)).astype(F)
    d = (d / ln[..., None]).astype(F)
    return np.broadcast_to(o, (height * width, 3)).copy(), d.reshape(-1, 3)


def _hash64(x):
    x = np.asarray(x, np.uint64).copy()
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def diffuse_rays(distance, triangle_index, va, vb, vc, width, height, near, tan_half_fov, cam_to_world, seed,
                 first_sample, num_samples, max_float):
    """Bounce rays of samples [first_sample, +num_samples): (num_samples * H * W, 8) float32, sample-major.
    distance / triangle_index: the primary hit records; va, vb, vc: (n, 3) triangle vertices."""
    o, d = primary_rays(width, height, near, tan_half_fov, cam_to_world)
    frame = width * height
    t = np.asarray(distance, F)
    hit = t != F(max_float)
    P = (o + (d * t[:, None]).astype(F)).astype(F)
    ti = np.asarray(triangle_index, np.int64)
    a, b, c = (np.asarray(v, F)[ti] for v in (va, vb, vc))
    n = _cross((b - a).astype(F), (c - a).astype(F))
    n2 = _dot(n, n)
    ok = hit & (n2 > 0)
    with np.errstate(all="ignore"):
        n = (n / np.sqrt(n2).astype(F)[:, None]).astype(F)
    n = np.where((_dot(n, d) > 0)[:, None], -n, n)
    out = np.zeros((num_samples, frame, 8), F)
    pix = np.arange(frame, dtype=np.uint64)
    for s in range(num_samples):
        smp = np.uint64((first_sample + s) & 0xFFFF)
        with np.errstate(over="ignore"):
            stream = _hash64(np.uint64(seed) ^ _hash64((pix << np.uint64(16)) | smp))
        u = n.copy()
        todo = np.ones(frame, bool)
        for k in range(16):
            with np.errstate(over="ignore"):
                bits = _hash64(stream + np.uint64(k))
            v = np.stack([(((bits >> np.uint64(sh)) & np.uint64(0x1FFFFF)).astype(np.uint32).astype(F) * F(2.0 ** -20)).astype(F) - F(1)
                          for sh in (0, 21, 42)], -1).astype(F)
            l2 = _dot(v, v)
            acc = todo & (l2 <= F(1)) & (l2 > F(1e-4))
            with np.errstate(all="ignore"):
                u[acc] = (v[acc] / np.sqrt(l2[acc]).astype(F)[:, None]).astype(F)
            todo &= ~acc
        dd = (n + u).astype(F)
        d2 = _dot(dd, dd)
        with np.errstate(all="ignore"):
            dn = (dd / np.sqrt(d2).astype(F)[:, None]).astype(F)
        dd = np.where((d2 < F(1e-8))[:, None], n, dn)
        out[s, :, 0:3] = np.where(ok[:, None], (P + (n * F(0.001)).astype(F)).astype(F), F(0))
        out[s, :, 4:7] = np.where(ok[:, None], dd, F(0))
    return out.reshape(-1, 8)
