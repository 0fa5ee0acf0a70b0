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
    cen = ((cen - F(whole_min)).astype(F) / (F(whole_max) - F(whole_min))).astype(F)
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
