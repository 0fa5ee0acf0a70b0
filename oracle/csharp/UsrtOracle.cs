// UsrtOracle.cs -- scalar C# statement of the oracle (TEST INFRASTRUCTURE; UNVERIFIED HERE).
//
// BASELINE.json's north_star asks for the oracle / CPU baseline as a scalar C# transliteration of the
// reference's kernels. This image has no .NET / Mono toolchain (dotnet, mono, mcs, csc: not found), so
// this file can be neither compiled nor run here or on the GPU box. What runs in tests and benchmarks
// is its function-for-function C++ twin, oracle/usrt_oracle.cpp; keep the two in step. Differences a
// porter must keep in mind (SURVEY.md 8a): C# promotes uint*int to long (HLSL wraps in 32 bits) -- use
// unchecked((int)(...)) as below; Math.Min/Max(float) propagate NaN unlike HLSL min/max -- the slab
// test uses FMin/FMax; float arithmetic must stay fp32 per operation (no double intermediates).
//
// Reference files restated: Assets/_Scripts/MeshBufferContainer.cs:32-83,123-169,
// Assets/_Shaders/Sorting/*.compute, Assets/_Shaders/BVH/BVH.compute:18-220,
// Assets/_Shaders/Raytracing/Raytracing.compute:23-176, Assets/_Shaders/Constants.cginc:1-54.
using System;
using System.Runtime.InteropServices;

namespace Usrt.Oracle
{
    [StructLayout(LayoutKind.Sequential, Pack = 4)] public struct Float3 { public float x, y, z; }
    [StructLayout(LayoutKind.Sequential, Size = 32)] public struct Aabb { public Float3 min; public float _d0; public Float3 max; public float _d1; }
    [StructLayout(LayoutKind.Sequential, Size = 24)] public struct InternalNode { public uint leftNode, leftNodeType, rightNode, rightNodeType, parent, index; }
    [StructLayout(LayoutKind.Sequential, Size = 8)] public struct LeafNode { public uint parent, index; }
    [StructLayout(LayoutKind.Sequential, Size = 16)] public struct RaycastResult { public float distance; public uint triangleIndex; public float u, v; }
    public struct Ray { public Float3 origin, dir, invDir; }

    public static class Oracle
    {
        public const uint Null = 0xFFFFFFFFu;
        public const uint InternalNodeType = 0, LeafNodeType = 1;          // Constants.cginc:17-18
        public static readonly float MaxFloat = (float)0x7F7FFFFF;          // Constants.cginc:7 (integer literal!)

        static float Sel(bool c, float a, float b) => c ? a : b;
        static float MinSel(float a, float b) => a < b ? a : b;            // finite inputs only
        static float MaxSel(float a, float b) => a > b ? a : b;
        static float FMin(float a, float b) => float.IsNaN(a) ? b : (float.IsNaN(b) ? a : (a < b ? a : b));   // HLSL min
        static float FMax(float a, float b) => float.IsNaN(a) ? b : (float.IsNaN(b) ? a : (a > b ? a : b));   // HLSL max

        // ---- K1: MeshBufferContainer.cs:32-83 --------------------------------------------------
        public static uint ExpandBits(uint v)
        {
            unchecked
            {
                v = (v * 0x00010001u) & 0xFF0000FFu;
                v = (v * 0x00000101u) & 0x0F00F00Fu;
                v = (v * 0x00000011u) & 0xC30C30C3u;
                v = (v * 0x00000005u) & 0x49249249u;
            }
            return v;
        }

        static uint Quantise(float x) => (uint)MinSel(MaxSel((float)(x * 1024.0f), 0.0f), 1023.0f);

        public static uint Morton3D(float x, float y, float z) =>
            unchecked(ExpandBits(Quantise(x)) * 4 + ExpandBits(Quantise(y)) * 2 + ExpandBits(Quantise(z)));

        public static void MortonAndAabb(Float3 a, Float3 b, Float3 c, float wholeMin, float wholeMax, out uint key, out Aabb box)
        {
            float mnx = (float)(MinSel(MinSel(a.x, b.x), c.x) - 0.001f), mxx = (float)(MaxSel(MaxSel(a.x, b.x), c.x) + 0.001f);
            float mny = (float)(MinSel(MinSel(a.y, b.y), c.y) - 0.001f), mxy = (float)(MaxSel(MaxSel(a.y, b.y), c.y) + 0.001f);
            float mnz = (float)(MinSel(MinSel(a.z, b.z), c.z) - 0.001f), mxz = (float)(MaxSel(MaxSel(a.z, b.z), c.z) + 0.001f);
            float extent = (float)(wholeMax - wholeMin);
            float cx = (float)((float)((float)((float)(mnx + mxx) * 0.5f) - wholeMin) / extent);
            float cy = (float)((float)((float)((float)(mny + mxy) * 0.5f) - wholeMin) / extent);
            float cz = (float)((float)((float)((float)(mnz + mxz) * 0.5f) - wholeMin) / extent);
            key = Morton3D(cx, cy, cz);
            box = new Aabb { min = new Float3 { x = mnx, y = mny, z = mnz }, max = new Float3 { x = mxx, y = mxy, z = mxz } };
        }

        // ---- K2: ComputeBufferSorter.cs:100-126 -- net contract: stable sort by key, 4 x 8-bit LSD ----
        public static void Sort(uint[] keys, uint[] values)
        {
            int n = keys.Length;
            var k2 = new uint[n]; var v2 = new uint[n];
            for (int bitOffset = 0; bitOffset < 32; bitOffset += 8)
            {
                var count = new int[257];
                for (int i = 0; i < n; i++) count[((keys[i] >> bitOffset) & 255u) + 1]++;
                for (int d = 0; d < 256; d++) count[d + 1] += count[d];
                for (int i = 0; i < n; i++) { int dst = count[(keys[i] >> bitOffset) & 255u]++; k2[dst] = keys[i]; v2[dst] = values[i]; }
                Array.Copy(k2, keys, n); Array.Copy(v2, values, n);
            }
        }

        // ---- K3: MeshBufferContainer.cs:154-169 ---------------------------------------------------
        public static void DistributeKeys(uint[] keys, uint trianglesLength)
        {
            if (trianglesLength == 0) return;
            uint newCurrentValue = 0, oldCurrentValue = keys[0];
            keys[0] = newCurrentValue;
            for (uint i = 1; i < trianglesLength; i++)
            {
                unchecked { newCurrentValue += Math.Max(keys[i] - oldCurrentValue, 1u); }
                oldCurrentValue = keys[i];
                keys[i] = newCurrentValue;
            }
        }

        // ---- K4: BVH.compute:18-149 -----------------------------------------------------------------
        static int Clz32(uint v) { if (v == 0) return 32; int n = 0; while ((v & 0x80000000u) == 0) { v <<= 1; n++; } return n; }
        static int Delta(uint[] codes, int x, int y, int n) => (x >= 0 && x <= n - 1 && y >= 0 && y <= n - 1) ? Clz32(codes[x] ^ codes[y]) : -1;

        public static void ConstructTree(uint[] codes, uint trianglesCount, InternalNode[] internalNodes, LeafNode[] leafNodes)
        {
            int n = (int)trianglesCount;
            for (int idx = 0; idx < n - 1; idx++)
            {
                int d = Math.Sign(Delta(codes, idx, idx + 1, n) - Delta(codes, idx, idx - 1, n));
                int dmin = Delta(codes, idx, idx - d, n);
                uint lmax = 2;
                while (Delta(codes, idx, unchecked((int)((uint)idx + lmax * (uint)d)), n) > dmin) lmax *= 2;   // 32-bit wrap as HLSL
                int l = 0;
                for (uint t = lmax / 2; t >= 1; t /= 2)
                    if (Delta(codes, idx, unchecked((int)((uint)idx + ((uint)l + t) * (uint)d)), n) > dmin) l += (int)t;
                int j = idx + l * d, first = Math.Min(idx, j), last = Math.Max(idx, j);
                int split;
                uint firstCode = codes[first], lastCode = codes[last];
                if (firstCode == lastCode) split = (first + last) >> 1;
                else
                {
                    int common = Clz32(firstCode ^ lastCode), step = last - first; split = first;
                    do
                    {
                        step = (step + 1) >> 1;
                        int cand = split + step;
                        if (cand < last && Clz32(firstCode ^ codes[cand]) > common) split = cand;
                    } while (step > 1);
                }
                internalNodes[idx].index = (uint)idx;
                if (split == first) { leafNodes[split] = new LeafNode { parent = (uint)idx, index = (uint)split }; internalNodes[idx].leftNode = (uint)split; internalNodes[idx].leftNodeType = LeafNodeType; }
                else { internalNodes[split].parent = (uint)idx; internalNodes[idx].leftNode = (uint)split; internalNodes[idx].leftNodeType = InternalNodeType; }
                if (split + 1 == last) { leafNodes[split + 1] = new LeafNode { parent = (uint)idx, index = (uint)(split + 1) }; internalNodes[idx].rightNode = (uint)(split + 1); internalNodes[idx].rightNodeType = LeafNodeType; }
                else { internalNodes[split + 1].parent = (uint)idx; internalNodes[idx].rightNode = (uint)(split + 1); internalNodes[idx].rightNodeType = InternalNodeType; }
            }
        }

        // ---- K5: BVH.compute:152-220 (serial emulation of the atomic climb) ---------------------------
        public static void ConstructBvh(uint trianglesCount, uint[] sortedIdx, Aabb[] triAabb, InternalNode[] nodes, LeafNode[] leaves, Aabb[] bvh)
        {
            var counters = new uint[trianglesCount];
            for (uint leaf = 0; leaf < trianglesCount; leaf++)
            {
                uint parent = leaves[leaf].parent;
                while (parent != Null)
                {
                    uint old = counters[parent]; if (old == 0) counters[parent] = 1;
                    if (old == 0) break;
                    var nd = nodes[parent];
                    Aabb l = nd.leftNodeType == InternalNodeType ? bvh[nd.leftNode] : triAabb[sortedIdx[nd.leftNode]];
                    Aabb r = nd.rightNodeType == InternalNodeType ? bvh[nd.rightNode] : triAabb[sortedIdx[nd.rightNode]];
                    bvh[parent] = new Aabb
                    {
                        min = new Float3 { x = MinSel(l.min.x, r.min.x), y = MinSel(l.min.y, r.min.y), z = MinSel(l.min.z, r.min.z) },
                        max = new Float3 { x = MaxSel(l.max.x, r.max.x), y = MaxSel(l.max.y, r.max.y), z = MaxSel(l.max.z, r.max.z) }
                    };
                    parent = nd.parent;
                }
            }
        }

        // ---- K6: Raytracing.compute:37-176 -----------------------------------------------------------
        static float Dot(Float3 a, Float3 b) => (float)((float)((float)(a.x * b.x) + (float)(a.y * b.y)) + (float)(a.z * b.z));
        static Float3 Sub(Float3 a, Float3 b) => new Float3 { x = (float)(a.x - b.x), y = (float)(a.y - b.y), z = (float)(a.z - b.z) };
        static Float3 Cross(Float3 a, Float3 b) => new Float3
        {
            x = (float)((float)(a.y * b.z) - (float)(a.z * b.y)),
            y = (float)((float)(a.z * b.x) - (float)(a.x * b.z)),
            z = (float)((float)(a.x * b.y) - (float)(a.y * b.x))
        };

        public static bool RayBox(Aabb b, Ray r)
        {
            float t1x = (float)((float)(b.min.x - r.origin.x) * r.invDir.x), t2x = (float)((float)(b.max.x - r.origin.x) * r.invDir.x);
            float t1y = (float)((float)(b.min.y - r.origin.y) * r.invDir.y), t2y = (float)((float)(b.max.y - r.origin.y) * r.invDir.y);
            float t1z = (float)((float)(b.min.z - r.origin.z) * r.invDir.z), t2z = (float)((float)(b.max.z - r.origin.z) * r.invDir.z);
            float tmin = FMax(FMin(t1x, t2x), FMax(FMin(t1y, t2y), FMin(t1z, t2z)));
            float tmax = FMin(FMax(t1x, t2x), FMin(FMax(t1y, t2y), FMax(t1z, t2z)));
            return tmax > tmin && tmax > 0;
        }

        public static bool RayTriangle(Ray r, Float3 v0, Float3 v1, Float3 v2, out float dist, out float u, out float v)
        {
            dist = MaxFloat; u = 0; v = 0;
            Float3 e1 = Sub(v1, v0), e2 = Sub(v2, v0), pvec = Cross(r.dir, e2);
            float det = Dot(e1, pvec);
            if (det < 1e-8f && det > -1e-8f) return false;
            float invDet = (float)(1.0f / det);
            Float3 tvec = Sub(r.origin, v0);
            u = (float)(Dot(tvec, pvec) * invDet);
            if (u < 0 || u > 1) return false;
            Float3 qvec = Cross(tvec, e1);
            v = (float)(Dot(r.dir, qvec) * invDet);
            if (v < 0 || (float)(u + v) > 1) return false;
            dist = (float)(Dot(e2, qvec) * invDet);
            return true;
        }

        public static RaycastResult Traverse(Ray ray, uint[] sortedIdx, Aabb[] triAabb, InternalNode[] nodes, LeafNode[] leaves, Aabb[] bvh, Float3[] va, Float3[] vb, Float3[] vc)
        {
            var best = new RaycastResult { distance = MaxFloat, triangleIndex = 0, u = 0, v = 0 };
            var stack = new uint[64]; int sp = 0; stack[sp++] = 0;
            while (sp != 0)
            {
                uint index = stack[--sp];
                if (!RayBox(bvh[index], ray)) continue;
                for (int side = 0; side < 2; side++)
                {
                    uint child = side == 0 ? nodes[index].leftNode : nodes[index].rightNode;
                    uint type = side == 0 ? nodes[index].leftNodeType : nodes[index].rightNodeType;
                    if (type == InternalNodeType) { stack[sp++] = child; continue; }
                    uint tri = sortedIdx[leaves[child].index];
                    if (!RayBox(triAabb[tri], ray)) continue;
                    if (RayTriangle(ray, va[tri], vb[tri], vc[tri], out float d, out float u, out float v) && d < best.distance)
                        best = new RaycastResult { distance = d, triangleIndex = tri, u = u, v = v };
                }
            }
            return best;
        }
    }
}
