// UsrtNative.cs -- P/Invoke host over libusrt_b200.so for .NET outside the Unity editor
// (BASELINE.json north_star (b)). UNVERIFIED HERE: this image has no .NET toolchain; the bindings are
// kept in lock-step with include/usrt.h by tests/test_abi.py checking the header, and the same call
// sequence is exercised from Python (unitysimpleraytracing_b200/host.py). See INTEGRATION.md.
//
// The classes keep the reference's names, constructor arguments and call order
// (Assets/_Scripts/RaytracingMeshDrawer.cs:30-54,76-84), so the reference's driver compiles against
// them with ComputeBuffer replaced by DeviceBuffer.
using System;
using System.Runtime.InteropServices;

namespace Usrt
{
    [StructLayout(LayoutKind.Sequential, Pack = 16, Size = 32)] public struct AABB { public float minX, minY, minZ, _dummy0, maxX, maxY, maxZ, _dummy1; }
    [StructLayout(LayoutKind.Sequential, Size = 24)] public struct InternalNode { public uint leftNode, leftNodeType, rightNode, rightNodeType, parent, index; }
    [StructLayout(LayoutKind.Sequential, Size = 8)] public struct LeafNode { public uint parent, index; }
    [StructLayout(LayoutKind.Sequential, Size = 16)] public struct RaycastResult { public float distance; public uint triangleIndex; public float u, v; }
    [StructLayout(LayoutKind.Sequential, Pack = 16, Size = 128)]
    public struct Triangle
    {
        public float ax, ay, az, _dummy0, bx, by, bz, _dummy1, cx, cy, cz, _dummy2;
        public float a_u, a_v, b_u, b_v, c_u, c_v, _dummy3a, _dummy3b;
        public float anx, any, anz, _dummy4, bnx, bny, bnz, _dummy5, cnx, cny, cnz, _dummy6;
    }

    public enum UsrtBuffer { Keys = 0, TriangleIndex = 1, TriangleData = 2, TriangleAABB = 3, BvhData = 4, LeafNodes = 5, InternalNodes = 6 }

    internal static class Native
    {
        const string Lib = "usrt_b200";   // libusrt_b200.so
        [DllImport(Lib)] public static extern int usrt_create(int device, uint capacity, out IntPtr ctx);
        [DllImport(Lib)] public static extern int usrt_destroy(IntPtr ctx);
        [DllImport(Lib)] public static extern IntPtr usrt_last_error(IntPtr ctx);
        [DllImport(Lib)] public static extern int usrt_sync(IntPtr ctx);
        [DllImport(Lib)] public static extern int usrt_set_world_bounds(IntPtr ctx, float wholeMin, float wholeMax);
        [DllImport(Lib)] public static extern uint usrt_triangles_length(IntPtr ctx);
        [DllImport(Lib)] public static extern int usrt_upload_triangles(IntPtr ctx, [In] Triangle[] tris, uint n);
        [DllImport(Lib)] public static extern int usrt_morton(IntPtr ctx);
        [DllImport(Lib)] public static extern int usrt_sort(IntPtr ctx);
        [DllImport(Lib)] public static extern int usrt_sort_pairs_host(IntPtr ctx, [In, Out] uint[] keys, [In, Out] uint[] values, ulong count);
        [DllImport(Lib)] public static extern int usrt_sort_pairs64_host(IntPtr ctx, [In, Out] ulong[] keys, [In, Out] uint[] values, ulong count);   // ComputeBufferSorter<ulong, uint>
        [DllImport(Lib)] public static extern int usrt_set_key_mode(IntPtr ctx, int mode);   // 0 reference, 1 index tie-break (no DistributeKeys), 2 63-bit Morton keys
        [DllImport(Lib)] public static extern int usrt_upload_positions(IntPtr ctx, [In] float[] positions12PerTriangle, uint n);   // first 48 bytes of every Triangle
        [DllImport(Lib)] public static extern int usrt_upload_positions_async(IntPtr ctx, IntPtr pinnedPositions, uint n);
        [DllImport(Lib)] public static extern int usrt_distribute_keys(IntPtr ctx);
        [DllImport(Lib)] public static extern int usrt_construct_tree(IntPtr ctx);
        [DllImport(Lib)] public static extern int usrt_construct_bvh(IntPtr ctx);
        [DllImport(Lib)] public static extern int usrt_rebuild(IntPtr ctx);
        [DllImport(Lib)] public static extern int usrt_trace_primary(IntPtr ctx, int width, int height, float near, float tanHalfFov,
                                                                    [In] float[] cameraToWorldRowMajor, int y0, int y1, [Out] RaycastResult[] hostOut);
        // frame pipeline over two contexts: page-locked buffers (e.g. cudaHostAlloc'ed, or Marshal-pinned + cudaHostRegister)
        [DllImport(Lib)] public static extern int usrt_upload_triangles_async(IntPtr ctx, IntPtr pinnedTriangles, uint n);
        // page-locked host memory without a CUDA binding on the C# side (cudaHostAlloc / cudaFreeHost behind the ABI)
        [DllImport(Lib)] public static extern int usrt_host_alloc(IntPtr ctx, ulong bytes, out IntPtr hostPtr);
        [DllImport(Lib)] public static extern int usrt_host_free(IntPtr ctx, IntPtr hostPtr);
        [DllImport(Lib)] public static extern int usrt_trace_primary_async(IntPtr ctx, int width, int height, float near, float tanHalfFov,
                                                                          [In] float[] cameraToWorldRowMajor, IntPtr pinnedHostOut);
        [DllImport(Lib)] public static extern int usrt_diffuse_rays_device(IntPtr ctx, int width, int height, float near, float tanHalfFov,
                                                                          [In] float[] cameraToWorldRowMajor, IntPtr devPrimaryHits, ulong seed,
                                                                          uint firstSample, uint numSamples, IntPtr devRaysOut);
        [DllImport(Lib)] public static extern int usrt_trace_rays_device(IntPtr ctx, IntPtr devRays, ulong numRays, IntPtr devOut);
        [DllImport(Lib)] public static extern int usrt_trace_rays(IntPtr ctx, [In] float[] rays, ulong numRays, [Out] RaycastResult[] hostOut);
        [DllImport(Lib)] public static extern int usrt_upload_texture(IntPtr ctx, [In] float[] rgba, int width, int height);
        [DllImport(Lib)] public static extern int usrt_shade(IntPtr ctx, IntPtr devOut, [Out] ushort[] hostOutRgba16f);
        [DllImport(Lib)] public static extern int usrt_upload_bvh(IntPtr ctx, uint n, [In] uint[] keys, [In] uint[] triangleIndex, [In] Triangle[] triangles,
                                                                 [In] AABB[] triangleAabb, [In] AABB[] bvhData, [In] LeafNode[] leafNodes, [In] InternalNode[] internalNodes);
        [DllImport(Lib)] public static extern int usrt_download(IntPtr ctx, int buffer, IntPtr hostDst, ulong count);
        [DllImport(Lib)] public static extern int usrt_count_corrupted_nodes(IntPtr ctx, out uint leaf, out uint inner);

        public static void Check(IntPtr ctx, int rc)
        {
            if (rc != 0) throw new InvalidOperationException($"usrt error {rc}: {Marshal.PtrToStringAnsi(usrt_last_error(ctx))}");
        }
    }

    /// Stands in for UnityEngine.ComputeBuffer: names one scene buffer of a context.
    public sealed class DeviceBuffer
    {
        internal readonly IntPtr Ctx; internal readonly UsrtBuffer Which;
        internal DeviceBuffer(IntPtr ctx, UsrtBuffer which) { Ctx = ctx; Which = which; }
        public void GetData<T>(T[] dst) where T : struct          // DataBuffer.cs:50-54
        {
            var h = GCHandle.Alloc(dst, GCHandleType.Pinned);
            try { Native.Check(Ctx, Native.usrt_download(Ctx, (int)Which, h.AddrOfPinnedObject(), (ulong)dst.Length)); }
            finally { h.Free(); }
        }
    }

    /// Assets/_Scripts/MeshBufferContainer.cs
    public sealed class MeshBufferContainer : IDisposable
    {
        readonly IntPtr _ctx;
        public MeshBufferContainer(Triangle[] mesh, uint capacity = 0, int device = 0)
        {
            if (Marshal.SizeOf(typeof(Triangle)) != 128 || Marshal.SizeOf(typeof(AABB)) != 32) throw new InvalidOperationException("struct layout");   // :98-106
            int rc = Native.usrt_create(device, Math.Max(capacity, Math.Max((uint)mesh.Length, 2u)), out _ctx);
            if (rc != 0) throw new InvalidOperationException("usrt_create failed (CUDA device required): " + rc);
            Native.Check(_ctx, Native.usrt_upload_triangles(_ctx, mesh, (uint)mesh.Length));   // :148-151
            Native.Check(_ctx, Native.usrt_morton(_ctx));                                      // :123-146, on the GPU
        }
        internal IntPtr Ctx => _ctx;
        public DeviceBuffer Keys => new DeviceBuffer(_ctx, UsrtBuffer.Keys);
        public DeviceBuffer TriangleIndex => new DeviceBuffer(_ctx, UsrtBuffer.TriangleIndex);
        public DeviceBuffer TriangleData => new DeviceBuffer(_ctx, UsrtBuffer.TriangleData);
        public DeviceBuffer TriangleAABB => new DeviceBuffer(_ctx, UsrtBuffer.TriangleAABB);
        public DeviceBuffer BvhData => new DeviceBuffer(_ctx, UsrtBuffer.BvhData);
        public DeviceBuffer BvhLeafNode => new DeviceBuffer(_ctx, UsrtBuffer.LeafNodes);
        public DeviceBuffer BvhInternalNode => new DeviceBuffer(_ctx, UsrtBuffer.InternalNodes);
        public uint TrianglesLength => Native.usrt_triangles_length(_ctx);
        public void DistributeKeys() => Native.Check(_ctx, Native.usrt_distribute_keys(_ctx));   // :154-169
        public void GetAllGpuData()                                                            // :171-196
        {
            Native.Check(_ctx, Native.usrt_count_corrupted_nodes(_ctx, out uint leaf, out uint inner));
            if (leaf != 0 || inner != 0) throw new InvalidOperationException($"LEAF/INTERNAL CORRUPTED {leaf}/{inner}");
        }
        public void Dispose() => Native.usrt_destroy(_ctx);
    }

    /// Assets/_Scripts/ComputeBufferSorter.cs (TKey = TValue = uint)
    public sealed class ComputeBufferSorter : IDisposable
    {
        readonly IntPtr _ctx;
        public ComputeBufferSorter(uint dataLength, DeviceBuffer keys, DeviceBuffer values)
        {
            if (keys.Ctx != values.Ctx || keys.Which != UsrtBuffer.Keys || values.Which != UsrtBuffer.TriangleIndex) throw new ArgumentException("Sort() is bound to the container's Keys / TriangleIndex");
            if (dataLength != Native.usrt_triangles_length(keys.Ctx)) throw new ArgumentException("dataLength != TrianglesLength");
            _ctx = keys.Ctx;
        }
        public void Sort() => Native.Check(_ctx, Native.usrt_sort(_ctx));                       // :100-126
        public void Dispose() { }
    }

    /// Assets/_Scripts/BVHConstructor.cs
    public sealed class BVHConstructor : IDisposable
    {
        readonly IntPtr _ctx;
        public BVHConstructor(uint trianglesCount, DeviceBuffer sortedMortonCodes, DeviceBuffer sortedTriangleIndices, DeviceBuffer triangleAABB,
                              DeviceBuffer internalNodes, DeviceBuffer leafNodes, DeviceBuffer bvhData)
        {
            _ctx = sortedMortonCodes.Ctx;
            if (trianglesCount != Native.usrt_triangles_length(_ctx)) throw new ArgumentException("trianglesCount != TrianglesLength");
        }
        public void ConstructTree() => Native.Check(_ctx, Native.usrt_construct_tree(_ctx));    // :61-64
        public void ConstructBVH() => Native.Check(_ctx, Native.usrt_construct_bvh(_ctx));      // :66-69
        public void Dispose() { }
    }

    /// The Awake()/Update() sequence of Assets/_Scripts/RaytracingMeshDrawer.cs:30-54,76-84.
    public sealed class RaytracingMeshDrawer : IDisposable
    {
        MeshBufferContainer _container; ComputeBufferSorter _sorter; BVHConstructor _bvh;
        public void Awake(Triangle[] mesh)
        {
            _container = new MeshBufferContainer(mesh);
            _sorter = new ComputeBufferSorter(_container.TrianglesLength, _container.Keys, _container.TriangleIndex);
            _sorter.Sort();
            _container.DistributeKeys();
            _bvh = new BVHConstructor(_container.TrianglesLength, _container.Keys, _container.TriangleIndex, _container.TriangleAABB,
                                      _container.BvhInternalNode, _container.BvhLeafNode, _container.BvhData);
            _bvh.ConstructTree();
            _bvh.ConstructBVH();
            _container.GetAllGpuData();
        }
        /// cameraFov = tan(fieldOfView * Deg2Rad / 2); near = camera.nearClipPlane; matrix row-major (m00,m01,..,m33).
        public RaycastResult[] Update(int screenWidth, int screenHeight, float near, float cameraFov, float[] cameraToWorldRowMajor)
        {
            var hits = new RaycastResult[screenWidth * screenHeight];
            Native.Check(_container.Ctx, Native.usrt_trace_primary(_container.Ctx, screenWidth, screenHeight, near, cameraFov, cameraToWorldRowMajor, 0, screenHeight, hits));
            return hits;
        }
        public void Dispose() { _sorter?.Dispose(); _bvh?.Dispose(); _container?.Dispose(); }
    }
}
