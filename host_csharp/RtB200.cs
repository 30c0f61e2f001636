// Shipped as source, NEVER COMPILED (no C# toolchain in the build image: dotnet, mono, mcs, csc are absent): the P/Invoke host for librt_b200.so.
// Compiled twin: ray_tracing_b200/host/RayComputeManager.cpp.  See INTEGRATION.md.
// Assets/Scripts/Tracer/RtB200.cs
using System;
using System.Runtime.InteropServices;
using Seb.AccelerationStructures;     // BVH (Assets/Scripts/Types/BVH.cs declares it in this namespace)

public static class RtB200
{
    const string Lib = "rt_b200";            // librt_b200.so next to the player / in Assets/Plugins/x86_64

    [DllImport(Lib)] public static extern int rtCreate(out IntPtr ctx, int device);
    [DllImport(Lib)] public static extern int rtDestroy(IntPtr ctx);
    [DllImport(Lib)] public static extern IntPtr rtLastError(IntPtr ctx);
    [DllImport(Lib)] public static extern int rtSetBuffer(IntPtr ctx, string name, BVH.Triangle[] data, int count, int stride);
    [DllImport(Lib)] public static extern int rtSetBuffer(IntPtr ctx, string name, BVH.Node[] data, int count, int stride);
    [DllImport(Lib)] public static extern int rtSetBuffer(IntPtr ctx, string name, RayTracingManager.MeshInfo[] data, int count, int stride);
    [DllImport(Lib)] public static extern int rtSetBuffer(IntPtr ctx, string name, RayTracingManager.Sphere[] data, int count, int stride);
    [DllImport(Lib)] public static extern int rtSetInt(IntPtr ctx, string name, int value);
    [DllImport(Lib)] public static extern int rtSetInts(IntPtr ctx, string name, int[] values, int n);
    [DllImport(Lib)] public static extern int rtSetFloat(IntPtr ctx, string name, float value);
    [DllImport(Lib)] public static extern int rtSetVector(IntPtr ctx, string name, ref UnityEngine.Vector4 value);
    [DllImport(Lib)] public static extern int rtSetMatrix(IntPtr ctx, string name, ref UnityEngine.Matrix4x4 value);   // column-major, as Unity stores it
    [DllImport(Lib)] public static extern int rtSetBool(IntPtr ctx, string name, int value);
    [DllImport(Lib)] public static extern int rtResize(IntPtr ctx, int width, int height);
    [DllImport(Lib)] public static extern int rtDispatch(IntPtr ctx, int kernelIndex, int gx, int gy, int gz);
    [DllImport(Lib)] public static extern int rtReadback(IntPtr ctx, string tex, float[] dst, UIntPtr bytes);
    [DllImport(Lib)] public static extern int rtSynchronize(IntPtr ctx);
    // optional: new BVH(verts, indices, normals, quality) built on the GPU — the same Triangles / Nodes arrays (BVH.cs:26)
    [DllImport(Lib)] public static extern int rtBuildBVH(IntPtr ctx, UnityEngine.Vector3[] verts, int vertCount, int[] indices, int indexCount, UnityEngine.Vector3[] normals,
                                                         int quality, [Out] BVH.Triangle[] tris, [Out] BVH.Node[] nodes, int nodeCapacity, out int nodeCount);

    // Display.shader:42-47 (tex / Frame, sRGB 8-bit): RayTraceDisplay.OnRenderImage and the screenshot of RayComputeManager.cs:106-111
    [DllImport(Lib)] public static extern int rtDisplay(IntPtr ctx, int useAccumulated, int frame, byte[] rgba, UIntPtr bytes);

    // pipelined forms: the copy of frame k travels while frame k+1 renders (dst pinned: GCHandle.Alloc(..., GCHandleType.Pinned))
    [DllImport(Lib)] public static extern int rtReadbackAsync(IntPtr ctx, string tex, IntPtr dst, UIntPtr bytes);
    [DllImport(Lib)] public static extern int rtDisplayAsync(IntPtr ctx, int useAccumulated, int frame, IntPtr rgba, UIntPtr bytes);
    [DllImport(Lib)] public static extern int rtReadbackWait(IntPtr ctx);

    // ---- extensions (no counterpart in the reference) ----
    [DllImport(Lib)] public static extern int rtGetVersion();
    [DllImport(Lib)] public static extern int rtSetOption(IntPtr ctx, string name, int value);           // "kernel", "tlas", "modelSkip", ... (rt_b200.h)
    [DllImport(Lib)] public static extern int rtSetStream(IntPtr ctx, IntPtr cudaStream);
    [StructLayout(LayoutKind.Sequential)]
    public struct Stats { public ulong rays, boxTests, triTests, sphereTests, dispatches; public double kernelMs; public ulong sphereBoxTests; public double exchangeMs; }
    [DllImport(Lib)] public static extern int rtGetStats(IntPtr ctx, out Stats stats);
    [DllImport(Lib)] public static extern int rtResetStats(IntPtr ctx);
    // multi-GPU inside the boundary.  A Unity host is ONE process: rtCreateMulti(out ctx, null, 8) instead of rtCreate, and nothing else changes —
    // every rtSet* / rtSetBuffer reaches all GPUs, rtDispatch(kernelRayTrace) traces tile r of 8 on GPU r and ends with the NCCL all-gather of the frame's
    // tiles, rtReadback / rtDisplay read GPU 0.  One process per GPU instead: rtGetUniqueId on rank 0, rtCommInit on every rank.
    [DllImport(Lib)] public static extern int rtCreateMulti(out IntPtr ctx, int[] devices, int nDevices);
    [DllImport(Lib)] public static extern int rtGetUniqueId(byte[] id128, UIntPtr bytes);
    [DllImport(Lib)] public static extern int rtCommInit(IntPtr ctx, byte[] id128, UIntPtr bytes, int rank, int worldSize);
    [DllImport(Lib)] public static extern int rtCommDestroy(IntPtr ctx);
    [DllImport(Lib)] public static extern int rtExchangeTiles(IntPtr ctx);
    // row bands by hand (option "exchange" = 0): the caller owns the collective between rtPackTile and rtUnpackTiles
    [DllImport(Lib)] public static extern int rtSetTile(IntPtr ctx, int rank, int worldSize, int bandRows);
    [DllImport(Lib)] public static extern int rtPackTile(IntPtr ctx);
    [DllImport(Lib)] public static extern int rtUnpackTiles(IntPtr ctx);
    [DllImport(Lib)] public static extern int rtGetDevicePointer(IntPtr ctx, string name, out IntPtr devPtr, out UIntPtr bytes);
    [DllImport(Lib)] public static extern int rtGetIpcHandles(IntPtr ctx, byte[] handles, UIntPtr bytes);
    [DllImport(Lib)] public static extern int rtSetPeers(IntPtr ctx, int nPeers, byte[] handles, UIntPtr bytes);

    public static void Check(IntPtr ctx, int rc)
    {
        if (rc != 0) throw new InvalidOperationException("rt_b200: " + Marshal.PtrToStringAnsi(rtLastError(ctx)));
    }
}
