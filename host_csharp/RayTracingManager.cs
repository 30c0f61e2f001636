// Shipped as source, NEVER COMPILED (no C# toolchain in the build image: dotnet, mono, mcs, csc are absent): the P/Invoke host for librt_b200.so.
// Compiled twin: ray_tracing_b200/host/RayComputeManager.cpp.  See INTEGRATION.md.
// Assets/Scripts/Tracer/RayTracingManager.cs — the reference's RayComputeManager with its ComputeShader calls swapped
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;
using Seb.AccelerationStructures;
using UnityEngine;

public class RayTracingManager : MonoBehaviour
{
    // identical public surface (RayComputeManager.cs:9-42)
    public bool rayTracingEnabled = true, accumulate = true;
    public BVH.Quality bvhQuality = BVH.Quality.High;
    [SerializeField, Range(0, 32)] int maxBounceCount = 4;
    [SerializeField] int numRaysPerPixel = 1;
    [SerializeField, Min(0)] float defocusStrength = 0, divergeStrength = 0.3f;
    [Min(0)] public float focusDistance = 1;
    public bool useSky;
    [SerializeField] float sunFocus = 500, sunIntensity = 10;
    [SerializeField] Color sunColor = Color.white;
    public Transform sunTransform;
    public Vector4 debugParams;
    public int numAccumulatedFrames, renderSeed;
    public Sphere[] spheres = Array.Empty<Sphere>();            // north-star extension
    public int gpuCount = 1;                                    // > 1: the image is row-tiled over that many GPUs of the box, one NCCL all-gather per frame inside rtDispatch

    [StructLayout(LayoutKind.Sequential)] public struct MeshInfo { public int NodeOffset, TriangleOffset; public Matrix4x4 WorldToLocalMatrix, LocalToWorldMatrix; public RayTracingMaterial Material; }   // 224 B
    [StructLayout(LayoutKind.Sequential)] public struct Sphere { public Vector3 centre; public float radius; public RayTracingMaterial material; }                                                // 104 B

    IntPtr ctx; MeshInfo[] meshInfo; Model[] models; bool hasBVH; int width, height;
    const int kernelRayTrace = 0, kernelResetAccumulated = 1;

    void OnEnable() { RtB200.Check(IntPtr.Zero, gpuCount > 1 ? RtB200.rtCreateMulti(out ctx, null, gpuCount) : RtB200.rtCreate(out ctx, 0)); hasBVH = false; renderSeed = new System.Random().Next(); ResetAccumulatedRender(); }
    void OnDestroy() { if (ctx != IntPtr.Zero) RtB200.rtDestroy(ctx); ctx = IntPtr.Zero; }
    void Update() { Render(); }

    public void ResetAccumulatedRender()
    {
        numAccumulatedFrames = 1;
        InitFrame();
        RtB200.Check(ctx, RtB200.rtDispatch(ctx, kernelResetAccumulated, (width + 7) / 8, (height + 7) / 8, 1));
    }

    public void Render()                                            // RenderFrame(), RayComputeManager.cs:84-95
    {
        if (!(Application.isPlaying && rayTracingEnabled)) return;
        InitFrame();
        RtB200.Check(ctx, RtB200.rtDispatch(ctx, kernelRayTrace, (width + 7) / 8, (height + 7) / 8, 1));
        if (accumulate) numAccumulatedFrames++;
    }

    void InitFrame()
    {
        width = Screen.width; height = Screen.height;
        RtB200.Check(ctx, RtB200.rtResize(ctx, width, height));
        RtB200.Check(ctx, RtB200.rtSetInts(ctx, "Resolution", new[] { width, height }, 2));
        models = FindObjectsByType<Model>(FindObjectsInactive.Exclude, FindObjectsSortMode.InstanceID);
        if (!hasBVH) { hasBVH = true; UploadMeshes(); }
        for (int i = 0; i < models.Length; i++)
        {
            meshInfo[i].WorldToLocalMatrix = models[i].transform.worldToLocalMatrix;
            meshInfo[i].LocalToWorldMatrix = models[i].transform.localToWorldMatrix;
            meshInfo[i].Material = models[i].material;
        }
        RtB200.Check(ctx, RtB200.rtSetBuffer(ctx, "ModelInfo", meshInfo, meshInfo.Length, 224));
        RtB200.Check(ctx, RtB200.rtSetInt(ctx, "modelCount", models.Length));
        RtB200.Check(ctx, RtB200.rtSetBuffer(ctx, "Spheres", spheres, spheres.Length, 104));
        Camera cam = Camera.main;
        float planeHeight = focusDistance * Mathf.Tan(cam.fieldOfView * 0.5f * Mathf.Deg2Rad) * 2;
        Vector4 view = new Vector3(planeHeight * cam.aspect, planeHeight, focusDistance);
        Matrix4x4 camMat = cam.transform.localToWorldMatrix;
        RtB200.Check(ctx, RtB200.rtSetVector(ctx, "ViewParams", ref view));
        RtB200.Check(ctx, RtB200.rtSetMatrix(ctx, "CamLocalToWorldMatrix", ref camMat));
        RtB200.rtSetInt(ctx, "Frame", numAccumulatedFrames); RtB200.rtSetInt(ctx, "UseSky", useSky ? 1 : 0);
        RtB200.rtSetInt(ctx, "MaxBounceCount", maxBounceCount); RtB200.rtSetInt(ctx, "NumRaysPerPixel", numRaysPerPixel);
        RtB200.rtSetFloat(ctx, "DefocusStrength", defocusStrength); RtB200.rtSetFloat(ctx, "DivergeStrength", divergeStrength);
        RtB200.rtSetFloat(ctx, "SunFocus", sunFocus); RtB200.rtSetFloat(ctx, "SunIntensity", sunIntensity);
        Vector4 sc = sunColor; Vector4 ds = sunTransform == null ? Vector3.down : -sunTransform.forward;
        RtB200.rtSetVector(ctx, "SunColour", ref sc); RtB200.rtSetVector(ctx, "dirToSun", ref ds);
        RtB200.rtSetInt(ctx, "renderSeed", renderSeed); RtB200.rtSetBool(ctx, "accumulate", accumulate ? 1 : 0);
    }

    void UploadMeshes()                                             // CreateAllMeshData + InitBVH, RayComputeManager.cs:143-161,206-236
    {
        var tris = new List<BVH.Triangle>(); var nodes = new List<BVH.Node>(); var infos = new List<MeshInfo>();
        var lookup = new Dictionary<Mesh, (int, int)>();
        foreach (Model m in models)
        {
            if (!lookup.ContainsKey(m.Mesh))
            {
                lookup.Add(m.Mesh, (nodes.Count, tris.Count));
                BVH bvh = new(m.Mesh.vertices, m.Mesh.triangles, m.Mesh.normals, bvhQuality);
                tris.AddRange(bvh.Triangles); nodes.AddRange(bvh.Nodes);
            }
            infos.Add(new MeshInfo { NodeOffset = lookup[m.Mesh].Item1, TriangleOffset = lookup[m.Mesh].Item2 });
        }
        meshInfo = infos.ToArray();
        RtB200.Check(ctx, RtB200.rtSetBuffer(ctx, "Triangles", tris.ToArray(), tris.Count, 72));
        RtB200.Check(ctx, RtB200.rtSetBuffer(ctx, "Nodes", nodes.ToArray(), nodes.Count, 32));
    }
}
