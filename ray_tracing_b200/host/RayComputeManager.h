// RayComputeManager.h — C++ host side above the C-ABI, mirroring the reference's C# MonoBehaviour
// Assets/Scripts/Tracer/RayComputeManager.cs (same public field names, ResetAccumulatedRender(),
// RenderFrame(), the same per-frame call sequence InitFrame -> InitTexturesAndBuffers / InitBVH /
// UpdateModels / UpdateCameraParams / SetShaderParams -> Dispatch).  The reference's toolchain (C# / Unity)
// does not exist in this image, so the host is C++; the C# P/Invoke version of the same class is shipped as
// source in INTEGRATION.md.  Unity engine objects the manager pulls from the scene (Screen, Camera.main,
// FindObjectsByType<Model>) are explicit members here.
//
// The manager talks to an implementation of include/rt_b200.h loaded at run time (RtApi): the product
// passes librt_b200.so; tests may pass the CPU oracle, which exports the same ABI.
#pragma once
#include "BVH.h"
#include "rt_b200.h"

#include <map>
#include <memory>
#include <string>
#include <vector>

namespace Seb {

struct Matrix4x4 { float m[16]; };            // column-major, like UnityEngine.Matrix4x4
struct Color { float r, g, b, a; };
struct Vector4 { float x, y, z, w; };

struct Mesh                                     // UnityEngine.Mesh: vertices / triangles / normals
{
    std::vector<Vector3> vertices;
    std::vector<int> triangles;
    std::vector<Vector3> normals;
};

struct Transform { Matrix4x4 localToWorldMatrix, worldToLocalMatrix; };

struct Model                                    // Assets/Scripts/Types/Model.cs: material + mesh + transform
{
    RtMaterial material;
    std::shared_ptr<Mesh> mesh;
    Transform transform;
    bool logBVHStats = false;
};

struct Camera { float fieldOfView = 60.0f; float aspect = 1.0f; Transform transform; };
struct ScreenInfo { int width = 0, height = 0; };

// Function table of the C-ABI, resolved with dlsym from the library given to the manager.
struct RtApi
{
    void* dl = nullptr;
    decltype(&rtCreate) Create = nullptr; decltype(&rtDestroy) Destroy = nullptr; decltype(&rtLastError) LastError = nullptr;
    decltype(&rtSetBuffer) SetBuffer = nullptr; decltype(&rtSetInt) SetInt = nullptr; decltype(&rtSetInts) SetInts = nullptr;
    decltype(&rtSetFloat) SetFloat = nullptr; decltype(&rtSetVector) SetVector = nullptr; decltype(&rtSetMatrix) SetMatrix = nullptr;
    decltype(&rtSetBool) SetBool = nullptr; decltype(&rtResize) Resize = nullptr; decltype(&rtDispatch) Dispatch = nullptr;
    decltype(&rtReadback) Readback = nullptr; decltype(&rtSynchronize) Synchronize = nullptr;
    decltype(&rtBuildBVH) BuildBVH = nullptr;   // optional: absent from older builds of the library
    decltype(&rtReadbackAsync) ReadbackAsync = nullptr; decltype(&rtReadbackWait) ReadbackWait = nullptr;   // optional: pipelined readback
    decltype(&rtCreateMulti) CreateMulti = nullptr;   // optional: several GPUs behind one context (the all-gather of tiles inside rtDispatch)
    std::string Load(const char* path);        // returns "" on success, else the error text
    void Unload();
};

class RayComputeManager
{
public:
    // [Header("Main Settings")]                                             RayComputeManager.cs:9-19
    bool rayTracingEnabled = true;
    bool accumulate = true;
    BVH::Quality bvhQuality = BVH::Quality::High;
    bool buildBVHOnDevice = false;              // not in the reference: build each mesh's BVH with rtBuildBVH (same buffers) instead of on the host
    int maxBounceCount = 4;
    int numRaysPerPixel = 1;
    float defocusStrength = 0;
    float divergeStrength = 0.3f;
    float focusDistance = 1;

    // [Header("Sky Settings")]                                              :21-27
    bool useSky = false;
    float sunFocus = 500;
    float sunIntensity = 10;
    Color sunColor = {1, 1, 1, 1};
    bool hasSunTransform = false;               // sunTransform == null -> dirToSun = Vector3.down (:176)
    Vector3 sunForward = {0, 0, 1};

    // [Header("Debug Settings")] / [Header("Info")]                         :29-42
    Vector4 debugParams = {0, 0, 0, 0};
    int numAccumulatedFrames = 0;
    int renderSeed = 0;
    bool randomizeSeedOnEnable = false;         // the reference always draws a fresh seed in OnEnable (:64); tests pin it

    // scene objects the reference finds through Unity
    ScreenInfo Screen;
    Camera mainCamera;
    std::vector<Model> models;                  // in the order FindObjectsByType(..., InstanceID) would return them (:118)
    std::vector<RtSphere> spheres;              // extension: analytic spheres (north-star Sphere buffer)

    // results of the last BVH build (for logging / tests)
    std::vector<BVH::BuildStats> bvhStats;
    std::string lastError;

    RayComputeManager(const char* backendLibrary, int device);
    // the same manager on several GPUs of this process: the image is row-tiled over `devices`, rtDispatch ends with one NCCL
    // all-gather of the frame's tiles (rtCreateMulti); nothing else in the call sequence changes
    RayComputeManager(const char* backendLibrary, const int* devices, int deviceCount);
    ~RayComputeManager();
    bool IsRendering() const { return rayTracingEnabled; }                  // :55 (Application.isPlaying is implied)

    int OnEnable();                              // :61-67
    int ResetAccumulatedRender();                // :69-76
    int Update() { return RenderFrame(); }       // :78-82 (HandleInput is UI, out of scope)
    int RenderFrame();                           // :84-95
    int OnDestroy();                             // :238-247
    void InvalidateBVH() { hasBVH = false; }     // meshes changed: rebuild on the next frame

    // readback of the two render textures (raytraceFrameTex / accumulatedResult, :53-54)
    int ReadFrame(float* dst, size_t bytes);
    int ReadAccumulated(float* dst, size_t bytes);
    int ReadAccumulatedAsync(float* dst, size_t bytes);      // returns at once; dst is valid after WaitReadback()
    int WaitReadback();
    RtContext* Context() const { return ctx; }
    const RtApi& Api() const { return api; }

private:
    struct MeshDataLists { std::vector<RtTriangle> triangles; std::vector<RtNode> nodes; std::vector<RtModel> meshInfo; };   // :249-254

    RtApi api;
    RtContext* ctx = nullptr;
    std::vector<RtModel> meshInfo;               // :49
    bool hasBVH = false;                         // :51
    int texWidth = 0, texHeight = 0;

    int InitFrame();                             // :115-124
    int InitTexturesAndBuffers();                // :126-141
    int InitBVH();                               // :143-161
    int SetShaderParams();                       // :163-181
    int UpdateCameraParams(const Camera& cam);   // :183-190
    int UpdateModels();                          // :192-204
    MeshDataLists CreateAllMeshData();           // :206-236
    int Check(int rc, const char* what);
};

} // namespace Seb
