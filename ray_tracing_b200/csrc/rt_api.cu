// rt_api.cu — the C-ABI of librt_b200.so (include/rt_b200.h): context, named uniforms, structured
// buffers, render targets, dispatch, readback, multi-GPU tile staging.  Host-side state only; all
// arithmetic of the path lives in rt_device.cuh / rt_kernel_*.cuh.
//
// The interface mirrors the ComputeShader calls RayComputeManager makes (each entry point
// cites its reference counterpart in the header).  There is deliberately no CPU path: every entry point
// that needs the GPU fails with RT_E_NO_DEVICE / RT_E_CUDA when CUDA is unavailable.
#include "../../include/rt_b200.h"
#include "rt_kernel_mega.cuh"
#include "rt_kernel_wave.cuh"
#include "rt_kernel_pool.cuh"
#include "rt_repack.cuh"
#include "rt_bvh_build.cuh"
#include "rt_comm.cuh"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace rtd;

namespace {

template <class T> struct DevBuf
{
    T* p = nullptr; size_t count = 0, cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n == count && p) return cudaSuccess;             // unchanged size re-uses the allocation (CH:86-92)
        if (n > cap || n == 0) { if (p) cudaFree(p); p = nullptr; cap = 0; if (n) { cudaError_t e = cudaMalloc(&p, n * sizeof(T)); if (e != cudaSuccess) { count = 0; return e; } cap = n; } }
        count = n; return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; count = cap = 0; }
};

struct EventPair { cudaEvent_t a, b; };

} // namespace

struct RtContext
{
    int device = 0;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    int numSMs = 0;
    unsigned long long dispatchPixels = 0;     // pixels of the RayTrace dispatch being prepared (this rank's tile): enters the automatic kernel choice
    std::string err;

    // uniforms (host copy)
    DevParams P;

    // reference-layout buffers
    DevBuf<RtNode> nodes; DevBuf<RtTriangle> tris; DevBuf<RtModel> models; DevBuf<RtSphere> spheres;
    std::vector<RtNode> hNodes; std::vector<RtModel> hModels; std::vector<RtSphere> hSpheres;   // host mirrors for the repack
    // repacked buffers
    RepackState repack;
    bool sceneDirty = true, modelsDirty = true, spheresDirty = true;

    // render targets
    DevBuf<float4> frame, accum, tileSend, tileRecv;
    DevBuf<uchar4> display;
    int width = 0, height = 0;
    int tileRank = 0, tileWorld = 1, bandRows = 1;

    // peers (fused tile exchange): IPC-mapped images of the other ranks
    std::vector<void*> peerBase;               // opened IPC mappings (to close)
    float4* peerFrame[RT_MAX_PEERS]; float4* peerAccum[RT_MAX_PEERS]; int nPeers = 0;

    // options
    int optKernel = -1, optCountStats = 0, optSmemPairs = -1, optPoolSlots = 0, optTailLanes = 16, optSortRays = 0, optForceExt = 0, optModelSkip = 1, optPairOrder = 0, optGridFit = 0, optL2Persist = 0, optTreeletPrefetch = 0, optZeroDefocus = 1, optTlas = -1;
    int l2PersistApplied = 0; const void* l2PersistBase = nullptr; size_t l2PersistBytes = 0;   // kernel -1 = automatic   // smemPairs -1 = automatic

    // exchange of finished tiles inside the ABI (rt_comm.cuh): a communicator over the ranks of a tiled render — one process per
    // GPU (rtCommInit) or several GPUs in this process (rtCreateMulti: this context leads, `followers` render the other tiles)
    NcclApi::Comm comm = nullptr; int commRank = 0, commWorld = 1;
    int optExchange = 1;
    std::vector<RtContext*> followers; RtContext* leader = nullptr;

    // pipelined readback (rtReadbackAsync / rtDisplayAsync): a snapshot on the dispatch stream, the device -> host copy on a stream of
    // its own, so that the next frame's kernel runs while the previous frame's result travels over PCIe
    cudaStream_t copyStream = nullptr; cudaEvent_t snapReady = nullptr, copyDone = nullptr; bool copyPending = false;
    DevBuf<float4> snap;

    // kernel 1 on small tiles: hand-off of a pixel's chain between sample chunks (rt_kernel_wave.cuh)
    DevBuf<unsigned long long> handoff; unsigned int chunkSerial = 0; int optSampleChunks = -1;
    DevBuf<float> poolCold;                    // kernel 2 (RT_POOL_COLD_GLOBAL builds): slot fields of the shade phase, one block per resident warp

    // rtBuildBVH: device arena kept between builds, pinned staging chunks for the copies of caller-owned arrays
    DevBuf<unsigned char> buildArena;
    static constexpr size_t STAGE_BYTES = 16u << 20;
    unsigned char* stageBuf[2] = {nullptr, nullptr}; cudaEvent_t stageEv[2] = {nullptr, nullptr};

    // counters / timing
    unsigned long long* dCounters = nullptr;   // 5
    unsigned int* dWork = nullptr;
    std::vector<EventPair> pending, pendingX, freeEvents;
    RtStats stats;
};

static std::string g_createErr = "";

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return failCuda(c, e_, #call); } while (0)

static int fail(RtContext* c, int code, const std::string& msg) { if (c) c->err = msg; else g_createErr = msg; return code; }
static int failCuda(RtContext* c, cudaError_t e, const char* what)
{
    return fail(c, RT_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

static int drainEvents(RtContext* c)
{
    for (auto& ev : c->pending)
    {
        CK(cudaEventSynchronize(ev.b));
        float ms = 0; CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
        c->stats.kernelMs += ms;
        c->freeEvents.push_back(ev);
    }
    c->pending.clear();
    for (auto& ev : c->pendingX)
    {
        CK(cudaEventSynchronize(ev.b));
        float ms = 0; CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
        c->stats.exchangeMs += ms;
        c->freeEvents.push_back(ev);
    }
    c->pendingX.clear();
    return RT_OK;
}

static void closePeers(RtContext* c)
{
    for (void* p : c->peerBase) cudaIpcCloseMemHandle(p);
    c->peerBase.clear(); c->nPeers = 0;
}

extern "C" {

static void destroyComm(RtContext* c);

int rtGetVersion(void) { return RT_B200_VERSION; }

int rtCreate(RtContext** out, int device)
{
    if (!out) return fail(nullptr, RT_E_INVALID, "rtCreate: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, RT_E_NO_DEVICE, std::string("rtCreate: no CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") + "); librt_b200 has no CPU path");
    if (device < 0 || device >= n) return fail(nullptr, RT_E_INVALID, "rtCreate: device ordinal out of range");
    RtContext* c = new RtContext();
    memset(&c->P, 0, sizeof(c->P));
    memset(&c->stats, 0, sizeof(c->stats));
    c->device = device;
    c->P.tileWorld = 1; c->P.bandRows = 1;
    const float ident[16] = {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1};
    memcpy(c->P.cam, ident, sizeof(ident));
    auto bail = [&](cudaError_t err, const char* what)
    {
        const std::string m = std::string(what) + ": " + cudaGetErrorString(err);
        if (c->dCounters) cudaFree(c->dCounters);
        if (c->dWork) cudaFree(c->dWork);
        if (c->ownStream) cudaStreamDestroy(c->ownStream);
        delete c;
        return fail(nullptr, RT_E_CUDA, m);
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    if (prop.major < 10) { delete c; return fail(nullptr, RT_E_NO_DEVICE, "rtCreate: device is not sm_100-class; this library carries sm_100a code only"); }
    c->numSMs = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    c->stream = c->ownStream;
    if ((e = cudaMalloc(&c->dCounters, 8 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc(&c->dWork, 64)) != cudaSuccess) return bail(e, "cudaMalloc");
    cudaMemset(c->dCounters, 0, 8 * sizeof(unsigned long long));
    cudaMemset(c->dWork, 0, 64);
    if ((e = wave_configure()) != cudaSuccess) return bail(e, "cudaFuncSetAttribute");
    if ((e = pool_configure()) != cudaSuccess) return bail(e, "cudaFuncSetAttribute");
    *out = c;
    return RT_OK;
}

int rtDestroy(RtContext* c)
{
    if (!c) return RT_E_INVALID;
    if (c->leader) return fail(c, RT_E_STATE, "rtDestroy: this context belongs to a group; destroy the context rtCreateMulti returned");
    {
        std::vector<RtContext*> members; members.swap(c->followers);
        destroyComm(c);
        for (RtContext* m : members) { destroyComm(m); m->leader = nullptr; rtDestroy(m); }
    }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->nodes.release(); c->tris.release(); c->models.release(); c->spheres.release();
    if (c->copyStream) { cudaStreamSynchronize(c->copyStream); cudaStreamDestroy(c->copyStream); }
    if (c->snapReady) cudaEventDestroy(c->snapReady);
    if (c->copyDone) cudaEventDestroy(c->copyDone);
    c->snap.release();
    c->buildArena.release(); c->handoff.release(); c->poolCold.release();
    for (int k = 0; k < 2; k++) { if (c->stageBuf[k]) cudaFreeHost(c->stageBuf[k]); if (c->stageEv[k]) cudaEventDestroy(c->stageEv[k]); }
    c->frame.release(); c->accum.release(); c->tileSend.release(); c->tileRecv.release(); c->display.release();
    c->repack.release();
    closePeers(c);
    for (auto& ev : c->pending) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    for (auto& ev : c->pendingX) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    for (auto& ev : c->freeEvents) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    if (c->dCounters) cudaFree(c->dCounters);
    if (c->dWork) cudaFree(c->dWork);
    if (c->ownStream) cudaStreamDestroy(c->ownStream);
    delete c;
    return RT_OK;
}

const char* rtLastError(const RtContext* c) { return c ? c->err.c_str() : g_createErr.c_str(); }

int rtSetStream(RtContext* c, void* s)
{
    if (!c) return RT_E_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));           // work already queued finishes before the switch
    c->stream = s ? (cudaStream_t)s : c->ownStream;
    return RT_OK;
}

static int one_rtSetBuffer(RtContext* c, const char* name, const void* data, int count, int stride)
{
    if (!c || !name || count < 0 || (count > 0 && !data)) return fail(c, RT_E_INVALID, "rtSetBuffer: bad argument");
    CK(cudaSetDevice(c->device));
    const std::string n(name);
    if (n == "Triangles")
    {
        if (stride != (int)sizeof(RtTriangle)) return fail(c, RT_E_INVALID, "rtSetBuffer: Triangles stride must be 72");
        CK(c->tris.ensure(count));
        if (count) CK(cudaMemcpyAsync(c->tris.p, data, (size_t)count * sizeof(RtTriangle), cudaMemcpyHostToDevice, c->stream));
        c->sceneDirty = true;
        return RT_OK;
    }
    if (n == "Nodes")
    {
        if (stride != (int)sizeof(RtNode)) return fail(c, RT_E_INVALID, "rtSetBuffer: Nodes stride must be 32");
        CK(c->nodes.ensure(count));
        c->hNodes.assign((const RtNode*)data, (const RtNode*)data + count);
        if (count) CK(cudaMemcpyAsync(c->nodes.p, c->hNodes.data(), (size_t)count * sizeof(RtNode), cudaMemcpyHostToDevice, c->stream));
        c->sceneDirty = true;
        return RT_OK;
    }
    if (n == "ModelInfo")
    {
        if (stride != (int)sizeof(RtModel)) return fail(c, RT_E_INVALID, "rtSetBuffer: ModelInfo stride must be 224");
        // the BVH repack depends only on the (nodeOffset, triOffset) pairs; matrices / materials change every frame (RCM:192-204)
        bool sameTopology = (size_t)count == c->hModels.size();
        const RtModel* m = (const RtModel*)data;
        // the reference re-sends ModelInfo every frame (RCM:192-204), mostly with the bytes of the frame before: nothing to do then
        if (sameTopology && count > 0 && c->models.count == (size_t)count && memcmp(m, c->hModels.data(), (size_t)count * sizeof(RtModel)) == 0) return RT_OK;
        for (int i = 0; sameTopology && i < count; i++)
            sameTopology = m[i].nodeOffset == c->hModels[i].nodeOffset && m[i].triOffset == c->hModels[i].triOffset;
        if (!sameTopology) c->sceneDirty = true;
        CK(c->models.ensure(count));
        c->hModels.assign(m, m + count);
        if (count) CK(cudaMemcpyAsync(c->models.p, c->hModels.data(), (size_t)count * sizeof(RtModel), cudaMemcpyHostToDevice, c->stream));
        c->modelsDirty = true;
        return RT_OK;
    }
    if (n == "Spheres")
    {
        if (stride != (int)sizeof(RtSphere)) return fail(c, RT_E_INVALID, "rtSetBuffer: Spheres stride must be 104");
        if (count > 0 && c->spheres.count == (size_t)count && c->hSpheres.size() == (size_t)count && memcmp(data, c->hSpheres.data(), (size_t)count * sizeof(RtSphere)) == 0) return RT_OK;
        CK(c->spheres.ensure(count));
        c->hSpheres.assign((const RtSphere*)data, (const RtSphere*)data + count);
        if (count) CK(cudaMemcpyAsync(c->spheres.p, c->hSpheres.data(), (size_t)count * sizeof(RtSphere), cudaMemcpyHostToDevice, c->stream));
        c->spheresDirty = true;
        return RT_OK;
    }
    return fail(c, RT_E_UNKNOWN_NAME, std::string("rtSetBuffer: unknown buffer ") + name);
}

static int one_rtSetInt(RtContext* c, const char* name, int v)
{
    if (!c || !name) return fail(c, RT_E_INVALID, "rtSetInt: bad argument");
    const std::string n(name);
    if (n == "Frame") c->P.Frame = v;
    else if (n == "UseSky") c->P.UseSky = v;
    else if (n == "MaxBounceCount") c->P.MaxBounceCount = v;
    else if (n == "NumRaysPerPixel") c->P.NumRaysPerPixel = v;
    else if (n == "renderSeed") c->P.renderSeed = v;
    else if (n == "modelCount")
    {
        // the pair plan, the mesh roots, the DevModel records and the TLAS are all sized for the count they were built with
        if (v != c->P.modelCount) c->sceneDirty = c->modelsDirty = true;
        c->P.modelCount = v;
    }
    else if (n == "triangleCount" || n == "visMode") { /* declared but never read by the shader (HL:24,120) */ }
    else return fail(c, RT_E_UNKNOWN_NAME, std::string("rtSetInt: unknown uniform ") + name);
    return RT_OK;
}

static int one_rtSetInts(RtContext* c, const char* name, const int* v, int n)
{
    if (!c || !name || !v) return fail(c, RT_E_INVALID, "rtSetInts: bad argument");
    if (std::string(name) == "Resolution")
    {
        if (n != 2 || v[0] <= 0 || v[1] <= 0) return fail(c, RT_E_INVALID, "rtSetInts: Resolution takes 2 positive values");
        c->P.W = (unsigned)v[0]; c->P.H = (unsigned)v[1];
        return RT_OK;
    }
    return fail(c, RT_E_UNKNOWN_NAME, std::string("rtSetInts: unknown uniform ") + name);
}

static int one_rtSetFloat(RtContext* c, const char* name, float v)
{
    if (!c || !name) return fail(c, RT_E_INVALID, "rtSetFloat: bad argument");
    const std::string n(name);
    if (n == "DefocusStrength") c->P.DefocusStrength = v;
    else if (n == "DivergeStrength") c->P.DivergeStrength = v;
    else if (n == "SunFocus") c->P.SunFocus = v;
    else if (n == "SunIntensity") c->P.SunIntensity = v;
    else if (n == "debugVisScale") { }
    else return fail(c, RT_E_UNKNOWN_NAME, std::string("rtSetFloat: unknown uniform ") + name);
    return RT_OK;
}

static int one_rtSetVector(RtContext* c, const char* name, const float v[4])
{
    if (!c || !name || !v) return fail(c, RT_E_INVALID, "rtSetVector: bad argument");
    const std::string n(name);
    if (n == "ViewParams") memcpy(c->P.ViewParams, v, 12);
    else if (n == "SunColour") memcpy(c->P.SunColour, v, 12);
    else if (n == "dirToSun") memcpy(c->P.dirToSun, v, 12);
    else if (n == "debugParams") { }
    else return fail(c, RT_E_UNKNOWN_NAME, std::string("rtSetVector: unknown uniform ") + name);
    return RT_OK;
}

static int one_rtSetMatrix(RtContext* c, const char* name, const float v[16])
{
    if (!c || !name || !v) return fail(c, RT_E_INVALID, "rtSetMatrix: bad argument");
    if (std::string(name) == "CamLocalToWorldMatrix") { memcpy(c->P.cam, v, 64); return RT_OK; }
    return fail(c, RT_E_UNKNOWN_NAME, std::string("rtSetMatrix: unknown uniform ") + name);
}

static int one_rtSetBool(RtContext* c, const char* name, int v)
{
    if (!c || !name) return fail(c, RT_E_INVALID, "rtSetBool: bad argument");
    if (std::string(name) == "accumulate") { c->P.accumulate = v != 0; return RT_OK; }
    return fail(c, RT_E_UNKNOWN_NAME, std::string("rtSetBool: unknown uniform ") + name);
}

static int one_rtResize(RtContext* c, int w, int h)
{
    if (!c || w <= 0 || h <= 0) return fail(c, RT_E_INVALID, "rtResize: bad size");
    CK(cudaSetDevice(c->device));
    if (w != c->width || h != c->height)
    {
        if (c->copyPending) { CK(cudaEventSynchronize(c->copyDone)); c->copyPending = false; }
        const size_t n = (size_t)w * h;
        CK(c->frame.ensure(n)); CK(c->accum.ensure(n));
        CK(cudaMemsetAsync(c->frame.p, 0, n * 16, c->stream));
        CK(cudaMemsetAsync(c->accum.p, 0, n * 16, c->stream));
        c->width = w; c->height = h;
        c->tileSend.release(); c->tileRecv.release();
        closePeers(c);                                      // peer mappings refer to images of the old size
    }
    c->P.W = (unsigned)w; c->P.H = (unsigned)h;
    return RT_OK;
}

int rtSetTile(RtContext* c, int rank, int world, int bandRows)
{
    if (!c || world < 1 || rank < 0 || rank >= world || bandRows < 1) return fail(c, RT_E_INVALID, "rtSetTile: bad argument");
    if (c->leader) return fail(c, RT_E_STATE, "rtSetTile: this context belongs to a group");
    if (!c->followers.empty())
    {
        // a group's ranks are its GPUs; only the band height is the caller's
        if (rank != 0 || world != (int)c->followers.size() + 1) return fail(c, RT_E_INVALID, "rtSetTile: a group renders as rank 0 of its own GPU count; only bandRows can change");
        for (RtContext* m : c->followers) { if (bandRows != m->bandRows) { m->tileSend.release(); m->tileRecv.release(); } m->bandRows = bandRows; }
    }
    if (c->comm && (rank != c->commRank || world != c->commWorld)) return fail(c, RT_E_INVALID, "rtSetTile: rank / world differ from the communicator's (rtCommInit)");
    if (rank != c->tileRank || world != c->tileWorld || bandRows != c->bandRows) { c->tileSend.release(); c->tileRecv.release(); }
    c->tileRank = rank; c->tileWorld = world; c->bandRows = bandRows;
    return RT_OK;
}

static int one_rtSetOption(RtContext* c, const char* name, int value)
{
    if (!c || !name) return fail(c, RT_E_INVALID, "rtSetOption: bad argument");
    const std::string n(name);
    if (n == "kernel") { if (value < -1 || value > 2) return fail(c, RT_E_INVALID, "rtSetOption: kernel must be -1 (auto), 0, 1 or 2"); c->optKernel = value; }
    else if (n == "countStats") c->optCountStats = value != 0;
    else if (n == "exchange") c->optExchange = value != 0;
    else if (n == "sampleChunks") { if (value < -1 || value > 64) return fail(c, RT_E_INVALID, "rtSetOption: sampleChunks must be -1 (automatic), 0 / 1 (whole pixels) or 2..64"); c->optSampleChunks = value; }
    else if (n == "smemNodes") c->optSmemPairs = value;
    else if (n == "modelSkip") c->optModelSkip = value != 0;
    else if (n == "tlas")
    {
        if (value < -1 || value > 1) return fail(c, RT_E_INVALID, "rtSetOption: tlas must be -1 (automatic), 0 or 1");
        if (value != c->optTlas) { c->optTlas = value; c->modelsDirty = true; }
    }
    else if (n == "extInstantiation") c->optForceExt = value != 0;
    else if (n == "sortRays") c->optSortRays = value != 0;
    else if (n == "gridFit") c->optGridFit = value != 0;
    else if (n == "l2Persist") c->optL2Persist = value != 0;
    else if (n == "zeroDefocusShortcut") c->optZeroDefocus = value != 0;      // only RT_SKIP_ZERO_DEFOCUS builds read the flag it controls
    else if (n == "treeletPrefetch")
    {
#ifdef RT_TREELET_PREFETCH
        c->optTreeletPrefetch = value != 0;
#else
        if (value) return fail(c, RT_E_INVALID, "rtSetOption: treeletPrefetch needs a library built with RT_TREELET_PREFETCH (the default kernels do not strip the flag bit)");
#endif
    }
    else if (n == "pairOrder") { if (value < 0 || value > 32) return fail(c, RT_E_INVALID, "rtSetOption: pairOrder must be 0 (breadth-first) or a treelet depth 1..32"); c->optPairOrder = value; }
    else if (n == "tailLanes") { if (value < 0 || value > 31) return fail(c, RT_E_INVALID, "rtSetOption: tailLanes must be in [0, 31]"); c->optTailLanes = value; }
    else if (n == "poolSlots") { if (value != 0 && value != 32 && value != 64 && value != 96) return fail(c, RT_E_INVALID, "rtSetOption: poolSlots must be 0 (automatic), 32, 64 or 96"); c->optPoolSlots = value; }
    else return fail(c, RT_E_UNKNOWN_NAME, std::string("rtSetOption: unknown option ") + name);
    return RT_OK;
}

// ---- the setters of a context that leads a group (rtCreateMulti) reach every GPU of the group ----
int rtSetBuffer(RtContext* c, const char* name, const void* data, int count, int stride)
{
    int rc = one_rtSetBuffer(c, name, data, count, stride);
    if (rc != RT_OK || !c) return rc;
    for (RtContext* m : c->followers) if ((rc = one_rtSetBuffer(m, name, data, count, stride)) != RT_OK) return fail(c, rc, "GPU " + std::to_string(m->device) + ": " + m->err);
    return RT_OK;
}

int rtSetInt(RtContext* c, const char* name, int v)
{
    int rc = one_rtSetInt(c, name, v);
    if (rc != RT_OK || !c) return rc;
    for (RtContext* m : c->followers) if ((rc = one_rtSetInt(m, name, v)) != RT_OK) return fail(c, rc, "GPU " + std::to_string(m->device) + ": " + m->err);
    return RT_OK;
}

int rtSetInts(RtContext* c, const char* name, const int* v, int n)
{
    int rc = one_rtSetInts(c, name, v, n);
    if (rc != RT_OK || !c) return rc;
    for (RtContext* m : c->followers) if ((rc = one_rtSetInts(m, name, v, n)) != RT_OK) return fail(c, rc, "GPU " + std::to_string(m->device) + ": " + m->err);
    return RT_OK;
}

int rtSetFloat(RtContext* c, const char* name, float v)
{
    int rc = one_rtSetFloat(c, name, v);
    if (rc != RT_OK || !c) return rc;
    for (RtContext* m : c->followers) if ((rc = one_rtSetFloat(m, name, v)) != RT_OK) return fail(c, rc, "GPU " + std::to_string(m->device) + ": " + m->err);
    return RT_OK;
}

int rtSetVector(RtContext* c, const char* name, const float v[4])
{
    int rc = one_rtSetVector(c, name, v);
    if (rc != RT_OK || !c) return rc;
    for (RtContext* m : c->followers) if ((rc = one_rtSetVector(m, name, v)) != RT_OK) return fail(c, rc, "GPU " + std::to_string(m->device) + ": " + m->err);
    return RT_OK;
}

int rtSetMatrix(RtContext* c, const char* name, const float v[16])
{
    int rc = one_rtSetMatrix(c, name, v);
    if (rc != RT_OK || !c) return rc;
    for (RtContext* m : c->followers) if ((rc = one_rtSetMatrix(m, name, v)) != RT_OK) return fail(c, rc, "GPU " + std::to_string(m->device) + ": " + m->err);
    return RT_OK;
}

int rtSetBool(RtContext* c, const char* name, int v)
{
    int rc = one_rtSetBool(c, name, v);
    if (rc != RT_OK || !c) return rc;
    for (RtContext* m : c->followers) if ((rc = one_rtSetBool(m, name, v)) != RT_OK) return fail(c, rc, "GPU " + std::to_string(m->device) + ": " + m->err);
    return RT_OK;
}

int rtResize(RtContext* c, int w, int h)
{
    int rc = one_rtResize(c, w, h);
    if (rc != RT_OK || !c) return rc;
    for (RtContext* m : c->followers) if ((rc = one_rtResize(m, w, h)) != RT_OK) return fail(c, rc, "GPU " + std::to_string(m->device) + ": " + m->err);
    return RT_OK;
}

int rtSetOption(RtContext* c, const char* name, int value)
{
    int rc = one_rtSetOption(c, name, value);
    if (rc != RT_OK || !c) return rc;
    for (RtContext* m : c->followers) if ((rc = one_rtSetOption(m, name, value)) != RT_OK) return fail(c, rc, "GPU " + std::to_string(m->device) + ": " + m->err);
    return RT_OK;
}

// Automatic choice (measured, profiles/): with meshes the pooled wavefront kernel wins (BVH traversal needs the ray queue);
// for sphere-only scenes every ray costs the same and one path per lane avoids the pool's shared-memory round trips.
// Small tiles (multi-GPU, small images).  A pixel's samples are one sequential chain, so with about one pixel per pool slot the pools
// have nothing to balance: the slots that get a second pixel run it on a mostly empty machine for another whole pixel-time.  One path
// per lane then finishes earlier although its throughput is lower.  Measured on one GPU over rank 0's tile of 2 / 4 / 8 of the default
// workload, kernel ms pooled vs one path per lane (profiles/r02_h_tile_ab_*): 256 spp 217.9 / 303.8, 127.5 / 156.6, 94.6 / 87.0 at
// 4.6 / 2.3 / 1.1 pixels per pool slot; 1 spp 1.30 / 1.43, 0.88 / 0.81, 0.68 / 0.55.  Hence: below 1.5 pixels per slot (2.5 in the
// 1-spp mode) the automatic choice is kernel 1.
#ifndef RT_SMALL_TILE_PIXELS_PER_SLOT
#define RT_SMALL_TILE_PIXELS_PER_SLOT 1.5
#endif
#ifndef RT_SMALL_TILE_PIXELS_PER_SLOT_1SPP
#define RT_SMALL_TILE_PIXELS_PER_SLOT_1SPP 2.5
#endif
static int effectiveKernel(const RtContext* c)
{
    if (c->optKernel >= 0) return c->optKernel;
    if (c->P.modelCount <= 0) return 1;
    const double slots = (double)c->numSMs * POOL_WARPS * 64.0;
    const double threshold = c->P.NumRaysPerPixel <= 1 ? RT_SMALL_TILE_PIXELS_PER_SLOT_1SPP : RT_SMALL_TILE_PIXELS_PER_SLOT;
    if (c->dispatchPixels > 0 && (double)c->dispatchPixels < threshold * slots) return 1;
    return 2;
}

// "l2Persist": ask the L2 to keep the node-pair records (persisting access-policy window on the dispatch stream; every ray walks
// them, the triangle stream is marked streaming by omission).  The scenes already hit L2 at 76-86 % (profiles/), so this is a
// candidate for the misses' latency, not for bandwidth; off by default (measured in round 2: +-1 %).
static void applyL2Persistence(RtContext* c)
{
#ifndef RT_SIMT_EMU
    const void* base = c->repack.pairs.p;
    const size_t bytes = c->repack.totalPairs * sizeof(NodePair);
    const int want = c->optL2Persist && base && bytes > 0;
    if (want == c->l2PersistApplied && (!want || (base == c->l2PersistBase && bytes == c->l2PersistBytes))) return;
    cudaStreamAttrValue attr; memset(&attr, 0, sizeof(attr));
    if (want)
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, c->device) != cudaSuccess || prop.persistingL2CacheMaxSize <= 0) return;
        size_t setAside = bytes < (size_t)prop.persistingL2CacheMaxSize ? bytes : (size_t)prop.persistingL2CacheMaxSize;
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setAside) != cudaSuccess) { cudaGetLastError(); return; }
        size_t window = bytes < (size_t)prop.accessPolicyMaxWindowSize ? bytes : (size_t)prop.accessPolicyMaxWindowSize;
        attr.accessPolicyWindow.base_ptr = const_cast<void*>(base);
        attr.accessPolicyWindow.num_bytes = window;
        attr.accessPolicyWindow.hitRatio = window > 0 ? (float)((double)setAside / (double)window > 1.0 ? 1.0 : (double)setAside / (double)window) : 0.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }
    if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) { cudaGetLastError(); return; }
    if (!want) cudaCtxResetPersistingL2Cache();
    c->l2PersistApplied = want; c->l2PersistBase = base; c->l2PersistBytes = bytes;
#else
    (void)c;
#endif
}

static int prepareScene(RtContext* c)
{
    if (c->P.modelCount < 0 || (size_t)c->P.modelCount > c->models.count) return fail(c, RT_E_STATE, "rtDispatch: modelCount exceeds the ModelInfo buffer");
    if (c->P.modelCount > 0 && (c->nodes.count == 0 || c->tris.count == 0)) return fail(c, RT_E_STATE, "rtDispatch: models without Nodes / Triangles buffers");
    for (int i = 0; i < c->P.modelCount; i++)
    {
        const RtModel& m = c->hModels[i];
        if (m.nodeOffset < 0 || (size_t)m.nodeOffset >= c->nodes.count || m.triOffset < 0 || (size_t)m.triOffset > c->tris.count)
            return fail(c, RT_E_STATE, "rtDispatch: model nodeOffset / triOffset out of range");
    }
    // shared-memory budget for the tree tops: what the selected kernel can afford next to its own shared state
    // automatic = 0: measured (profiles/r01_sweeps.log), staging tree tops never beat leaving that shared memory to L1
    int budget = c->optSmemPairs < 0 ? 0 : c->optSmemPairs;
    const int kernelSel = effectiveKernel(c);
    const bool treelets = c->optTreeletPrefetch && kernelSel != 0;      // two-level treelets with flagged roots; no shared-memory staging with it
    const int pairOrder = treelets ? 2 : c->optPairOrder;
    if (treelets) budget = 0;
    if (budget > 3072) budget = 3072;                                   // planScene's own clamp (192 KB), so that budgetUsed compares equal
    if (kernelSel == 2) { const int mx = pool_max_smem_pairs(c->optPoolSlots ? c->optPoolSlots : 96, (int)c->spheres.count); if (budget > mx) budget = mx; }
    else if (kernelSel == 0) budget = 0;
    if (budget != c->repack.budgetUsed || pairOrder != c->repack.orderUsed || treelets != c->repack.flagsUsed) c->sceneDirty = true;
    if (c->sceneDirty)
    {
        std::string msg;
        cudaError_t e = c->repack.buildScene(c->hNodes, c->hModels, c->P.modelCount, c->tris.p, c->tris.count, budget, c->stream, msg, pairOrder, treelets);
        if (e != cudaSuccess) return failCuda(c, e, "repack scene");
        if (!msg.empty()) return fail(c, RT_E_STATE, "rtDispatch: " + msg);
        c->sceneDirty = false; c->modelsDirty = true;
    }
    {
        // box containing every possible ray origin: the camera (with its defocus disc) and all geometry; it sizes the padding of
        // the models' world boxes and of the sphere accelerator, which are rebuilt when the box outgrows the one they were built for
        float lo[3], hi[3];
        const float jitter = fabsf(c->P.DefocusStrength) / (float)(c->P.W ? c->P.W : 1u) *
                             (fabsf(c->P.cam[0]) + fabsf(c->P.cam[1]) + fabsf(c->P.cam[2]) + fabsf(c->P.cam[4]) + fabsf(c->P.cam[5]) + fabsf(c->P.cam[6]));
        for (int a = 0; a < 3; a++) { lo[a] = c->P.cam[12 + a] - jitter; hi[a] = c->P.cam[12 + a] + jitter; }
        for (const RtSphere& s : c->hSpheres) for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], s.centre[a] - fabsf(s.radius)); hi[a] = fmaxf(hi[a], s.centre[a] + fabsf(s.radius)); }
        for (int i = 0; i < c->P.modelCount; i++)
        {
            const RtModel& m = c->hModels[i];
            double mlo[3], mhi[3];
            if (!RepackState::worldBoxOfModel(m, c->repack.roots[std::make_pair(m.nodeOffset, m.triOffset)], mlo, mhi)) { for (int a = 0; a < 3; a++) { lo[a] = -INFINITY; hi[a] = INFINITY; } break; }
            for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], std::nextafterf((float)mlo[a], -INFINITY)); hi[a] = fmaxf(hi[a], std::nextafterf((float)mhi[a], INFINITY)); }
        }
        if (!c->modelsDirty && !c->repack.modelBoundCovers(lo, hi)) c->modelsDirty = true;
        if (!c->spheresDirty && c->repack.sphBvh && !c->repack.sphereBoundCovers(lo, hi)) c->spheresDirty = true;
        if (c->modelsDirty || c->spheresDirty)      // some room, so that a moving camera does not rebuild every frame
            for (int a = 0; a < 3; a++) { const float ext = 0.25f * (hi[a] - lo[a]) + 0.01f; lo[a] -= ext; hi[a] += ext; }
        if (c->modelsDirty)
        {
            cudaError_t e = c->repack.buildModels(c->hModels, c->hNodes, c->P.modelCount, c->stream, c->optTlas, lo, hi);
            if (e != cudaSuccess) return failCuda(c, e, "repack models");
            c->modelsDirty = false;
        }
        if (c->spheresDirty)
        {
            cudaError_t e = c->repack.buildSpheres(c->hSpheres, lo, hi, c->stream);
            if (e != cudaSuccess) return failCuda(c, e, "repack spheres");
            c->spheresDirty = false;
        }
    }
    return RT_OK;
}

// One GPU's share of a dispatch: the RayTrace kernel over this context's tile (the whole image without rtSetTile).
static int dispatchLocal(RtContext* c, int kernelIndex, int gx, int gy, int gz)
{
    if (!c || gx < 0 || gy < 0 || gz < 0) return fail(c, RT_E_INVALID, "rtDispatch: bad argument");
    if (c->width == 0) return fail(c, RT_E_STATE, "rtDispatch: rtResize has not been called");
    if ((int)c->P.W != c->width || (int)c->P.H != c->height) return fail(c, RT_E_STATE, "rtDispatch: Resolution does not match the render textures");
    CK(cudaSetDevice(c->device));
    const unsigned W = c->P.W, H = c->P.H;
    unsigned limX = (unsigned long long)gx * 8ull < W ? (unsigned)gx * 8u : W;
    unsigned limY = (unsigned long long)gy * 8ull < H ? (unsigned)gy * 8u : H;
    if (gz == 0) limX = limY = 0;

    if (kernelIndex == RT_KERNEL_RESET_ACCUMULATED)            // RC:26-32
    {
        if (limX == 0 || limY == 0) return RT_OK;
        CK(cudaMemset2DAsync(c->accum.p, (size_t)W * 16, 0, (size_t)limX * 16, limY, c->stream));
        return RT_OK;
    }
    if (kernelIndex != RT_KERNEL_RAYTRACE) return fail(c, RT_E_INVALID, "rtDispatch: kernelIndex must be 0 (RayTrace) or 1 (ResetAccumulated)");
    if (limX == 0 || limY == 0) return RT_OK;
    if (c->P.NumRaysPerPixel < 0) return fail(c, RT_E_STATE, "rtDispatch: NumRaysPerPixel is negative");
    if (c->P.MaxBounceCount > 200 || W > 65535u || H > 65535u) return fail(c, RT_E_STATE, "rtDispatch: MaxBounceCount > 200 or a resolution above 65535 is not supported");
    if (c->P.MaxBounceCount < 0) return fail(c, RT_E_STATE, "rtDispatch: MaxBounceCount is negative (the reference clamps it to [0, 32], RCM:15)");

    {
        unsigned long long rows = 0;
        for (unsigned int y0 = 0, b = 0; y0 < limY; y0 += (unsigned int)c->bandRows, b++)
            if ((int)(b % (unsigned int)c->tileWorld) == c->tileRank) rows += (limY - y0) < (unsigned int)c->bandRows ? (limY - y0) : (unsigned int)c->bandRows;
        c->dispatchPixels = rows * limX;
    }
    int rc = prepareScene(c);
    if (rc != RT_OK) return rc;
    applyL2Persistence(c);

    DevParams P = c->P;
    P.limX = limX; P.limY = limY;
    P.tileRank = c->tileRank; P.tileWorld = c->tileWorld; P.bandRows = c->bandRows; P.countStats = c->optCountStats;
    P.sphereCount = (int)c->spheres.count;
    P.Nodes = c->nodes.p; P.Triangles = c->tris.p; P.ModelInfo = c->models.p; P.Spheres = c->spheres.p;
    P.pairs = c->repack.pairs.p; P.triGeom = c->repack.triGeom.p; P.triNormals = c->repack.triNormals.p;
    P.models = c->repack.models.p; P.spheres = c->repack.spheres.p;
    P.sphPairs = c->repack.sphPairs.p; P.sphLeaves = c->repack.sphLeaves.p; P.sphBvh = c->repack.sphBvh;
    P.sphRootStart = c->repack.sphRootStart; P.sphRootCount = c->repack.sphRootCount; P.smemPairs = c->repack.smemPairs; P.tailLanes = c->optTailLanes; P.sortRays = c->optSortRays; P.gridFit = c->optGridFit;
    P.FrameRender = c->frame.p; P.AccumulatedRender = c->accum.p;
    P.counters = c->dCounters; P.workCounter = c->dWork;
    P.nPeers = c->nPeers; P.forceExt = c->optForceExt; P.modelSkip = c->optModelSkip;
    P.tlasPairs = c->repack.tlasPairs.p; P.tlasLeaves = c->repack.tlasLeaves.p; P.tlas = c->repack.tlas;
    P.tlasRootStart = c->repack.tlasRootStart; P.tlasRootCount = c->repack.tlasRootCount;
    {
        // camOrigin = M * (0,0,0,1) as SetupPixel computes it (left to right, unfused): exact zero components and non-finite axes rule the shortcut out
        bool ok = c->P.DefocusStrength == 0.0f && c->optZeroDefocus != 0;
        for (int a = 0; a < 3 && ok; a++)
        {
            const float o = ((P.cam[a] * 0.0f + P.cam[4 + a] * 0.0f) + P.cam[8 + a] * 0.0f) + P.cam[12 + a] * 1.0f;
            ok = o != 0.0f && std::isfinite(o) && std::isfinite(P.cam[a]) && std::isfinite(P.cam[4 + a]);
        }
        P.zeroDefocus = ok ? 1 : 0;
    }
    for (int k = 0; k < c->nPeers; k++) { P.peerFrame[k] = c->peerFrame[k]; P.peerAccum[k] = c->peerAccum[k]; }

    EventPair ev;
    if (!c->freeEvents.empty()) { ev = c->freeEvents.back(); c->freeEvents.pop_back(); }
    else { CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b)); }
    if (c->pending.size() >= 512) { rc = drainEvents(c); if (rc != RT_OK) return rc; }

    int kernel = effectiveKernel(c);
    if (c->P.NumRaysPerPixel == 0) kernel = 0;       // 0 samples: the per-pixel kernel reproduces the reference's 0/0 directly
    // sample chunks (small tiles): a pixel's chain may change lanes / slots between samples; the hand-off buffers belong to the context
    // (kernel 1 only: in the pooled kernel the same hand-off measured 14 % slower than whole pixels on rank 0's tile of 4 and 5 % slower on
    // the tile of 8 — profiles/r02_k_tile_ab_pool_sample_chunks.jsonl — and was taken out again)
    P.chunks = kernel == 1 ? wave_chunks(c->optSampleChunks, c->numSMs, c->dispatchPixels, P.NumRaysPerPixel, P.modelCount) : 1;
    if (c->dispatchPixels == 0) P.chunks = 1;        // a rank whose tile is empty (an image lower than the band pattern)
    if (P.chunks > 1)
    {
        unsigned long long rows = limX ? c->dispatchPixels / limX : 0;
        const size_t jobs = (size_t)((limX + 7u) / 8u) * (size_t)((rows + 3ull) / 4ull) * 32u;
        // four self-validating 64-bit words per pixel job; the tag carries a serial number of the dispatch, so nothing has to be cleared
        // between frames (only a fresh allocation and the wrap of the 24-bit serial zero the buffer: tag 0 is never expected)
        const size_t before = c->handoff.count;
        CK(c->handoff.ensure(4 * jobs));
        c->chunkSerial = (c->chunkSerial + 1u) & 0xffffffu;
        if (c->handoff.count != before || c->chunkSerial == 0u)
        {
            CK(cudaMemsetAsync(c->handoff.p, 0, c->handoff.count * sizeof(unsigned long long), c->stream));
            if (c->chunkSerial == 0u) c->chunkSerial = 1u;
        }
        P.handoff = c->handoff.p; P.handoffFlags = nullptr; P.chunkSerial = (int)c->chunkSerial;
    }
    if (kernel == 0)
    {
        CK(cudaEventRecord(ev.a, c->stream));
        dim3 grid((limX + 7) / 8, (limY + 7) / 8, 1), block(8, 8, 1);
        RT_LAUNCH(grid, block, 0, c->stream, k_raytrace_mega, P);
        CK(cudaEventRecord(ev.b, c->stream));
    }
    else if (kernel == 1)
    {
        cudaError_t e = wave_launch(P, c->numSMs, c->stream, ev.a, ev.b);
        if (e != cudaSuccess) return failCuda(c, e, "wavefront launch");
    }
    else
    {
        if (POOL_COLD_WORDS > 0)
        {
            CK(c->poolCold.ensure((size_t)c->numSMs * POOL_WARPS * POOL_COLD_WORDS * 96));      // sized for the largest pools
            P.poolCold = c->poolCold.p;
        }
        cudaError_t e = pool_launch(P, c->optPoolSlots, c->numSMs, c->stream, ev.a, ev.b);
        if (e != cudaSuccess) return failCuda(c, e, "pool wavefront launch");
    }
    CK(cudaGetLastError());
    c->pending.push_back(ev);
    c->stats.dispatches++;
    return RT_OK;
}

int rtReadback(RtContext* c, const char* tex, float* dst, size_t bytes)
{
    if (!c || !tex || !dst) return fail(c, RT_E_INVALID, "rtReadback: bad argument");
    const std::string n(tex);
    const float4* src = n == "FrameRender" ? c->frame.p : n == "AccumulatedRender" ? c->accum.p : nullptr;
    if (n != "FrameRender" && n != "AccumulatedRender") return fail(c, RT_E_UNKNOWN_NAME, std::string("rtReadback: unknown texture ") + tex);
    if (!src) return fail(c, RT_E_STATE, "rtReadback: rtResize has not been called");
    if (bytes != (size_t)c->width * c->height * 16) return fail(c, RT_E_INVALID, "rtReadback: bytes must equal W*H*16");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return RT_OK;
}

int rtDisplay(RtContext* c, int useAccumulated, int Frame, uint8_t* dst, size_t bytes)
{
    if (!c || !dst) return fail(c, RT_E_INVALID, "rtDisplay: bad argument");
    if (!c->frame.p) return fail(c, RT_E_STATE, "rtDisplay: rtResize has not been called");
    const size_t n = (size_t)c->width * c->height;
    if (bytes != n * 4) return fail(c, RT_E_INVALID, "rtDisplay: bytes must equal W*H*4");
    CK(cudaSetDevice(c->device));
    if (c->copyPending) { CK(cudaEventSynchronize(c->copyDone)); c->copyPending = false; }
    CK(c->display.ensure(n));
    RT_LAUNCH((unsigned)((n + 255) / 256), 256, 0, c->stream, k_display, useAccumulated ? c->accum.p : c->frame.p, c->display.p, n, (float)Frame);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dst, c->display.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return RT_OK;
}

// ---- pipelined readback -------------------------------------------------------------------------------------------------------
static int copyPipeline(RtContext* c)
{
    if (!c->copyStream) CK(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
    if (!c->snapReady) CK(cudaEventCreateWithFlags(&c->snapReady, cudaEventDisableTiming));
    if (!c->copyDone) CK(cudaEventCreateWithFlags(&c->copyDone, cudaEventDisableTiming));
    if (c->copyPending) CK(cudaStreamWaitEvent(c->stream, c->copyDone, 0));      // the snapshot buffer is free once the copy before has read it
    return RT_OK;
}
static int copyOut(RtContext* c, void* dst, const void* src, size_t bytes)
{
    CK(cudaEventRecord(c->snapReady, c->stream));
    CK(cudaStreamWaitEvent(c->copyStream, c->snapReady, 0));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->copyStream));
    CK(cudaEventRecord(c->copyDone, c->copyStream));
    c->copyPending = true;
    return RT_OK;
}

int rtReadbackAsync(RtContext* c, const char* tex, float* dst, size_t bytes)
{
    if (!c || !tex || !dst) return fail(c, RT_E_INVALID, "rtReadbackAsync: bad argument");
    const std::string n(tex);
    if (n != "FrameRender" && n != "AccumulatedRender") return fail(c, RT_E_UNKNOWN_NAME, std::string("rtReadbackAsync: unknown texture ") + tex);
    const float4* src = n == "FrameRender" ? c->frame.p : c->accum.p;
    if (!src) return fail(c, RT_E_STATE, "rtReadbackAsync: rtResize has not been called");
    if (bytes != (size_t)c->width * c->height * 16) return fail(c, RT_E_INVALID, "rtReadbackAsync: bytes must equal W*H*16");
    CK(cudaSetDevice(c->device));
    int rc = copyPipeline(c); if (rc != RT_OK) return rc;
    CK(c->snap.ensure((size_t)c->width * c->height));
    CK(cudaMemcpyAsync(c->snap.p, src, bytes, cudaMemcpyDeviceToDevice, c->stream));       // the texture as it is after the work queued so far
    return copyOut(c, dst, c->snap.p, bytes);
}

int rtDisplayAsync(RtContext* c, int useAccumulated, int Frame, uint8_t* dst, size_t bytes)
{
    if (!c || !dst) return fail(c, RT_E_INVALID, "rtDisplayAsync: bad argument");
    if (!c->frame.p) return fail(c, RT_E_STATE, "rtDisplayAsync: rtResize has not been called");
    const size_t n = (size_t)c->width * c->height;
    if (bytes != n * 4) return fail(c, RT_E_INVALID, "rtDisplayAsync: bytes must equal W*H*4");
    CK(cudaSetDevice(c->device));
    int rc = copyPipeline(c); if (rc != RT_OK) return rc;
    CK(c->display.ensure(n));
    RT_LAUNCH((unsigned)((n + 255) / 256), 256, 0, c->stream, k_display, useAccumulated ? c->accum.p : c->frame.p, c->display.p, n, (float)Frame);
    CK(cudaGetLastError());
    return copyOut(c, dst, c->display.p, bytes);
}

int rtReadbackWait(RtContext* c)
{
    if (!c) return RT_E_INVALID;
    if (!c->copyPending) return RT_OK;
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->copyDone));
    c->copyPending = false;
    return RT_OK;
}

int rtSynchronize(RtContext* c)
{
    if (!c) return RT_E_INVALID;
    if (c->copyPending) { CK(cudaSetDevice(c->device)); CK(cudaEventSynchronize(c->copyDone)); c->copyPending = false; }
    for (RtContext* m : c->followers) { CK(cudaSetDevice(m->device)); CK(cudaStreamSynchronize(m->stream)); }
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return RT_OK;
}

// ---- multi-GPU tile staging -----------------------------------------------------------------------------------------------

static size_t tileRows(const RtContext* c, int rank)
{
    size_t rows = 0;
    for (int y0 = 0, b = 0; y0 < c->height; y0 += c->bandRows, b++)
        if (b % c->tileWorld == rank) rows += (size_t)((c->height - y0) < c->bandRows ? (c->height - y0) : c->bandRows);
    return rows;
}

static int ensureTileBuffers(RtContext* c)
{
    if (c->width == 0) return fail(c, RT_E_STATE, "tile staging: rtResize has not been called");
    // every rank must contribute the same count to the all-gather: size by the largest share (rank 0's)
    const size_t per = tileRows(c, 0) * (size_t)c->width * 2;    // frame + accumulated
    CK(c->tileSend.ensure(per));
    CK(c->tileRecv.ensure(per * c->tileWorld));
    return RT_OK;
}

int rtPackTile(RtContext* c)
{
    if (!c) return RT_E_INVALID;
    CK(cudaSetDevice(c->device));
    int rc = ensureTileBuffers(c); if (rc != RT_OK) return rc;
    launch_pack_tile(c->frame.p, c->accum.p, c->tileSend.p, c->width, c->height, c->tileRank, c->tileWorld, c->bandRows,
                     (int)tileRows(c, 0), c->stream);
    CK(cudaGetLastError());
    return RT_OK;
}

int rtUnpackTiles(RtContext* c)
{
    if (!c) return RT_E_INVALID;
    CK(cudaSetDevice(c->device));
    int rc = ensureTileBuffers(c); if (rc != RT_OK) return rc;
    launch_unpack_tiles(c->tileRecv.p, c->frame.p, c->accum.p, c->width, c->height, c->tileWorld, c->bandRows,
                        (int)tileRows(c, 0), c->stream);
    CK(cudaGetLastError());
    return RT_OK;
}


// ---- exchange of finished tiles inside the ABI (rtCommInit / rtCreateMulti) -----------------------------------------------------

// true when a RayTrace dispatch on this context is followed by the all-gather of the frame's tiles
static bool exchanges(const RtContext* c) { return c->optExchange && c->tileWorld > 1 && (c->comm || !c->followers.empty() || c->leader); }

static int beginExchangeTiming(RtContext* c, EventPair& ev)
{
    if (!c->freeEvents.empty()) { ev = c->freeEvents.back(); c->freeEvents.pop_back(); }
    else { CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b)); }
    CK(cudaEventRecord(ev.a, c->stream));
    return RT_OK;
}

static int ncclFail(RtContext* c, NcclApi* api, int r, const char* what)
{
    return fail(c, RT_E_CUDA, std::string(what) + ": " + (api && api->GetErrorString ? api->GetErrorString(r) : "NCCL error") + " (" + std::to_string(r) + ")");
}

// pack -> ncclAllGather -> unpack on the streams of `n` contexts (one per GPU; n > 1 only for a single-process group, whose
// collectives are issued as one NCCL group).  Under the SIMT interpreter the group's all-gather is the copy it stands for.
static int exchangeTiles(RtContext* const* cs, int n)
{
    RtContext* lead = cs[0];
    std::vector<EventPair> evs((size_t)n);
    for (int i = 0; i < n; i++)
    {
        RtContext* c = cs[i];
        CK(cudaSetDevice(c->device));
        int rc = beginExchangeTiming(c, evs[i]); if (rc != RT_OK) return rc;
        rc = rtPackTile(c); if (rc != RT_OK) return i ? fail(lead, rc, c->err) : rc;
    }
#ifdef RT_SIMT_EMU
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++)
        memcpy(cs[i]->tileRecv.p + (size_t)j * cs[j]->tileSend.count, cs[j]->tileSend.p, cs[j]->tileSend.count * sizeof(float4));
#else
    std::string why;
    NcclApi* api = NcclApi::get(why);
    if (!api) return fail(lead, RT_E_STATE, "tile exchange: " + why);
    int r;
    if (n > 1 && (r = api->GroupStart()) != NcclApi::Success) return ncclFail(lead, api, r, "ncclGroupStart");
    for (int i = 0; i < n; i++)
    {
        RtContext* c = cs[i];
        if (!c->comm) return fail(lead, RT_E_STATE, "tile exchange: no communicator (rtCommInit)");
        CK(cudaSetDevice(c->device));
        if ((r = api->AllGather(c->tileSend.p, c->tileRecv.p, c->tileSend.count * sizeof(float4), NcclApi::Char, c->comm, c->stream)) != NcclApi::Success)
            return ncclFail(lead, api, r, "ncclAllGather");
    }
    if (n > 1 && (r = api->GroupEnd()) != NcclApi::Success) return ncclFail(lead, api, r, "ncclGroupEnd");
#endif
    for (int i = 0; i < n; i++)
    {
        RtContext* c = cs[i];
        CK(cudaSetDevice(c->device));
        int rc = rtUnpackTiles(c); if (rc != RT_OK) return i ? fail(lead, rc, c->err) : rc;
        CK(cudaEventRecord(evs[i].b, c->stream));
        c->pendingX.push_back(evs[i]);
    }
    { RtContext* c = lead; CK(cudaSetDevice(c->device)); }
    return RT_OK;
}

int rtDispatch(RtContext* c, int kernelIndex, int gx, int gy, int gz)
{
    if (!c) return RT_E_INVALID;
    if (c->leader) return fail(c, RT_E_STATE, "rtDispatch: this context belongs to a group; dispatch on the context rtCreateMulti returned");
    int rc = dispatchLocal(c, kernelIndex, gx, gy, gz);
    if (rc != RT_OK) return rc;
    for (RtContext* m : c->followers)
        if ((rc = dispatchLocal(m, kernelIndex, gx, gy, gz)) != RT_OK) return fail(c, rc, "GPU " + std::to_string(m->device) + ": " + m->err);
    if (!c->followers.empty()) CK(cudaSetDevice(c->device));
    const bool traced = kernelIndex == RT_KERNEL_RAYTRACE && gz != 0 && gx > 0 && gy > 0;
    if (!traced || !exchanges(c)) return RT_OK;
    std::vector<RtContext*> all(1, c);
    all.insert(all.end(), c->followers.begin(), c->followers.end());
    return exchangeTiles(all.data(), (int)all.size());
}

int rtExchangeTiles(RtContext* c)
{
    if (!c) return RT_E_INVALID;
    if (c->leader) return fail(c, RT_E_STATE, "rtExchangeTiles: call it on the context rtCreateMulti returned");
    if (c->tileWorld <= 1) return RT_OK;
    if (!c->comm && c->followers.empty()) return fail(c, RT_E_STATE, "rtExchangeTiles: no communicator (rtCommInit / rtCreateMulti)");
    std::vector<RtContext*> all(1, c);
    all.insert(all.end(), c->followers.begin(), c->followers.end());
    return exchangeTiles(all.data(), (int)all.size());
}

int rtGetUniqueId(void* id, size_t bytes)
{
    if (!id || bytes != RT_UNIQUE_ID_BYTES) return fail(nullptr, RT_E_INVALID, "rtGetUniqueId: id must hold 128 bytes");
    std::string why;
    NcclApi* api = NcclApi::get(why);
    if (!api) return fail(nullptr, RT_E_STATE, "rtGetUniqueId: " + why);
    NcclApi::UniqueId u;
    const int r = api->GetUniqueId(&u);
    if (r != NcclApi::Success) return ncclFail(nullptr, api, r, "ncclGetUniqueId");
    memcpy(id, &u, RT_UNIQUE_ID_BYTES);
    return RT_OK;
}

static void destroyComm(RtContext* c)
{
    if (!c->comm) return;
    std::string why;
    NcclApi* api = NcclApi::get(why);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (api) api->CommDestroy(c->comm);
    c->comm = nullptr; c->commRank = 0; c->commWorld = 1;
}

int rtCommInit(RtContext* c, const void* id, size_t bytes, int rank, int world)
{
    if (!c || !id || bytes != RT_UNIQUE_ID_BYTES || world < 1 || rank < 0 || rank >= world) return fail(c, RT_E_INVALID, "rtCommInit: bad argument");
    if (c->leader || !c->followers.empty()) return fail(c, RT_E_STATE, "rtCommInit: the context belongs to a single-process group (rtCreateMulti), which has its communicator");
    std::string why;
    NcclApi* api = NcclApi::get(why);
    if (!api) return fail(c, RT_E_STATE, "rtCommInit: " + why);
    CK(cudaSetDevice(c->device));
    destroyComm(c);
    NcclApi::UniqueId u; memcpy(&u, id, RT_UNIQUE_ID_BYTES);
    const int r = api->CommInitRank(&c->comm, world, u, rank);
    if (r != NcclApi::Success) { c->comm = nullptr; return ncclFail(c, api, r, "ncclCommInitRank"); }
    c->commRank = rank; c->commWorld = world;
    return rtSetTile(c, rank, world, c->tileWorld == world && c->tileRank == rank ? c->bandRows : 8);
}

int rtCommDestroy(RtContext* c)
{
    if (!c) return RT_E_INVALID;
    if (c->leader || !c->followers.empty()) return fail(c, RT_E_STATE, "rtCommDestroy: a group's communicator lives as long as the group (rtDestroy)");
    destroyComm(c);
    return rtSetTile(c, 0, 1, c->bandRows);
}

int rtCreateMulti(RtContext** out, const int* devices, int nDevices)
{
    if (!out) return fail(nullptr, RT_E_INVALID, "rtCreateMulti: out is NULL");
    *out = nullptr;
    if (nDevices < 1 || nDevices > 64) return fail(nullptr, RT_E_INVALID, "rtCreateMulti: between 1 and 64 devices");
    for (int i = 0; i < nDevices; i++) for (int j = 0; j < i; j++)
        if (devices && devices[i] == devices[j]) return fail(nullptr, RT_E_INVALID, "rtCreateMulti: a device is listed twice (NCCL needs one rank per GPU)");
    std::vector<RtContext*> cs;
    auto undo = [&](int code) { for (RtContext* m : cs) { destroyComm(m); m->leader = nullptr; m->followers.clear(); rtDestroy(m); } return code; };
    for (int i = 0; i < nDevices; i++)
    {
        RtContext* m = nullptr;
        const int rc = rtCreate(&m, devices ? devices[i] : i);
        if (rc != RT_OK) return undo(rc);
        cs.push_back(m);
    }
    if (nDevices > 1)
    {
#ifndef RT_SIMT_EMU
        std::string why;
        NcclApi* api = NcclApi::get(why);
        if (!api) { fail(nullptr, RT_E_STATE, "rtCreateMulti: " + why); return undo(RT_E_STATE); }
        NcclApi::UniqueId u;
        int r = api->GetUniqueId(&u);
        if (r == NcclApi::Success) r = api->GroupStart();
        for (int i = 0; i < nDevices && r == NcclApi::Success; i++)
        {
            cudaSetDevice(cs[i]->device);
            r = api->CommInitRank(&cs[i]->comm, nDevices, u, i);
        }
        if (r == NcclApi::Success) r = api->GroupEnd();
        if (r != NcclApi::Success) { ncclFail(nullptr, api, r, "rtCreateMulti: NCCL communicator"); return undo(RT_E_CUDA); }
#endif
        for (int i = 0; i < nDevices; i++)
        {
            cs[i]->commRank = i; cs[i]->commWorld = nDevices;
            cs[i]->tileRank = i; cs[i]->tileWorld = nDevices; cs[i]->bandRows = 8;
            if (i) { cs[i]->leader = cs[0]; cs[0]->followers.push_back(cs[i]); }
        }
        cudaSetDevice(cs[0]->device);
    }
    *out = cs[0];
    return RT_OK;
}

int rtGetDevicePointer(RtContext* c, const char* name, void** devPtr, size_t* bytes)
{
    if (!c || !name || !devPtr || !bytes) return fail(c, RT_E_INVALID, "rtGetDevicePointer: bad argument");
    const std::string n(name);
    CK(cudaSetDevice(c->device));
    if (n == "FrameRender") { *devPtr = c->frame.p; *bytes = c->frame.count * 16; }
    else if (n == "AccumulatedRender") { *devPtr = c->accum.p; *bytes = c->accum.count * 16; }
    else if (n == "TileSend" || n == "TileRecv")
    {
        int rc = ensureTileBuffers(c); if (rc != RT_OK) return rc;
        if (n == "TileSend") { *devPtr = c->tileSend.p; *bytes = c->tileSend.count * 16; }
        else { *devPtr = c->tileRecv.p; *bytes = c->tileRecv.count * 16; }
    }
    else return fail(c, RT_E_UNKNOWN_NAME, std::string("rtGetDevicePointer: unknown object ") + name);
    if (!*devPtr) return fail(c, RT_E_STATE, "rtGetDevicePointer: object not allocated yet (call rtResize)");
    return RT_OK;
}

// One exported image: the IPC handle of the allocation that contains it + its byte offset inside that allocation
// (cudaMalloc may sub-allocate small buffers from a larger block; the handle always names the whole block).
struct RtIpcImage { cudaIpcMemHandle_t handle; unsigned long long offset; };
static_assert(sizeof(RtIpcImage) == 72, "IPC blob layout");

static cudaError_t allocationBase(const void* p, unsigned long long* offset)
{
    typedef int (*GetRangeFn)(unsigned long long*, size_t*, unsigned long long);
    static GetRangeFn fn = nullptr;
    if (!fn)
    {
        void* sym = nullptr; cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &qr);
        if (e != cudaSuccess) return e;
        if (!sym) return cudaErrorNotSupported;
        fn = (GetRangeFn)sym;
    }
    unsigned long long base = 0; size_t size = 0;
    if (fn(&base, &size, (unsigned long long)(uintptr_t)p) != 0) return cudaErrorInvalidValue;
    *offset = (unsigned long long)(uintptr_t)p - base;
    return cudaSuccess;
}

int rtGetIpcHandles(RtContext* c, void* handles, size_t bytes)
{
    if (!c || !handles || bytes != 2 * sizeof(RtIpcImage)) return fail(c, RT_E_INVALID, "rtGetIpcHandles: handles must hold 2 x 72 bytes");
    if (!c->frame.p || !c->accum.p) return fail(c, RT_E_STATE, "rtGetIpcHandles: rtResize has not been called");
    CK(cudaSetDevice(c->device));
    RtIpcImage h[2];
    memset(h, 0, sizeof(h));
    CK(cudaIpcGetMemHandle(&h[0].handle, c->frame.p));
    CK(allocationBase(c->frame.p, &h[0].offset));
    CK(cudaIpcGetMemHandle(&h[1].handle, c->accum.p));
    CK(allocationBase(c->accum.p, &h[1].offset));
    memcpy(handles, h, sizeof(h));
    return RT_OK;
}

int rtSetPeers(RtContext* c, int nPeers, const void* handles, size_t bytes)
{
    if (!c || nPeers < 0 || nPeers > RT_MAX_PEERS) return fail(c, RT_E_INVALID, "rtSetPeers: at most 7 peers");
    if (nPeers > 0 && (!handles || bytes != (size_t)nPeers * 2 * sizeof(RtIpcImage))) return fail(c, RT_E_INVALID, "rtSetPeers: handles must hold nPeers x 2 x 72 bytes");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    closePeers(c);
    const RtIpcImage* h = (const RtIpcImage*)handles;
    for (int k = 0; k < nPeers; k++)
    {
        void* base[2] = {nullptr, nullptr};
        for (int j = 0; j < 2; j++)
        {
            // both images of a peer may live in the same allocation: open every distinct handle once
            if (j == 1 && memcmp(&h[2 * k].handle, &h[2 * k + 1].handle, sizeof(cudaIpcMemHandle_t)) == 0) { base[1] = base[0]; continue; }
            CK(cudaIpcOpenMemHandle(&base[j], h[2 * k + j].handle, cudaIpcMemLazyEnablePeerAccess));
            c->peerBase.push_back(base[j]);
        }
        c->peerFrame[k] = (float4*)((char*)base[0] + h[2 * k].offset);
        c->peerAccum[k] = (float4*)((char*)base[1] + h[2 * k + 1].offset);
    }
    c->nPeers = nPeers;
    return RT_OK;
}

// Host <-> device copies of large caller-owned (pageable) arrays through two pinned chunks: the CPU copies chunk k + 1 into / out of
// pinned memory while the DMA engine moves chunk k, instead of the driver's own slower staging of pageable memory.
// memcpy of one staging chunk on a few host threads: one core moves 5-10 GB/s (less into pages that are touched for the first time),
// the DMA engine 25-50 GB/s
static void hostCopy(void* dst, const void* src, size_t n)
{
#ifndef RT_SIMT_EMU
    const unsigned hw = std::thread::hardware_concurrency();
    const int parts = n < (1u << 20) || hw < 4 ? 1 : 4;
    if (parts > 1)
    {
        std::thread th[3];
        const size_t each = ((n / parts) + 4095) & ~(size_t)4095;
        for (int k = 1; k < parts; k++)
        {
            const size_t off = each * k, len = off >= n ? 0 : (k == parts - 1 ? n - off : (each < n - off ? each : n - off));
            th[k - 1] = std::thread([=]() { if (len) memcpy((unsigned char*)dst + off, (const unsigned char*)src + off, len); });
        }
        memcpy(dst, src, each < n ? each : n);
        for (int k = 1; k < parts; k++) th[k - 1].join();
        return;
    }
#endif
    memcpy(dst, src, n);
}

static int stagingReady(RtContext* c)
{
    for (int k = 0; k < 2; k++)
    {
        if (!c->stageBuf[k]) CK(cudaMallocHost((void**)&c->stageBuf[k], RtContext::STAGE_BYTES));
        if (!c->stageEv[k]) CK(cudaEventCreateWithFlags(&c->stageEv[k], cudaEventDisableTiming));
    }
    return RT_OK;
}
static int stagedH2D(RtContext* c, void* dst, const void* src, size_t bytes)
{
    int rc = stagingReady(c); if (rc != RT_OK) return rc;
    size_t off = 0; int k = 0;
    while (off < bytes)
    {
        const size_t n = bytes - off < RtContext::STAGE_BYTES ? bytes - off : RtContext::STAGE_BYTES;
        CK(cudaEventSynchronize(c->stageEv[k]));                                   // the copy that last used this chunk has read it
        hostCopy(c->stageBuf[k], (const unsigned char*)src + off, n);
        CK(cudaMemcpyAsync((unsigned char*)dst + off, c->stageBuf[k], n, cudaMemcpyHostToDevice, c->stream));
        CK(cudaEventRecord(c->stageEv[k], c->stream));
        off += n; k ^= 1;
    }
    return RT_OK;
}
static int stagedD2H(RtContext* c, void* dst, const void* src, size_t bytes)
{
    int rc = stagingReady(c); if (rc != RT_OK) return rc;
    CK(cudaEventSynchronize(c->stageEv[0])); CK(cudaEventSynchronize(c->stageEv[1]));
    size_t off = 0, prevOff = 0, prevN = 0; int k = 0;
    while (off < bytes || prevN)
    {
        size_t n = 0;
        if (off < bytes)
        {
            n = bytes - off < RtContext::STAGE_BYTES ? bytes - off : RtContext::STAGE_BYTES;
            CK(cudaMemcpyAsync(c->stageBuf[k], (const unsigned char*)src + off, n, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaEventRecord(c->stageEv[k], c->stream));
        }
        if (prevN)                                                                  // while chunk k travels, hand chunk k ^ 1 to the caller
        {
            CK(cudaEventSynchronize(c->stageEv[k ^ 1]));
            hostCopy((unsigned char*)dst + prevOff, c->stageBuf[k ^ 1], prevN);
        }
        prevOff = off; prevN = n; off += n; k ^= 1;
    }
    return RT_OK;
}

// BVH(verts, indices, normals, quality) of the reference (BVH.cs:26, called per mesh from RayComputeManager.cs:209-232), built on the GPU:
// the same Nodes / Triangles the host builder returns (rt_bvh_build.cuh).  Host arrays in, host arrays out.
int rtBuildBVH(RtContext* c, const float* verts, int vertCount, const int* indices, int indexCount, const float* normals, int quality,
               RtTriangle* outTris, RtNode* outNodes, int nodeCapacity, int* outNodeCount)
{
    if (!c || !verts || !indices || !normals || !outTris || !outNodes || !outNodeCount) return fail(c, RT_E_INVALID, "rtBuildBVH: bad argument");
    if (quality < 0 || quality > 2) return fail(c, RT_E_INVALID, "rtBuildBVH: quality must be 0 (Low), 1 (High) or 2 (Disabled)");
    if (indexCount <= 0 || indexCount % 3 != 0) return fail(c, RT_E_INVALID, "rtBuildBVH: index count must be a positive multiple of 3");
    if (vertCount <= 0) return fail(c, RT_E_INVALID, "rtBuildBVH: no vertices");
    for (int i = 0; i < indexCount; i++)
        if (indices[i] < 0 || indices[i] >= vertCount) return fail(c, RT_E_INVALID, "rtBuildBVH: vertex index out of range");
    const int triCount = indexCount / 3;
    if (nodeCapacity < 2 * triCount + 1) return fail(c, RT_E_INVALID, "rtBuildBVH: outNodes must hold 2 * triangles + 1 entries");
    CK(cudaSetDevice(c->device));
    // ONE device allocation for inputs, outputs and the build's temporaries, kept by the context between builds (grow-only;
    // RayComputeManager builds every mesh of a scene in turn), and two pinned staging chunks for the host <-> device copies.
    const bool timing = getenv("RT_B200_BVH_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto ms = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
    const size_t bV = BvhScratch::align((size_t)vertCount * 12), bI = BvhScratch::align((size_t)indexCount * 4),
                 bNodes = BvhScratch::align(((size_t)2 * triCount + 1) * sizeof(RtNode)), bTris = BvhScratch::align((size_t)triCount * sizeof(RtTriangle));
    const size_t need = 2 * bV + bI + bNodes + bTris + BvhScratch::bytesFor(triCount);
    if (need > c->buildArena.count) { CK(cudaStreamSynchronize(c->stream)); CK(c->buildArena.ensure(need)); }
    unsigned char* base = c->buildArena.p;
    float* dV = reinterpret_cast<float*>(base); float* dN = reinterpret_cast<float*>(base + bV); int* dI = reinterpret_cast<int*>(base + 2 * bV);
    RtNode* dNodes = reinterpret_cast<RtNode*>(base + 2 * bV + bI); RtTriangle* dTris = reinterpret_cast<RtTriangle*>(base + 2 * bV + bI + bNodes);
    BvhScratch scratch; scratch.base = base + 2 * bV + bI + bNodes + bTris; scratch.cap = c->buildArena.count - (2 * bV + bI + bNodes + bTris);
    const double tAlloc = ms(t0);
    const auto t1 = std::chrono::steady_clock::now();
    int rc;
    if ((rc = stagedH2D(c, dV, verts, (size_t)vertCount * 12)) != RT_OK || (rc = stagedH2D(c, dN, normals, (size_t)vertCount * 12)) != RT_OK ||
        (rc = stagedH2D(c, dI, indices, (size_t)indexCount * 4)) != RT_OK) return rc;
    const double tUp = ms(t1);
    const auto t2 = std::chrono::steady_clock::now();
    BvhBuildResult res; std::string msg;
    cudaError_t e = bvh_build_device(dV, dN, dI, triCount, quality, dNodes, dTris, c->stream, res, msg, scratch);
    if (e != cudaSuccess) return msg.empty() ? failCuda(c, e, "rtBuildBVH") : fail(c, RT_E_STATE, "rtBuildBVH: " + msg);
    const double tBuild = ms(t2);
    const auto t3 = std::chrono::steady_clock::now();
    if ((rc = stagedD2H(c, outNodes, dNodes, (size_t)res.nodeCount * sizeof(RtNode))) != RT_OK || (rc = stagedD2H(c, outTris, dTris, (size_t)triCount * sizeof(RtTriangle))) != RT_OK) return rc;
    if (timing) fprintf(stderr, "[rtBuildBVH] %d triangles, %d nodes, %d levels: arena %.2f ms, upload %.2f ms (%.1f MB), build %.2f ms, download %.2f ms (%.1f MB), total %.2f ms\n",
                        triCount, res.nodeCount, res.levels, tAlloc, tUp, ((size_t)vertCount * 24 + (size_t)indexCount * 4) / 1e6, tBuild, ms(t3),
                        ((size_t)res.nodeCount * sizeof(RtNode) + (size_t)triCount * sizeof(RtTriangle)) / 1e6, ms(t0));
    *outNodeCount = res.nodeCount;
    return RT_OK;
}

/* Device-free test hook (not part of include/rt_b200.h): the breadth-first pair layout the upload would build.
 * pairsOut: capacity pairCap records of 64 bytes; rootsOut: 2 ints (start, count) per model.  Returns the number of pair
 * records, or a negative RT_E_* code with the reason in msg. */
int rtxPlanPairsOrdered(const RtNode* nodes, int nodeCount, const RtModel* models, int modelCount, int triCount, int smemBudget, int treeletDepth,
                        void* pairsOut, int pairCap, int* smemPairsOut, int* rootsOut, char* msg, int msgCap);
int rtxPlanPairs(const RtNode* nodes, int nodeCount, const RtModel* models, int modelCount, int triCount, int smemBudget,
                 void* pairsOut, int pairCap, int* smemPairsOut, int* rootsOut, char* msg, int msgCap)
{
    return rtxPlanPairsOrdered(nodes, nodeCount, models, modelCount, triCount, smemBudget, 0, pairsOut, pairCap, smemPairsOut, rootsOut, msg, msgCap);
}
int rtxPlanPairsOrdered(const RtNode* nodes, int nodeCount, const RtModel* models, int modelCount, int triCount, int smemBudget, int treeletDepth,
                        void* pairsOut, int pairCap, int* smemPairsOut, int* rootsOut, char* msg, int msgCap)
{
    if (!nodes || !models || nodeCount <= 0 || modelCount < 0 || !pairsOut || !rootsOut) return RT_E_INVALID;
    std::vector<RtNode> n(nodes, nodes + nodeCount);
    std::vector<RtModel> m(models, models + modelCount);
    for (int i = 0; i < modelCount; i++)
        if (m[i].nodeOffset < 0 || m[i].nodeOffset >= nodeCount) { if (msg && msgCap > 0) snprintf(msg, msgCap, "model nodeOffset out of range"); return RT_E_STATE; }
    RepackState st;
    std::vector<NodePair> out; std::string why;
    st.planScene(n, m, modelCount, (size_t)triCount, smemBudget, out, why, treeletDepth);
    if (!why.empty()) { if (msg && msgCap > 0) snprintf(msg, msgCap, "%s", why.c_str()); return RT_E_STATE; }
    if ((int)out.size() > pairCap) return RT_E_INVALID;
    if (!out.empty()) memcpy(pairsOut, out.data(), out.size() * sizeof(NodePair));
    if (smemPairsOut) *smemPairsOut = st.smemPairs;
    for (int i = 0; i < modelCount; i++)
    {
        const MeshRoot& r = st.roots[std::make_pair(m[i].nodeOffset, m[i].triOffset)];
        rootsOut[2 * i] = r.rootStart; rootsOut[2 * i + 1] = r.rootCount;
    }
    return (int)out.size();
}

/* Device-free test hook (not part of include/rt_b200.h): the TLAS the upload would build over `modelCount` world boxes
 * (6 floats each: min xyz, max xyz).  pairsOut: capacity pairCap records of 64 bytes; leavesOut: modelCount model indices in leaf
 * order; rootOut: 2 ints (start, count).  Returns the number of pair records or a negative RT_E_* code. */
int rtxPlanTlas(const float* boxes, int modelCount, void* pairsOut, int pairCap, int* leavesOut, int* rootOut)
{
    if (!boxes || modelCount < 1 || modelCount > RT_TLAS_MAX_MODELS || !pairsOut || !leavesOut || !rootOut) return RT_E_INVALID;
    std::vector<DevModel> dm(modelCount);
    for (int i = 0; i < modelCount; i++)
    {
        memset(&dm[i], 0, sizeof(DevModel));
        dm[i].wmin[0] = boxes[6 * i]; dm[i].wmin[1] = boxes[6 * i + 1]; dm[i].wmin[2] = boxes[6 * i + 2];
        dm[i].wmaxx = boxes[6 * i + 3]; dm[i].wmaxy = boxes[6 * i + 4]; dm[i].wmaxz = boxes[6 * i + 5];
    }
    std::vector<NodePair> out; std::vector<int> order;
    RepackState::planTlas(dm, modelCount, out, order, rootOut[0], rootOut[1]);
    if ((int)out.size() > pairCap) return RT_E_INVALID;
    if (!out.empty()) memcpy(pairsOut, out.data(), out.size() * sizeof(NodePair));
    memcpy(leavesOut, order.data(), order.size() * sizeof(int));
    return (int)out.size();
}

static int one_rtGetStats(RtContext* c, RtStats* out);
int rtGetStats(RtContext* c, RtStats* out)
{
    // a group reports the work of all its GPUs: counters summed, times = the slowest GPU's (they run side by side)
    int rc = one_rtGetStats(c, out);
    if (rc != RT_OK || !c) return rc;
    for (RtContext* m : c->followers)
    {
        RtStats s;
        if ((rc = one_rtGetStats(m, &s)) != RT_OK) return fail(c, rc, "GPU " + std::to_string(m->device) + ": " + m->err);
        out->rays += s.rays; out->boxTests += s.boxTests; out->triTests += s.triTests; out->sphereTests += s.sphereTests; out->sphereBoxTests += s.sphereBoxTests;
        if (s.kernelMs > out->kernelMs) out->kernelMs = s.kernelMs;
        if (s.exchangeMs > out->exchangeMs) out->exchangeMs = s.exchangeMs;
    }
    if (!c->followers.empty()) CK(cudaSetDevice(c->device));
    return RT_OK;
}
static int one_rtGetStats(RtContext* c, RtStats* out)
{
    if (!c || !out) return fail(c, RT_E_INVALID, "rtGetStats: bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    int rc = drainEvents(c); if (rc != RT_OK) return rc;
    unsigned long long h[5];
    CK(cudaMemcpy(h, c->dCounters, sizeof(h), cudaMemcpyDeviceToHost));
    c->stats.rays = h[0]; c->stats.boxTests = h[1]; c->stats.triTests = h[2]; c->stats.sphereTests = h[3]; c->stats.sphereBoxTests = h[4];
    *out = c->stats;
    return RT_OK;
}

static int one_rtResetStats(RtContext* c);
int rtResetStats(RtContext* c)
{
    int rc = one_rtResetStats(c);
    if (rc != RT_OK || !c) return rc;
    for (RtContext* m : c->followers) if ((rc = one_rtResetStats(m)) != RT_OK) return fail(c, rc, m->err);
    if (!c->followers.empty()) CK(cudaSetDevice(c->device));
    return RT_OK;
}
static int one_rtResetStats(RtContext* c)
{
    if (!c) return RT_E_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    int rc = drainEvents(c); if (rc != RT_OK) return rc;
    CK(cudaMemset(c->dCounters, 0, 8 * sizeof(unsigned long long)));
    memset(&c->stats, 0, sizeof(c->stats));
    return RT_OK;
}

} // extern "C"
