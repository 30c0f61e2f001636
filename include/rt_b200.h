/* rt_b200.h — C-ABI of librt_b200.so, the B200 (sm_100a) replacement for the reference's
 * RayTrace / ResetAccumulated compute kernels.
 *
 * The boundary it replaces is the Unity ComputeShader / ComputeBuffer / RenderTexture API exactly
 * as RayComputeManager drives it through ComputeHelper.  Every entry point below
 * names the reference call it stands in for (paths relative to the reference repo root:
 *   RCM = Assets/Scripts/Tracer/RayComputeManager.cs, CH = Assets/Scripts/Helpers/ComputeHelper.cs,
 *   RC  = Assets/Scripts/Tracer/RayCompute.compute,   HL = Assets/Scripts/Tracer/RayCommon.hlsl).
 *
 * Conventions
 *   - plain C, no torch / CUDA types in any signature (streams and device pointers travel as void*).
 *   - every function returns 0 on success, a negative RT_E_* code on failure; rtLastError() gives text.
 *     Nothing throws or exits across the ABI.
 *   - copy-in semantics: host arrays are copied before the call returns (ComputeBuffer.SetData
 *     semantics, CH:83-153); the caller may mutate them afterwards.
 *   - single caller thread per context (Unity main thread, RCM:78); rtDispatch is asynchronous
 *     w.r.t. the CPU, rtReadback / rtSynchronize synchronise.
 *   - there is NO CPU fallback: without a CUDA device rtCreate fails with RT_E_NO_DEVICE.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>
#include "rt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RtContext RtContext;

enum {
    RT_OK              =  0,
    RT_E_INVALID       = -1,  /* bad argument (NULL, negative count, wrong stride, wrong size) */
    RT_E_UNKNOWN_NAME  = -2,  /* uniform / buffer / texture / option name not part of the kernel's interface */
    RT_E_NO_DEVICE     = -3,  /* no usable CUDA device — there is no CPU path */
    RT_E_CUDA          = -4,  /* a CUDA runtime call failed; see rtLastError */
    RT_E_STATE         = -5   /* call sequence error (e.g. dispatch before rtResize / buffers) */
};

enum { RT_KERNEL_RAYTRACE = 0, RT_KERNEL_RESET_ACCUMULATED = 1 };  /* RC:1-2, RCM:57-58 */

/* ---- lifetime ------------------------------------------------------------------------------- */

/* Replaces: the ComputeShader asset reference + lazily created GPU objects (RCM:34,126-161).
 * device = CUDA ordinal this context renders on (one context per GPU / per process). */
int rtCreate(RtContext** out, int device);

/* Replaces: OnDestroy → ComputeHelper.Release(buffers, textures) (RCM:238-247, CH:209-247). */
int rtDestroy(RtContext* ctx);

/* Text of the last error on this context (never NULL). ctx may be NULL → last rtCreate error. */
const char* rtLastError(const RtContext* ctx);

/* ---- structured buffers ------------------------------------------------------------------------ */

/* Replaces: ComputeHelper.CreateStructuredBuffer(ref buf, data) + cs.SetBuffer(kernel, name, buf)
 * (RCM:151-160,201-202; CH:83-153).  name ∈ {"Triangles"(72), "Nodes"(32), "ModelInfo"(224),
 * "Spheres"(104, extension)}; stride is validated against the element size in rt_types.h exactly as
 * Unity validates it against the HLSL struct.  The allocation is re-used when count is unchanged
 * (CH:86-92).  count == 0 (data may be NULL) empties the buffer. */
int rtSetBuffer(RtContext* ctx, const char* name, const void* data, int count, int stride);

/* ---- uniforms (same names as the HLSL globals, HL:5-26,120-121, RC:7-8) --------------------------- */

int rtSetInt   (RtContext* ctx, const char* name, int value);               /* cs.SetInt    RCM:156,165-169,178-179,203 */
int rtSetInts  (RtContext* ctx, const char* name, const int* values, int n);/* cs.SetInts   RCM:139 ("Resolution")      */
int rtSetFloat (RtContext* ctx, const char* name, float value);             /* cs.SetFloat  RCM:170-174                 */
int rtSetVector(RtContext* ctx, const char* name, const float value[4]);    /* cs.SetVector RCM:140,175-176,188         */
int rtSetMatrix(RtContext* ctx, const char* name, const float value[16]);   /* cs.SetMatrix RCM:189 (column-major)      */
int rtSetBool  (RtContext* ctx, const char* name, int value);               /* cs.SetBool   RCM:180 ("accumulate")      */

/* ---- render targets ------------------------------------------------------------------------------ */

/* Replaces: CreateRenderTexture(ref raytraceFrameTex / accumulatedResult, w, h, R32G32B32A32_SFloat)
 * + SetTexture ×3 + SetInts("Resolution") (RCM:126-139, CH:303-323).  Re-allocates only when the size
 * changes; new textures are zero-filled (a fresh RenderTexture), unchanged size keeps their content. */
int rtResize(RtContext* ctx, int width, int height);

/* Replaces: cs.Dispatch(kernelIndex, groupsX, groupsY, groupsZ) (CH:25-32, RCM:75,90).
 * Thread groups are 8×8×1 (RC:10,26): pixels with x >= min(W, 8·groupsX) or y >= min(H, 8·groupsY)
 * are left untouched, like threads that were never launched.  Asynchronous. */
int rtDispatch(RtContext* ctx, int kernelIndex, int groupsX, int groupsY, int groupsZ);

/* Replaces: reading the RenderTexture back (ComputeHelper.ReadbackData / AsyncGPUReadback; the
 * reference itself only samples it on-GPU in Display.shader).  tex ∈ {"FrameRender",
 * "AccumulatedRender"}; dst receives W·H float4 in [y][x] order (row 0 = bottom of the view plane),
 * bytes must equal W·H·16.  Synchronises. */
int rtReadback(RtContext* ctx, const char* tex, float* dst, size_t bytes);

/* Replaces: RayTraceDisplay.OnRenderImage -> Graphics.Blit(rt, destination, displayMat) with Display.shader
 * (Assets/Scripts/Tracer/RayTraceDisplay.cs:9-23, Display.shader:42-47): col = tex / Frame, where tex is the accumulated
 * image and Frame = numAccumulatedFrames when accumulating, else the frame image and Frame = 1.  The quotient is then
 * encoded like Unity's linear-colour-space back buffer: clamp to [0,1], sRGB transfer function, 8 bits per channel.
 * dst receives W*H RGBA8 pixels in [y][x] order (row 0 = bottom), bytes must equal W*H*4.  Synchronises. */
int rtDisplay(RtContext* ctx, int useAccumulated, int Frame, uint8_t* dst, size_t bytes);

/* Pipelined forms of rtReadback / rtDisplay for hosts that read every frame: the call snapshots the texture (or encodes the
 * display image) on the dispatch stream as it is after the work queued so far, starts the device -> host copy on a stream of
 * its own and returns; the next rtDispatch runs while the copy is in flight.  dst (pinned memory for a real overlap) is valid
 * after rtReadbackWait (or rtSynchronize).  One copy in flight per context: a second call waits — on the device — for the first
 * copy to have read the snapshot. */
int rtReadbackAsync(RtContext* ctx, const char* tex, float* dst, size_t bytes);
int rtDisplayAsync(RtContext* ctx, int useAccumulated, int Frame, uint8_t* dst, size_t bytes);
int rtReadbackWait(RtContext* ctx);

/* Block until all work queued on this context has finished. */
int rtSynchronize(RtContext* ctx);

/* ---- extensions (no counterpart in the reference) ------------------------------------------------------- */

/* Run the context's work on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL = the
 * context's own stream).  Lets a host that owns streams (e.g. torch) time and order the dispatches. */
int rtSetStream(RtContext* ctx, void* cudaStream);

/* Row-band tiling for multi-GPU rendering: the image is cut into bands of bandRows
 * rows; band b belongs to rank (b mod worldSize).  After this call rtDispatch(RAYTRACE) traces only
 * this rank's bands (pixels keep their GLOBAL index, so seeds and results equal the 1-GPU run) and
 * writes them into the textures at their global position.  rank 0 / worldSize 1 = whole image. */
int rtSetTile(RtContext* ctx, int rank, int worldSize, int bandRows);

/* Gather / scatter staging for the per-frame all-gather of finished tiles.
 * rtPackTile   : copies this rank's bands of FrameRender and AccumulatedRender into the contiguous
 *                "TileSend" buffer ([frame bands | accumulated bands]).
 * rtUnpackTiles: scatters "TileRecv" (worldSize × TileSend, rank-major — what ncclAllGather produces)
 *                back into the two full-size textures.  Both are asynchronous on the context stream. */
int rtPackTile(RtContext* ctx);
int rtUnpackTiles(RtContext* ctx);

/* Fused tile exchange (alternative to rtPackTile / all-gather / rtUnpackTiles): every rank exports CUDA-IPC handles of
 * its two images (rtGetIpcHandles: 2 x 72 bytes = {cudaIpcMemHandle_t of the allocation, byte offset inside it} for
 * FrameRender and AccumulatedRender), the host
 * exchanges them, and each rank maps the images of the other ranks (rtSetPeers: nPeers x 2 x 72 bytes, at most 7 peers).
 * From then on the RayTrace kernel stores every finished pixel into its own images AND into every peer's, over NVLink,
 * as part of the kernel that produced it; the host only needs a stream-ordered barrier across ranks after the dispatch
 * before any rank reads its images.  rtResize with a new size and rtSetPeers(0) drop the mappings. */
int rtGetIpcHandles(RtContext* ctx, void* handles, size_t bytes);
int rtSetPeers(RtContext* ctx, int nPeers, const void* handles, size_t bytes);

/* ---- multi-GPU inside the boundary -----------------------------------------------------------------------
 * The reference host is ONE object issuing ONE dispatch per frame (RayComputeManager.RenderFrame, RCM:84-95); it must reach all
 * GPUs of the box without hand-rolling a collective.  Two ways, same exchange (pack this GPU's row bands -> one ncclAllGather
 * over NVLink -> scatter into both full-size textures, issued by rtDispatch(RT_KERNEL_RAYTRACE) itself on the context's stream):
 *
 *  (a) one process, several GPUs — what a Unity / C# host is:
 *        rtCreateMulti(&ctx, devices, n)   devices = n CUDA ordinals (NULL = 0 .. n-1).  The returned context LEADS a group: every
 *        rtSet* / rtSetBuffer / rtResize / rtSetOption call on it reaches all n GPUs (each keeps the whole scene),
 *        rtDispatch traces tile r of n on GPU r and exchanges, rtReadback / rtDisplay read GPU devices[0], whose textures are
 *        complete after the exchange; rtGetStats sums the counters of the group (times: the slowest GPU).  rtDestroy frees all.
 *        n = 1 is rtCreate.  The call sequence of RayComputeManager does not change at all.
 *  (b) one process per GPU (torchrun, mpirun, N copies of a C++ host):
 *        rank 0: rtGetUniqueId(id, 128); the host hands the 128 bytes to the other ranks by any means (file, socket, MPI,
 *        torch.distributed store); every rank: rtCreate(&ctx, localDevice) + rtCommInit(ctx, id, 128, rank, world)  — collective,
 *        returns when all ranks have joined; it also makes the context render tile `rank` of `world` (rtSetTile, bands of 8 rows;
 *        call rtSetTile afterwards with the same rank / world to change bandRows).  rtCommDestroy returns to one GPU.
 *
 * NCCL is loaded at run time (the libnccl.so.2 already in the process, else the loader's; RT_B200_NCCL_LIB overrides): the
 * library has no link-time NCCL dependency and single-GPU hosts never load it.  Option "exchange" = 0 leaves the exchange to the
 * caller (rtExchangeTiles, or rtPackTile / own collective / rtUnpackTiles as in round 1). */
#define RT_UNIQUE_ID_BYTES 128
int rtCreateMulti(RtContext** out, const int* devices, int nDevices);
int rtGetUniqueId(void* id, size_t bytes);
int rtCommInit(RtContext* ctx, const void* id, size_t bytes, int rank, int worldSize);
int rtCommDestroy(RtContext* ctx);
/* The exchange alone (pack -> all-gather -> unpack), asynchronous on the context's stream(s).  rtDispatch calls it. */
int rtExchangeTiles(RtContext* ctx);

/* Device address + size of a named device object, for zero-copy plumbing (NCCL, torch views).
 * name ∈ {"FrameRender","AccumulatedRender","TileSend","TileRecv"}. */
int rtGetDevicePointer(RtContext* ctx, const char* name, void** devPtr, size_t* bytes);

/* Replaces `new BVH(mesh.vertices, mesh.triangles, mesh.normals, quality)` (BVH.cs:26; called once per shared mesh from
 * RayComputeManager.CreateAllMeshData, RayComputeManager.cs:209-232) by a build on the GPU that takes the reference's decisions
 * node for node — the same candidate planes, costs, split test and partition order — and therefore returns the Nodes and
 * Triangles buffers the reference's builder returns (mesh-relative indices, children adjacent, left subtree first, triangles
 * in leaf order).  verts / normals: vertCount x 3 floats; indices: 3 per triangle; quality: 0 Low, 1 High, 2 Disabled
 * (BVH.cs:11-16).  outTris: indexCount / 3 entries; outNodes: nodeCapacity >= 2 * triangles + 1 entries; *outNodeCount
 * receives the number of nodes written. */
int rtBuildBVH(RtContext* ctx, const float* verts, int vertCount, const int* indices, int indexCount, const float* normals, int quality,
               RtTriangle* outTris, RtNode* outNodes, int nodeCapacity, int* outNodeCount);

/* Tuning / instrumentation switches.  name ∈
 *   "kernel"      0 = reference-shaped per-pixel megakernel, 1 = persistent threads (one path per lane),
 *                 2 = persistent-thread wavefront with per-warp path pools, sorting and ray compaction,
 *                 -1 = automatic (default): 2 when the scene has meshes, 1 for sphere-only scenes and for tiles so small that
 *                 a pool slot gets about one pixel (below 1.5 pixels per slot; 2.5 with NumRaysPerPixel = 1: multi-GPU tiles)
 *   "countStats"  1 = also count box / triangle tests (HL:254,271) — slower, off by default
 *   "sampleChunks"  kernel 1: a pixel's NumRaysPerPixel samples as this many consecutive jobs, the RNG state and the running sum
 *                 handed from the lane that finishes a chunk to the lane that takes the next (same samples in the same order:
 *                 same bits).  -1 (default) = automatic: mesh scenes on small tiles only (more than one and fewer than four
 *                 pixels per resident lane: multi-GPU tiles), about eight chunk-rounds per launch; 0 / 1 = whole pixels; 2..64 forced
 *   "exchange"    1 (default) = rtDispatch(RAYTRACE) on a context with a communicator (rtCommInit / rtCreateMulti) ends with the
 *                 all-gather of the frame's tiles; 0 = the caller exchanges (rtExchangeTiles or its own collective)
 *   "smemNodes"   number of top-of-tree node pairs staged in shared memory by TMA bulk copy (0 = off; -1 = automatic, which
 *                 is currently 0: measured, the staging never beat leaving that shared memory to L1)
 *   "poolSlots"   paths per warp pool of kernel 2: 32, 64 or 96; 0 (default) = automatic, which is 64 (measured)
 *   "modelSkip"   1 (default) = kernels 1 and 2 skip a model when the ray misses its padded world-space box or enters it beyond
 *                 the closest hit so far (exact: such a model cannot change the result); 0 = walk every model like the
 *                 reference.  Always 0 when "countStats" = 1, so that the test counts equal the reference's
 *   "tlas"        kernels 1 and 2, with "modelSkip" = 1: a tree over the models' padded world boxes (the per-frame model loop of
 *                 RayCommon.hlsl:347 for many-Model scenes).  One walk per ray segment marks the models the ray can reach; they
 *                 are then processed in buffer order against the running result, exactly like the linear loop, so the output
 *                 is unchanged.  Rebuilt on the host whenever ModelInfo is re-sent.  -1 (default) = automatic: used above 128
 *                 models; 0 = off; 1 = on for any model count (up to 4096 models; above that the linear test is used)
 *   "extInstantiation"  1 = launch the kernel instantiation that carries the extensions (peer stores, sphere accelerator)
 *                 even when none is active — for testing that instantiation on one GPU
 *   "sortRays"    kernel 2: 1 = group each warp's ray queue by direction octant before tracing, 0 = slot order (default;
 *                 measured: the grouping changes throughput by -3 % .. +1.5 %)
 *   "tailLanes"   kernel 2 leaves its trace phase when the ray queue is empty and at most this many lanes still trace
 *   "gridFit"     kernels 1 and 2: 1 = shrink the persistent grid so that every lane (pool slot) gets a whole number of pixels
 *                 — fewer, fully occupied rounds instead of a last round of mostly empty warps when the image is small for the
 *                 machine (multi-GPU tiles); 0 = one CTA set filling the machine (default; measured in round 2: the fit is 2-20 % slower, kept as an option)
 *   "l2Persist"   1 = persisting L2 access-policy window over the node-pair records on the dispatch stream (0 = default; not yet
 *                 measured)
 *   "treeletPrefetch"  1 = two-level treelet layout with flagged treelet roots and an L1 prefetch of both possible next records
 *                 (only in a library built with -DRT_TREELET_PREFETCH; the default build returns RT_E_INVALID; measured in round 2: -1 ... -30 %)
 *   "pairOrder"   order of the repacked node-pair records inside a mesh: 0 = breadth-first (default), d = 1..32 = treelets of d
 *                 levels laid out depth-first (1 = plain pre-order: child A's record follows its parent's).  Layout only: the
 *                 traversal visits the same nodes in the same order; measured in round 2: +-0.4 %
 * Unknown names return RT_E_UNKNOWN_NAME. */
int rtSetOption(RtContext* ctx, const char* name, int value);

typedef struct RtStats {
    uint64_t rays;          /* CalculateRayCollision calls (HL:487) since the last rtResetStats          */
    uint64_t boxTests;      /* stats[1] (HL:271), only when "countStats" = 1                               */
    uint64_t triTests;      /* stats[0] (HL:254), only when "countStats" = 1                               */
    uint64_t sphereTests;   /* RaySphere calls (extension), only when "countStats" = 1                     */
    uint64_t dispatches;    /* RAYTRACE dispatches since the last reset                                    */
    double   kernelMs;      /* CUDA-event time of the RAYTRACE kernels since the last reset (sum)          */
    uint64_t sphereBoxTests;/* box tests of the sphere accelerator (Spheres buffers above 64 entries), "countStats" = 1;
                               with the accelerator, sphereTests counts the sphere tests actually made               */
    double   exchangeMs;    /* CUDA-event time of the per-frame tile exchanges (pack + ncclAllGather + unpack) issued by
                               rtDispatch / rtExchangeTiles since the last reset (sum)                              */
} RtStats;

/* Synchronises, then reports the counters accumulated since the last rtResetStats. */
int rtGetStats(RtContext* ctx, RtStats* out);
int rtResetStats(RtContext* ctx);

/* ABI version of this header: (major << 16) | minor. */
int rtGetVersion(void);
#define RT_B200_VERSION ((1 << 16) | 1)

#ifdef __cplusplus
}
#endif

#endif /* RT_B200_H */
