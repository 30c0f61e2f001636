// rt_bvh_build.cuh — the reference's BVH builder (Assets/Scripts/Types/BVH.cs:26-318) as a level-synchronous GPU build.
//
// Same tree, not a similar one: every node takes the decision BVH.cs takes — the same candidate planes (up to 5 per axis,
// BVH.cs:183-250), the same cost (BVH.cs:313-318) evaluated on the same left / right boxes and counts (BVH.cs:253-311), the
// same `cost < parentCost && depth < 32` test (BVH.cs:103), the same in-place partition (BVH.cs:118-147) — so Nodes and
// Triangles come out in the order the reference's recursion (and host/BVH.cpp) produces.  What differs is the schedule:
//
//   * one level of the tree per round; all nodes of the level in parallel, one thread per TRIANGLE POSITION (a node's
//     triangles are a contiguous range, so most warps work for one node) — segmented warp reductions (shuffles) fold the
//     left / right boxes and counts of a candidate plane; a node that fits inside one warp stores its result directly, a
//     larger one combines warps with atomicMin / atomicMax on order-preserving 64-bit keys (value, position in the sequence):
//     min / max / count are exact in any order, so the cost is the serial EvaluateSplit's bit for bit, and the key's position part
//     makes the reduction return what the reference's sequential strict comparison (`if (tri.MinX < xMin) xMin = tri.MinX`,
//     BVH.cs:53-58,285-300) keeps among equal values: the FIRST one — which only shows when +0 and -0 meet in one bound;
//   * the reference's partition is a sequential swap loop whose ORDER is part of the result.  Its outcome has a closed form:
//     triangles with centre < plane keep their order at the front; a position x behind them keeps its triangle if that one
//     belongs right, otherwise it receives the triangle found by following q -> start + (#left before q) from q = start +
//     (#left before x) until q holds a right-hand triangle (each swap moves the right-hand triangle at the left block's
//     frontier to the scan position, and nothing behind the final frontier ever moves again).  With a prefix sum of the
//     left flags every position is resolved independently;
//   * node indices are assigned afterwards: the recursion appends [left, right] + everything below left + everything below
//     right, so a node's block size is 2 + size(left) + size(right) (bottom-up by level) and indices follow top-down.
//
// HBM-bound integer / byte work: per level one pass over the 40-byte build records per candidate plane plus the partition
// pass; no tensor cores, nothing to contract.  Not measured on the GPU yet (written after the round's GPU budget was spent);
// bit-exact against host/BVH.cpp on the SIMT interpreter build (tests/test_simt_kernels.py).
#pragma once
#include "rt_device.cuh"
#include <vector>

namespace rtd {

struct BuildTriD { float cx, cy, cz, minX, minY, minZ, maxX, maxY, maxZ; int index; };   // BVH.cs:459-496, 40 bytes

enum : int { BN_ACTIVE = 0, BN_LEAF = 1, BN_INNER = 2, BN_SPLIT_NOW = 3 };

struct BNode
{
    float bmin[3], bmax[3];
    int   start, count;
    int   left;                 // id of the first child (second = left + 1) once split
    int   depth, state;
    int   numCand;              // candidate planes of this node (0: cannot split)
    float parentCost;
    // best candidate so far (BVH.cs:232-246: first strictly smaller cost wins)
    float bestCost, bestPos; int bestAxis, bestNL;
    int   picked;               // some candidate was strictly cheaper than float.MaxValue (else ChooseSplit returns axis 0, pos 0: BVH.cs:204-206,248)
    float bestL[6], bestR[6];   // min xyz, max xyz of the two sides
    // numbering
    int   blockSize, index, blockStart;
};

struct CandAcc { unsigned long long lmin[3], lmax[3], rmin[3], rmax[3]; int nL, nR; };

// order-preserving integer image of a float (NaNs are filtered before: KeyMin / KeyMax)
RT_DI unsigned int f2ord(float f) { const unsigned int b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
RT_DI float ord2f(unsigned int u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
#define RT_FLT_MAX 3.402823466e+38f

// Keys of a running minimum / maximum that behave like the reference's sequential loops: `if (v < best) best = v` keeps the first
// of equal values, and equal values with different bits exist (+0 and -0).  High word: the value with -0 folded onto +0; low word:
// the position in the sequence (ascending for the minimum, descending for the maximum, so that min / max of the keys prefers the
// earlier element) and, in bit 0, whether the element was a negative zero.  Position < 2^31 (one per triangle).
// A NaN element never replaces the running value (`NaN < best` is false) and the running value is never NaN: it is skipped.
RT_DI unsigned long long KeyMin(float f, unsigned int pos)
{
    if (f != f) return ((unsigned long long)f2ord(RT_FLT_MAX) << 32) | 0xfffffffeull;             // = RT_KEY_MIN_EMPTY
    const bool zero = f == 0.0f;
    return ((unsigned long long)f2ord(zero ? 0.0f : f) << 32) | ((unsigned long long)pos << 1) | (unsigned long long)(zero && (__float_as_uint(f) >> 31));
}
RT_DI unsigned long long KeyMax(float f, unsigned int pos)
{
    if (f != f) return (unsigned long long)f2ord(-RT_FLT_MAX) << 32;                                // = RT_KEY_MAX_EMPTY
    const bool zero = f == 0.0f;
    return ((unsigned long long)f2ord(zero ? 0.0f : f) << 32) | ((unsigned long long)(0x7fffffffu - pos) << 1) | (unsigned long long)(zero && (__float_as_uint(f) >> 31));
}
RT_DI float KeyValue(unsigned long long k)
{
    const float v = ord2f((unsigned int)(k >> 32));
    return (v == 0.0f && (k & 1ull)) ? -0.0f : v;
}
// what an empty sequence leaves: float.MaxValue / float.MinValue (BVH.cs:37-42,258-270), later than every real element
#define RT_KEY_MIN_EMPTY KeyMin(RT_FLT_MAX, 0x7fffffffu)
#define RT_KEY_MAX_EMPTY KeyMax(-RT_FLT_MAX, 0x7fffffffu)

RT_DI float NodeCostD(float x, float y, float z, int numTriangles)      // BVH.cs:313-318
{
    if (numTriangles == 0) return 0.0f;
    const float area = (x * y + x * z) + y * z;
    return area * __int2float_rn(numTriangles);
}

// numSplitTests of one axis (BVH.cs:214-216): Mathf.CeilToInt(size / maxAxis * maxSplitTests) clamped to [1, maxSplitTests]
RT_DI int SplitTestsOnAxis(float sizeAxis, float maxAxis, int maxSplitTests)
{
    const float ratio = (sizeAxis / maxAxis) * __int2float_rn(maxSplitTests);
    if (ratio != ratio) return 1;
    int n = (int)ceilf(ratio);
    if (!(ratio < 1e9f)) n = maxSplitTests;
    if (n < 1) n = 1;
    if (n > maxSplitTests) n = maxSplitTests;
    return n;
}

// candidate plane c of a node, in the reference's enumeration order; false when the node has fewer candidates
RT_DI bool CandidatePlane(const BNode& nd, int quality, int c, int& axis, float& pos, int* total = nullptr)
{
    const float size[3] = {nd.bmax[0] - nd.bmin[0], nd.bmax[1] - nd.bmin[1], nd.bmax[2] - nd.bmin[2]};
    if (nd.count <= 1) { if (total) *total = 0; return false; }
    if (quality == 0)                                                      // Quality.Low: the middle of the longest axis (BVH.cs:191-202)
    {
        if (total) *total = 1;
        if (c != 0) return false;
        axis = (size[0] > size[1] && size[0] > size[2]) ? 0 : (size[1] > size[2] ? 1 : 2);
        pos = nd.bmin[axis] + size[axis] * 0.5f;
        return true;
    }
    const int maxSplitTests = nd.count < 10 ? 3 : 5;
    float maxAxis = size[0];
    if (size[1] > maxAxis) maxAxis = size[1];
    if (size[2] > maxAxis) maxAxis = size[2];
    int base = 0; bool found = false;
    for (int a = 0; a < 3; a++)
    {
        const int n = SplitTestsOnAxis(size[a], maxAxis, maxSplitTests);
        if (!found && c < base + n)
        {
            const int i = c - base;
            const float splitT = __int2float_rn(i + 1) / (__int2float_rn(n) + 1.0f);
            axis = a; pos = nd.bmin[a] + size[a] * splitT; found = true;
        }
        base += n;
    }
    if (total) *total = base;
    return found;
}

RT_DI float TriCentre(const BuildTriD& t, int axis) { return axis == 0 ? t.cx : (axis == 1 ? t.cy : t.cz); }

// ---- kernels ---------------------------------------------------------------------------------------------------------------

// BVH.cs:44-59: build records + the root box (warp-reduced, then one set of atomics per warp)
__global__ void k_bvh_init(const float* __restrict__ verts, const int* __restrict__ indices, int triCount, BuildTriD* __restrict__ tris,
                           int* __restrict__ posNode, unsigned long long* __restrict__ rootBox /* 6 keys: min xyz, max xyz */)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {RT_FLT_MAX, RT_FLT_MAX, RT_FLT_MAX}, hi[3] = {-RT_FLT_MAX, -RT_FLT_MAX, -RT_FLT_MAX};
    if (k < triCount)
    {
        const int i = 3 * k;
        const float* a = verts + 3 * (size_t)indices[i]; const float* b = verts + 3 * (size_t)indices[i + 1]; const float* c = verts + 3 * (size_t)indices[i + 2];
        BuildTriD t;
        t.cx = ((a[0] + b[0]) + c[0]) / 3.0f; t.cy = ((a[1] + b[1]) + c[1]) / 3.0f; t.cz = ((a[2] + b[2]) + c[2]) / 3.0f;
        for (int d = 0; d < 3; d++)
        {
            lo[d] = a[d] < b[d] ? (a[d] < c[d] ? a[d] : c[d]) : (b[d] < c[d] ? b[d] : c[d]);     // BVH.cs:488-493
            hi[d] = a[d] > b[d] ? (a[d] > c[d] ? a[d] : c[d]) : (b[d] > c[d] ? b[d] : c[d]);
        }
        t.minX = lo[0]; t.minY = lo[1]; t.minZ = lo[2]; t.maxX = hi[0]; t.maxY = hi[1]; t.maxZ = hi[2];
        t.index = i;
        tris[k] = t;
        posNode[k] = 0;
    }
    for (int d = 0; d < 3; d++)
    {
        unsigned long long mn = k < triCount ? KeyMin(lo[d], (unsigned int)k) : RT_KEY_MIN_EMPTY, mx = k < triCount ? KeyMax(hi[d], (unsigned int)k) : RT_KEY_MAX_EMPTY;
        for (int off = 16; off > 0; off >>= 1)
        {
            const unsigned long long omn = __shfl_down_sync(0xffffffffu, mn, off), omx = __shfl_down_sync(0xffffffffu, mx, off);
            if (omn < mn) mn = omn;
            if (omx > mx) mx = omx;
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(rootBox + d, mn); atomicMax(rootBox + 3 + d, mx); }
    }
}

__global__ void k_bvh_root(BNode* nodes, const unsigned long long* rootBox, int triCount, int* nodeCounter)
{
    if (blockIdx.x * blockDim.x + threadIdx.x != 0) return;
    BNode r; memset(&r, 0, sizeof(r));
    for (int d = 0; d < 3; d++) { r.bmin[d] = KeyValue(rootBox[d]); r.bmax[d] = KeyValue(rootBox[3 + d]); }
    r.start = 0; r.count = triCount; r.left = -1; r.depth = 0; r.state = BN_ACTIVE;
    nodes[0] = r;
    *nodeCounter = 1;
}

RT_DI void ResetAcc(CandAcc& a)
{
    for (int d = 0; d < 3; d++) { a.lmin[d] = a.rmin[d] = RT_KEY_MIN_EMPTY; a.lmax[d] = a.rmax[d] = RT_KEY_MAX_EMPTY; }
    a.nL = a.nR = 0;
}

// start of a level: every node of the level gets its candidate count, the cost of not splitting, an empty best and accumulator
__global__ void k_bvh_level_begin(BNode* nodes, CandAcc* acc, int levelStart, int levelCount, int quality)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= levelCount) return;
    BNode& nd = nodes[levelStart + s];
    int axis; float pos; int total = 0;
    CandidatePlane(nd, quality, 0, axis, pos, &total);
    if (quality == 2) total = 0;                                           // Quality.Disabled: one leaf (BVH.cs:62-66)
    nd.numCand = total;
    nd.parentCost = NodeCostD(nd.bmax[0] - nd.bmin[0], nd.bmax[1] - nd.bmin[1], nd.bmax[2] - nd.bmin[2], nd.count);
    nd.bestCost = quality == 0 ? __uint_as_float(0x7f800000u) : RT_FLT_MAX;   // Low takes its single candidate whatever it costs; High starts from float.MaxValue
    nd.bestAxis = 0; nd.bestPos = 0.0f; nd.bestNL = 0; nd.picked = 0;
    ResetAcc(acc[s]);
}

// one candidate plane for every triangle position: BVH.cs:253-311, folded per node
__global__ void k_bvh_evaluate(const BNode* __restrict__ nodes, const BuildTriD* __restrict__ tris, const int* __restrict__ posNode, int triCount,
                               CandAcc* acc, int levelStart, int quality, int cand)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int lane = threadIdx.x & 31u;
    int node = -1;
    unsigned long long v[12]; int nL = 0, nR = 0;
    for (int d = 0; d < 3; d++) { v[d] = v[6 + d] = RT_KEY_MIN_EMPTY; v[3 + d] = v[9 + d] = RT_KEY_MAX_EMPTY; }
    int nodeStart = 0, nodeCount = 0;
    if (x < triCount)
    {
        node = posNode[x];
        if (node >= 0)
        {
            const BNode& nd = nodes[node];
            int axis; float pos;
            // cand = -1: the plane ChooseSplit returns when no candidate was cheaper than float.MaxValue (costs that overflowed to
            // inf or are NaN) — axis 0, position 0 (BVH.cs:204-206).  Split() then partitions by it whatever it is (BVH.cs:101-147).
            bool use;
            if (cand < 0) { use = nd.state == BN_ACTIVE && quality == 1 && nd.numCand > 0 && !nd.picked; axis = 0; pos = 0.0f; }
            else use = nd.state == BN_ACTIVE && cand < nd.numCand && CandidatePlane(nd, quality, cand, axis, pos);
            if (use)
            {
                const BuildTriD t = tris[x];
                const bool left = TriCentre(t, axis) < pos;
                const unsigned int px = (unsigned int)x;                      // the reference's loop runs over ascending positions (BVH.cs:274)
                const unsigned long long key[6] = {KeyMin(t.minX, px), KeyMin(t.minY, px), KeyMin(t.minZ, px), KeyMax(t.maxX, px), KeyMax(t.maxY, px), KeyMax(t.maxZ, px)};
#pragma unroll
                for (int k = 0; k < 6; k++) { if (left) v[k] = key[k]; else v[6 + k] = key[k]; }     // (constant indices: v stays in registers)
                if (left) nL = 1; else nR = 1;
                nodeStart = nd.start; nodeCount = nd.count;
            }
            else node = -1;
        }
    }
    if (__ballot_sync(0xffffffffu, node >= 0) == 0u) return;       // no position of this warp belongs to a node that is being evaluated (deep levels: most triangles are final)
    // segmented reduction over runs of equal node id (a node's positions are contiguous).  A round in which no lane has its
    // off-neighbour in its own run ends the reduction: runs are contiguous, so no longer offset can match either (small nodes
    // of the deep levels need one or two rounds, not five).
    for (int off = 1; off < 32; off <<= 1)
    {
        const int other = __shfl_down_sync(0xffffffffu, node, off);
        const bool take = (lane + (unsigned)off < 32u) && other == node && node >= 0;
        if (!__any_sync(0xffffffffu, take)) break;
        for (int k = 0; k < 12; k++)
        {
            const unsigned long long ov = __shfl_down_sync(0xffffffffu, v[k], off);
            const bool isMin = (k % 6) < 3;
            if (take && (isMin ? ov < v[k] : ov > v[k])) v[k] = ov;
        }
        const int oL = __shfl_down_sync(0xffffffffu, nL, off), oR = __shfl_down_sync(0xffffffffu, nR, off);
        if (take) { nL += oL; nR += oR; }
    }
    const int prev = __shfl_up_sync(0xffffffffu, node, 1);
    const bool head = node >= 0 && (lane == 0 || prev != node);
    if (head)
    {
        CandAcc& a = acc[node - levelStart];
        const int warpFirst = x - (int)lane;
        const bool whole = nodeStart >= warpFirst && nodeStart + nodeCount <= warpFirst + 32;     // the node lives inside this warp
        if (whole)
        {
            for (int d = 0; d < 3; d++) { a.lmin[d] = v[d]; a.lmax[d] = v[3 + d]; a.rmin[d] = v[6 + d]; a.rmax[d] = v[9 + d]; }
            a.nL = nL; a.nR = nR;
        }
        else
        {
            for (int d = 0; d < 3; d++) { atomicMin(&a.lmin[d], v[d]); atomicMax(&a.lmax[d], v[3 + d]); atomicMin(&a.rmin[d], v[6 + d]); atomicMax(&a.rmax[d], v[9 + d]); }
            atomicAdd(&a.nL, nL); atomicAdd(&a.nR, nR);
        }
    }
}

// cost of the candidate just evaluated; keep it if strictly cheaper than the best so far (BVH.cs:232-246); empty the accumulator
__global__ void k_bvh_best(BNode* nodes, CandAcc* acc, int levelStart, int levelCount, int quality, int cand)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= levelCount) return;
    BNode& nd = nodes[levelStart + s];
    if (nd.state != BN_ACTIVE || cand >= nd.numCand) return;
    if (cand < 0 && (quality != 1 || nd.numCand <= 0 || nd.picked)) return;
    CandAcc& a = acc[s];
    float L[6], R[6];
    for (int d = 0; d < 3; d++) { L[d] = KeyValue(a.lmin[d]); L[3 + d] = KeyValue(a.lmax[d]); R[d] = KeyValue(a.rmin[d]); R[3 + d] = KeyValue(a.rmax[d]); }
    const float costA = NodeCostD(L[3] - L[0], L[4] - L[1], L[5] - L[2], a.nL);
    const float costB = NodeCostD(R[3] - R[0], R[4] - R[1], R[5] - R[2], a.nR);
    const float cost = costA + costB;
    if (cand < 0)
    {
        // no candidate won: the split — if float.MaxValue < parentCost lets it happen — is by (axis 0, pos 0); its sides were just evaluated
        nd.bestAxis = 0; nd.bestPos = 0.0f; nd.bestNL = a.nL;
        for (int k = 0; k < 6; k++) { nd.bestL[k] = L[k]; nd.bestR[k] = R[k]; }
    }
    else if (cost < nd.bestCost || quality == 0)
    {
        int axis; float pos;
        CandidatePlane(nd, quality, cand, axis, pos);
        nd.bestCost = cost; nd.bestAxis = axis; nd.bestPos = pos; nd.bestNL = a.nL; nd.picked = 1;
        for (int k = 0; k < 6; k++) { nd.bestL[k] = L[k]; nd.bestR[k] = R[k]; }
    }
    ResetAcc(a);
}

// BVH.cs:103: split or leaf; children are allocated in pairs (their final indices are assigned by the numbering pass)
__global__ void k_bvh_decide(BNode* nodes, int levelStart, int levelCount, int* nodeCounter, int nodeCapacity, int* overflow)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= levelCount) return;
    BNode& nd = nodes[levelStart + s];
    if (nd.state != BN_ACTIVE) return;
    const float cost = nd.numCand > 0 ? nd.bestCost : __uint_as_float(0x7f800000u);
    if (cost < nd.parentCost && nd.depth < 32)
    {
        const int id = atomicAdd(nodeCounter, 2);
        if (id + 2 > nodeCapacity) { *overflow = 1; nd.state = BN_LEAF; return; }
        BNode l; memset(&l, 0, sizeof(l)); BNode r = l;
        for (int d = 0; d < 3; d++) { l.bmin[d] = nd.bestL[d]; l.bmax[d] = nd.bestL[3 + d]; r.bmin[d] = nd.bestR[d]; r.bmax[d] = nd.bestR[3 + d]; }
        l.start = nd.start; l.count = nd.bestNL; r.start = nd.start + nd.bestNL; r.count = nd.count - nd.bestNL;
        l.left = r.left = -1; l.depth = r.depth = nd.depth + 1; l.state = r.state = BN_ACTIVE;
        nodes[id] = l; nodes[id + 1] = r;
        nd.left = id; nd.state = BN_SPLIT_NOW;
    }
    else nd.state = BN_LEAF;
}

__global__ void k_bvh_flags(const BNode* __restrict__ nodes, const BuildTriD* __restrict__ tris, const int* __restrict__ posNode, int triCount, int* __restrict__ flags)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= triCount) return;
    const int node = posNode[x];
    int f = 0;
    if (node >= 0)
    {
        const BNode& nd = nodes[node];
        if (nd.state == BN_SPLIT_NOW) f = TriCentre(tris[x], nd.bestAxis) < nd.bestPos ? 1 : 0;
    }
    flags[x] = f;
}

// ---- exclusive prefix sum of the flags (three small kernels) -----------------------------------------------------------------
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

RT_DI int BlockExclusiveScan(int v, int* sh /* SCAN_THREADS */, int& total)
{
    const int t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (int off = 1; off < SCAN_THREADS; off <<= 1)
    {
        const int add = t >= off ? sh[t - off] : 0;
        __syncthreads();
        sh[t] += add;
        __syncthreads();
    }
    total = sh[SCAN_THREADS - 1];
    const int incl = sh[t];
    __syncthreads();
    return incl - v;
}

__global__ void k_scan_tiles(const int* __restrict__ in, int n, int* __restrict__ out, int* __restrict__ tileSums)
{
    __shared__ int sh[SCAN_THREADS];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS], sum = 0;
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = base + k < n ? in[base + k] : 0; sum += v[k]; }
    int total;
    int run = BlockExclusiveScan(sum, sh, total);
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = run; run += v[k]; }
    if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

__global__ void k_scan_sums(int* tileSums, int numTiles)                 // one CTA: exclusive scan of the tile totals, in place
{
    __shared__ int sh[SCAN_THREADS];
    int carry = 0;
    for (int base = 0; base < numTiles; base += SCAN_THREADS)
    {
        const int i = base + threadIdx.x;
        const int v = i < numTiles ? tileSums[i] : 0;
        int total;
        const int ex = BlockExclusiveScan(v, sh, total);
        if (i < numTiles) tileSums[i] = carry + ex;
        carry += total;
    }
}

__global__ void k_scan_add(int* out, int n, const int* __restrict__ tileSums)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tileSums[i / SCAN_TILE];
}

// the reference's swap loop (BVH.cs:118-147) in closed form; see the header of this file
__global__ void k_bvh_partition(const BNode* __restrict__ nodes, const BuildTriD* __restrict__ src, const int* __restrict__ posNode, const int* __restrict__ flags,
                                const int* __restrict__ prefix, int triCount, BuildTriD* __restrict__ dst, int* __restrict__ dstNode, int* __restrict__ inconsistent)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= triCount) return;
    const int node = posNode[x];
    if (node < 0) { dst[x] = src[x]; dstNode[x] = -1; return; }
    const BNode& nd = nodes[node];
    if (nd.state != BN_SPLIT_NOW) { dst[x] = src[x]; dstNode[x] = -1; return; }      // became a leaf: its triangles are final
    const int s = nd.start, g0 = prefix[s], nL = nd.bestNL;
    if (flags[x])
    {
        const int d = s + (prefix[x] - g0);                                // left: stable, at the front
        dst[d] = src[x]; dstNode[d] = nd.left;
    }
    if (x >= s + nL)
    {
        int q = x;
        if (flags[x])
        {
            // every step moves q to a strictly smaller position as long as nL is the number of set flags of the node (it is: both come
            // from the same comparison); the bound turns a broken invariant into an error code instead of a kernel that never ends
            q = s + (prefix[x] - g0);
            int steps = 0;
            while (flags[q] && steps <= nd.count) { q = s + (prefix[q] - g0); steps++; }
            if (steps > nd.count) { *inconsistent = 1; return; }
        }
        dst[x] = src[q]; dstNode[x] = nd.left + 1;
    }
}

__global__ void k_bvh_finish_level(BNode* nodes, int levelStart, int levelCount)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < levelCount && nodes[levelStart + s].state == BN_SPLIT_NOW) nodes[levelStart + s].state = BN_INNER;
}

// ---- numbering and output ----------------------------------------------------------------------------------------------------
__global__ void k_bvh_sizes(BNode* nodes, int levelStart, int levelCount)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= levelCount) return;
    BNode& nd = nodes[levelStart + s];
    nd.blockSize = nd.state == BN_INNER ? 2 + nodes[nd.left].blockSize + nodes[nd.left + 1].blockSize : 0;
}
__global__ void k_bvh_indices(BNode* nodes, int levelStart, int levelCount)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= levelCount) return;
    BNode& nd = nodes[levelStart + s];
    if (levelStart == 0) { nd.index = 0; nd.blockStart = 1; }
    if (nd.state != BN_INNER) return;
    BNode& l = nodes[nd.left]; BNode& r = nodes[nd.left + 1];
    l.index = nd.blockStart; r.index = nd.blockStart + 1;
    l.blockStart = nd.blockStart + 2; r.blockStart = nd.blockStart + 2 + l.blockSize;
}
__global__ void k_bvh_emit_nodes(const BNode* __restrict__ nodes, int nodeCount, RtNode* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nodeCount) return;
    const BNode& nd = nodes[i];
    RtNode o;
    for (int d = 0; d < 3; d++) { o.boundsMin[d] = nd.bmin[d]; o.boundsMax[d] = nd.bmax[d]; }
    if (nd.state == BN_INNER) { o.startIndex = nodes[nd.left].index; o.triangleCount = i == 0 ? -1 : 0; }   // (the reference leaves the root's count at -1)
    else { o.startIndex = nd.start; o.triangleCount = nd.count; }
    out[nd.index] = o;
}
__global__ void k_bvh_emit_tris(const BuildTriD* __restrict__ tris, int triCount, const float* __restrict__ verts, const float* __restrict__ normals,
                                const int* __restrict__ indices, RtTriangle* __restrict__ out)          // BVH.cs:69-80
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= triCount) return;
    const int base = tris[i].index;
    RtTriangle o;
    const float* p[3] = {verts + 3 * (size_t)indices[base], verts + 3 * (size_t)indices[base + 1], verts + 3 * (size_t)indices[base + 2]};
    const float* n[3] = {normals + 3 * (size_t)indices[base], normals + 3 * (size_t)indices[base + 1], normals + 3 * (size_t)indices[base + 2]};
    for (int d = 0; d < 3; d++)
    {
        o.posA[d] = p[0][d]; o.posB[d] = p[1][d]; o.posC[d] = p[2][d];
        o.normA[d] = n[0][d]; o.normB[d] = n[1][d]; o.normC[d] = n[2][d];
    }
    out[i] = o;
}

// ---- host driver -------------------------------------------------------------------------------------------------------------
struct BvhBuildResult { int nodeCount = 0, levels = 0; };

// Scratch of one build, carved from ONE allocation the caller keeps between builds (a cudaMalloc / cudaFree pair per temporary cost
// more than the kernels: 11 of each per build, every cudaFree a device-wide synchronisation).
struct BvhScratch
{
    unsigned char* base = nullptr; size_t cap = 0, off = 0;
    static size_t align(size_t n) { return (n + 255) & ~(size_t)255; }
    static size_t bytesFor(int triCount)
    {
        const size_t n = (size_t)triCount, numTiles = (n + SCAN_TILE - 1) / SCAN_TILE;
        return 2 * align(n * sizeof(BuildTriD)) + 2 * align(n * sizeof(int)) + 2 * align(n * sizeof(int)) + align(numTiles * sizeof(int)) + align(3 * sizeof(int))
             + align(6 * sizeof(unsigned long long)) + align((2 * n + 1) * sizeof(BNode)) + align((n + 1) * sizeof(CandAcc));
    }
    template <class T> T* take(size_t count) { T* p = reinterpret_cast<T*>(base + off); off += align(count * sizeof(T)); return p; }
};

// Device buffers in, device buffers out (dNodesOut: capacity 2 * triCount + 1, dTrisOut: triCount).  Everything on `stream`;
// one 12-byte read-back per level tells the host how many nodes the next level has.  scratch: BvhScratch::bytesFor(triCount) bytes.
inline cudaError_t bvh_build_device(const float* dVerts, const float* dNormals, const int* dIndices, int triCount, int quality,
                                    RtNode* dNodesOut, RtTriangle* dTrisOut, cudaStream_t stream, BvhBuildResult& res, std::string& msg, BvhScratch scratch)
{
    cudaError_t e = cudaSuccess;
    const int nodeCapacity = 2 * triCount + 1;
    const int numTiles = (triCount + SCAN_TILE - 1) / SCAN_TILE;
    if (!scratch.base || scratch.cap < BvhScratch::bytesFor(triCount)) { msg = "BVH build: scratch too small (internal error)"; return cudaErrorInvalidValue; }
    scratch.off = 0;
    BuildTriD* tris[2]; int* posNode[2];
    for (int k = 0; k < 2; k++) { tris[k] = scratch.take<BuildTriD>((size_t)triCount); posNode[k] = scratch.take<int>((size_t)triCount); }
    int* flags = scratch.take<int>((size_t)triCount);
    int* prefix = scratch.take<int>((size_t)triCount);
    int* tileSums = scratch.take<int>((size_t)numTiles);
    int* counters = scratch.take<int>(3);                                   // [0] nodes allocated  [1] capacity exceeded  [2] partition invariant broken
    unsigned long long* rootBox = scratch.take<unsigned long long>(6);
    BNode* nodes = scratch.take<BNode>((size_t)nodeCapacity);
    CandAcc* acc = scratch.take<CandAcc>((size_t)triCount + 1);            // one slot per node of a level; every node owns at least one triangle
    auto cleanup = [&]() { };
#define RT_BVH_CK(call) do { e = (call); if (e != cudaSuccess) { cleanup(); return e; } } while (0)
    {
        static const unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};      // above / below every key
        RT_BVH_CK(cudaMemcpyAsync(rootBox, init, sizeof(init), cudaMemcpyHostToDevice, stream));
        RT_BVH_CK(cudaMemsetAsync(counters, 0, 3 * sizeof(int), stream));
    }
    const unsigned int T = 256, gridTri = (unsigned int)((triCount + T - 1) / T);
    RT_LAUNCH(gridTri, T, 0, stream, k_bvh_init, dVerts, dIndices, triCount, tris[0], posNode[0], rootBox);
    RT_LAUNCH(1, 32, 0, stream, k_bvh_root, nodes, rootBox, triCount, counters);

    std::vector<int> levelStart(1, 0);
    int cur = 0, nodeCount = 1, levelBegin = 0;
    const int maxCand = quality == 1 ? 15 : (quality == 0 ? 1 : 0);
    for (int level = 0; level <= 33; level++)
    {
        const int levelCount = nodeCount - levelBegin;
        if (levelCount <= 0) break;
        const unsigned int gridLvl = (unsigned int)((levelCount + T - 1) / T);
        RT_LAUNCH(gridLvl, T, 0, stream, k_bvh_level_begin, nodes, acc, levelBegin, levelCount, quality);
        for (int c = 0; c < maxCand; c++)
        {
            RT_LAUNCH(gridTri, T, 0, stream, k_bvh_evaluate, nodes, tris[cur], posNode[cur], triCount, acc, levelBegin, quality, c);
            RT_LAUNCH(gridLvl, T, 0, stream, k_bvh_best, nodes, acc, levelBegin, levelCount, quality, c);
        }
        if (quality == 1)
        {
            // nodes none of whose candidates was cheaper than float.MaxValue (overflowed or NaN costs): the reference still splits them,
            // by axis 0 / position 0, when float.MaxValue < parentCost.  Threads of every other node leave this pass at once.
            RT_LAUNCH(gridTri, T, 0, stream, k_bvh_evaluate, nodes, tris[cur], posNode[cur], triCount, acc, levelBegin, quality, -1);
            RT_LAUNCH(gridLvl, T, 0, stream, k_bvh_best, nodes, acc, levelBegin, levelCount, quality, -1);
        }
        RT_LAUNCH(gridLvl, T, 0, stream, k_bvh_decide, nodes, levelBegin, levelCount, counters, nodeCapacity, counters + 1);
        RT_LAUNCH(gridTri, T, 0, stream, k_bvh_flags, nodes, tris[cur], posNode[cur], triCount, flags);
        RT_LAUNCH((unsigned int)numTiles, SCAN_THREADS, 0, stream, k_scan_tiles, flags, triCount, prefix, tileSums);
        RT_LAUNCH(1, SCAN_THREADS, 0, stream, k_scan_sums, tileSums, numTiles);
        RT_LAUNCH(gridTri, T, 0, stream, k_scan_add, prefix, triCount, tileSums);
        RT_LAUNCH(gridTri, T, 0, stream, k_bvh_partition, nodes, tris[cur], posNode[cur], flags, prefix, triCount, tris[cur ^ 1], posNode[cur ^ 1], counters + 2);
        RT_LAUNCH(gridLvl, T, 0, stream, k_bvh_finish_level, nodes, levelBegin, levelCount);
        cur ^= 1;
        int h[3] = {0, 0, 0};
        RT_BVH_CK(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
        RT_BVH_CK(cudaStreamSynchronize(stream));
        if (h[1]) { msg = "BVH node capacity exceeded"; cleanup(); return cudaErrorInvalidValue; }
        if (h[2]) { msg = "BVH partition: left count and flags disagree (internal error)"; cleanup(); return cudaErrorInvalidValue; }
        levelBegin = nodeCount; nodeCount = h[0];
        levelStart.push_back(levelBegin);
        res.levels = level + 1;
    }
    // numbering: block sizes bottom-up, indices top-down (levelStart[k] .. levelStart[k + 1])
    levelStart.push_back(nodeCount);
    const int numLevels = (int)levelStart.size() - 1;
    for (int k = numLevels - 1; k >= 0; k--)
    {
        const int cnt = levelStart[k + 1] - levelStart[k];
        if (cnt > 0) RT_LAUNCH((unsigned int)((cnt + T - 1) / T), T, 0, stream, k_bvh_sizes, nodes, levelStart[k], cnt);
    }
    for (int k = 0; k < numLevels; k++)
    {
        const int cnt = levelStart[k + 1] - levelStart[k];
        if (cnt > 0) RT_LAUNCH((unsigned int)((cnt + T - 1) / T), T, 0, stream, k_bvh_indices, nodes, levelStart[k], cnt);
    }
    RT_LAUNCH((unsigned int)((nodeCount + T - 1) / T), T, 0, stream, k_bvh_emit_nodes, nodes, nodeCount, dNodesOut);
    RT_LAUNCH(gridTri, T, 0, stream, k_bvh_emit_tris, tris[cur], triCount, dVerts, dNormals, dIndices, dTrisOut);
    RT_BVH_CK(cudaGetLastError());
    RT_BVH_CK(cudaStreamSynchronize(stream));
#undef RT_BVH_CK
    cleanup();
    res.nodeCount = nodeCount;
    return cudaSuccess;
}

} // namespace rtd
