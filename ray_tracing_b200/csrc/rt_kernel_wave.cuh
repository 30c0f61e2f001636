// rt_kernel_wave.cuh — kernel variant 1: persistent threads, one path per lane (the default for sphere-only scenes; the
// pooled wavefront kernel, rt_kernel_pool.cuh, is the default when the scene has meshes).
//
// Shape (this implementation's own; the arithmetic per path is the reference's, see rt_device.cuh):
//   * persistent CTAs (a multiple of the SM count) pull pixel jobs from one global counter.  A lane that
//     finishes its pixel is refilled in place: the lanes that need work are found with a warp ballot, ONE
//     atomicAdd per warp reserves a contiguous job range, and the lanes take consecutive jobs by popcount
//     rank (warp-ballot compaction of the free slots).  Jobs are ordered in 8x4-pixel tiles so a freshly
//     filled warp traces a coherent 32-pixel block.
//   * a pixel's samples stay on one lane, in order, because the reference threads ONE rng state through all
//     NumRaysPerPixel samples of a pixel (HL:552,563-579) and sums them in order (HL:578): this keeps
//     results bit-identical while every lane always has a live path segment to intersect (no idle lanes
//     after early path termination, which is what the megakernel wastes).
//   * every loop iteration is one wavefront step for the warp: [refill] -> [generate] -> [intersect] ->
//     [shade]; all live lanes run the intersection phase together.
//   * the top of every BVH (the first `smemPairs` breadth-first NodePair records, contiguous by
//     construction) is staged into shared memory once per CTA by a TMA bulk copy (cp.async.bulk +
//     mbarrier complete_tx); deeper records and triangles are read with 128-bit vector loads from the
//     repacked 64-byte / 48-byte aligned streams.  Spheres are staged in shared memory too.
//   * traversal visits exactly the nodes and triangles the reference visits, in the same order, with the
//     same push-time-only culling (HL:243-283) — so closest hits, tie winners and the box/triangle test
//     counts equal the oracle's.  What is removed is the re-read of a popped node (the stack carries the
//     node's (start,count)) and the push/pop round trip of the near child.
#pragma once
#include "rt_device.cuh"

namespace rtd {

constexpr int WAVE_THREADS = 128;
constexpr int WAVE_MAX_SMEM_SPHERES = 256;
constexpr int WAVE_STACK = 64;

struct WaveSmemHeader
{
    unsigned long long mbar;
    unsigned int pad[14];
};

#ifndef RT_SIMT_EMU      // (the test-only host build, tests/simt, supplies its own versions of these six helpers)
RT_DI uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

RT_DI void mbar_init(uint32_t mbar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory"); }
RT_DI void mbar_expect_tx(uint32_t mbar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory"); }
RT_DI void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
RT_DI bool mbar_try_wait(uint32_t mbar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    return ok != 0;
}
RT_DI void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif

struct NodeRef { int start, count; };

// RT_TREELET_PREFETCH (compiled out by default; option "treeletPrefetch"; measured in round 2: -1 ... -30 %): with records laid out as two-level
// treelets — [record of a node, records of its inner children] contiguous, "pairOrder" = 2 — a reference to a treelet's first
// record carries bit 30.  Fetching such a record also asks L1 for the one or two records behind it, i.e. for BOTH possible next
// steps of the descent, before the box tests have decided which one it is: every second level of a descent then finds its
// record in L1 instead of paying another trip to L2 — the dependent-fetch chain that stalls the mesh scenes (profiles/) is halved.
constexpr int TREELET_ROOT_BIT = 0x40000000;

// Fetch one 64-byte pair record: from the shared-memory copy of the tree tops (STAGED instantiations only: the test for it was 4.7 %
// of the warp instructions of the plain pooled kernel, for an option that measured neutral-to-negative twice — profiles/r02_e_*), or from L2 / HBM.
template <bool STAGED>
RT_DI void LoadPair(const DevParams& P, const float4* __restrict__ smemPairs, int idx, float4& q0, float4& q1, float4& q2, float4& q3)
{
#ifdef RT_TREELET_PREFETCH
    const bool treeletRoot = (idx & TREELET_ROOT_BIT) != 0;
    idx &= ~TREELET_ROOT_BIT;
#endif
    if (STAGED && idx < P.smemPairs)
    {
        const float4* p = smemPairs + (size_t)idx * 4;
        q0 = p[0]; q1 = p[1]; q2 = p[2]; q3 = p[3];
    }
    else
    {
        const float4* p = reinterpret_cast<const float4*>(P.pairs + idx);
        ldg_record64(p, q0, q1, q2, q3);
#if defined(RT_TREELET_PREFETCH) && !defined(RT_SIMT_EMU)
        if (treeletRoot)
        {
            asm volatile("prefetch.global.L1 [%0];" :: "l"(p + 4));      // the records of the inner children (or, for a one-child
            asm volatile("prefetch.global.L1 [%0];" :: "l"(p + 8));      // treelet, the start of the next treelet: harmless)
        }
#endif
    }
}

// HL:234-287 on the repacked streams.  Same visiting order and culling as the reference.
template <bool STATS, bool STAGED>
RT_DI void TraverseMesh(const DevParams& P, const float4* __restrict__ smemPairs, f3 pos, f3 dir, f3 invDir, float rayLength,
                        NodeRef root, bool cullBackface,
                        float& bestDst, int& bestTri, float& bestU, float& bestV, float& bestDet, Counters& cnt)
{
    bestDst = rayLength; bestTri = -1;
    NodeRef stack[WAVE_STACK];
    int stackCount = 0;
    NodeRef cur = root;
    for (;;)
    {
        if (cur.count > 0)
        {
            const float4* g = reinterpret_cast<const float4*>(P.triGeom + cur.start);
            for (int i = 0; i < cur.count; i++, g += TRI_GEOM_F4)
            {
                float4 g0, g1, g2;
                ldg_trigeom(g, g0, g1, g2);
                float dst, u, v, det;
                const bool didHit = RayTriangleCore(pos, dir, make_f3(g0.x, g0.y, g0.z), make_f3(g0.w, g1.x, g1.y), make_f3(g1.z, g1.w, g2.x),
                                                    make_f3(g2.y, g2.z, g2.w), cullBackface, dst, u, v, det);
                if (STATS) cnt.tri++;
                if (didHit && dst < bestDst) { bestDst = dst; bestTri = cur.start + i; bestU = u; bestV = v; bestDet = det; }
            }
            if (stackCount == 0) return;
            cur = stack[--stackCount];
        }
        else
        {
            float4 q0, q1, q2, q3;
            LoadPair<STAGED>(P, smemPairs, cur.start, q0, q1, q2, q3);
            const float dstA = RayBoundingBoxDst(pos, invDir, make_f3(q0.x, q0.y, q0.z), make_f3(q0.w, q1.x, q1.y));
            const float dstB = RayBoundingBoxDst(pos, invDir, make_f3(q2.x, q2.y, q2.z), make_f3(q2.w, q3.x, q3.y));
            if (STATS) cnt.box += 2;
            NodeRef a, b;
            a.start = __float_as_int(q1.z); a.count = __float_as_int(q1.w);
            b.start = __float_as_int(q3.z); b.count = __float_as_int(q3.w);
            const bool isNearestA = dstA <= dstB;
            const float dstNear = isNearestA ? dstA : dstB;
            const float dstFar = isNearestA ? dstB : dstA;
            const NodeRef nearRef = isNearestA ? a : b;
            const NodeRef farRef = isNearestA ? b : a;
            // reference: push far, push near, pop (= near).  Equivalent: push far, continue with near.
            if (dstFar < bestDst && stackCount < WAVE_STACK) stack[stackCount++] = farRef;
            if (dstNear < bestDst) cur = nearRef;
            else { if (stackCount == 0) return; cur = stack[--stackCount]; }
        }
    }
}

// HL:335-374 (+ sphere extension) on the repacked streams
template <bool STATS, bool EXT, bool TLAS>
RT_DI Hit Intersect(const DevParams& P, const float4* __restrict__ smemPairs, const DevSphere* __restrict__ smemSpheres,
                    f3 rayPos, f3 rayDir, Counters& cnt)
{
    Hit result;
    result.dst = inf32(); result.isBackface = false; result.normal = splat3(0.0f); result.pos = splat3(0.0f); result.material = nullptr;
    cnt.rays++;

    int bestSphere = -1; bool bestInside = false;
    if (EXT && P.sphBvh)
    {
        int idx = 0x7fffffff, flag = 0;
        TraverseSpheres(P, rayPos, rayDir, result.dst, idx, bestInside, flag, cnt, STATS);
        if (idx != 0x7fffffff) bestSphere = idx;
    }
    else
    for (int i = 0; i < P.sphereCount; i++)
    {
        float cx, cy, cz, r2;
        if (i < WAVE_MAX_SMEM_SPHERES) { const DevSphere s = smemSpheres[i]; cx = s.cx; cy = s.cy; cz = s.cz; r2 = s.r2; }
        else { const float4 s0 = __ldg(reinterpret_cast<const float4*>(P.spheres + i)); cx = s0.x; cy = s0.y; cz = s0.z; r2 = __ldg(&P.spheres[i].r2); }
        float dst; bool inside;
        if (STATS) cnt.sph++;
        if (RaySphereCore(rayPos, rayDir, make_f3(cx, cy, cz), r2, dst, inside) && dst < result.dst)
        {
            result.dst = dst; bestSphere = i; bestInside = inside;
        }
    }
    if (bestSphere >= 0)
    {
        // HL:319-320 for the winner only (position / normal of a sphere hit depend on nothing but the winner)
        f3 centre;
        if (bestSphere < WAVE_MAX_SMEM_SPHERES) { const DevSphere s = smemSpheres[bestSphere]; centre = make_f3(s.cx, s.cy, s.cz); }
        else centre = make_f3(P.spheres[bestSphere].cx, P.spheres[bestSphere].cy, P.spheres[bestSphere].cz);
        result.isBackface = bestInside;
        result.pos = rayPos + rayDir * result.dst;
        result.normal = normalize3(result.pos - centre) * (bestInside ? -1.0f : 1.0f);
        result.material = &P.Spheres[bestSphere].material;
    }

    const f3 rayInv = rcp3(rayDir);
    // many-model scenes (<TLAS> kernels only): one walk of the TLAS marks the models this ray can reach at all; the loop below
    // then jumps from marked model to marked model, still in buffer order and still re-testing each against the running result
    unsigned int tlasMask[TLAS ? RT_TLAS_WORDS : 1];
    const bool useTlas = TLAS && !STATS && P.modelSkip && P.tlas;
    if (useTlas) TlasCollect(P, rayPos, rayInv, result.dst, tlasMask);
    for (int i = 0; i < P.modelCount; i++)
    {
        if (useTlas) { i = TlasNext(tlasMask, i, P.modelCount); if (i >= P.modelCount) break; }
        const DevModel* m = P.models + i;
        const float4* mr = reinterpret_cast<const float4*>(m);
        // the instrumented build walks every model like the reference (identical test counts); otherwise models the ray cannot
        // reach before its current best hit are skipped — they could not have changed the result
        if (!STATS && P.modelSkip && ModelOutOfReach(mr, rayPos, rayInv, result.dst)) continue;
        float w2l[12];
        {
            const float4 r0 = __ldg(mr), r1 = __ldg(mr + 1), r2 = __ldg(mr + 2);
            w2l[0] = r0.x; w2l[1] = r0.y; w2l[2] = r0.z; w2l[3] = r0.w;
            w2l[4] = r1.x; w2l[5] = r1.y; w2l[6] = r1.z; w2l[7] = r1.w;
            w2l[8] = r2.x; w2l[9] = r2.y; w2l[10] = r2.z; w2l[11] = r2.w;
        }
        const int4 meta = __ldg(reinterpret_cast<const int4*>(mr + 6));     // rootStart, rootCount, cullBackface, matIndex
        const f3 localPos = mul_rm(w2l, rayPos, 1.0f);
        const f3 localDir = mul_rm(w2l, rayDir, 0.0f);
        const f3 invDir = rcp3(localDir);
        NodeRef root; root.start = meta.x; root.count = meta.y;
        float dst, u, v, det; int tri;
        TraverseMesh<STATS, EXT>(P, smemPairs, localPos, localDir, invDir, result.dst, root, meta.z != 0, dst, tri, u, v, det, cnt);
        if (dst < result.dst)
        {
            const float4* nq = reinterpret_cast<const float4*>(P.triNormals + tri);
            const float4 n0 = ldg_tri(nq), n1 = ldg_tri(nq + 1), n2 = ldg_tri(nq + 2);
            const f3 n = TriangleSmoothNormal(make_f3(n0.x, n0.y, n0.z), make_f3(n0.w, n1.x, n1.y), make_f3(n1.z, n1.w, n2.x), u, v, det);
            float l2w[12];
            {
                const float4 r0 = __ldg(mr + 3), r1 = __ldg(mr + 4), r2 = __ldg(mr + 5);
                l2w[0] = r0.x; l2w[1] = r0.y; l2w[2] = r0.z; l2w[3] = r0.w;
                l2w[4] = r1.x; l2w[5] = r1.y; l2w[6] = r1.z; l2w[7] = r1.w;
                l2w[8] = r2.x; l2w[9] = r2.y; l2w[10] = r2.z; l2w[11] = r2.w;
            }
            result.isBackface = det < 0.0f;
            result.dst = dst;
            result.normal = normalize3(mul_rm(l2w, n, 0.0f));
            result.pos = rayPos + rayDir * dst;
            result.material = &P.ModelInfo[meta.w].material;
        }
    }
    return result;
}

#if defined(RT_SIMT_PROFILE) && defined(RT_SIMT_EMU)
#define WAVE_PROF(i, v) do { if ((threadIdx.x & 31u) == 0) simt::prof_add(i, (unsigned long long)(v)); } while (0)
#define WAVE_PROF_LANES(i, pred) do { const unsigned m_ = __ballot_sync(0xffffffffu, (pred)); if ((threadIdx.x & 31u) == 0) { simt::prof_add(i, (unsigned long long)__popc(m_)); simt::prof_add((i) + 20, m_ ? 1u : 0u); } } while (0)
#define WAVE_PROF_LANE(i, pred) do { if (pred) simt::prof_add(i, 1); } while (0)
#else
#define WAVE_PROF(i, v) do { } while (0)
#define WAVE_PROF_LANES(i, pred) do { } while (0)
#define WAVE_PROF_LANE(i, pred) do { } while (0)
#endif
#ifndef RT_WAVE_MINBLOCKS
#define RT_WAVE_MINBLOCKS 6      // <= 85 registers: 24 warps per SM (measured: 31.0 ms vs 37.1 ms at 16 warps on config 2)
#endif
// The kernel body.  TLAS = the many-model instantiation (its own kernel, k_raytrace_wave_tlas, so that the kernels measured in
// round 1 keep their code: the TLAS walk and its mask exist only there).
template <bool STATS, bool EXT, bool TLAS, bool CHUNKED>
RT_DI void wave_body(const DevParams& P, const unsigned int totalJobs, const unsigned int tilesX, const unsigned int ownedRows)
{
    RT_DYNAMIC_SMEM(smemRaw);
    WaveSmemHeader* hdr = reinterpret_cast<WaveSmemHeader*>(smemRaw);
    float4* smemPairs = reinterpret_cast<float4*>(smemRaw + sizeof(WaveSmemHeader));
    DevSphere* smemSpheres = reinterpret_cast<DevSphere*>(smemRaw + sizeof(WaveSmemHeader) + (size_t)P.smemPairs * sizeof(NodePair));

    // ---- stage the tree tops (TMA bulk copy) and the spheres ------------------------------------------------------------
    const uint32_t mbar = smem_u32(&hdr->mbar);
    if (threadIdx.x == 0)
    {
        mbar_init(mbar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (P.smemPairs > 0 && threadIdx.x == 0)
    {
        const uint32_t bytes = (uint32_t)P.smemPairs * (uint32_t)sizeof(NodePair);
        mbar_expect_tx(mbar, bytes);
        // chunks of <= 32 KB; all complete on the same mbarrier phase
        uint32_t off = 0;
        while (off < bytes)
        {
            const uint32_t n = (bytes - off) < 32768u ? (bytes - off) : 32768u;
            tma_bulk_g2s(smem_u32(smemPairs) + off, reinterpret_cast<const unsigned char*>(P.pairs) + off, n, mbar);
            off += n;
        }
    }
    {
        const int nS = P.sphereCount < WAVE_MAX_SMEM_SPHERES ? P.sphereCount : WAVE_MAX_SMEM_SPHERES;
        for (int i = threadIdx.x; i < nS; i += WAVE_THREADS) smemSpheres[i] = P.spheres[i];
    }
    if (P.smemPairs > 0) { while (!mbar_try_wait(mbar, 0)) { } }
    __syncthreads();

    // ---- persistent wavefront loop ------------------------------------------------------------------------------------------
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int laneMaskLt = (1u << lane) - 1u;
    Counters cnt; cnt.rays = cnt.box = cnt.tri = cnt.sph = cnt.sbox = 0;

    bool exhausted = false;          // no more jobs for this lane
    bool havePixel = false;
    bool pathActive = false;
    PixelSetup px; px.camOrigin = px.focusPoint = px.camRight = px.camUp = splat3(0.0f); px.rngState = 0;
    uint32_t rngState = 0;
    f3 totalIncomingLight = splat3(0.0f);
    int sample = 0, bounce = 0;
    size_t pixelOffset = 0;
    PathState ray; ray.pos = ray.dir = ray.transmittance = ray.totalLight = splat3(0.0f);
    // Sample chunks (small tiles, P.chunks > 1).  A pixel's samples are one sequential chain, but the chain may change lanes between
    // samples: job j is chunk j / totalJobs of pixel job j % totalJobs — samples [c R / C, (c + 1) R / C) — so every chunk 0 is handed
    // out before any chunk 1, and the lane that finishes a chunk leaves the RNG state and the running sum for the lane that takes the
    // next one.  With 2.3 pixels per lane (8 GPUs, 1080p) the last round of whole pixels runs on a quarter-full machine; in chunks of
    // an eighth the same work is 18.2 rounds.  A lane whose predecessor chunk is still running holds its job and idles (it keeps
    // taking part in the warp's collectives); the predecessor was handed out a whole pass over the tile earlier, so that is rare.
    // (CHUNKED is a template parameter: with the hand-off code merely present the whole-pixel kernel ran 7 % slower on config 2)
    const int chunks = CHUNKED && P.chunks > 1 && totalJobs > 0u ? P.chunks : 1;
    const unsigned int chunkJobs = totalJobs * (unsigned int)chunks;
    int chunk = 0, sampleEnd = P.NumRaysPerPixel; unsigned int pixJob = 0; bool waiting = false;

    for (;;)
    {
        // [refill] lanes without a pixel take the next jobs (ballot + one atomic per warp)
        const bool need = !havePixel && !exhausted;
        const unsigned int needMask = __ballot_sync(0xffffffffu, need);
        if (needMask)
        {
            unsigned int base = 0;
            const int leader = __ffs(needMask) - 1;
            if ((int)lane == leader) base = atomicAdd(P.workCounter, (unsigned int)__popc(needMask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (need)
            {
                const unsigned int jobAll = base + (unsigned int)__popc(needMask & laneMaskLt);
                if (jobAll >= chunkJobs) exhausted = true;
                else
                {
                    const unsigned int job = CHUNKED && chunks > 1 ? jobAll % totalJobs : jobAll;
                    const unsigned int tile = job >> 5, l = job & 31u;
                    const unsigned int x = (tile % tilesX) * 8u + (l & 7u);
                    const unsigned int r = (tile / tilesX) * 4u + (l >> 3);
                    if (x < P.limX && r < ownedRows)
                    {
                        const unsigned int y = ((r / (unsigned int)P.bandRows) * (unsigned int)P.tileWorld + (unsigned int)P.tileRank) * (unsigned int)P.bandRows
                                             + (r % (unsigned int)P.bandRows);
                        px = SetupPixel(P, x, y);
                        rngState = px.rngState;
                        totalIncomingLight = splat3(0.0f);
                        sample = 0; sampleEnd = P.NumRaysPerPixel;
                        pixelOffset = (size_t)y * P.W + x;
                        havePixel = true; pathActive = false;
                        if (CHUNKED && chunks > 1)
                        {
                            chunk = (int)(jobAll / totalJobs); pixJob = job;
                            sample = (int)(((long long)chunk * P.NumRaysPerPixel) / chunks);
                            sampleEnd = (int)(((long long)(chunk + 1) * P.NumRaysPerPixel) / chunks);
                            waiting = chunk > 0;          // the state at `sample` comes from the lane that ran the chunk before
                        }
                    }
                }
            }
        }
        if (__ballot_sync(0xffffffffu, havePixel) == 0u)
        {
            if (__ballot_sync(0xffffffffu, !exhausted) == 0u) break;
            continue;
        }
        if (CHUNKED && chunks > 1 && havePixel && waiting)
        {
            // four self-validating words: {tag, value} with tag = this dispatch's serial and the number of chunks completed
            const volatile unsigned long long* h = reinterpret_cast<const volatile unsigned long long*>(P.handoff + 4ull * pixJob);
            const unsigned long long w0 = h[0], w1 = h[1], w2 = h[2], w3 = h[3];
            const unsigned int tag = ((unsigned int)P.chunkSerial << 8) | (unsigned int)chunk;
            if ((unsigned int)(w0 >> 32) == tag && (unsigned int)(w1 >> 32) == tag && (unsigned int)(w2 >> 32) == tag && (unsigned int)(w3 >> 32) == tag)
            {
                rngState = (uint32_t)w0;
                totalIncomingLight = make_f3(__uint_as_float((unsigned int)w1), __uint_as_float((unsigned int)w2), __uint_as_float((unsigned int)w3));
                waiting = false;
            }
        }

        // (interpreter-only schedule profile, see rt_kernel_pool.cuh / tools/simt_schedule_profile.py)
        WAVE_PROF(20, 1); WAVE_PROF_LANES(21, havePixel && !pathActive && sample < P.NumRaysPerPixel); WAVE_PROF_LANES(22, pathActive || (havePixel && sample < P.NumRaysPerPixel));
        // [generate] start the pixel's next sample
        if (havePixel && !(CHUNKED && waiting) && !pathActive && sample < (CHUNKED ? sampleEnd : P.NumRaysPerPixel))
        {
            GenerateCameraRay(P, px, rngState, ray);
            bounce = 0;
            pathActive = true;
        }

        // [intersect] + [shade]
        if (pathActive)
        {
            const Hit hit = Intersect<STATS, EXT, TLAS>(P, smemPairs, smemSpheres, ray.pos, ray.dir, cnt);
            WAVE_PROF_LANE(23, !(hit.dst < inf32())); WAVE_PROF_LANE(24, hit.dst < inf32() && hit.material->flag == RT_MATERIAL_GLASS);
            const bool cont = ShadeSegment(P, hit, ray, rngState);
            bounce++;
            if (!cont || bounce > P.MaxBounceCount)
            {
                totalIncomingLight = totalIncomingLight + ray.totalLight;        // HL:578
                sample++;
                pathActive = false;
            }
        }

        // [write] pixel finished (RC:18-23) — or, in chunks, this part of its chain: leave the state for the next chunk's lane
        if (havePixel && !(CHUNKED && waiting) && !pathActive && sample >= (CHUNKED ? sampleEnd : P.NumRaysPerPixel))
        {
            if (!CHUNKED || sampleEnd >= P.NumRaysPerPixel)
            {
                const f3 pixelCol = totalIncomingLight / __int2float_rn(P.NumRaysPerPixel);
                WritePixel<EXT>(P, pixelOffset, pixelCol.x, pixelCol.y, pixelCol.z);
            }
            else
            {
                volatile unsigned long long* h = reinterpret_cast<volatile unsigned long long*>(P.handoff + 4ull * pixJob);
                const unsigned long long tag = (unsigned long long)(((unsigned int)P.chunkSerial << 8) | (unsigned int)(chunk + 1)) << 32;
                h[0] = tag | rngState; h[1] = tag | __float_as_uint(totalIncomingLight.x);
                h[2] = tag | __float_as_uint(totalIncomingLight.y); h[3] = tag | __float_as_uint(totalIncomingLight.z);
            }
            havePixel = false;
        }
    }

    // ---- counters ---------------------------------------------------------------------------------------------------------------
    unsigned int r = __reduce_add_sync(0xffffffffu, cnt.rays);
    if (lane == 0) atomicAdd(P.counters + 0, (unsigned long long)r);
    if (STATS)
    {
        const unsigned int b = __reduce_add_sync(0xffffffffu, cnt.box), t = __reduce_add_sync(0xffffffffu, cnt.tri), s = __reduce_add_sync(0xffffffffu, cnt.sph), sb = __reduce_add_sync(0xffffffffu, cnt.sbox);
        if (lane == 0) { atomicAdd(P.counters + 1, (unsigned long long)b); atomicAdd(P.counters + 2, (unsigned long long)t); atomicAdd(P.counters + 3, (unsigned long long)s); atomicAdd(P.counters + 4, (unsigned long long)sb); }
    }
}

template <bool STATS, bool EXT, bool CHUNKED = false>
__global__ void __launch_bounds__(WAVE_THREADS, RT_WAVE_MINBLOCKS) k_raytrace_wave(const __grid_constant__ DevParams P, const unsigned int totalJobs,
                                                               const unsigned int tilesX, const unsigned int ownedRows)
{
    wave_body<STATS, EXT, false, CHUNKED>(P, totalJobs, tilesX, ownedRows);
}

// many-model scenes: every extension + the TLAS walk (never instrumented: the counting build walks every model like the reference)
__global__ void __launch_bounds__(WAVE_THREADS, RT_WAVE_MINBLOCKS) k_raytrace_wave_tlas(const __grid_constant__ DevParams P, const unsigned int totalJobs,
                                                                    const unsigned int tilesX, const unsigned int ownedRows)
{
    wave_body<false, true, true, true>(P, totalJobs, tilesX, ownedRows);      // (the many-model kernel carries the chunk hand-off: it is never the Cornell-box kernel)
}

// "gridFit": a pixel's samples cannot be split (one RNG chain), so a persistent lane works through whole pixels.  When the image
// gives every lane only a few (8 GPUs, config 2: 2.3 pixels per lane) the last round runs with mostly empty warps.  With the
// option on, the grid is shrunk so that pixels / lanes is just under a whole number k = ceil(pixels / maxLanes): the same k
// rounds, every one of them with full warps, on fewer resident warps.  Scheduling only; off by default (measured in round 2: 2-20 % slower than the full grid).
inline unsigned int fit_persistent_grid(int enabled, unsigned int grid, unsigned int lanesPerCta, unsigned int totalJobs)
{
    if (!enabled || grid == 0 || totalJobs == 0) return grid;
    const unsigned long long lanes = (unsigned long long)grid * lanesPerCta;
    const unsigned long long k = (totalJobs + lanes - 1) / lanes;                    // rounds at full size
    const unsigned long long fitLanes = (totalJobs + k - 1) / k;                     // lanes that keep k rounds full
    const unsigned long long fit = (fitLanes + lanesPerCta - 1) / lanesPerCta;
    return fit < grid ? (unsigned int)(fit ? fit : 1) : grid;
}

// Sample chunks per pixel for kernel 1 (wave_body): automatic = mesh scenes on tiles where whole pixels quantise badly — more than one and
// fewer than four pixels per resident lane — and then about eight chunk-rounds per launch.  forced > 0 overrides (tests); always <= samples.
inline int wave_chunks(int forced, int numSMs, unsigned long long pixels, int samples, int modelCount)
{
    if (samples < 2) return 1;
    int c = 1;
    if (forced > 0) c = forced;
    else if (forced < 0 && modelCount > 0)
    {
        // measured on rank 0's tile of 8 of the default workload (2.3 pixels per lane; profiles/r02_j_tile_ab_sample_chunks.jsonl): whole pixels
        // 89.3 ms, 4 chunks 79.9, 8: 80.4, 16: 81.9, 32: 84.0.  On the sphere-only Cornell box a sample is too cheap for the hand-off
        // (4.93 ms whole pixels, 5.21 with 2 chunks, 5.32 with 4), so sphere-only scenes keep whole pixels.
        const double lanes = (double)numSMs * RT_WAVE_MINBLOCKS * WAVE_THREADS;
        const double perLane = (double)pixels / lanes;
        if (perLane > 1.0 && perLane < 4.0) { c = (int)(8.0 / perLane + 0.999); if (c > 8) c = 8; }
    }
    if (c > samples) c = samples;
    return c < 1 ? 1 : c;
}

template <bool S, bool X> inline cudaError_t wave_configure_one()
{
    return cudaFuncSetAttribute(k_raytrace_wave<S, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
}
inline cudaError_t wave_configure()
{
    cudaError_t e;
    if ((e = wave_configure_one<false, false>()) != cudaSuccess) return e;
    if ((e = wave_configure_one<true, false>()) != cudaSuccess) return e;
    if ((e = wave_configure_one<false, true>()) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_raytrace_wave_tlas, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_raytrace_wave<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_raytrace_wave<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024)) != cudaSuccess) return e;
    return wave_configure_one<true, true>();
}

template <bool S, bool X, bool T = false, bool C = false> inline cudaError_t wave_launch_one(const DevParams& P, int numSMs, size_t smemBytes, unsigned totalJobs, unsigned tilesX, unsigned ownedRows,
                                                             cudaStream_t stream, cudaEvent_t evA, cudaEvent_t evB)
{
    int ctasPerSM = 0;
    cudaError_t e = T ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSM, k_raytrace_wave_tlas, WAVE_THREADS, smemBytes)
                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSM, k_raytrace_wave<S, X, C>, WAVE_THREADS, smemBytes);
    if (e != cudaSuccess) return e;
    if (ctasPerSM < 1) return cudaErrorInvalidConfiguration;
    unsigned int grid = (unsigned int)(numSMs * ctasPerSM);                  // persistent: a multiple of the SM count
    const unsigned int warpsNeeded = (totalJobs + 31u) / 32u;      // (whole pixels: a chunk of a pixel cannot run beside the chunk before it)
    const unsigned int ctasNeeded = (warpsNeeded + (WAVE_THREADS / 32) - 1) / (WAVE_THREADS / 32);
    if (grid > ctasNeeded) grid = ctasNeeded ? ctasNeeded : 1;
    grid = fit_persistent_grid(P.gridFit, grid, WAVE_THREADS, totalJobs);
    if ((e = cudaMemsetAsync(P.workCounter, 0, sizeof(unsigned int), stream)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(evA, stream)) != cudaSuccess) return e;
    if (T) RT_LAUNCH(grid, WAVE_THREADS, smemBytes, stream, k_raytrace_wave_tlas, P, totalJobs, tilesX, ownedRows);
    else RT_LAUNCH(grid, WAVE_THREADS, smemBytes, stream, RT_K(k_raytrace_wave<S, X, C>), P, totalJobs, tilesX, ownedRows);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return cudaEventRecord(evB, stream);
}

inline cudaError_t wave_launch(const DevParams& P, int numSMs, cudaStream_t stream, cudaEvent_t evA, cudaEvent_t evB)
{
    // rows of this rank's bands inside the dispatched region
    unsigned int ownedRows = 0;
    for (unsigned int y0 = 0, b = 0; y0 < P.limY; y0 += (unsigned int)P.bandRows, b++)
        if ((int)(b % (unsigned int)P.tileWorld) == P.tileRank) ownedRows += (P.limY - y0) < (unsigned int)P.bandRows ? (P.limY - y0) : (unsigned int)P.bandRows;
    const unsigned int tilesX = (P.limX + 7u) / 8u;
    const unsigned int tileRows = (ownedRows + 3u) / 4u;
    const unsigned long long jobs64 = (unsigned long long)tilesX * tileRows * 32ull;
    if (jobs64 >= 0xffff0000ull) return cudaErrorInvalidValue;
    const unsigned int totalJobs = (unsigned int)jobs64;

    const size_t smemBytes = sizeof(WaveSmemHeader) + (size_t)P.smemPairs * sizeof(NodePair) + (size_t)WAVE_MAX_SMEM_SPHERES * sizeof(DevSphere);
    const bool ext = P.nPeers > 0 || P.sphBvh != 0 || P.forceExt != 0 || P.smemPairs > 0;   // extensions (peer stores, sphere accelerator, staged tree tops) compiled into their own instantiation
    // sample chunks (P.chunks > 1, small tiles): the hand-off lives in instantiations of its own (the extension family and the TLAS kernel)
    if (P.chunks > 1 && !(P.tlas && P.modelSkip && !P.countStats))
        return P.countStats ? wave_launch_one<true, true, false, true>(P, numSMs, smemBytes, totalJobs, tilesX, ownedRows, stream, evA, evB)
                            : wave_launch_one<false, true, false, true>(P, numSMs, smemBytes, totalJobs, tilesX, ownedRows, stream, evA, evB);
    if (P.countStats) return ext ? wave_launch_one<true, true>(P, numSMs, smemBytes, totalJobs, tilesX, ownedRows, stream, evA, evB)
                                 : wave_launch_one<true, false>(P, numSMs, smemBytes, totalJobs, tilesX, ownedRows, stream, evA, evB);
    if (P.tlas && P.modelSkip) return wave_launch_one<false, true, true>(P, numSMs, smemBytes, totalJobs, tilesX, ownedRows, stream, evA, evB);
    return ext ? wave_launch_one<false, true>(P, numSMs, smemBytes, totalJobs, tilesX, ownedRows, stream, evA, evB)
               : wave_launch_one<false, false>(P, numSMs, smemBytes, totalJobs, tilesX, ownedRows, stream, evA, evB);
}

} // namespace rtd
