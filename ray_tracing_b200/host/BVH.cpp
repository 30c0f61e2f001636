// BVH.cpp — see BVH.h.  Splitting rules follow Assets/Scripts/Types/BVH.cs:26-318 of the reference.
#include "BVH.h"

#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <stdexcept>

namespace Seb {

namespace {

struct Box
{
    float lo[3], hi[3];
    Box() { for (int a = 0; a < 3; a++) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } }    // float.MaxValue / float.MinValue (BVH.cs:37-42)
    template <class T> void grow(const T& t)
    {
        if (t.minX < lo[0]) lo[0] = t.minX;
        if (t.minY < lo[1]) lo[1] = t.minY;
        if (t.minZ < lo[2]) lo[2] = t.minZ;
        if (t.maxX > hi[0]) hi[0] = t.maxX;
        if (t.maxY > hi[1]) hi[1] = t.maxY;
        if (t.maxZ > hi[2]) hi[2] = t.maxZ;
    }
};

inline float min3(float a, float b, float c) { return a < b ? (a < c ? a : c) : (b < c ? b : c); }   // BVH.cs:488-490
inline float max3(float a, float b, float c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }   // BVH.cs:491-493

inline RtNode makeNode(const Box& b, int start, int count)
{
    RtNode n;
    for (int a = 0; a < 3; a++) { n.boundsMin[a] = b.lo[a]; n.boundsMax[a] = b.hi[a]; }
    n.startIndex = start; n.triangleCount = count;
    return n;
}

} // namespace

void BVH::BuildStats::RecordNode(int depth, bool isLeaf, int triCount)
{
    TotalNodeCount++;
    if (!isLeaf) return;
    LeafNodeCount++;
    LeafDepthSum += depth;
    if (depth < LeafDepthMin) LeafDepthMin = depth;
    if (depth > LeafDepthMax) LeafDepthMax = depth;
    TriangleCount += triCount;
    if (triCount > LeafMaxTriCount) LeafMaxTriCount = triCount;
    if (triCount < LeafMinTriCount) LeafMinTriCount = triCount;
}

std::string BVH::BuildStats::ToString() const
{
    char buf[512];
    const char* q = quality == Quality::Low ? "Low" : quality == Quality::High ? "High" : "Disabled";
    snprintf(buf, sizeof(buf),
             "Time (BVH): %d ms (quality = %s)\nTriangles: %d\nNode Count: %d\nLeaf Count: %d\nLeaf Depth:\n - Min: %d\n - Max: %d\n - Mean: %.4g\n"
             "Leaf Tris:\n - Min: %d\n - Max: %d\n - Mean: %.4g\n",
             TimeMs, q, TriangleCount, TotalNodeCount, LeafNodeCount, LeafDepthMin, LeafDepthMax,
             LeafNodeCount ? LeafDepthSum / (float)LeafNodeCount : 0.0f, LeafMinTriCount, LeafMaxTriCount,
             LeafNodeCount ? TriangleCount / (float)LeafNodeCount : 0.0f);
    return buf;
}

BVH::BVH(const Vector3* verts, int vertCount, const int* indices, int indexCount, const Vector3* normals, Quality q)
    : quality(q)
{
    const auto t0 = std::chrono::steady_clock::now();
    stats.quality = q;
    if (indexCount <= 0 || indexCount % 3 != 0) throw std::invalid_argument("BVH: index count must be a positive multiple of 3");
    for (int i = 0; i < indexCount; i++)
        if (indices[i] < 0 || indices[i] >= vertCount) throw std::invalid_argument("BVH: vertex index out of range");

    const int triCount = indexCount / 3;
    buildTris.resize(triCount);
    Box all;
    for (int i = 0; i < indexCount; i += 3)                       // BVH.cs:44-59
    {
        const Vector3 a = verts[indices[i]], b = verts[indices[i + 1]], c = verts[indices[i + 2]];
        BuildTri t;
        t.cx = (a.x + b.x + c.x) / 3; t.cy = (a.y + b.y + c.y) / 3; t.cz = (a.z + b.z + c.z) / 3;
        t.minX = min3(a.x, b.x, c.x); t.minY = min3(a.y, b.y, c.y); t.minZ = min3(a.z, b.z, c.z);
        t.maxX = max3(a.x, b.x, c.x); t.maxY = max3(a.y, b.y, c.y); t.maxZ = max3(a.z, b.z, c.z);
        t.index = i;
        buildTris[i / 3] = t;
        all.grow(t);
    }

    Nodes.reserve(256);
    AddNode(makeNode(all, -1, -1));                               // BVH.cs:61
    if (quality == Quality::Disabled) { Nodes[0].startIndex = 0; Nodes[0].triangleCount = triCount; }
    else Split(0, 0, triCount, 0);

    Triangles.resize(triCount);                                   // BVH.cs:69-80: leaf order
    for (int i = 0; i < triCount; i++)
    {
        const int base = buildTris[i].index;
        RtTriangle& o = Triangles[i];
        const Vector3* p[3] = {&verts[indices[base]], &verts[indices[base + 1]], &verts[indices[base + 2]]};
        const Vector3* n[3] = {&normals[indices[base]], &normals[indices[base + 1]], &normals[indices[base + 2]]};
        o.posA[0] = p[0]->x; o.posA[1] = p[0]->y; o.posA[2] = p[0]->z;
        o.posB[0] = p[1]->x; o.posB[1] = p[1]->y; o.posB[2] = p[1]->z;
        o.posC[0] = p[2]->x; o.posC[1] = p[2]->y; o.posC[2] = p[2]->z;
        o.normA[0] = n[0]->x; o.normA[1] = n[0]->y; o.normA[2] = n[0]->z;
        o.normB[0] = n[1]->x; o.normB[1] = n[1]->y; o.normB[2] = n[1]->z;
        o.normC[0] = n[2]->x; o.normC[1] = n[2]->y; o.normC[2] = n[2]->z;
    }
    std::vector<BuildTri>().swap(buildTris);
    stats.TimeMs = (int)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
}

int BVH::AddNode(const RtNode& n) { Nodes.push_back(n); return (int)Nodes.size() - 1; }

float BVH::NodeCost(float x, float y, float z, int numTriangles)  // BVH.cs:313-318
{
    if (numTriangles == 0) return 0;
    const float area = x * y + x * z + y * z;
    return area * numTriangles;
}

float BVH::EvaluateSplit(int splitAxis, float splitPos, int start, int count) const   // BVH.cs:253-311
{
    Box left, right;
    int numOnLeft = 0, numOnRight = 0;
    for (int i = start; i < start + count; i++)
    {
        const BuildTri& t = buildTris[i];
        const float c = splitAxis == 0 ? t.cx : splitAxis == 1 ? t.cy : t.cz;
        if (c < splitPos) { left.grow(t); numOnLeft++; }
        else { right.grow(t); numOnRight++; }
    }
    const float costA = NodeCost(left.hi[0] - left.lo[0], left.hi[1] - left.lo[1], left.hi[2] - left.lo[2], numOnLeft);
    const float costB = NodeCost(right.hi[0] - right.lo[0], right.hi[1] - right.lo[1], right.hi[2] - right.lo[2], numOnRight);
    return costA + costB;
}

BVH::SplitChoice BVH::ChooseSplit(const RtNode& node, int start, int count) const     // BVH.cs:183-250
{
    if (count <= 1) return SplitChoice{0, 0.0f, INFINITY};
    const float size[3] = {node.boundsMax[0] - node.boundsMin[0], node.boundsMax[1] - node.boundsMin[1], node.boundsMax[2] - node.boundsMin[2]};

    if (quality == Quality::Low)
    {
        const int axis = size[0] > size[1] && size[0] > size[2] ? 0 : size[1] > size[2] ? 1 : 2;
        const float pos = node.boundsMin[axis] + size[axis] * 0.5f;
        return SplitChoice{axis, pos, EvaluateSplit(axis, pos, start, count)};
    }

    SplitChoice best{0, 0.0f, FLT_MAX};
    const int maxSplitTests = count < 10 ? 3 : 5;
    float maxAxis = size[0];                                      // Mathf.Max(sizeX, sizeY, sizeZ)
    if (size[1] > maxAxis) maxAxis = size[1];
    if (size[2] > maxAxis) maxAxis = size[2];

    for (int axis = 0; axis < 3; axis++)
    {
        const float ratio = size[axis] / maxAxis * maxSplitTests;
        // Mathf.CeilToInt; a NaN ratio (degenerate node, 0/0) ends up at the lower clamp like C#'s (int)NaN = int.MinValue
        int numSplitTests = std::isnan(ratio) ? 1 : (int)std::ceil((double)ratio);
        if (numSplitTests < 1) numSplitTests = 1;
        if (numSplitTests > maxSplitTests) numSplitTests = maxSplitTests;
        for (int i = 0; i < numSplitTests; i++)
        {
            const float splitT = (i + 1) / (numSplitTests + 1.0f);
            const float splitPos = node.boundsMin[axis] + size[axis] * splitT;
            const float cost = EvaluateSplit(axis, splitPos, start, count);
            if (cost < best.cost) { best.cost = cost; best.pos = splitPos; best.axis = axis; }
        }
    }
    return best;
}

void BVH::Split(int parentIndex, int triGlobalStart, int triNum, int depth)           // BVH.cs:89-181
{
    const int MaxDepth = 32;
    const RtNode parent = Nodes[parentIndex];
    const float parentCost = NodeCost(parent.boundsMax[0] - parent.boundsMin[0], parent.boundsMax[1] - parent.boundsMin[1],
                                      parent.boundsMax[2] - parent.boundsMin[2], triNum);
    const SplitChoice split = ChooseSplit(parent, triGlobalStart, triNum);

    if (split.cost < parentCost && depth < MaxDepth)
    {
        Box left, right;
        int numOnLeft = 0;
        for (int i = triGlobalStart; i < triGlobalStart + triNum; i++)
        {
            const BuildTri t = buildTris[i];
            const float c = split.axis == 0 ? t.cx : split.axis == 1 ? t.cy : t.cz;
            if (c < split.pos)
            {
                left.grow(t);
                // swap into the left partition (BVH.cs:138-141)
                const BuildTri other = buildTris[triGlobalStart + numOnLeft];
                buildTris[triGlobalStart + numOnLeft] = t;
                buildTris[i] = other;
                numOnLeft++;
            }
            else right.grow(t);
        }
        const int numOnRight = triNum - numOnLeft;
        const int childIndexLeft = AddNode(makeNode(left, triGlobalStart, 0));
        const int childIndexRight = AddNode(makeNode(right, triGlobalStart + numOnLeft, 0));
        Nodes[parentIndex].startIndex = childIndexLeft;
        stats.RecordNode(depth, false);
        Split(childIndexLeft, triGlobalStart, numOnLeft, depth + 1);
        Split(childIndexRight, triGlobalStart + numOnLeft, numOnRight, depth + 1);
    }
    else
    {
        Nodes[parentIndex].startIndex = triGlobalStart;
        Nodes[parentIndex].triangleCount = triNum;
        stats.RecordNode(depth, true, triNum);
    }
}

} // namespace Seb
