/* rt_types.h — the blittable buffer-element layouts that cross the drop-in boundary.
 *
 * These are the exact byte layouts the reference's C# host uploads with
 * ComputeBuffer.SetData (strides from Marshal.SizeOf, ComputeHelper.cs:56-59) and that its
 * HLSL kernel reads as StructuredBuffers.  Little-endian, tightly packed, FP32 / INT32.
 *
 *   RtNode      32 B   BVH.cs:432-457            == RayCommon.hlsl:87-95   (BVHNode)
 *   RtTriangle  72 B   BVH.cs:579-598            == RayCommon.hlsl:49-53   (Triangle)
 *   RtMaterial  88 B   RayTracingMaterial.cs:14-27 == RayCommon.hlsl:64-76
 *   RtModel    224 B   RayComputeManager.cs:256-263 (MeshInfo) == RayCommon.hlsl:78-85 (Model)
 *   RtSphere   104 B   extension mandated by BASELINE.json north_star ("Sphere structured buffer");
 *                      intersect semantics = RaySphere, RayCommon.hlsl:289-332
 *
 * Matrices are Unity Matrix4x4 = column-major in memory (m00,m10,m20,m30,m01,...), which is also
 * HLSL's default packing; mul(M, v) is the ordinary M·v.
 */
#ifndef RT_TYPES_H
#define RT_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#pragma pack(push, 4)

typedef struct RtNode {
    float   boundsMin[3];
    float   boundsMax[3];
    int32_t startIndex;     /* leaf: first triangle (mesh-relative); inner: first child (mesh-relative), second = +1 */
    int32_t triangleCount;  /* > 0  <=> leaf (RayCommon.hlsl:246) */
} RtNode;

typedef struct RtTriangle {
    float posA[3], posB[3], posC[3];
    float normA[3], normB[3], normC[3];
} RtTriangle;

enum { RT_MATERIAL_DEFAULT = 0, RT_MATERIAL_CHECKERED = 1, RT_MATERIAL_GLASS = 2 }; /* RayTracingMaterial.cs:7-12 */

typedef struct RtMaterial {
    float   diffuseCol[4];
    float   emissionCol[4];
    float   specularCol[4];
    float   absorption[4];
    float   absorptionStrength;   /* C# name: absorptionMultiplier */
    float   emissionStrength;
    float   smoothness;
    float   specularProbability;
    float   ior;
    int32_t flag;
} RtMaterial;

typedef struct RtModel {
    int32_t    nodeOffset;
    int32_t    triOffset;
    float      worldToLocal[16];  /* column-major */
    float      localToWorld[16];  /* column-major */
    RtMaterial material;
} RtModel;

typedef struct RtSphere {
    float      centre[3];
    float      radius;
    RtMaterial material;
} RtSphere;

#pragma pack(pop)

#ifdef __cplusplus
}
static_assert(sizeof(RtNode) == 32, "BVHNode stride");
static_assert(sizeof(RtTriangle) == 72, "Triangle stride");
static_assert(sizeof(RtMaterial) == 88, "RayTracingMaterial stride");
static_assert(sizeof(RtModel) == 224, "Model/MeshInfo stride");
static_assert(sizeof(RtSphere) == 104, "Sphere stride");
#endif

#endif /* RT_TYPES_H */
