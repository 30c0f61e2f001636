// render_cornell.cpp — a C++ host driving the C-ABI exactly as the reference's C# RayComputeManager drives its
// ComputeShader: build the scene, OnEnable(), RenderFrame() x N, read the accumulated image, show it.
//
//   g++ -std=c++17 -O2 -Iinclude -Iray_tracing_b200/host examples/render_cornell.cpp \
//       ray_tracing_b200/host/RayComputeManager.cpp ray_tracing_b200/host/BVH.cpp -ldl -o render_cornell
//   ./render_cornell ray_tracing_b200/librt_b200.so 8 out.ppm        (backend library, frames, output)
//
// The backend is loaded at run time: librt_b200.so on a B200; the test suite points it at the CPU oracle.
#include "RayComputeManager.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace Seb;

static RtMaterial Material(float r, float g, float b, float specularProbability = 0.0f, float smoothness = 0.0f)
{
    RtMaterial m;
    memset(&m, 0, sizeof(m));
    const float white[4] = {1, 1, 1, 1};
    m.diffuseCol[0] = r; m.diffuseCol[1] = g; m.diffuseCol[2] = b; m.diffuseCol[3] = 1;
    memcpy(m.emissionCol, white, 16); memcpy(m.specularCol, white, 16);       // RayTracingMaterial.SetDefaultValues
    m.specularProbability = specularProbability; m.smoothness = smoothness; m.ior = 1;
    return m;
}

static RtSphere Sphere(float x, float y, float z, float radius, const RtMaterial& m)
{
    RtSphere s; s.centre[0] = x; s.centre[1] = y; s.centre[2] = z; s.radius = radius; s.material = m; return s;
}

int main(int argc, char** argv)
{
    const char* backend = argc > 1 ? argv[1] : "ray_tracing_b200/librt_b200.so";
    const int frames = argc > 2 ? atoi(argv[2]) : 4;
    const char* out = argc > 3 ? argv[3] : nullptr;

    RayComputeManager mgr(backend, 0);
    if (!mgr.Context()) { fprintf(stderr, "cannot start the ray tracer: %s\n", mgr.lastError.c_str()); return 2; }

    // the 9-sphere Cornell box of BASELINE config 1 (same numbers as ray_tracing_b200/scenes.py)
    const float R = 1000.0f;
    RtMaterial light = Material(0, 0, 0); light.emissionCol[0] = 1; light.emissionCol[1] = 0.95f; light.emissionCol[2] = 0.85f; light.emissionStrength = 4;
    RtMaterial glass = Material(1, 1, 1, 1, 1); glass.flag = RT_MATERIAL_GLASS; glass.ior = 1.6f;
    glass.absorption[0] = 0.1f; glass.absorption[1] = 0.4f; glass.absorption[2] = 0.4f; glass.absorption[3] = 1; glass.absorptionStrength = 0.5f;
    mgr.spheres = {
        Sphere(-R - 2, 2, 0, R, Material(0.75f, 0.25f, 0.25f)), Sphere(R + 2, 2, 0, R, Material(0.25f, 0.25f, 0.75f)),
        Sphere(0, -R, 0, R, Material(0.75f, 0.75f, 0.75f)), Sphere(0, R + 4, 0, R, Material(0.75f, 0.75f, 0.75f)),
        Sphere(0, 2, R + 2, R, Material(0.75f, 0.75f, 0.75f)), Sphere(0, 2, -R - 7, R, Material(0.1f, 0.1f, 0.1f)),
        Sphere(-0.9f, 0.8f, 0.6f, 0.8f, Material(0.95f, 0.95f, 0.95f, 1, 1)), Sphere(0.9f, 0.8f, -0.4f, 0.8f, glass),
        Sphere(0, 13.95f, 0, 10, light)};

    mgr.Screen.width = 256; mgr.Screen.height = 256;
    mgr.mainCamera.fieldOfView = 60; mgr.mainCamera.aspect = 1;
    const float camL2W[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 2, -5.5f, 1};       // column-major translation (0, 2, -5.5)
    memcpy(mgr.mainCamera.transform.localToWorldMatrix.m, camL2W, 64);
    mgr.maxBounceCount = 4; mgr.numRaysPerPixel = 1; mgr.renderSeed = 12345;

    if (mgr.OnEnable() != RT_OK) { fprintf(stderr, "OnEnable: %s\n", mgr.lastError.c_str()); return 3; }
    for (int f = 0; f < frames; f++)
        if (mgr.RenderFrame() != RT_OK) { fprintf(stderr, "RenderFrame: %s\n", mgr.lastError.c_str()); return 3; }

    std::vector<float> acc((size_t)256 * 256 * 4);
    if (mgr.ReadAccumulated(acc.data(), acc.size() * 4) != RT_OK) { fprintf(stderr, "readback: %s\n", mgr.lastError.c_str()); return 3; }
    double sum = 0; for (size_t i = 0; i < acc.size(); i += 4) sum += acc[i] + acc[i + 1] + acc[i + 2];
    printf("frames=%d numAccumulatedFrames=%d alpha=%g mean_rgb=%.6f\n", frames, mgr.numAccumulatedFrames, acc[3], sum / (3.0 * 256 * 256 * frames));

    if (out)
    {
        FILE* f = fopen(out, "wb");
        if (!f) return 4;
        fprintf(f, "P6 256 256 255\n");
        for (int y = 255; y >= 0; y--) for (int x = 0; x < 256; x++) for (int c = 0; c < 3; c++)
        {
            const float v = acc[((size_t)y * 256 + x) * 4 + c] / (float)frames;
            fputc((int)(255.0f * powf(fminf(fmaxf(v, 0.0f), 1.0f), 1.0f / 2.2f) + 0.5f), f);
        }
        fclose(f);
    }
    return 0;
}
