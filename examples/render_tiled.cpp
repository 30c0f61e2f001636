// render_tiled.cpp — a C++ host rendering ONE image on several GPUs through the C-ABI alone (no Python, no torch in the data plane):
// the image is row-tiled over the GPUs and every RenderFrame ends with one NCCL all-gather of the frame's tiles, issued by
// rtDispatch itself (include/rt_b200.h, "multi-GPU inside the boundary").  Two forms:
//
//   one process, N GPUs (what a Unity host is):      render_tiled <lib> <out.bin> --gpus N [--frames F]
//   N processes, one GPU each (N copies of this):    render_tiled <lib> <out.bin> --rank R --world N --id-file /tmp/id [--frames F]
//        rank 0 draws the NCCL id (rtGetUniqueId) and writes it to the id file, the others wait for the file — any transport
//        would do; RANK / WORLD_SIZE / LOCAL_RANK from the environment are used when the flags are absent (torchrun, mpirun).
//   --gpus 1 (default) is the plain single-GPU run the others must equal byte for byte.
//
//   g++ -std=c++17 -O2 -Iinclude -Iray_tracing_b200/host examples/render_tiled.cpp ray_tracing_b200/host/RayComputeManager.cpp \
//       ray_tracing_b200/host/BVH.cpp -ldl -pthread -o render_tiled
//
// Rank 0 (or the single process) writes the accumulated float4 image to <out.bin> and prints one line with its checksum.
#include "RayComputeManager.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <thread>
#include <vector>

using namespace Seb;

static RtMaterial Material(float r, float g, float b, float specularProbability = 0.0f, float smoothness = 0.0f)
{
    RtMaterial m;
    memset(&m, 0, sizeof(m));
    const float white[4] = {1, 1, 1, 1};
    m.diffuseCol[0] = r; m.diffuseCol[1] = g; m.diffuseCol[2] = b; m.diffuseCol[3] = 1;
    memcpy(m.emissionCol, white, 16); memcpy(m.specularCol, white, 16);
    m.specularProbability = specularProbability; m.smoothness = smoothness; m.ior = 1;
    return m;
}

static RtSphere Sphere(float x, float y, float z, float radius, const RtMaterial& m)
{
    RtSphere s; s.centre[0] = x; s.centre[1] = y; s.centre[2] = z; s.radius = radius; s.material = m; return s;
}

// a torus of nu x nv quads (two triangles each) with analytic normals: the mesh the BVH is built over
static Mesh Torus(int nu, int nv, float R, float r)
{
    Mesh m;
    for (int i = 0; i < nu; i++) for (int j = 0; j < nv; j++)
    {
        const float u = 6.2831853f * (float)i / (float)nu, v = 6.2831853f * (float)j / (float)nv;
        const float cu = cosf(u), su = sinf(u), cv = cosf(v), sv = sinf(v);
        m.vertices.push_back(Vector3{(R + r * cv) * cu, r * sv, (R + r * cv) * su});
        m.normals.push_back(Vector3{cv * cu, sv, cv * su});
    }
    for (int i = 0; i < nu; i++) for (int j = 0; j < nv; j++)
    {
        const int a = i * nv + j, b = ((i + 1) % nu) * nv + j, c = ((i + 1) % nu) * nv + (j + 1) % nv, d = i * nv + (j + 1) % nv;
        const int idx[6] = {a, c, b, a, d, c};
        m.triangles.insert(m.triangles.end(), idx, idx + 6);
    }
    return m;
}

static int argInt(int argc, char** argv, const char* flag, const char* env, int fallback)
{
    for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], flag)) return atoi(argv[i + 1]);
    const char* e = env ? getenv(env) : nullptr;
    return e ? atoi(e) : fallback;
}
static const char* argStr(int argc, char** argv, const char* flag, const char* fallback)
{
    for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], flag)) return argv[i + 1];
    return fallback;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: render_tiled <librt_b200.so> <out.bin> [--gpus N | --rank R --world N --id-file F] [--frames F] [--size WxH]\n"); return 1; }
    const char* backend = argv[1];
    const char* out = argv[2];
    const int gpus = argInt(argc, argv, "--gpus", nullptr, 1);
    const int world = argInt(argc, argv, "--world", "WORLD_SIZE", 1);
    const int rank = argInt(argc, argv, "--rank", "RANK", 0);
    const int local = argInt(argc, argv, "--device", "LOCAL_RANK", rank);
    const int frames = argInt(argc, argv, "--frames", nullptr, 3);
    const char* idFile = argStr(argc, argv, "--id-file", "/tmp/render_tiled.id");
    int W = 640, H = 360;
    sscanf(argStr(argc, argv, "--size", "640x360"), "%dx%d", &W, &H);

    // ---- the context: one GPU, several GPUs of this process, or this process's GPU of a multi-process render ----
    std::vector<int> devices; for (int i = 0; i < gpus; i++) devices.push_back(i);
    RayComputeManager* mgrp = gpus > 1 ? new RayComputeManager(backend, devices.data(), gpus) : new RayComputeManager(backend, world > 1 ? local : 0);
    RayComputeManager& mgr = *mgrp;
    if (!mgr.Context()) { fprintf(stderr, "cannot start the ray tracer: %s\n", mgr.lastError.c_str()); return 2; }
    if (world > 1)
    {
        void* dl = dlopen(backend, RTLD_NOW | RTLD_LOCAL);
        auto getId = reinterpret_cast<decltype(&rtGetUniqueId)>(dl ? dlsym(dl, "rtGetUniqueId") : nullptr);
        auto commInit = reinterpret_cast<decltype(&rtCommInit)>(dl ? dlsym(dl, "rtCommInit") : nullptr);
        auto lastError = reinterpret_cast<decltype(&rtLastError)>(dl ? dlsym(dl, "rtLastError") : nullptr);
        if (!getId || !commInit) { fprintf(stderr, "%s lacks rtGetUniqueId / rtCommInit\n", backend); return 2; }
        unsigned char id[RT_UNIQUE_ID_BYTES];
        if (rank == 0)
        {
            if (getId(id, sizeof(id)) != RT_OK) { fprintf(stderr, "rtGetUniqueId: %s\n", lastError(nullptr)); return 2; }
            const std::string tmp = std::string(idFile) + ".tmp";
            FILE* f = fopen(tmp.c_str(), "wb"); if (!f) return 2;
            fwrite(id, 1, sizeof(id), f); fclose(f);
            rename(tmp.c_str(), idFile);                                  // atomic: readers see the whole id or nothing
        }
        else
        {
            FILE* f = nullptr;
            for (int tries = 0; tries < 3000 && !(f = fopen(idFile, "rb")); tries++) std::this_thread::sleep_for(std::chrono::milliseconds(10));
            if (!f || fread(id, 1, sizeof(id), f) != sizeof(id)) { fprintf(stderr, "rank %d: no id in %s\n", rank, idFile); return 2; }
            fclose(f);
        }
        if (commInit(mgr.Context(), id, sizeof(id), rank, world) != RT_OK) { fprintf(stderr, "rtCommInit: %s\n", lastError(mgr.Context())); return 2; }
    }

    // ---- scene: a glass torus and a diffuse torus over a BVH, a checker floor sphere, a mirror ball, a light; sky on ----
    const std::shared_ptr<Mesh> torus = std::make_shared<Mesh>(Torus(160, 64, 1.0f, 0.35f));      // 20,480 triangles, shared by both models
    auto place = [&](float x, float y, float z, float s, float tiltDeg, const RtMaterial& mat) {
        Model m; m.mesh = torus; m.material = mat;
        const float c = cosf(tiltDeg * 0.01745329f), sn = sinf(tiltDeg * 0.01745329f);
        const float l2w[16] = {s, 0, 0, 0,  0, s * c, s * sn, 0,  0, -s * sn, s * c, 0,  x, y, z, 1};          // column-major: scale, tilt about X, translate
        const float is = 1.0f / s;
        const float w2l[16] = {is, 0, 0, 0,  0, is * c, -is * sn, 0,  0, is * sn, is * c, 0,
                               -is * x, -is * (c * y + sn * z), -is * (-sn * y + c * z), 1};
        memcpy(m.transform.localToWorldMatrix.m, l2w, 64); memcpy(m.transform.worldToLocalMatrix.m, w2l, 64);
        mgr.models.push_back(m);
    };
    RtMaterial glass = Material(1, 1, 1, 1, 1); glass.flag = RT_MATERIAL_GLASS; glass.ior = 1.5f;
    glass.absorption[0] = 0.2f; glass.absorption[1] = 0.6f; glass.absorption[2] = 0.3f; glass.absorption[3] = 1; glass.absorptionStrength = 0.8f;
    place(-1.1f, 1.1f, 0.4f, 1.0f, 35.0f, glass);
    place(1.3f, 0.8f, -0.2f, 0.8f, -60.0f, Material(0.85f, 0.45f, 0.2f, 0.15f, 0.7f));
    RtMaterial light = Material(0, 0, 0); light.emissionStrength = 6;
    RtMaterial floor = Material(0.8f, 0.8f, 0.8f); floor.flag = RT_MATERIAL_CHECKERED; floor.emissionCol[0] = 0.2f; floor.emissionCol[1] = 0.25f; floor.emissionCol[2] = 0.3f; floor.specularCol[0] = 2.0f;
    mgr.spheres = {Sphere(0, -1000, 0, 1000, floor), Sphere(0.1f, 0.5f, 1.6f, 0.5f, Material(0.95f, 0.95f, 0.95f, 1, 1)), Sphere(0, 9, 2, 3, light)};
    mgr.useSky = true;
    mgr.Screen.width = W; mgr.Screen.height = H;
    mgr.mainCamera.fieldOfView = 55; mgr.mainCamera.aspect = (float)W / (float)H;
    const float camL2W[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 1.6f, -4.5f, 1};
    memcpy(mgr.mainCamera.transform.localToWorldMatrix.m, camL2W, 64);
    mgr.maxBounceCount = 6; mgr.numRaysPerPixel = 2; mgr.renderSeed = 12345;

    if (mgr.OnEnable() != RT_OK) { fprintf(stderr, "OnEnable: %s\n", mgr.lastError.c_str()); return 3; }
    const auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; f++)
        if (mgr.RenderFrame() != RT_OK) { fprintf(stderr, "RenderFrame: %s\n", mgr.lastError.c_str()); return 3; }

    std::vector<float> acc((size_t)W * H * 4);
    if (mgr.ReadAccumulated(acc.data(), acc.size() * 4) != RT_OK) { fprintf(stderr, "readback: %s\n", mgr.lastError.c_str()); return 3; }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rank == 0)
    {
        unsigned long long h = 1469598103934665603ull;                        // FNV-1a over the image bytes
        const unsigned char* b = reinterpret_cast<const unsigned char*>(acc.data());
        for (size_t i = 0; i < acc.size() * 4; i++) { h ^= b[i]; h *= 1099511628211ull; }
        FILE* f = fopen(out, "wb"); if (!f) return 4;
        fwrite(acc.data(), 4, acc.size(), f); fclose(f);
        printf("render_tiled: %dx%d frames=%d gpus=%d world=%d alpha=%g fnv1a=%016llx wall_ms=%.1f\n", W, H, frames, gpus, world, acc[3], h, ms);
    }
    delete mgrp;
    return 0;
}
