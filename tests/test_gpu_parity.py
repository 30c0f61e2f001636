"""GPU parity tests (run on the B200 with `-m gpu`): the CUDA path, called through the C-ABI by the host manager,
against the CPU oracle on the same seeded inputs.  The bar is BIT-EXACT float32 output (the arithmetic contract
pins every operation), which is far inside BASELINE.json's "per-pixel RGB within 1e-4"; the tolerance test at the
end states that bound explicitly on the per-frame-averaged RGB.
"""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import CUDA_LIB, GOLDEN, ORACLE_LIB, assert_bit_equal, render
import ray_tracing_b200 as rt
from ray_tracing_b200 import capi, scenes

pytestmark = pytest.mark.gpu

KERNELS = [0, 1, 2]     # 0 = reference-shaped megakernel, 1 = persistent threads, 2 = pooled wavefront (the product default)


def _same_counters(a, b, keys=("rays", "boxTests", "triTests", "sphereTests")):
    for k in keys:
        assert a[k] == b[k], f"{k}: {a[k]} vs {b[k]}"


@pytest.mark.parametrize("kernel", KERNELS)
def test_config1_cornell_bitwise_and_golden(kernel):
    """BASELINE config 1: 9-sphere Cornell box, 256x256, 4 bounces, 1 spp — full frame, two accumulated frames."""
    sc = scenes.cornell_spheres(256, 256, 4, 1)
    fo, ao, so = render(ORACLE_LIB, sc, frames=2, want_stats=True)
    fg, ag, sg = render(CUDA_LIB, sc, frames=2, options={"kernel": kernel, "countStats": 1}, want_stats=True)
    assert_bit_equal(fg, fo, "FrameRender")
    assert_bit_equal(ag, ao, "AccumulatedRender")
    _same_counters(sg, so)
    fix = json.load(open(os.path.join(GOLDEN, "cornell_c1.json")))
    assert hashlib.sha256(ag.tobytes()).hexdigest() == fix["accum_sha256"]
    assert sg["rays"] == fix["rays"]


@pytest.mark.parametrize("kernel", KERNELS)
def test_mesh_scene_bitwise_with_traversal_counters(kernel):
    """BVH traversal (3 models: knot + room + light): same pixels AND the same number of box / triangle tests as the
    reference traversal order (RayCommon.hlsl:254,271) — the wavefront kernel visits exactly the oracle's nodes."""
    sc = scenes.knot_room(160, 90, max_bounces=5, rays_per_pixel=3, nu=200, nv=12)
    fo, ao, so = render(ORACLE_LIB, sc, frames=2, want_stats=True)
    fg, ag, sg = render(CUDA_LIB, sc, frames=2, options={"kernel": kernel, "countStats": 1}, want_stats=True)
    assert_bit_equal(fg, fo, "FrameRender")
    assert_bit_equal(ag, ao, "AccumulatedRender")
    _same_counters(sg, so)
    assert sg["boxTests"] > 0 and sg["triTests"] > 0
    # the non-instrumented build skips models whose padded world box the ray cannot reach: same pixels
    fn, an = render(CUDA_LIB, sc, frames=2, options={"kernel": kernel})
    assert_bit_equal(an, ao, "AccumulatedRender without instrumentation (model skipping active)")


@pytest.mark.parametrize("kernel", KERNELS)
def test_glass_mesh_and_checker_bitwise(kernel):
    """Glass branch (refraction, absorption, no back-face culling) on a mesh + a checkered floor (flag 1)."""
    sc = scenes.knot_room(128, 72, max_bounces=10, rays_per_pixel=2, nu=150, nv=10, glass=True)
    sc.models[1].material["flag"] = scenes.MAT_CHECKER
    sc.models[1].material["emissionCol"] = (0.1, 0.3, 0.1, 1.0)
    fo, ao = render(ORACLE_LIB, sc, frames=1)
    fg, ag = render(CUDA_LIB, sc, frames=1, options={"kernel": kernel})
    assert_bit_equal(fg, fo, "FrameRender")


@pytest.mark.parametrize("kernel", KERNELS)
def test_soup_with_spheres_sky_defocus_bitwise(kernel):
    """Config-5 shape, small: three triangle-soup models + 300 spheres (more than the 256 staged in shared memory),
    sky + sun, depth of field (DefocusStrength) — covers GetEnvironmentLight (pow, smoothstep) and the global-memory
    sphere path."""
    sc = scenes.random_soup(96, 96, max_bounces=6, rays_per_pixel=2, triangles=20000, spheres=300)
    sc.settings.update(defocusStrength=20.0, focusDistance=20.0)
    fo, ao, so = render(ORACLE_LIB, sc, frames=1, want_stats=True)
    fg, ag, sg = render(CUDA_LIB, sc, frames=1, options={"kernel": kernel, "countStats": 1}, want_stats=True)
    assert_bit_equal(fg, fo, "FrameRender")
    # 300 spheres go through the sphere accelerator: same pixels, same mesh traversal, fewer sphere tests than the linear scan
    _same_counters(sg, so, keys=("rays", "boxTests", "triTests"))
    assert 0 < sg["sphereTests"] < so["sphereTests"] and sg["sphereBoxTests"] > 0


@pytest.mark.parametrize("kernel", KERNELS)
def test_sphere_accelerator_keeps_first_index_on_ties(kernel):
    """Large Spheres buffers are searched through a BVH over padded boxes; the winner must be the sphere a linear scan
    with strict `<` keeps: exact duplicates (equal dst) resolve to the LOWEST buffer index, whatever the tree order."""
    rng = np.random.RandomState(3)
    n = 400
    sph = np.zeros(n, dtype=scenes.SPHERE_DTYPE)
    sph["centre"] = rng.uniform(-4, 4, (n, 3)).astype(np.float32) + np.float32([0, 0, 9])
    sph["radius"] = rng.uniform(0.15, 0.6, n).astype(np.float32)
    for i in range(n):
        sph["material"][i] = scenes.material(diffuse=tuple(rng.uniform(0.1, 1, 3)), emission=tuple(rng.uniform(0, 1, 3)), emissionStrength=1.0,
                                             specularProbability=0.0)
    # duplicates: sphere 7 copied to 250 and 399, sphere 300 copied to 11 — different materials, identical geometry
    for src, dst in ((7, 250), (7, 399), (300, 11)):
        sph["centre"][dst] = sph["centre"][src]; sph["radius"][dst] = sph["radius"][src]
    sc = scenes.Scene(name="dups", width=160, height=120, spheres=sph, settings=dict(maxBounceCount=3, numRaysPerPixel=2))
    fo, _ = render(ORACLE_LIB, sc)
    fg, _ = render(CUDA_LIB, sc, options={"kernel": kernel})
    assert_bit_equal(fg, fo, "duplicate spheres")


def test_config5_ten_thousand_spheres_bitwise():
    """BASELINE config 5's Sphere buffer: 10,000 spheres (+ 20k triangles, sky).  Oracle = linear scan of all 10,000 per ray;
    GPU = sphere accelerator.  Bit-identical frames."""
    sc = scenes.random_soup(96, 96, max_bounces=8, rays_per_pixel=2, triangles=20000, spheres=10000)
    fo, _, so = render(ORACLE_LIB, sc, frames=1, want_stats=True)
    fg, _, sg = render(CUDA_LIB, sc, frames=1, options={"countStats": 1}, want_stats=True)
    assert_bit_equal(fg, fo, "10k spheres")
    _same_counters(sg, so, keys=("rays", "boxTests", "triTests"))
    assert so["sphereTests"] == so["rays"] * 10000 and sg["sphereTests"] < so["sphereTests"] // 100


@pytest.mark.parametrize("kernel", [1, 2])
def test_extension_instantiation_alone_is_equivalent(kernel):
    """The <EXT = true> kernel instantiation (used with peers and/or the sphere accelerator) must equal the plain one when no
    extension is active — the case of a multi-GPU job with fused tile exchange on a scene without a large Spheres buffer."""
    for sc in (scenes.knot_room(160, 90, max_bounces=5, rays_per_pixel=2, nu=120, nv=10),
               scenes.random_soup(64, 64, max_bounces=5, rays_per_pixel=2, triangles=5000, spheres=20)):
        fo, ao, so = render(ORACLE_LIB, sc, frames=2, want_stats=True)
        fg, ag, sg = render(CUDA_LIB, sc, frames=2, options={"kernel": kernel, "extInstantiation": 1, "countStats": 1}, want_stats=True)
        assert_bit_equal(ag, ao, f"{sc.name} EXT instantiation, kernel {kernel}")
        _same_counters(sg, so)


@pytest.mark.parametrize("kernel", [1, 2])
def test_many_instanced_models_with_world_bounds_skipping(kernel):
    """Twelve models sharing two meshes (instancing through nodeOffset / triOffset), scattered, rotated, non-uniformly scaled,
    some glass: most rays miss most models, which the non-instrumented kernels skip by their padded world boxes."""
    rng = np.random.RandomState(9)
    meshes = [scenes.knot_mesh(nu=60, nv=8), scenes.room_mesh()]
    models = [scenes.ModelDesc(1, np.eye(4), np.eye(4), scenes.material(diffuse=(0.7, 0.7, 0.7), specularProbability=0.0))]
    for i in range(11):
        l2w, w2l = scenes.trs(position=(rng.uniform(-2, 2), rng.uniform(0.4, 3.4), rng.uniform(-1, 2)), euler_deg=tuple(rng.uniform(0, 360, 3)),
                              scale=tuple(rng.uniform(0.08, 0.22, 3)))
        mat = (scenes.material(flag=scenes.MAT_GLASS, ior=1.4, smoothness=0.9, specularProbability=0.9) if i % 3 == 0 else
               scenes.material(diffuse=tuple(rng.uniform(0.2, 0.9, 3)), emission=(1, 1, 1), emissionStrength=float(i % 4 == 1) * 3.0, specularProbability=0.1))
        models.append(scenes.ModelDesc(0, l2w, w2l, mat))
    sc = scenes.Scene(name="instances", width=160, height=90, meshes=meshes, models=models, cam_local_to_world=scenes.trs(position=(0, 1.9, -5.67))[0],
                      fov=54.5, settings=dict(maxBounceCount=6, numRaysPerPixel=2))
    fo, ao = render(ORACLE_LIB, sc, frames=2)
    fg, ag = render(CUDA_LIB, sc, frames=2, options={"kernel": kernel})
    assert_bit_equal(ag, ao, "instanced models")


def test_shared_memory_staging_does_not_change_results():
    """TMA-staged tree tops: 0, a few, many pairs in shared memory — identical output."""
    sc = scenes.knot_room(96, 54, max_bounces=4, rays_per_pixel=2, nu=200, nv=12)
    ref, _ = render(CUDA_LIB, sc, options={"kernel": 1, "smemNodes": 0})
    for kernel in (1, 2):
        for n in (0, 1, 7, 64, 1024, 3000):
            img, _ = render(CUDA_LIB, sc, options={"kernel": kernel, "smemNodes": n})
            assert_bit_equal(img, ref, f"kernel={kernel} smemNodes={n}")


@pytest.mark.parametrize("slots", [32, 64, 96])
def test_pool_sizes_do_not_change_results(slots):
    """Paths per warp pool of the wavefront kernel: scheduling only, identical output and counters."""
    sc = scenes.knot_room(128, 72, max_bounces=6, rays_per_pixel=3, nu=150, nv=10, glass=True)
    fo, ao, so = render(ORACLE_LIB, sc, frames=1, want_stats=True)
    for sort_rays in (0, 1):
        fg, ag, sg = render(CUDA_LIB, sc, frames=1, options={"kernel": 2, "poolSlots": slots, "countStats": 1, "sortRays": sort_rays, "tailLanes": 4 + 9 * sort_rays},
                            want_stats=True)
        assert_bit_equal(fg, fo, f"poolSlots={slots} sortRays={sort_rays}")
        _same_counters(sg, so)


def test_bvh_quality_modes_on_gpu():
    """Low-quality tree and no tree at all (one big leaf, BVH.cs:62-66) through the same kernels."""
    for q in (0, 2):
        sc = scenes.knot_room(64, 36, max_bounces=3, rays_per_pixel=1, nu=60, nv=8)
        sc.settings["bvhQuality"] = q
        fo, _ = render(ORACLE_LIB, sc)
        for kernel in (1, 2):
            fg, _ = render(CUDA_LIB, sc, options={"kernel": kernel})
            assert_bit_equal(fg, fo, f"bvhQuality={q} kernel={kernel}")


def test_config2_size_sparse_pixels_against_oracle():
    """BASELINE config 2 shape (1920x1080, 8 bounces) at 4 spp: the oracle traces a seeded sparse subset of 4096 pixels
    (exact: pixels are independent and seeded from their global index), compared bitwise with the GPU's full frame."""
    sc = scenes.cornell_spheres(1920, 1080, 8, 4)
    fg, ag = render(CUDA_LIB, sc, frames=1)
    rng = np.random.RandomState(7)
    xy = np.stack([rng.randint(0, 1920, 4096), rng.randint(0, 1080, 4096)], axis=1).astype(np.int32)
    xy[:8] = [[0, 0], [1919, 0], [0, 1079], [1919, 1079], [1918, 1079], [1919, 1078], [960, 540], [1, 0]]   # corners: quirk Q1
    mgr = rt.RayComputeManager(ORACLE_LIB)
    scenes.apply(sc, mgr)
    mgr.OnEnable()
    # uniforms were set by InitFrame inside OnEnable: Frame = 1, as in the GPU's first frame
    L = C.CDLL(ORACLE_LIB)
    out = np.empty((4096, 4), dtype=np.float32)
    rc = L.orRenderPixels(C.c_void_p(mgr.context.handle.value), xy.ctypes.data_as(C.c_void_p), 4096, out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert_bit_equal(fg[xy[:, 1], xy[:, 0]], out, "sparse pixels of the 1080p frame")


def _sparse_oracle(sc, n, seed):
    """Oracle values of n seeded random pixels of the scene's first frame (Frame = 1)."""
    rng = np.random.RandomState(seed)
    xy = np.stack([rng.randint(0, sc.width, n), rng.randint(0, sc.height, n)], axis=1).astype(np.int32)
    mgr = rt.RayComputeManager(ORACLE_LIB)
    scenes.apply(sc, mgr)
    mgr.OnEnable()
    L = C.CDLL(ORACLE_LIB)
    out = np.empty((n, 4), dtype=np.float32)
    assert L.orRenderPixels(C.c_void_p(mgr.context.handle.value), xy.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p)) == 0
    return xy, out


def test_config4_shape_deep_bvh_sparse_pixels():
    """BASELINE config 4 shape: ~871k triangles merged into ONE mesh (1.7M-node BVH), glass, 12 bounces.  The GPU renders
    the frame at reduced resolution; 2048 seeded pixels are checked bitwise against the oracle."""
    sc = scenes.knot_cluster(960, 540, max_bounces=12, rays_per_pixel=1)
    assert sc.triangle_count > 870000
    fg, _ = render(CUDA_LIB, sc, frames=1)
    xy, out = _sparse_oracle(sc, 2048, 21)
    assert_bit_equal(fg[xy[:, 1], xy[:, 0]], out, "sparse pixels, 871k-triangle scene")


def test_config5_shape_million_triangles_sparse_pixels():
    """BASELINE config 5 shape: 1,000,000 random triangles in three models (opaque / glass / emissive) + spheres, sky on,
    16 bounces; 2048 seeded pixels bitwise against the oracle."""
    sc = scenes.random_soup(768, 768, max_bounces=16, rays_per_pixel=1, triangles=1_000_000, spheres=10_000)
    fg, _ = render(CUDA_LIB, sc, frames=1)
    xy, out = _sparse_oracle(sc, 2048, 22)
    assert_bit_equal(fg[xy[:, 1], xy[:, 0]], out, "sparse pixels, 1M-triangle scene")


def test_full_size_properties():
    """Size-independent properties at config-2 size: determinism, megakernel == wavefront kernel, alpha = frame count,
    FrameRender alpha = 1, accumulated = sum of frames."""
    sc = scenes.cornell_spheres(1920, 1080, 8, 2)
    f1, a1 = render(CUDA_LIB, sc, frames=2, options={"kernel": 2})
    f1b, a1b = render(CUDA_LIB, sc, frames=2, options={"kernel": 2})
    assert_bit_equal(a1, a1b, "two runs of the wavefront kernel")
    f0, a0 = render(CUDA_LIB, sc, frames=2, options={"kernel": 0})
    assert_bit_equal(a1, a0, "wavefront vs megakernel")
    fp, ap = render(CUDA_LIB, sc, frames=2, options={"kernel": 1})
    assert_bit_equal(a1, ap, "wavefront vs persistent-threads kernel")
    assert np.all(a1[..., 3] == 2.0) and np.all(f1[..., 3] == 1.0)
    first, _ = render(CUDA_LIB, sc, frames=1)
    assert_bit_equal(a1[..., :3], first[..., :3] + f1[..., :3], "SUM buffer = frame 1 + frame 2 (RayCompute.compute:22)")


def test_row_band_tiles_reassemble_to_the_single_gpu_image():
    """rtSetTile / rtPackTile / rtUnpackTiles: two contexts act as ranks 0 and 1 of a 2-GPU job on one device; the
    concatenation of their TileSend buffers (what ncclAllGather would deliver) unpacks to the 1-GPU image."""
    import torch
    sc = scenes.cornell_spheres(200, 150, 4, 2)
    fref, aref = render(CUDA_LIB, sc, frames=2)
    world, band = 2, 8
    mgrs, sends = [], []
    for r in range(world):
        m = rt.RayComputeManager(CUDA_LIB)
        scenes.apply(sc, m)
        m.context.set_tile(r, world, band)
        m.OnEnable(); m.RenderFrame(); m.RenderFrame()
        ctx = m.context
        ctx.pack_tile(); ctx.synchronize()
        ptr, nbytes = ctx.device_pointer("TileSend")
        t = torch.as_tensor(_Dev(ptr, nbytes // 4), device="cuda:0")
        sends.append(t.clone())
        mgrs.append(m)
    gathered = torch.cat(sends)
    ctx0 = mgrs[0].context
    ptr, nbytes = ctx0.device_pointer("TileRecv")
    recv = torch.as_tensor(_Dev(ptr, nbytes // 4), device="cuda:0")
    recv.copy_(gathered)
    torch.cuda.synchronize()
    ctx0.unpack_tiles(); ctx0.synchronize()
    assert_bit_equal(mgrs[0].raytraceFrameTex, fref, "FrameRender reassembled")
    assert_bit_equal(mgrs[0].accumulatedResult, aref, "AccumulatedRender reassembled")


class _Dev:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


def test_cuda_abi_error_behaviour():
    ctx = capi.RtLib(CUDA_LIB).create(0)
    with pytest.raises(capi.RtError) as e:
        ctx.set_float("Gamma", 2.2)
    assert e.value.code == capi.RT_E_UNKNOWN_NAME
    with pytest.raises(capi.RtError) as e:
        ctx.set_buffer_raw("Triangles", (C.c_char * 144)(), 2, 64)
    assert e.value.code == capi.RT_E_INVALID
    with pytest.raises(capi.RtError) as e:
        ctx.dispatch(0, 1, 1, 1)
    assert e.value.code == capi.RT_E_STATE
    ctx.resize(32, 16)
    with pytest.raises(capi.RtError) as e:
        ctx.readback("AccumulatedRender", np.empty((1, 1, 4), dtype=np.float32))
    assert e.value.code == capi.RT_E_INVALID
    # a model pointing outside the node buffer is refused, not traced
    models = np.zeros(1, dtype=capi.MODEL_DTYPE)
    models["nodeOffset"] = 5
    ctx.set_buffer("ModelInfo", models)
    ctx.set_buffer("Nodes", np.zeros(1, dtype=capi.NODE_DTYPE))
    ctx.set_buffer("Triangles", np.zeros(1, dtype=capi.TRIANGLE_DTYPE))
    ctx.set_int("modelCount", 1)
    ctx.set_int("NumRaysPerPixel", 1)
    with pytest.raises(capi.RtError) as e:
        ctx.dispatch(0, 4, 2, 1)
    assert e.value.code == capi.RT_E_STATE
    ctx.destroy()


def test_nan_pixels_are_nan_on_both_sides():
    """Quirk Q4 territory: whatever makes a pixel NaN in the reference arithmetic (log(0) in Box-Muller, ...) must make it NaN
    on the GPU too.  Forced here with a NaN sun intensity and the sky on: every path that escapes to the sky picks up
    NaN * 0 = NaN (RayCommon.hlsl:179-181), every other path stays finite."""
    sc = scenes.random_soup(48, 48, max_bounces=3, rays_per_pixel=1, triangles=2000, spheres=8)
    sc.settings["sunIntensity"] = float("nan")
    fo, _ = render(ORACLE_LIB, sc)
    assert 0.2 < np.isnan(fo[..., 0]).mean() < 1.0
    for kernel in KERNELS:
        fg, _ = render(CUDA_LIB, sc, options={"kernel": kernel})
        assert np.array_equal(np.isnan(fg), np.isnan(fo)), f"kernel {kernel}"
        assert_bit_equal(fg, fo, f"kernel {kernel}")


def test_stated_tolerance_on_averaged_rgb():
    """BASELINE.json north_star: per-pixel RGB within 1e-4 of the reference at matched seed.  Stated on the
    per-frame-averaged RGB (SUM.rgb / SUM.a); NaN pixels (quirk Q4) must be NaN in both."""
    sc = scenes.knot_room(160, 90, max_bounces=8, rays_per_pixel=4, nu=120, nv=10)
    _, ao = render(ORACLE_LIB, sc, frames=3)
    _, ag = render(CUDA_LIB, sc, frames=3)
    avg_o, avg_g = ao[..., :3] / ao[..., 3:4], ag[..., :3] / ag[..., 3:4]
    assert np.array_equal(np.isnan(avg_o), np.isnan(avg_g))
    assert np.nanmax(np.abs(avg_o - avg_g)) <= 1e-4


def _animated_run(lib, options=None):
    """Three frames with per-frame changes, as the reference's UpdateModels / SetShaderParams allow (RayComputeManager.cs:163-204):
    a model moves and changes material, the sphere list changes, the camera moves, then the screen is resized and the
    accumulation restarted."""
    sc = scenes.knot_room(112, 64, max_bounces=4, rays_per_pixel=2, nu=90, nv=8)
    sc.spheres = np.zeros(3, dtype=scenes.SPHERE_DTYPE)
    for i in range(3):
        sc.spheres[i] = scenes._sphere((-1.5 + 1.5 * i, 0.5, -1.0), 0.45, scenes.material(diffuse=(0.9, 0.4 + 0.2 * i, 0.2), specularProbability=0.0))
    mgr = rt.RayComputeManager(lib)
    scenes.apply(sc, mgr)
    for k, v in (options or {}).items():
        mgr.context.set_option(k, v)
    mgr.OnEnable()
    out = []
    mgr.RenderFrame()
    mgr.set_model_transform(0, *scenes.trs(position=(0.6, 1.4, 0.2), euler_deg=(10, 80, 30), scale=(0.35, 0.5, 0.4)))     # knot moves, non-uniform scale
    mgr.set_model_material(0, scenes.material(flag=scenes.MAT_GLASS, ior=1.45, smoothness=0.9, specularProbability=0.9))
    mgr.RenderFrame()
    sph = sc.spheres[:2].copy()
    sph["centre"][0] = (30.0, 0.5, 40.0)                             # a sphere far outside the old scene bounds
    mgr.set_spheres(sph)
    mgr.set_camera(48.0, scenes.trs(position=(0.4, 2.1, -6.3), euler_deg=(3, -4, 0))[0])
    mgr.maxBounceCount = 6
    mgr.RenderFrame()
    out.append(mgr.accumulatedResult.copy())
    mgr.set_screen(80, 48)                                            # Screen size change: new textures, accumulation restarts
    mgr.ResetAccumulatedRender()
    mgr.RenderFrame()
    out.append(mgr.accumulatedResult.copy())
    mgr.OnDestroy()
    return out


@pytest.mark.parametrize("kernel", KERNELS)
def test_per_frame_updates_models_spheres_camera_resize(kernel):
    ref = _animated_run(ORACLE_LIB)
    got = _animated_run(CUDA_LIB, {"kernel": kernel})
    assert ref[0].shape == (64, 112, 4) and ref[1].shape == (48, 80, 4)
    assert_bit_equal(got[0], ref[0], "after three frames with changing models / spheres / camera")
    assert_bit_equal(got[1], ref[1], "after resize + ResetAccumulatedRender")
    assert np.all(ref[0][..., 3] == 3.0) and np.all(ref[1][..., 3] == 1.0)


def test_sphere_accelerator_with_extreme_radii_and_distant_camera():
    """Padding of the sphere accelerator scales with (D^2 + r^2) / r: mix of tiny and huge spheres seen from far away, where the
    reference's own cancellation error is largest."""
    rng = np.random.RandomState(17)
    n = 1500
    sph = np.zeros(n, dtype=scenes.SPHERE_DTYPE)
    sph["centre"] = rng.uniform(-60, 60, (n, 3)).astype(np.float32)
    sph["radius"] = np.exp(rng.uniform(np.log(0.02), np.log(25.0), n)).astype(np.float32)
    for i in range(n):
        glass = i % 5 == 0
        sph["material"][i] = (scenes.material(flag=scenes.MAT_GLASS, ior=1.5, smoothness=1.0, specularProbability=1.0) if glass else
                              scenes.material(diffuse=tuple(rng.uniform(0.2, 1, 3)), emission=tuple(rng.uniform(0, 1, 3)), emissionStrength=float(i % 7 == 0),
                                              specularProbability=0.2, smoothness=0.5))
    sc = scenes.Scene(name="radii", width=128, height=96, spheres=sph, cam_local_to_world=scenes.trs(position=(0, 0, -260.0))[0], fov=40.0,
                      settings=dict(maxBounceCount=6, numRaysPerPixel=2, useSky=True), sun_forward=(0.2, -0.7, 0.6))
    fo, _ = render(ORACLE_LIB, sc, frames=1)
    for kernel in KERNELS:
        fg, _ = render(CUDA_LIB, sc, frames=1, options={"kernel": kernel})
        assert_bit_equal(fg, fo, f"kernel {kernel}")
