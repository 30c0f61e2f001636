"""The CUDA kernels' own source, executed on the CPU by the SIMT interpreter of tests/simt (test infrastructure, see
tests/simt/shim/cuda_runtime.h): rt_api.cu with every kernel it includes is compiled by g++ against a stand-in for
<cuda_runtime.h> in which a CTA's threads are fibers and warp collectives really rendezvous.  What runs is the product's
code — the pooled wavefront state machine, ballot compaction, shared-memory pools, the persistent work queue, the repack,
the C-ABI host logic — so these tests check, without a GPU, everything about the kernels except how nvcc lowers them and
how fast they are.  The bodies are the GPU parity tests themselves (tests/test_gpu_parity.py), pointed at the interpreter
build instead of librt_b200.so; the same tests run on the B200 under `-m gpu`.

The product never loads the interpreter build (tests/test_host.py::test_product_does_not_reference_the_simt_build).
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import ORACLE_LIB, REPO, assert_bit_equal, render, seed_forcing_draw as _seed_forcing_draw
import ray_tracing_b200 as rt
from ray_tracing_b200 import scenes
import test_gpu_parity as G
import test_zz_gpu_round2_candidates as Z

sys.path.insert(0, os.path.join(REPO, "tests", "simt"))
import build as simt_build   # noqa: E402


@pytest.fixture(scope="session")
def simt_lib():
    return simt_build.build()


@pytest.fixture(autouse=True)
def _point_parity_tests_at_the_interpreter(monkeypatch, simt_lib):
    monkeypatch.setattr(G, "CUDA_LIB", simt_lib)
    monkeypatch.setattr(Z, "CUDA_LIB", simt_lib)


# Left to the GPU run: the three full-size cases (minutes of interpretation) and the one that needs torch CUDA tensors.
_GPU_ONLY = {"test_config4_shape_deep_bvh_sparse_pixels", "test_config5_shape_million_triangles_sparse_pixels", "test_full_size_properties",
             "test_row_band_tiles_reassemble_to_the_single_gpu_image", "test_config2_size_sparse_pixels_against_oracle"}
for _name in sorted(dir(G)):
    if _name.startswith("test_") and _name not in _GPU_ONLY:
        globals()["test_simt_" + _name[5:]] = getattr(G, _name)


# the TLAS tests of the round-2 GPU file run here too (the other tests of that file have interpreter versions of their own below)
for _name in sorted(dir(Z)):
    if _name.startswith(("test_tlas_", "test_golden_", "test_model_", "test_moving_")):
        globals()["test_simt_" + _name[5:]] = getattr(Z, _name)


def test_simt_row_band_tiles_reassemble(simt_lib):
    """rtSetTile / rtPackTile / rtUnpackTiles with three 'ranks' (bands of 8 rows, 150 rows: ranks own 56 / 48 / 46 rows, so the
    padding of the short ranks is exercised): the concatenation of the TileSend buffers — what the all-gather delivers — unpacks
    to the single-context image.  Device pointers of the interpreter build are host pointers."""
    sc = scenes.cornell_spheres(200, 150, 4, 2)
    fref, aref = render(simt_lib, sc, frames=2)
    world, band = 3, 8
    mgrs, sends = [], []
    for r in range(world):
        m = rt.RayComputeManager(simt_lib)
        scenes.apply(sc, m)
        m.context.set_tile(r, world, band)
        m.OnEnable(); m.RenderFrame(); m.RenderFrame()
        m.context.pack_tile(); m.context.synchronize()
        ptr, nbytes = m.context.device_pointer("TileSend")
        sends.append(C.string_at(ptr, nbytes))
        mgrs.append(m)
    ctx0 = mgrs[0].context
    ptr, nbytes = ctx0.device_pointer("TileRecv")
    blob = b"".join(sends)
    assert len(blob) == nbytes
    C.memmove(ptr, blob, nbytes)
    ctx0.unpack_tiles(); ctx0.synchronize()
    assert_bit_equal(mgrs[0].raytraceFrameTex, fref, "FrameRender reassembled")
    assert_bit_equal(mgrs[0].accumulatedResult, aref, "AccumulatedRender reassembled")
    for m in mgrs:
        m.OnDestroy()


def test_simt_corner_pixels_of_a_ragged_frame(simt_lib):
    """Quirk Q1 (seed aliasing of the last column / row) and partial 8x8 groups at a size that is not a multiple of 8, all kernels."""
    sc = scenes.cornell_spheres(75, 41, 5, 3)
    fo, ao = render(ORACLE_LIB, sc, frames=2)
    for kernel in (0, 1, 2):
        fg, ag = render(simt_lib, sc, frames=2, options={"kernel": kernel})
        assert_bit_equal(ag, ao, f"kernel {kernel}")


def test_simt_tail_lanes_and_ray_sorting_are_schedule_only(simt_lib):
    """Every scheduling knob of the pooled kernel — when the trace phase hands over to shading, the queue order, the pool size,
    how many CTAs / 'SMs' share the work queue — leaves pixels and traversal counters untouched."""
    sc = scenes.knot_room(72, 40, max_bounces=6, rays_per_pixel=2, nu=80, nv=8, glass=True)
    fo, ao, so = render(ORACLE_LIB, sc, frames=1, want_stats=True)
    for tail in (0, 1, 16, 31):
        for sort_rays in (0, 1):
            fg, ag, sg = render(simt_lib, sc, frames=1, options={"kernel": 2, "tailLanes": tail, "sortRays": sort_rays, "countStats": 1, "poolSlots": 32 + 32 * (tail % 3)},
                                want_stats=True)
            assert_bit_equal(fg, fo, f"tailLanes={tail} sortRays={sort_rays}")
            for k in ("rays", "boxTests", "triTests"):
                assert sg[k] == so[k]


def _equal_to_oracle(lib, sc, frames=1, kernels=(0, 1, 2), extra=()):
    fo, ao, so = render(ORACLE_LIB, sc, frames=frames, want_stats=True)
    for k in kernels:
        for opts in ({"kernel": k}, {"kernel": k, "countStats": 1}) + tuple(dict(o, kernel=k) for o in extra):
            fg, ag, sg = render(lib, sc, frames=frames, options=opts, want_stats=True)
            assert_bit_equal(ag, ao, f"{sc.name} {opts}")
            assert_bit_equal(fg, fo, f"{sc.name} frame {opts}")
            if opts.get("countStats"):
                assert all(sg[key] == so[key] for key in ("rays", "boxTests", "triTests")), (sc.name, opts)


@pytest.mark.parametrize("size", [(1, 1), (2, 1), (7, 3), (9, 9), (33, 5), (64, 1), (1, 64)])
def test_simt_tiny_and_ragged_images(simt_lib, size):
    _equal_to_oracle(simt_lib, scenes.cornell_spheres(size[0], size[1], 3, 2), frames=2)
    _equal_to_oracle(simt_lib, scenes.knot_room(size[0], size[1], 3, 2, nu=20, nv=6))


@pytest.mark.parametrize("bounces,spp", [(0, 1), (0, 5), (1, 1), (32, 1), (3, 17)])
def test_simt_bounce_and_sample_extremes(simt_lib, bounces, spp):
    _equal_to_oracle(simt_lib, scenes.cornell_spheres(24, 16, bounces, spp))
    _equal_to_oracle(simt_lib, scenes.knot_room(24, 16, bounces, spp, nu=30, nv=6, glass=True))


@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 255, 256, 257, 300])
def test_simt_sphere_count_thresholds(simt_lib, n):
    """64 / 65: linear scan vs sphere accelerator; 256 / 257: spheres staged in shared memory vs read from global memory."""
    _equal_to_oracle(simt_lib, scenes.random_soup(24, 24, max_bounces=3, rays_per_pixel=1, triangles=300, spheres=n))


def test_simt_degenerate_scenes(simt_lib):
    """Nothing to hit (with and without sky), a mesh whose root is a leaf, 150 instanced models, a huge defocus disc."""
    sc = scenes.cornell_spheres(16, 8, 3, 1); sc.spheres = sc.spheres[:0]
    _equal_to_oracle(simt_lib, sc)
    sc.settings["useSky"] = True
    _equal_to_oracle(simt_lib, sc)
    one = scenes.MeshDesc(np.float32([[-1, 0, 3], [1, 0, 3], [0, 1.5, 3]]), np.int32([0, 1, 2]), np.float32([[0, 0, -1]] * 3))
    mat = scenes.material(diffuse=(0.8, 0.3, 0.3), emission=(1, 1, 1), emissionStrength=2.0)
    _equal_to_oracle(simt_lib, scenes.Scene(name="tri", width=24, height=16, meshes=[one], models=[scenes.ModelDesc(0, np.eye(4), np.eye(4), mat)],
                                            settings=dict(maxBounceCount=2, numRaysPerPixel=2, useSky=True)))
    rng = np.random.RandomState(0)
    models = []
    for i in range(150):
        l2w, w2l = scenes.trs(position=tuple(rng.uniform(-3, 3, 3) + np.array([0, 0, 6])), euler_deg=tuple(rng.uniform(0, 360, 3)), scale=tuple(rng.uniform(0.2, 0.6, 3)))
        models.append(scenes.ModelDesc(i % 2, l2w, w2l, scenes.material(diffuse=tuple(rng.uniform(0.2, 1, 3)), flag=scenes.MAT_GLASS if i % 7 == 0 else 0, ior=1.4,
                                                                         emission=(1, 1, 1), emissionStrength=float(i % 5 == 0))))
    _equal_to_oracle(simt_lib, scenes.Scene(name="many", width=32, height=20, meshes=[scenes.knot_mesh(nu=16, nv=5), one], models=models,
                                            settings=dict(maxBounceCount=4, numRaysPerPixel=1, useSky=True)), kernels=(1, 2))
    sc = scenes.cornell_spheres(20, 12, 3, 2); sc.settings.update(defocusStrength=500.0, focusDistance=3.0, divergeStrength=0.0)
    _equal_to_oracle(simt_lib, sc)
    _equal_to_oracle(simt_lib, scenes.knot_room(40, 24, 5, 2, nu=40, nv=6, glass=True), kernels=(2,),
                     extra=[{"poolSlots": 32, "tailLanes": 31}, {"poolSlots": 96, "tailLanes": 0, "sortRays": 1}, {"smemNodes": 100000}, {"modelSkip": 0}, {"gridFit": 1, "pairOrder": 1}])


def test_simt_random_value_exactly_zero_and_exactly_one(simt_lib):
    """Quirk Q4 made to happen: RandomValue is inclusive at both ends.  PCG's output is 0 only from post-step state 0 and 2^32 - 1 only
    from 515875080 (exhaustive search), and the seed arithmetic is invertible, so a seed can put either value at a chosen draw of a
    chosen pixel.  (a) 0 as the rho draw of the first diffuse hit: log(0) = -inf, rho = +inf, the random direction becomes
    (inf * 0 = NaN, 0, 0), the bounce direction NaN; the ray then misses everything and the sky — whose smoothstep / max are
    NaN-ignoring — answers with the ground colour: a FINITE pixel, equal on both sides.  (b) 1.0 as the roulette draw:
    `rv >= p` ends the path even for p = 1."""
    W, H, x, y = 48, 32, 5, 7
    sc = scenes.cornell_spheres(W, H, 4, 1)
    sc.settings["useSky"] = True
    for post_state, draw in ((0, 7), (515875080, 12)):             # draws 1-4 camera jitter, 5 specular test, 6 theta_x, 7 rho_x, ..., 12 roulette
        sc.settings["renderSeed"] = _seed_forcing_draw(post_state, draw, y * W + x)
        fo, _ = render(ORACLE_LIB, sc)
        if post_state == 0:
            wall = fo[y, x, :3] / np.float32([0.35, 0.3, 0.35])     # ground colour of GetEnvironmentLight times the colour of the wall that was hit
            assert np.all(np.isfinite(fo[y, x])) and np.all(wall > 0) and np.all(wall <= 1.0 + 1e-6)
        for kernel in (0, 1, 2):
            fg, _ = render(simt_lib, sc, options={"kernel": kernel})
            assert_bit_equal(fg, fo, f"forced draw {draw} kernel {kernel}")


@pytest.mark.parametrize("order", ["1", "2"])
def test_simt_results_do_not_depend_on_the_lane_schedule(simt_lib, monkeypatch, order):
    """The interpreter visits the threads of a CTA in descending or freshly shuffled order every round instead of ascending: a
    shared-memory race between lanes that one order happens to hide (reader scheduled after writer) surfaces under another."""
    monkeypatch.setenv("RT_SIMT_ORDER", order)
    for sc in (scenes.cornell_spheres(40, 24, 4, 2), scenes.knot_room(56, 32, max_bounces=5, rays_per_pixel=2, nu=60, nv=8, glass=True),
               scenes.random_soup(32, 32, max_bounces=4, rays_per_pixel=2, triangles=4000, spheres=300)):
        fo, ao, so = render(ORACLE_LIB, sc, frames=2, want_stats=True)
        for opts in ({"kernel": 0}, {"kernel": 1, "smemNodes": 40}, {"kernel": 2, "countStats": 1}, {"kernel": 2, "poolSlots": 32, "sortRays": 1, "smemNodes": 100}):
            fg, ag, sg = render(simt_lib, sc, frames=2, options=opts, want_stats=True)
            assert_bit_equal(ag, ao, f"order {order} {sc.name} {opts}")
    from ray_tracing_b200 import capi
    gpu = capi.RtLib(simt_lib).create(0)
    m = scenes.knot_mesh(nu=60, nv=8)
    th, nh, _ = rt.build_bvh(m.vertices, m.indices, m.normals, 1)
    tg, ng = gpu.build_bvh(m.vertices, m.indices, m.normals, 1)
    assert np.array_equal(ng.view(np.uint8), nh.view(np.uint8)) and np.array_equal(tg.view(np.uint8), th.view(np.uint8))
    gpu.destroy()


def _bvh_meshes():
    rng = np.random.RandomState(5)
    knot = scenes.knot_mesh(nu=160, nv=10)
    shuffled = scenes.MeshDesc(knot.vertices, knot.indices.reshape(-1, 3)[rng.permutation(knot.triangle_count)].reshape(-1), knot.normals)
    soup = max(scenes.random_soup(16, 16, 2, 1, triangles=9000, spheres=1).meshes, key=lambda m: m.triangle_count)
    one = scenes.MeshDesc(np.float32([[0, 0, 0], [1, 0, 0], [0, 1, 0]]), np.int32([0, 1, 2]), np.float32([[0, 0, 1]] * 3))
    flat = scenes.MeshDesc(np.float32([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]]), np.int32([0, 1, 2] * 40 + [1, 3, 2] * 40), np.float32([[0, 0, 1]] * 4))   # 80 coincident triangles: no plane separates them
    return [("knot", knot), ("knot, shuffled triangle order", shuffled), ("soup", soup), ("room", scenes.room_mesh()), ("one triangle", one), ("coincident", flat)]


def test_simt_device_bvh_build_returns_the_reference_builders_buffers(simt_lib):
    """rtBuildBVH (level-synchronous GPU build, csrc/rt_bvh_build.cuh) against the host builder and the oracle's restatement of BVH.cs:
    Nodes and Triangles byte for byte — node order, bounds, leaf ranges AND the triangle order the reference's swap partition
    leaves behind — for all three quality modes, meshes in authored and in shuffled triangle order, degenerate inputs."""
    from ray_tracing_b200 import capi
    gpu, orc = capi.RtLib(simt_lib).create(0), capi.RtLib(ORACLE_LIB).create(0)
    for name, m in _bvh_meshes():
        for q in (1, 0, 2):
            th, nh, _ = rt.build_bvh(m.vertices, m.indices, m.normals, q)
            to, no = orc.build_bvh(m.vertices, m.indices, m.normals, q)
            tg, ng = gpu.build_bvh(m.vertices, m.indices, m.normals, q)
            assert np.array_equal(no.view(np.uint8), nh.view(np.uint8)) and np.array_equal(to.view(np.uint8), th.view(np.uint8)), f"oracle vs host: {name} q={q}"
            assert len(ng) == len(nh) and np.array_equal(ng.view(np.uint8), nh.view(np.uint8)), f"device nodes: {name} q={q}"
            assert np.array_equal(tg.view(np.uint8), th.view(np.uint8)), f"device triangle order: {name} q={q}"
    with pytest.raises(capi.RtError):
        gpu.build_bvh(np.zeros((3, 3), np.float32), np.int32([0, 1, 7]), np.zeros((3, 3), np.float32))      # index out of range
    gpu.destroy(); orc.destroy()


def test_simt_manager_can_build_on_the_device(simt_lib):
    """buildBVHOnDevice: the host manager takes its Nodes / Triangles from rtBuildBVH instead of the host builder — same frame."""
    sc = scenes.knot_room(64, 36, max_bounces=4, rays_per_pixel=2, nu=60, nv=8)
    fo, ao = render(ORACLE_LIB, sc, frames=2)
    mgr = rt.RayComputeManager(simt_lib)
    scenes.apply(sc, mgr)
    mgr.buildBVHOnDevice = 1
    mgr.OnEnable(); mgr.RenderFrame(); mgr.RenderFrame()
    assert_bit_equal(mgr.accumulatedResult, ao, "device-built BVH")
    mgr.OnDestroy()


def test_simt_grid_fit_is_schedule_only(simt_lib, monkeypatch):
    """"gridFit" changes how many persistent CTAs share the pixel queue (a whole number of pixels per lane), nothing else — checked with
    several machine sizes so that the fitted grid is really smaller than the full one (e.g. 3 'SMs' x 2 CTAs x 128 lanes = 768 lanes for
    60 x 33 = 1980 pixels + padding: 3 rounds -> 5 CTAs instead of 6)."""
    scs = [scenes.cornell_spheres(60, 33, 4, 3), scenes.knot_room(60, 33, max_bounces=4, rays_per_pixel=2, nu=60, nv=8)]
    for sms in ("1", "3", "5"):
        monkeypatch.setenv("RT_SIMT_SMS", sms)
        for sc in scs:
            fo, ao = render(ORACLE_LIB, sc, frames=2)
            for kernel in (1, 2):
                fg, ag = render(simt_lib, sc, frames=2, options={"kernel": kernel, "gridFit": 1, "poolSlots": 32})
                assert_bit_equal(ag, ao, f"{sc.name} gridFit kernel={kernel} SMs={sms}")


def test_simt_pair_record_order_is_layout_only(simt_lib):
    """"pairOrder" (treelets in depth-first order instead of breadth-first records): same pixels, same traversal counters, with
    and without tree tops staged in shared memory, shared and unshared meshes."""
    scs = [scenes.knot_room(72, 40, max_bounces=5, rays_per_pixel=2, nu=100, nv=10, glass=True),
           scenes.random_soup(48, 48, max_bounces=5, rays_per_pixel=2, triangles=8000, spheres=12)]
    for sc in scs:
        fo, ao, so = render(ORACLE_LIB, sc, frames=1, want_stats=True)
        for order in (1, 2, 3, 6):
            for kernel, smem in ((2, 0), (2, 300), (1, 0), (1, 64)):
                fg, ag, sg = render(simt_lib, sc, frames=1, options={"kernel": kernel, "pairOrder": order, "smemNodes": smem, "countStats": 1}, want_stats=True)
                assert_bit_equal(fg, fo, f"{sc.name} pairOrder={order} kernel={kernel} smemNodes={smem}")
                assert all(sg[k] == so[k] for k in ("rays", "boxTests", "triTests"))


@pytest.mark.skipif(not os.environ.get("RT_SIMT_VARIANTS"), reason="four extra interpreter builds (~3 min): set RT_SIMT_VARIANTS=1")
@pytest.mark.parametrize("defines", [("RT_STACK_TOP_REG",), ("RT_CACHE_RAYINV",), ("RT_LEAF_REPEAT=2",), ("RT_SPHERE_SKIP_SQRT",), ("RT_TREELET_PREFETCH",), ("RT_SMEM_STACK=2",), ("RT_SMEM_STACK=2", "RT_BRANCHLESS_POP"), ("RT_SMEM_STACK=2", "RT_PUSH_PREDICATED"), ("RT_PUSH_PREDICATED",), ("RT_POOL_COLD_GLOBAL",), ("RT_POOL_COLD_GLOBAL", "RT_SMEM_STACK=16"), ("RT_SMEM_STACK=8", "RT_BRANCHLESS_POP", "RT_LDG256", "RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_LEAF_REPEAT=2", "RT_SPHERE_SAH_DEPTH=14"), ("RT_SKIP_ZERO_DEFOCUS", "RT_GLASS_OUT_OF_LINE"), ("RT_VOTE_WL=3", "RT_VOTE_WN=2"), ("RT_VOTE_WI=2", "RT_VOTE_WL=7", "RT_VOTE_WN=4", "RT_LEAF_REPEAT=2"), ("RT_SMEM_STACK=8", "RT_PREFETCH_CUR"), ("RT_LDG256", "RT_TRI_PAD64"),
                                     ("RT_STACK_TOP_REG", "RT_CACHE_RAYINV", "RT_LEAF_REPEAT=2", "RT_INNER_REPEAT=1")])
def test_simt_compile_time_variants_are_bit_exact(defines, tmp_path):
    """The A/B candidates of tools/round2_sweep.sh change scheduling / instruction selection only: same pixels, same counters."""
    lib = simt_build.build(force=True, defines=defines, out=str(tmp_path / "variant.so"))
    flat = scenes.knot_room(48, 27, max_bounces=3, rays_per_pixel=1, nu=40, nv=6)
    flat.settings["bvhQuality"] = 2
    for sc, frames in ((scenes.knot_room(96, 54, max_bounces=6, rays_per_pixel=2, nu=120, nv=10, glass=True), 2),
                       (scenes.random_soup(64, 64, max_bounces=6, rays_per_pixel=2, triangles=20000, spheres=300), 1),
                       (scenes.cornell_spheres(64, 48, 5, 3), 2), (flat, 1)):
        fo, ao, so = render(ORACLE_LIB, sc, frames=frames, want_stats=True)
        for opts in ({"kernel": 2, "countStats": 1}, {"kernel": 2, "poolSlots": 32, "tailLanes": 3}, {"kernel": 1}, {"kernel": 0}):
            if "RT_TREELET_PREFETCH" in defines:
                opts = dict(opts, treeletPrefetch=1, smemNodes=64)            # flagged treelet roots (the staging request is overridden)
            fg, ag, sg = render(lib, sc, frames=frames, options=opts, want_stats=True)
            assert_bit_equal(ag, ao, f"{defines} {sc.name} {opts}")
            if opts.get("countStats"):
                assert all(sg[k] == so[k] for k in ("rays", "boxTests", "triTests"))
    # the TLAS kernels of the variant (their model step is the one RT_CACHE_RAYINV and the stack variants touch)
    many = scenes.instanced_knots(64, 36, max_bounces=4, rays_per_pixel=2, instances=70)
    fo, ao = render(ORACLE_LIB, many, frames=1)
    for opts in ({"kernel": 2}, {"kernel": 1}):
        if "RT_TREELET_PREFETCH" in defines:
            opts = dict(opts, treeletPrefetch=1)
        fg, ag = render(lib, many, frames=1, options=opts)
        assert_bit_equal(ag, ao, f"{defines} TLAS {opts}")


def test_simt_default_build_refuses_flagged_treelet_roots(simt_lib):
    """The default kernels do not strip the treelet-root flag from a record index, so the option that sets it must be refused."""
    from ray_tracing_b200 import capi
    ctx = capi.RtLib(simt_lib).create(0)
    with pytest.raises(capi.RtError) as e:
        ctx.set_option("treeletPrefetch", 1)
    assert e.value.code == capi.RT_E_INVALID
    ctx.set_option("treeletPrefetch", 0)
    ctx.destroy()
