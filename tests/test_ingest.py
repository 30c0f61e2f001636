"""Asset ingestion and display — the rows next to the hot path: the OBJ loader that feeds the BVH
builder, and the Display.shader / screenshot path behind the accumulated image."""
import os

import numpy as np
import pytest

from conftest import CUDA_LIB, ORACLE_LIB, render
import ray_tracing_b200 as rt
from ray_tracing_b200 import scenes

OBJ = """# two faces: a quad with v/vt/vn corners and a triangle with v//vn corners
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 2 0 1
vn 0 0 1
vn 0 1 0
vt 0 0
f 1/1/1 2/1/1 3/1/1 4/1/1
f 2//2 5//2 3//2
"""


def test_obj_loader_quads_handedness_and_vertex_split(tmp_path):
    p = tmp_path / "m.obj"
    p.write_text(OBJ)
    v, idx, n = rt.load_obj(str(p), unity_handedness=False)
    assert idx.size == 9                                          # quad -> 2 triangles, + 1
    assert v.shape == (7, 3)                                      # vertices 2 and 3 are used with two different normals
    assert idx[:6].tolist() == [0, 1, 2, 0, 2, 3]                 # fan triangulation
    vu, iu, nu = rt.load_obj(str(p), unity_handedness=True)
    assert np.array_equal(vu[:, 0], -v[:, 0]) and np.array_equal(vu[:, 1:], v[:, 1:])     # X negated
    assert iu[:3].tolist() == [0, 2, 1]                           # winding reversed
    # orientation is preserved by the pair of flips: cross(B-A, C-A) keeps pointing along the (mirrored) authored normal
    a, b, c = vu[iu[0]], vu[iu[1]], vu[iu[2]]
    assert np.dot(np.cross(b - a, c - a), nu[iu[0]]) > 0           # geometric face vector still agrees with the (mirrored) normal
    a0, b0, c0 = v[idx[0]], v[idx[1]], v[idx[2]]
    assert np.dot(np.cross(b0 - a0, c0 - a0), n[idx[0]]) > 0
    tris, nodes, st = rt.build_bvh(vu, iu, nu)
    assert st["TriangleCount"] == 3


def test_obj_loader_without_normals_and_errors(tmp_path):
    p = tmp_path / "n.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    v, idx, n = rt.load_obj(str(p), unity_handedness=False)
    assert np.allclose(n, [[0, 0, 1]] * 3)                        # area-weighted face normal
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(ValueError):
        rt.load_obj(str(bad))
    with pytest.raises(ValueError):
        rt.load_obj(str(tmp_path / "missing.obj"))


@pytest.mark.skipif(not os.path.exists("/root/reference/Assets/Graphics/Dragon_80K.obj"), reason="reference assets not mounted")
def test_reference_dragon_loads_and_traces(oracle_path):
    v, idx, n = rt.load_obj("/root/reference/Assets/Graphics/Dragon_80K.obj")
    assert idx.size // 3 == 87130
    tris, nodes, st = rt.build_bvh(v, idx, n)
    assert st["TriangleCount"] == 87130 and st["LeafDepthMax"] <= 32
    mesh = scenes.MeshDesc(v, idx, n)
    l2w, w2l = scenes.trs(position=(-0.58, 1.37, 0.09), euler_deg=(0, 59.33, 0), scale=(4.98, 4.98, 4.98))   # Glass Dragon.unity:2075-2088
    sc = scenes.Scene(name="dragon", width=48, height=27, meshes=[mesh],
                      models=[scenes.ModelDesc(0, l2w, w2l, scenes.material(emission=(1, 1, 1), emissionStrength=1.0, specularProbability=0.0))],
                      cam_local_to_world=scenes.trs(position=(0, 1.9, -5.67))[0], fov=54.5, settings=dict(maxBounceCount=1))
    frame, _ = render(oracle_path, sc)
    lit = frame[..., :3].sum(-1) > 0
    assert 0.05 < lit.mean() < 0.6                                # the dragon covers part of the view


def test_display_divides_by_frame_and_encodes_srgb(oracle_path):
    sc = scenes.cornell_spheres(32, 24, 3, 2)
    mgr = rt.RayComputeManager(oracle_path)
    scenes.apply(sc, mgr)
    mgr.OnEnable()
    for _ in range(3):
        mgr.RenderFrame()
    acc = mgr.accumulatedResult
    disp = rt.RayTraceDisplay(mgr)
    img = disp.OnRenderImage()
    assert img.shape == (24, 32, 4) and img.dtype == np.uint8
    # reference semantics incl. its off-by-one: sum of 3 frames / numAccumulatedFrames (= 4)   
    lin = np.clip(acc[::-1, :, :3] / 4.0, 0, 1).astype(np.float64)
    srgb = np.where(lin <= 0.0031308, 12.92 * lin, 1.055 * np.power(lin, 1 / 2.4) - 0.055)
    assert np.abs(img[..., :3].astype(np.int32) - np.floor(srgb * 255 + 0.5).astype(np.int32)).max() <= 1
    assert np.all(img[..., 3] == np.floor(np.clip(3 / 4.0, 0, 1) ** (1 / 2.4) * 1.055 * 255 - 0.055 * 255 + 0.5))   # alpha = 3/4
    p = os.path.join(os.path.dirname(oracle_path), "_screenshot_test.png")
    try:
        disp.save_screenshot(p)
        assert open(p, "rb").read(8) == b"\x89PNG\r\n\x1a\n"
    finally:
        if os.path.exists(p):
            os.remove(p)


@pytest.mark.gpu
def test_display_matches_oracle_on_gpu():
    sc = scenes.knot_room(96, 54, max_bounces=4, rays_per_pixel=2, nu=80, nv=8)
    outs = []
    for lib in (ORACLE_LIB, CUDA_LIB):
        mgr = rt.RayComputeManager(lib)
        scenes.apply(sc, mgr)
        mgr.OnEnable()
        mgr.RenderFrame(); mgr.RenderFrame()
        outs.append(rt.RayTraceDisplay(mgr).OnRenderImage())
        mgr.accumulate = False
        outs.append(rt.RayTraceDisplay(mgr).OnRenderImage())       # not accumulating: the frame image with Frame = 1
        mgr.OnDestroy()
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[1], outs[3])


# ---- the reference's serialized scenes (Unity YAML) ---------------------------------------------------------------------------

SCENES_DIR = "/root/reference/Assets/Scenes"


def test_builtin_unity_meshes():
    from ray_tracing_b200 import unity_scene
    cube, quad = unity_scene.builtin_cube(), unity_scene.builtin_quad()
    assert cube.triangle_count == 12 and cube.vertices.shape == (24, 3) and np.abs(cube.vertices).max() == 0.5
    for t in range(12):                                            # outward-facing: cross(B-A, C-A) along the vertex normal and away from the centre
        a, b, c = (cube.vertices[cube.indices[3 * t + k]] for k in range(3))
        n = np.cross(b - a, c - a)
        assert np.dot(n, cube.normals[cube.indices[3 * t]]) > 0 and np.dot(n, a + b + c) > 0
    assert quad.triangle_count == 2 and np.allclose(quad.normals, [[0, 0, -1]] * 4)


@pytest.mark.skipif(not os.path.isdir(SCENES_DIR), reason="reference scenes not mounted")
def test_reference_scenes_load_with_the_serialized_settings(oracle_path):
    from ray_tracing_b200 import unity_scene
    expect = {
        "Glass Dragon": dict(models=11, bounces=10, fov=54.5, cam=(0.0, 1.9, -5.67), diverge=1.5, defocus=0.0),
        "Glass Balls": dict(models=17, bounces=10, fov=60.0, cam=(0.0, 1.99, -5.895), diverge=1.5, defocus=0.0),
        "Sphere Refract": dict(models=10, bounces=32, fov=38.0, cam=(0.0, 1.28, -11.65), diverge=1.5, defocus=100.0),
    }
    for name, e in expect.items():
        sc = unity_scene.load_unity_scene(os.path.join(SCENES_DIR, name + ".unity"), width=64, height=36)
        assert len(sc.models) == e["models"] and sc.settings["maxBounceCount"] == e["bounces"] and sc.fov == e["fov"]
        assert np.allclose(sc.cam_local_to_world[:3, 3], e["cam"], atol=1e-3)
        assert sc.settings["divergeStrength"] == e["diverge"] and sc.settings["defocusStrength"] == e["defocus"]
    assert sc.settings["focusDistance"] == pytest.approx(5.3)      # Sphere Refract: depth of field
    # Glass Dragon: the dragon mesh is the shipped OBJ, glass, as serialized (Glass Dragon.unity:1988-2073)
    sc = unity_scene.load_unity_scene(os.path.join(SCENES_DIR, "Glass Dragon.unity"), width=64, height=36)
    dragon = [m for m in sc.models if sc.meshes[m.mesh].triangle_count == 87130]
    assert len(dragon) == 1 and int(dragon[0].material["flag"]) == scenes.MAT_GLASS and float(dragon[0].material["ior"]) == 1.5
    assert np.allclose(np.linalg.norm(dragon[0].local_to_world[:3, 0]), 4.98, atol=1e-2)
    sc.settings["numRaysPerPixel"] = 16
    frame, _ = render(oracle_path, sc)
    assert np.isfinite(frame).all() and (frame[..., :3].sum(-1) > 0).mean() > 0.1
    # scenes that need binary .fbx meshes say so instead of rendering something else
    with pytest.raises(NotImplementedError):
        unity_scene.load_unity_scene(os.path.join(SCENES_DIR, "Text.unity"))
