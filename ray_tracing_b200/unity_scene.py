"""Reader for the reference's serialized scenes (Assets/Scenes/*.unity, Unity YAML) — SURVEY.md §8f #2.

It recovers exactly what RayComputeManager pulls from the Unity scene graph at run time: the manager's inspector
fields (RayComputeManager.cs:9-42), Camera.main (field of view + transform), and every active `Model` component with its
material, its mesh and its transform (RayComputeManager.cs:118,192-204).  Meshes are resolved by asset guid through the
`.meta` files next to the `.obj` assets (loaded with host/ObjLoader.cpp; normals recomputed when the `.meta` says
`normalImportMode: 1`, as for Icosphere.obj) or built here for Unity's built-in Cube / Quad
(fileID 10202 / 10210 of the built-in resources, not part of any repository).  `.fbx` meshes (Text, Water) are read by
fbx_mesh.py (binary FBX container, Unity's handedness / normal import settings) and picked by the sub-asset fileID the
scene stores.  Any other mesh reference raises unless `skip_unsupported=True`.

Not recoverable from the YAML: the run-time instance-ID order of FindObjectsByType (it only decides exact-tie winners
between models, RayCommon.hlsl:362) — file order is used — and `renderSeed`, which the reference re-rolls in OnEnable
(RayComputeManager.cs:64); the serialized value is used so renders are reproducible.
"""
from __future__ import annotations

import os
import re
from typing import Optional

import numpy as np

from . import scenes
from .manager import load_obj
from .fbx_mesh import load_fbx_meshes, _calculated_normals, _meta_settings

MODEL_SCRIPT_GUID = "cacf7f4e77ad8814ca6868b309a77322"       # Assets/Scripts/Types/Model.cs.meta
MANAGER_SCRIPT_GUID = "5a097d4e14022bb47ae63bb730d39172"     # Assets/Scripts/Tracer/RayComputeManager.cs.meta
BUILTIN_GUID = "0000000000000000e000000000000000"
BUILTIN_CUBE, BUILTIN_QUAD = 10202, 10210


def _parse_documents(path: str) -> dict:
    import yaml
    text = open(path, "r", encoding="utf-8").read()
    docs = {}
    for m in re.finditer(r"^--- !u!(\d+) &(-?\d+)( stripped)?\n(.*?)(?=^--- !u!|\Z)", text, flags=re.S | re.M):
        class_id, file_id, body = int(m.group(1)), int(m.group(2)), m.group(4)
        try:
            data = yaml.safe_load(body)
        except yaml.YAMLError:
            continue
        if isinstance(data, dict) and len(data) == 1:
            (kind, fields), = data.items()
            docs[file_id] = (class_id, kind, fields or {})
    return docs


def _quat_to_matrix(q) -> np.ndarray:
    x, y, z, w = (float(q[k]) for k in "xyzw")
    n = x * x + y * y + z * z + w * w
    s = 2.0 / n if n > 0 else 0.0
    return np.array([[1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                     [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                     [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]])


def _vec(d, keys="xyz"):
    return np.array([float(d[k]) for k in keys])


def builtin_cube() -> scenes.MeshDesc:
    """Unity's built-in Cube: unit size, 24 vertices (4 per face, flat normals), 12 triangles facing outwards."""
    faces = []
    for axis in range(3):
        for sign in (-1.0, 1.0):
            n = np.zeros(3); n[axis] = sign
            u = np.zeros(3); u[(axis + 1) % 3] = 1.0
            v = np.cross(n, u)
            p = [0.5 * n - 0.5 * u - 0.5 * v, 0.5 * n + 0.5 * u - 0.5 * v, 0.5 * n + 0.5 * u + 0.5 * v, 0.5 * n - 0.5 * u + 0.5 * v]
            faces.append(scenes.quad_mesh(*p))          # cross(p1-p0, p2-p0) = u x v = n: faces outwards
    return scenes.merge_meshes(faces)


def builtin_quad() -> scenes.MeshDesc:
    """Unity's built-in Quad: unit size in the XY plane, facing -Z."""
    return scenes.quad_mesh((-0.5, -0.5, 0.0), (-0.5, 0.5, 0.0), (0.5, 0.5, 0.0), (0.5, -0.5, 0.0))     # cross = (0, 0, -1)


def _guid_table(graphics_dir: str) -> dict:
    table = {}
    for f in os.listdir(graphics_dir):
        if f.endswith(".meta"):
            m = re.search(r"^guid: ([0-9a-f]{32})", open(os.path.join(graphics_dir, f)).read(), flags=re.M)
            if m:
                table[m.group(1)] = os.path.join(graphics_dir, f[:-5])
    return table


def load_unity_scene(scene_path: str, graphics_dir: Optional[str] = None, width: Optional[int] = None, height: Optional[int] = None,
                     skip_unsupported: bool = False) -> scenes.Scene:
    docs = _parse_documents(scene_path)
    if graphics_dir is None:
        graphics_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(scene_path))), "Graphics")
    guids = _guid_table(graphics_dir) if os.path.isdir(graphics_dir) else {}

    game_objects = {fid: f for fid, (cid, _, f) in docs.items() if cid == 1}
    transforms = {fid: f for fid, (cid, _, f) in docs.items() if cid == 4}
    transform_of_go = {int(f["m_GameObject"]["fileID"]): fid for fid, f in transforms.items()}
    world_cache: dict = {}

    def world_matrix(tid: int) -> np.ndarray:
        if tid in world_cache:
            return world_cache[tid]
        t = transforms[tid]
        local = np.eye(4)
        local[:3, :3] = _quat_to_matrix(t["m_LocalRotation"]) * _vec(t["m_LocalScale"])[None, :]
        local[:3, 3] = _vec(t["m_LocalPosition"])
        parent = int(t.get("m_Father", {}).get("fileID", 0))
        m = world_matrix(parent) @ local if parent in transforms else local
        world_cache[tid] = m
        return m

    def active(go_id: int) -> bool:
        if not int(game_objects.get(go_id, {}).get("m_IsActive", 1)):
            return False
        tid = transform_of_go.get(go_id)
        parent = int(transforms[tid].get("m_Father", {}).get("fileID", 0)) if tid in transforms else 0
        return active(int(transforms[parent]["m_GameObject"]["fileID"])) if parent in transforms else True

    manager = next((f for _, (cid, _, f) in docs.items() if cid == 114 and f.get("m_Script", {}).get("guid") == MANAGER_SCRIPT_GUID), None)
    if manager is None:
        raise ValueError("no RayComputeManager component in " + scene_path)
    cameras = [(fid, f) for fid, (cid, _, f) in docs.items() if cid == 20 and int(f.get("m_Enabled", 1)) and active(int(f["m_GameObject"]["fileID"]))]
    if not cameras:
        raise ValueError("no active camera in " + scene_path)
    main = next((c for c in cameras if game_objects[int(c[1]["m_GameObject"]["fileID"])].get("m_TagString") == "MainCamera"), cameras[0])[1]
    cam_world = world_matrix(transform_of_go[int(main["m_GameObject"]["fileID"])])

    mesh_filters = {int(f["m_GameObject"]["fileID"]): f for _, (cid, _, f) in docs.items() if cid == 33}
    meshes, mesh_index, models = [], {}, []
    fbx_cache: dict = {}
    for fid, (cid, _, f) in docs.items():
        if cid != 114 or f.get("m_Script", {}).get("guid") != MODEL_SCRIPT_GUID or not int(f.get("m_Enabled", 1)):
            continue
        go = int(f["m_GameObject"]["fileID"])
        if not active(go) or go not in mesh_filters:
            continue
        mref = mesh_filters[go]["m_Mesh"]
        key = (str(mref.get("guid", "")), int(mref.get("fileID", 0)))
        if key not in mesh_index:
            if key[0] == BUILTIN_GUID and key[1] == BUILTIN_CUBE:
                mesh = builtin_cube()
            elif key[0] == BUILTIN_GUID and key[1] == BUILTIN_QUAD:
                mesh = builtin_quad()
            elif key[0] in guids and guids[key[0]].lower().endswith(".obj"):
                mesh = scenes.MeshDesc(*load_obj(guids[key[0]]))
                mode, angle = _meta_settings(guids[key[0]])
                if mode == 1:
                    # `normalImportMode: 1` in the .meta (Icosphere.obj): Unity discards the file's normals and recomputes them
                    corners = mesh.vertices[mesh.indices].astype(np.float64)
                    tri = np.arange(len(corners)).reshape(-1, 3)
                    normals = _calculated_normals(corners, tri, angle)
                    mesh = scenes.MeshDesc(np.ascontiguousarray(corners, dtype=np.float32), np.arange(len(corners), dtype=np.int32),
                                           np.ascontiguousarray(normals, dtype=np.float32))
            elif key[0] in guids and guids[key[0]].lower().endswith(".fbx"):
                path = guids[key[0]]
                if path not in fbx_cache:
                    fbx_cache[path] = load_fbx_meshes(path)
                if key[1] not in fbx_cache[path]:
                    raise NotImplementedError(f"mesh fileID {key[1]} of '{game_objects[go].get('m_Name')}' is not a Mesh sub-asset of {path} "
                                              f"(it holds {sorted(n for n, _ in fbx_cache[path].values())})")
                mesh = fbx_cache[path][key[1]][1]
            elif skip_unsupported:
                continue
            else:
                raise NotImplementedError(f"mesh {key} of '{game_objects[go].get('m_Name')}' is not an .obj / built-in Cube / Quad "
                                          f"({guids.get(key[0], 'unknown asset')})")
            mesh_index[key] = len(meshes)
            meshes.append(mesh)
        mat = f["material"]
        col = lambda c: (float(c["r"]), float(c["g"]), float(c["b"]))
        m = scenes.material(diffuse=col(mat["diffuseCol"]), emission=col(mat["emissionCol"]), emissionStrength=float(mat["emissionStrength"]),
                            specular=col(mat["specularCol"]), smoothness=float(mat["smoothness"]), specularProbability=float(mat["specularProbability"]),
                            ior=float(mat["ior"]), flag=int(mat["flag"]), absorption=col(mat["absorption"]),
                            absorptionStrength=float(mat["absorptionMultiplier"]))
        for name in ("diffuseCol", "emissionCol", "specularCol", "absorption"):       # keep the serialized alpha too
            m[name][3] = float(mat[name]["a"])
        l2w = world_matrix(transform_of_go[go])
        models.append(scenes.ModelDesc(mesh_index[key], l2w, np.linalg.inv(l2w), m))

    size = manager.get("screenSize", {"x": 1920, "y": 1080})
    settings = dict(maxBounceCount=int(manager["maxBounceCount"]), numRaysPerPixel=int(manager["numRaysPerPixel"]),
                    defocusStrength=float(manager["defocusStrength"]), divergeStrength=float(manager["divergeStrength"]),
                    focusDistance=float(manager["focusDistance"]), useSky=bool(int(manager["useSky"])), sunFocus=float(manager["sunFocus"]),
                    sunIntensity=float(manager["sunIntensity"]), renderSeed=int(manager["renderSeed"]), accumulate=bool(int(manager["accumulate"])),
                    bvhQuality=int(manager["bvhQuality"]))
    sun_forward = None
    sun_id = int(manager.get("sunTransform", {}).get("fileID", 0))
    if sun_id in transforms:
        sun_forward = tuple(world_matrix(sun_id)[:3, :3] @ np.array([0.0, 0.0, 1.0]))
    cam = np.eye(4)
    cam[:3, :3] = cam_world[:3, :3] / np.linalg.norm(cam_world[:3, :3], axis=0, keepdims=True)     # Camera ignores scale
    cam[:3, 3] = cam_world[:3, 3]
    return scenes.Scene(name=os.path.splitext(os.path.basename(scene_path))[0], width=width or int(size["x"]), height=height or int(size["y"]),
                        meshes=meshes, models=models, cam_local_to_world=cam, fov=float(main["field of view"]), settings=settings,
                        sun_forward=sun_forward)
