#!/usr/bin/env python
"""Writes tests/fixtures/unity_project/: a small, self-authored Unity project skeleton in the serialization the reference's scenes
use (Unity YAML: GameObject 1 / Transform 4 / Camera 20 / MeshFilter 33 / MonoBehaviour 114 documents; `.obj` + `.meta` assets
resolved by guid) — so that the ingestion path of ray_tracing_b200/unity_scene.py has a scene that travels to
the GPU box.  Shapes follow Assets/Scenes/Glass Dragon.unity:133-171 (manager block) and :1988-2073 (a Model with its MeshFilter);
every number in it is this script's.

Contents: a RayComputeManager block (10 bounces, 1 ray per pixel, sky off), a tagged main camera under a rotated parent, a room of
five scaled built-in Cubes (one checker floor), a built-in Quad emitter, one OBJ mesh ("Blob": a displaced subdivided octahedron,
`v//vn` triangles plus a band of `v/vt/vn` quads) instanced three times — glass, glossy, diffuse — under a rotated and non-uniformly
scaled parent, one inactive Model and one disabled Model component that must be skipped, and a sun transform.

    python tests/fixtures/make_unity_fixture.py        (the output is committed; re-running reproduces it byte for byte)
"""
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "unity_project", "Assets")
MODEL_GUID = "cacf7f4e77ad8814ca6868b309a77322"        # guid of Model.cs in the reference project (the reader keys on it)
MANAGER_GUID = "5a097d4e14022bb47ae63bb730d39172"      # guid of RayComputeManager.cs
BUILTIN = "0000000000000000e000000000000000"
BLOB_GUID = "7e57f1c7a5e44b0b9d2a6c3f18b200aa"


def blob_obj():
    """Subdivided octahedron pushed out by a few sine bumps; faces as v//vn triangles, the equator band as v/vt/vn quads."""
    n = 6
    verts, norms, faces = [], [], []
    rings = 2 * n + 1
    def point(i, j):
        # latitude ring i (0 .. 2n), j around (0 .. 4*max(1, w) - 1), octahedral parametrisation mapped to the sphere
        th = math.pi * i / (2 * n)
        seg = 4 * n
        ph = 2 * math.pi * j / seg
        x, y, z = math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph)
        r = 1.0 + 0.12 * math.sin(3 * ph) * math.sin(2 * th) + 0.08 * math.cos(5 * th)
        return (r * x, r * y, r * z), (x, y, z)
    seg = 4 * n
    index = {}
    for i in range(rings):
        for j in range(seg):
            if (i == 0 or i == rings - 1) and j > 0:
                index[(i, j)] = index[(i, 0)]
                continue
            p, nn = point(i, j)
            index[(i, j)] = len(verts) + 1
            verts.append(p); norms.append(nn)
    lines = ["# Blob.obj - test asset written by tests/fixtures/make_unity_fixture.py", "o Blob"]
    lines += ["v %.6f %.6f %.6f" % p for p in verts]
    lines += ["vt %.4f %.4f" % ((k % 7) / 7.0, (k % 5) / 5.0) for k in range(len(verts))]
    lines += ["vn %.6f %.6f %.6f" % q for q in norms]
    for i in range(rings - 1):
        for j in range(seg):
            a, b, c, d = index[(i, j)], index[(i, (j + 1) % seg)], index[(i + 1, (j + 1) % seg)], index[(i + 1, j)]
            if i == 0:
                faces.append("f %d//%d %d//%d %d//%d" % (a, a, c, c, d, d))
            elif i == rings - 2:
                faces.append("f %d//%d %d//%d %d//%d" % (a, a, b, b, d, d))
            elif i in (n - 1, n):
                faces.append("f %d/%d/%d %d/%d/%d %d/%d/%d %d/%d/%d" % (a, a, a, b, b, b, c, c, c, d, d, d))       # quads with texture coordinates
            else:
                faces.append("f %d//%d %d//%d %d//%d" % (a, a, b, b, c, c))
                faces.append("f %d//%d %d//%d %d//%d" % (a, a, c, c, d, d))
    return "\n".join(lines + faces) + "\n"


class Scene:
    def __init__(self):
        self.docs = []
        self.next_id = 100000

    def fid(self):
        self.next_id += 137
        return self.next_id

    def game_object(self, name, components, active=1, tag="Untagged"):
        go = self.fid()
        comp = "".join("  - component: {fileID: %d}\n" % c for c in components)
        self.docs.append("--- !u!1 &%d\nGameObject:\n  m_ObjectHideFlags: 0\n  serializedVersion: 6\n  m_Component:\n%s  m_Layer: 0\n  m_Name: %s\n"
                         "  m_TagString: %s\n  m_IsActive: %d\n" % (go, comp, name, tag, active))
        return go

    def transform(self, tid, go, pos, quat, scale, father=0, children=()):
        ch = "".join("  - {fileID: %d}\n" % c for c in children) or "  []\n"
        if children:
            ch = "  m_Children:\n" + ch
        else:
            ch = "  m_Children: []\n"
        self.docs.append("--- !u!4 &%d\nTransform:\n  m_ObjectHideFlags: 0\n  m_GameObject: {fileID: %d}\n  serializedVersion: 2\n"
                         "  m_LocalRotation: {x: %.7g, y: %.7g, z: %.7g, w: %.7g}\n  m_LocalPosition: {x: %.7g, y: %.7g, z: %.7g}\n"
                         "  m_LocalScale: {x: %.7g, y: %.7g, z: %.7g}\n  m_ConstrainProportionsScale: 0\n%s  m_Father: {fileID: %d}\n"
                         % ((tid, go) + tuple(quat) + tuple(pos) + tuple(scale) + (ch, father)))

    def mesh_filter(self, fid, go, guid, mesh_file_id):
        self.docs.append("--- !u!33 &%d\nMeshFilter:\n  m_ObjectHideFlags: 0\n  m_GameObject: {fileID: %d}\n  m_Mesh: {fileID: %d, guid: %s, type: %d}\n"
                         % (fid, go, mesh_file_id, guid, 0 if guid == BUILTIN else 3))

    def model(self, fid, go, mat, mesh_filter, enabled=1):
        c = lambda v: "{r: %.7g, g: %.7g, b: %.7g, a: 1}" % tuple(v)
        self.docs.append("--- !u!114 &%d\nMonoBehaviour:\n  m_ObjectHideFlags: 0\n  m_GameObject: {fileID: %d}\n  m_Enabled: %d\n  m_EditorHideFlags: 0\n"
                         "  m_Script: {fileID: 11500000, guid: %s, type: 3}\n  m_Name: \n  m_EditorClassIdentifier: \n  material:\n"
                         "    diffuseCol: %s\n    emissionCol: %s\n    specularCol: %s\n    absorption: %s\n    absorptionMultiplier: %.7g\n"
                         "    emissionStrength: %.7g\n    smoothness: %.7g\n    specularProbability: %.7g\n    ior: %.7g\n    flag: %d\n"
                         "  meshFilter: {fileID: %d}\n  meshRenderer: {fileID: 0}\n  logBVHStats: 0\n  materialObjectID: 0\n"
                         % (fid, go, enabled, MODEL_GUID, c(mat["diffuse"]), c(mat.get("emission", (0, 0, 0))), c(mat.get("specular", (1, 1, 1))),
                            c(mat.get("absorption", (0, 0, 0))), mat.get("absorptionMultiplier", 0), mat.get("emissionStrength", 0), mat.get("smoothness", 0),
                            mat.get("specularProbability", 0), mat.get("ior", 1), mat.get("flag", 0), mesh_filter))

    def add_model(self, name, mesh, mat, pos, quat=(0, 0, 0, 1), scale=(1, 1, 1), father=0, active=1, enabled=1):
        tid, mf, mb = self.fid(), self.fid(), self.fid()
        go = self.game_object(name, [tid, mf, mb], active=active)
        self.transform(tid, go, pos, quat, scale, father)
        self.mesh_filter(mf, go, *mesh)
        self.model(mb, go, mat, mf, enabled)
        return tid

    def text(self):
        return "%YAML 1.1\n%TAG !u! tag:unity3d.com,2011:\n" + "".join(self.docs)


def axis_angle(axis, deg):
    s, c = math.sin(math.radians(deg) / 2), math.cos(math.radians(deg) / 2)
    return (axis[0] * s, axis[1] * s, axis[2] * s, c)


def main():
    os.makedirs(os.path.join(ROOT, "Scenes"), exist_ok=True)
    os.makedirs(os.path.join(ROOT, "Graphics"), exist_ok=True)
    open(os.path.join(ROOT, "Graphics", "Blob.obj"), "w").write(blob_obj())
    open(os.path.join(ROOT, "Graphics", "Blob.obj.meta"), "w").write(
        "fileFormatVersion: 2\nguid: %s\nModelImporter:\n  serializedVersion: 22200\n  tangentSpace:\n    normalSmoothAngle: 60\n    normalImportMode: 0\n" % BLOB_GUID)
    cube, quad, blob = (BUILTIN, 10202), (BUILTIN, 10210), (BLOB_GUID, 4300000)
    S = Scene()
    # manager (Glass Dragon.unity:133-171 shape)
    t_mgr, mb_mgr = S.fid(), S.fid()
    go_mgr = S.game_object("Ray Tracer", [t_mgr, mb_mgr])
    S.transform(t_mgr, go_mgr, (0, 0, 0), (0, 0, 0, 1), (1, 1, 1))
    t_sun = S.fid()
    S.docs.append("--- !u!114 &%d\nMonoBehaviour:\n  m_ObjectHideFlags: 0\n  m_GameObject: {fileID: %d}\n  m_Enabled: 1\n  m_EditorHideFlags: 0\n"
                  "  m_Script: {fileID: 11500000, guid: %s, type: 3}\n  m_Name: \n  m_EditorClassIdentifier: \n  rayTracingEnabled: 1\n  accumulate: 1\n"
                  "  bvhQuality: 1\n  maxBounceCount: 10\n  numRaysPerPixel: 1\n  defocusStrength: 0\n  divergeStrength: 1.5\n  focusDistance: 1\n  useSky: 0\n"
                  "  sunFocus: 400\n  sunIntensity: 8\n  sunColor: {r: 1, g: 0.95, b: 0.9, a: 1}\n  sunTransform: {fileID: %d}\n  screenshotName: fixture\n"
                  "  debugParams: {x: 0, y: 0, z: 0, w: 0}\n  numAccumulatedFrames: 12\n  renderSeed: 20260923\n  screenSize: {x: 480, y: 270}\n"
                  % (mb_mgr, go_mgr, MANAGER_GUID, t_sun))
    go_sun = S.game_object("Sun", [t_sun])
    S.transform(t_sun, go_sun, (0, 5, 0), axis_angle((1, 0, 0), 50), (1, 1, 1))
    # camera under a rotated rig
    t_rig, t_cam, cam = S.fid(), S.fid(), S.fid()
    go_rig = S.game_object("Rig", [t_rig])
    go_cam = S.game_object("Main Camera", [t_cam, cam], tag="MainCamera")
    S.transform(t_rig, go_rig, (0.3, 1.9, -5.4), axis_angle((0, 1, 0), -4), (1, 1, 1), children=[t_cam])
    S.transform(t_cam, go_cam, (0, 0, 0), axis_angle((1, 0, 0), 3), (1, 1, 1), father=t_rig)
    S.docs.append("--- !u!20 &%d\nCamera:\n  m_ObjectHideFlags: 0\n  m_GameObject: {fileID: %d}\n  m_Enabled: 1\n  serializedVersion: 2\n  field of view: 52\n"
                  "  orthographic: 0\n  near clip plane: 0.3\n  far clip plane: 1000\n" % (cam, go_cam))
    # room: scaled cubes (non-uniform scale -> quirk Q6, normals through localToWorld)
    white = dict(diffuse=(0.82, 0.82, 0.8))
    S.add_model("Floor", cube, dict(diffuse=(0.75, 0.75, 0.78), emission=(0.18, 0.2, 0.24), specular=(2.5, 1, 1), flag=1), (0, -0.075, 0), scale=(5.5, 0.15, 5.5))
    S.add_model("Ceiling", cube, white, (0, 4.075, 0), quat=axis_angle((0, 0, 1), 90), scale=(0.15, 5.5, 5.5))
    S.add_model("Back", cube, white, (0, 2, 2.825), scale=(5.5, 4.3, 0.15))
    S.add_model("Left", cube, dict(diffuse=(0.8, 0.22, 0.18)), (-2.825, 2, 0), scale=(0.15, 4.3, 5.5))
    S.add_model("Right", cube, dict(diffuse=(0.2, 0.65, 0.25)), (2.825, 2, 0), scale=(0.15, 4.3, 5.5))
    S.add_model("Light", quad, dict(diffuse=(0, 0, 0), emission=(1, 0.96, 0.9), emissionStrength=14), (0, 3.99, 0.2), quat=axis_angle((1, 0, 0), -90), scale=(1.8, 1.4, 1))
    # the OBJ mesh, three instances under a rotated, non-uniformly scaled parent
    t_group = S.fid()
    go_group = S.game_object("Blobs", [t_group])
    kids = [
        S.add_model("Blob Glass", blob, dict(diffuse=(1, 1, 1), specular=(1, 1, 1), absorption=(0.15, 0.5, 0.35), absorptionMultiplier=1.2, smoothness=0.9,
                                            specularProbability=0.9, ior=1.5, flag=2), (-1.25, 1.05, 0.1), quat=axis_angle((0, 1, 0), 25), scale=(0.9, 0.9, 0.9), father=t_group),
        S.add_model("Blob Glossy", blob, dict(diffuse=(0.9, 0.6, 0.15), smoothness=0.8, specularProbability=0.25), (1.2, 0.85, -0.3),
                    quat=axis_angle((0.6, 0, 0.8), 40), scale=(0.7, 0.75, 0.7), father=t_group),
        S.add_model("Blob Diffuse", blob, dict(diffuse=(0.35, 0.45, 0.85)), (0.1, 0.55, 1.1), scale=(0.5, 0.5, 0.5), father=t_group),
        S.add_model("Blob Hidden", blob, dict(diffuse=(1, 0, 1)), (0, 2.5, 0), father=t_group, active=0),
        S.add_model("Blob Disabled", blob, dict(diffuse=(1, 0, 1)), (0, 2.5, 0.5), father=t_group, enabled=0),
    ]
    S.transform(t_group, go_group, (0, 0.02, 0.2), axis_angle((0, 1, 0), 12), (1.0, 1.1, 1.0), children=kids)
    open(os.path.join(ROOT, "Scenes", "Fixture.unity"), "w").write(S.text())
    print("wrote", ROOT)


if __name__ == "__main__":
    main()
