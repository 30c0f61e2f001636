"""Deterministic synthetic scenes for the BASELINE.json configurations.

Nothing here reads the reference repository: the GPU box does not have it.  Meshes named after reference
assets that cannot travel (Dragon_80K.obj, 87,130 triangles) are replaced by procedural meshes of the same
size class (a knotted tube of 87,120 triangles), as recorded in DESIGN.md.

A scene is a plain description (spheres, meshes, models, camera, manager settings); `apply()` pushes it
into a RayComputeManager exactly as a Unity scene would populate the reference's manager.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .capi import MATERIAL_DTYPE, SPHERE_DTYPE

MAT_DEFAULT, MAT_CHECKER, MAT_GLASS = 0, 1, 2


def material(diffuse=(1, 1, 1), emission=(1, 1, 1), emissionStrength=0.0, specular=(1, 1, 1), smoothness=0.0,
             specularProbability=1.0, ior=1.0, flag=MAT_DEFAULT, absorption=(0, 0, 0), absorptionStrength=0.0) -> np.ndarray:
    """RayTracingMaterial with the reference's defaults (RayTracingMaterial.cs:29-38)."""
    m = np.zeros((), dtype=MATERIAL_DTYPE)
    m["diffuseCol"] = (*diffuse, 1.0)
    m["emissionCol"] = (*emission, 1.0)
    m["specularCol"] = (*specular, 1.0)
    m["absorption"] = (*absorption, 1.0)
    m["absorptionStrength"] = absorptionStrength
    m["emissionStrength"] = emissionStrength
    m["smoothness"] = smoothness
    m["specularProbability"] = specularProbability
    m["ior"] = ior
    m["flag"] = flag
    return m


def trs(position=(0, 0, 0), euler_deg=(0, 0, 0), scale=(1, 1, 1)):
    """Unity-style TRS (rotation order Z, X, Y like Quaternion.Euler) -> (localToWorld, worldToLocal), float64."""
    ex, ey, ez = np.deg2rad(np.asarray(euler_deg, dtype=np.float64))
    cx, sx, cy, sy, cz, sz = np.cos(ex), np.sin(ex), np.cos(ey), np.sin(ey), np.cos(ez), np.sin(ez)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    r = ry @ rx @ rz
    m = np.eye(4)
    m[:3, :3] = r * np.asarray(scale, dtype=np.float64)[None, :]
    m[:3, 3] = position
    return m, np.linalg.inv(m)


@dataclass
class MeshDesc:
    vertices: np.ndarray      # (n, 3) float32
    indices: np.ndarray       # (3*t,) int32
    normals: np.ndarray       # (n, 3) float32

    @property
    def triangle_count(self) -> int:
        return self.indices.size // 3


@dataclass
class ModelDesc:
    mesh: int
    local_to_world: np.ndarray
    world_to_local: np.ndarray
    material: np.ndarray


@dataclass
class Scene:
    name: str
    width: int
    height: int
    spheres: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=SPHERE_DTYPE))
    meshes: list = field(default_factory=list)
    models: list = field(default_factory=list)
    cam_local_to_world: np.ndarray = field(default_factory=lambda: np.eye(4))
    fov: float = 60.0
    settings: dict = field(default_factory=dict)     # RayComputeManager fields
    sun_forward: Optional[tuple] = None

    @property
    def triangle_count(self) -> int:
        return sum(self.meshes[m.mesh].triangle_count for m in self.models)


def apply(scene: Scene, mgr, width: Optional[int] = None, height: Optional[int] = None) -> None:
    """Populate a RayComputeManager from a scene description (does not call OnEnable)."""
    mgr.set_screen(width or scene.width, height or scene.height)
    mgr.set_camera(scene.fov, scene.cam_local_to_world)
    ids = [mgr.add_mesh(m.vertices, m.indices, m.normals) for m in scene.meshes]
    for md in scene.models:
        mgr.add_model(ids[md.mesh], md.local_to_world, md.world_to_local, md.material)
    mgr.set_spheres(scene.spheres)
    defaults = dict(maxBounceCount=4, numRaysPerPixel=1, defocusStrength=0.0, divergeStrength=0.3, focusDistance=1.0,
                    useSky=False, sunFocus=500.0, sunIntensity=10.0, renderSeed=12345, accumulate=True, bvhQuality=1)
    defaults.update(scene.settings)
    for k, v in defaults.items():
        setattr(mgr, k, v)
    if scene.sun_forward is not None:
        mgr.set_sun((1, 1, 1, 1), scene.sun_forward)


# ---------------------------------------------------------------------------------------------------------------------
# c1 / c2: 9-sphere Cornell box (smallpt layout through the Sphere buffer; wall radius 1e3)
# ---------------------------------------------------------------------------------------------------------------------

def _sphere(centre, radius, mat) -> np.ndarray:
    s = np.zeros((), dtype=SPHERE_DTYPE)
    s["centre"] = centre
    s["radius"] = radius
    s["material"] = mat
    return s


def cornell_spheres(width=256, height=256, max_bounces=4, rays_per_pixel=1) -> Scene:
    R = 1000.0
    grey, red, blue = (0.75, 0.75, 0.75), (0.75, 0.25, 0.25), (0.25, 0.25, 0.75)
    sph = [
        _sphere((-R - 2.0, 2.0, 0.0), R, material(diffuse=red, specularProbability=0.0)),            # left wall   x = -2
        _sphere((R + 2.0, 2.0, 0.0), R, material(diffuse=blue, specularProbability=0.0)),            # right wall  x = +2
        _sphere((0.0, -R, 0.0), R, material(diffuse=grey, specularProbability=0.0)),                 # floor       y = 0
        _sphere((0.0, R + 4.0, 0.0), R, material(diffuse=grey, specularProbability=0.0)),            # ceiling     y = 4
        _sphere((0.0, 2.0, R + 2.0), R, material(diffuse=grey, specularProbability=0.0)),            # back wall   z = +2
        _sphere((0.0, 2.0, -R - 7.0), R, material(diffuse=(0.1, 0.1, 0.1), specularProbability=0.0)),  # wall behind the camera z = -7
        _sphere((-0.9, 0.8, 0.6), 0.8, material(diffuse=(0.95, 0.95, 0.95), smoothness=1.0, specularProbability=1.0)),   # mirror
        _sphere((0.9, 0.8, -0.4), 0.8, material(flag=MAT_GLASS, ior=1.6, smoothness=1.0, specularProbability=1.0,
                                                 absorption=(0.1, 0.4, 0.4), absorptionStrength=0.5)),               # glass
        _sphere((0.0, 13.95, 0.0), 10.0, material(diffuse=(0, 0, 0), emission=(1, 0.95, 0.85), emissionStrength=4.0, specularProbability=0.0)),   # light: cap of radius ~1 poking through the ceiling
    ]
    cam, _ = trs(position=(0.0, 2.0, -5.5))
    return Scene(name="cornell9", width=width, height=height, spheres=np.array(sph, dtype=SPHERE_DTYPE),
                 cam_local_to_world=cam, fov=60.0,
                 settings=dict(maxBounceCount=max_bounces, numRaysPerPixel=rays_per_pixel))


# ---------------------------------------------------------------------------------------------------------------------
# procedural meshes
# ---------------------------------------------------------------------------------------------------------------------

def quad_mesh(p0, p1, p2, p3) -> MeshDesc:
    """Two triangles (p0,p1,p2), (p0,p2,p3); geometric normal = cross(p1-p0, p2-p0) on every vertex."""
    v = np.array([p0, p1, p2, p3], dtype=np.float64)
    n = np.cross(v[1] - v[0], v[2] - v[0])
    n = n / np.linalg.norm(n)
    return MeshDesc(v.astype(np.float32), np.array([0, 1, 2, 0, 2, 3], dtype=np.int32), np.tile(n, (4, 1)).astype(np.float32))


def merge_meshes(meshes) -> MeshDesc:
    vs, ns, idx, base = [], [], [], 0
    for m in meshes:
        vs.append(m.vertices); ns.append(m.normals); idx.append(m.indices + base); base += m.vertices.shape[0]
    return MeshDesc(np.concatenate(vs), np.concatenate(idx).astype(np.int32), np.concatenate(ns))


def transform_mesh(m: MeshDesc, l2w: np.ndarray) -> MeshDesc:
    v = (m.vertices.astype(np.float64) @ l2w[:3, :3].T + l2w[:3, 3]).astype(np.float32)
    nmat = np.linalg.inv(l2w[:3, :3]).T
    n = m.normals.astype(np.float64) @ nmat.T
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return MeshDesc(v, m.indices.copy(), n.astype(np.float32))


def room_mesh(half_width=2.75, height=4.0, depth_front=-3.0, depth_back=2.75) -> MeshDesc:
    """Five inward-facing quads: floor, ceiling, back, left, right (10 triangles), like the reference's Cornell rooms."""
    x0, x1, y0, y1, z0, z1 = -half_width, half_width, 0.0, height, depth_front, depth_back
    quads = [
        quad_mesh((x0, y0, z0), (x0, y0, z1), (x1, y0, z1), (x1, y0, z0)),     # floor, normal +y
        quad_mesh((x0, y1, z0), (x1, y1, z0), (x1, y1, z1), (x0, y1, z1)),     # ceiling, normal -y
        quad_mesh((x0, y0, z1), (x0, y1, z1), (x1, y1, z1), (x1, y0, z1)),     # back wall, normal -z
        quad_mesh((x0, y0, z0), (x0, y1, z0), (x0, y1, z1), (x0, y0, z1)),     # left wall, normal +x
        quad_mesh((x1, y0, z0), (x1, y0, z1), (x1, y1, z1), (x1, y1, z0)),     # right wall, normal -x
    ]
    return merge_meshes(quads)


def knot_mesh(nu=1210, nv=36, p=2, q=3, tube=0.28, bump=0.06) -> MeshDesc:
    """A (p,q) torus-knot tube with a rippled surface: 2*nu*nv triangles (default 87,120), smooth vertex normals.
    Stands in for the reference's Dragon_80K.obj (87,130 triangles): thin, winding, self-occluding geometry that
    produces a deep, irregular BVH."""
    u = np.linspace(0.0, 2.0 * np.pi, nu, endpoint=False)
    def centre(t):
        r = 2.0 + np.cos(q * t)
        return np.stack([r * np.cos(p * t), r * np.sin(p * t), -np.sin(q * t)], axis=-1)
    h = 1e-4
    c = centre(u)
    tangent = centre(u + h) - centre(u - h)
    tangent /= np.linalg.norm(tangent, axis=1, keepdims=True)
    accel = centre(u + h) - 2.0 * c + centre(u - h)
    normal = accel - tangent * np.sum(accel * tangent, axis=1, keepdims=True)
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    binormal = np.cross(tangent, normal)
    v = np.linspace(0.0, 2.0 * np.pi, nv, endpoint=False)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    rad = tube * (1.0 + bump / tube * np.sin(9.0 * uu) * np.cos(4.0 * vv))
    pos = c[:, None, :] + rad[..., None] * (np.cos(vv)[..., None] * normal[:, None, :] + np.sin(vv)[..., None] * binormal[:, None, :])
    verts = pos.reshape(-1, 3)
    iu, iv = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (iu * nv + iv).reshape(-1)
    b = (((iu + 1) % nu) * nv + iv).reshape(-1)
    cidx = (((iu + 1) % nu) * nv + (iv + 1) % nv).reshape(-1)
    d = (iu * nv + (iv + 1) % nv).reshape(-1)
    tris = np.concatenate([np.stack([a, b, cidx], axis=1), np.stack([a, cidx, d], axis=1)], axis=0)
    # orientation: make cross(B-A, C-A) point away from the tube centre line
    fa, fb, fc = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    fn = np.cross(fb - fa, fc - fa)
    outward = fa - c[tris[:, 0] // nv]
    if np.mean(np.sum(fn * outward, axis=1) > 0) < 0.5:
        tris = tris[:, [0, 2, 1]]
        fn = -fn
    normals = np.zeros_like(verts)
    for k in range(3):
        np.add.at(normals, tris[:, k], fn)
    normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    return MeshDesc(verts.astype(np.float32), tris.reshape(-1).astype(np.int32), normals.astype(np.float32))


# ---------------------------------------------------------------------------------------------------------------------
# c3: "~70k-triangle mesh in a Cornell room" — knot (87,120 tris) + room (10) + ceiling light (2); three models
# ---------------------------------------------------------------------------------------------------------------------

def knot_room(width=1920, height=1080, max_bounces=8, rays_per_pixel=1, nu=1210, nv=36, glass=False) -> Scene:
    meshes = [knot_mesh(nu, nv), room_mesh(), quad_mesh((-0.8, 3.98, -0.4), (0.8, 3.98, -0.4), (0.8, 3.98, 1.2), (-0.8, 3.98, 1.2))]
    knot_l2w, knot_w2l = trs(position=(0.0, 1.75, 0.4), euler_deg=(20.0, 35.0, 0.0), scale=(0.5, 0.5, 0.5))
    ident = np.eye(4)
    knot_mat = (material(flag=MAT_GLASS, ior=1.5, smoothness=0.85, specularProbability=0.888,
                         absorption=(0.914, 0.791, 0.247), absorptionStrength=1.5) if glass
                else material(diffuse=(0.85, 0.55, 0.3), smoothness=0.6, specularProbability=0.1))
    models = [
        ModelDesc(0, knot_l2w, knot_w2l, knot_mat),
        ModelDesc(1, ident, ident, material(diffuse=(0.78, 0.78, 0.78), specularProbability=0.0)),
        ModelDesc(2, ident, ident, material(diffuse=(0, 0, 0), emission=(1, 1, 1), emissionStrength=15.0, specularProbability=0.0)),
    ]
    cam, _ = trs(position=(0.0, 1.9, -5.67))
    return Scene(name="knot_room", width=width, height=height, meshes=meshes, models=models, cam_local_to_world=cam, fov=54.5,
                 settings=dict(maxBounceCount=max_bounces, numRaysPerPixel=rays_per_pixel))


def instanced_knots(width=1920, height=1080, max_bounces=8, rays_per_pixel=1, instances=23, seed=9) -> Scene:
    """Many Model components sharing one mesh (instancing through nodeOffset / triOffset, RayComputeManager.cs:209-232) in a
    room: the shape of the reference's shipped scenes (10 - 28 models each)."""
    rng = np.random.RandomState(seed)
    meshes = [knot_mesh(nu=160, nv=24), room_mesh(), quad_mesh((-0.8, 3.98, -0.4), (0.8, 3.98, -0.4), (0.8, 3.98, 1.2), (-0.8, 3.98, 1.2))]
    ident = np.eye(4)
    models = [ModelDesc(1, ident, ident, material(diffuse=(0.78, 0.78, 0.78), specularProbability=0.0)),
              ModelDesc(2, ident, ident, material(diffuse=(0, 0, 0), emission=(1, 1, 1), emissionStrength=15.0, specularProbability=0.0))]
    for i in range(instances):
        l2w, w2l = trs(position=(rng.uniform(-2.2, 2.2), rng.uniform(0.4, 3.4), rng.uniform(-1.5, 2.2)), euler_deg=tuple(rng.uniform(0, 360, 3)),
                       scale=tuple(rng.uniform(0.07, 0.16, 3)))
        mat = (material(flag=MAT_GLASS, ior=1.5, smoothness=0.9, specularProbability=0.9, absorption=(0.6, 0.3, 0.1), absorptionStrength=1.0) if i % 3 == 0 else
               material(diffuse=tuple(rng.uniform(0.2, 0.9, 3)), smoothness=0.5, specularProbability=0.1))
        models.append(ModelDesc(0, l2w, w2l, mat))
    cam, _ = trs(position=(0.0, 1.9, -5.67))
    return Scene(name="instanced_knots", width=width, height=height, meshes=meshes, models=models, cam_local_to_world=cam, fov=54.5,
                 settings=dict(maxBounceCount=max_bounces, numRaysPerPixel=rays_per_pixel))


# ---------------------------------------------------------------------------------------------------------------------
# c4: "~870k triangles, deep BVH" — ten transformed copies of the knot merged into ONE mesh (871,200 tris), glass
# ---------------------------------------------------------------------------------------------------------------------

def knot_cluster(width=3840, height=2160, max_bounces=12, rays_per_pixel=1, copies=10, nu=1210, nv=36) -> Scene:
    base = knot_mesh(nu, nv)
    rng = np.random.RandomState(4)
    parts = []
    for i in range(copies):
        pos = (rng.uniform(-1.8, 1.8), rng.uniform(0.6, 3.2), rng.uniform(-1.0, 1.8))
        rot = tuple(rng.uniform(0.0, 360.0, 3))
        s = rng.uniform(0.18, 0.3)
        l2w, _ = trs(pos, rot, (s, s, s))
        parts.append(transform_mesh(base, l2w))
    meshes = [merge_meshes(parts), room_mesh(), quad_mesh((-0.8, 3.98, -0.4), (0.8, 3.98, -0.4), (0.8, 3.98, 1.2), (-0.8, 3.98, 1.2))]
    ident = np.eye(4)
    models = [
        ModelDesc(0, ident, ident, material(flag=MAT_GLASS, ior=1.5, smoothness=0.85, specularProbability=0.888,
                                            absorption=(0.914, 0.791, 0.247), absorptionStrength=1.5)),
        ModelDesc(1, ident, ident, material(diffuse=(0.78, 0.78, 0.78), specularProbability=0.0)),
        ModelDesc(2, ident, ident, material(diffuse=(0, 0, 0), emission=(1, 1, 1), emissionStrength=15.0, specularProbability=0.0)),
    ]
    cam, _ = trs(position=(0.0, 1.9, -5.67))
    return Scene(name="knot_cluster", width=width, height=height, meshes=meshes, models=models, cam_local_to_world=cam, fov=54.5,
                 settings=dict(maxBounceCount=max_bounces, numRaysPerPixel=rays_per_pixel))


# ---------------------------------------------------------------------------------------------------------------------
# c5: 1M random triangles (3 models: 90 % diffuse, 5 % glass, 5 % emissive) + 10k spheres, sky on
# ---------------------------------------------------------------------------------------------------------------------

def random_soup(width=4096, height=4096, max_bounces=16, rays_per_pixel=1, triangles=1_000_000, spheres=10_000, seed=5) -> Scene:
    rng = np.random.RandomState(seed)

    def soup(n):
        centre = rng.uniform(-10.0, 10.0, (n, 3))
        e1 = rng.uniform(-0.15, 0.15, (n, 3))
        e2 = rng.uniform(-0.15, 0.15, (n, 3))
        v = np.stack([centre, centre + e1, centre + e2], axis=1).reshape(-1, 3)
        fn = np.cross(e1, e2)
        fn /= np.maximum(np.linalg.norm(fn, axis=1, keepdims=True), 1e-20)
        return MeshDesc(v.astype(np.float32), np.arange(3 * n, dtype=np.int32), np.repeat(fn, 3, axis=0).astype(np.float32))

    n_glass = triangles // 20
    n_emit = triangles // 20
    n_diff = triangles - n_glass - n_emit
    meshes = [soup(n_diff), soup(n_glass), soup(n_emit)]
    ident = np.eye(4)
    models = [
        ModelDesc(0, ident, ident, material(diffuse=(0.8, 0.8, 0.8), specularProbability=0.02, smoothness=0.9)),
        ModelDesc(1, ident, ident, material(flag=MAT_GLASS, ior=1.5, smoothness=1.0, specularProbability=1.0)),
        ModelDesc(2, ident, ident, material(diffuse=(0, 0, 0), emission=(1.0, 0.8, 0.6), emissionStrength=4.0)),
    ]
    sph = np.zeros(spheres, dtype=SPHERE_DTYPE)
    if spheres:
        sph["centre"] = rng.uniform(-10.0, 10.0, (spheres, 3)).astype(np.float32)
        sph["radius"] = rng.uniform(0.05, 0.2, spheres).astype(np.float32)
        cols = rng.uniform(0.2, 0.95, (spheres, 3))
        for i in range(spheres):
            sph["material"][i] = material(diffuse=tuple(cols[i]), smoothness=float(i % 3 == 0), specularProbability=0.5)
    cam, _ = trs(position=(0.0, 0.0, -22.0))
    return Scene(name="random_soup", width=width, height=height, spheres=sph, meshes=meshes, models=models,
                 cam_local_to_world=cam, fov=60.0,
                 settings=dict(maxBounceCount=max_bounces, numRaysPerPixel=rays_per_pixel, useSky=True),
                 sun_forward=(0.3, -0.8, 0.5))
