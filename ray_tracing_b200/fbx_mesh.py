"""Reader for binary FBX meshes (Assets/Graphics/Text.fbx, Water.fbx of the reference) — SURVEY.md §8f #2, asset ingestion.

What the reference's manager sees of an .fbx is what Unity's ModelImporter turned it into: one `Mesh` per FBX geometry, named
after its node, referenced from the scene by `{fileID, guid}`.  This module reproduces that hand-over:

  * container: "Kaydara FBX Binary" 7.x node records (32-bit offsets below version 7500, 64-bit from 7500), typed scalar
    properties, typed arrays (raw or zlib-deflated), strings; parsed into a plain (name, properties, children) tree;
  * geometry: `Vertices`, `PolygonVertexIndex` (polygon end = bitwise-complemented index), `LayerElementNormal`
    (ByPolygonVertex / ByVertice x Direct / IndexToDirect); polygons with more than three corners are fan-triangulated
    (`keepQuads: 0` in the .meta; both files hold triangles and convex quads only);
  * Unity conventions, as host/ObjLoader.cpp applies them to .obj: X negated on positions and normals and the winding
    reversed (FBX is right-handed, Unity left-handed); no unit scaling (`useFileScale: 0`, `globalScale: 1` in both .meta
    files) — the node's own Lcl Scaling / Rotation live in the scene's Transform, not in the mesh;
  * normals: imported from the file (`normalImportMode: 0`, Water.fbx) or recomputed (`normalImportMode: 1`, Text.fbx):
    per corner, the area-and-angle weighted sum of the face normals around that position whose face is within the
    smoothing angle (`normalSmoothAngle: 60`) of the corner's own face (`normalCalculationMode: 4`).  Unity's exact
    weights are not published; this is the documented behaviour, not a bit-level claim;
  * the scene's `m_Mesh.fileID` of a model-importer sub-asset is XXH64("Type:Mesh->" + meshName + "0") as a signed 64-bit
    integer (`fileIdsGeneration: 2`) — found by matching the ten Text.fbx ids and the Water.fbx id of the shipped scenes
    (tests/test_ingest.py pins them), so meshes are resolved by id, not by guessing from GameObject names.

Pure numpy / zlib; used by unity_scene.load_unity_scene.  Nothing here is on the render path.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

from . import scenes

_M = (1 << 64) - 1
_P1, _P2, _P3, _P4, _P5 = 11400714785074694791, 14029467366897019727, 1609587929392839161, 9650029242287828579, 2870177450012600261


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (64 - r))) & _M


def _round(acc: int, inp: int) -> int:
    return (_rotl((acc + inp * _P2) & _M, 31) * _P1) & _M


def xxh64(data: bytes, seed: int = 0) -> int:
    """XXH64 (public specification), unsigned."""
    n, i = len(data), 0
    if n >= 32:
        v = [(seed + _P1 + _P2) & _M, (seed + _P2) & _M, seed & _M, (seed - _P1) & _M]
        while i <= n - 32:
            for k in range(4):
                v[k] = _round(v[k], int.from_bytes(data[i + 8 * k:i + 8 * k + 8], "little"))
            i += 32
        h = (_rotl(v[0], 1) + _rotl(v[1], 7) + _rotl(v[2], 12) + _rotl(v[3], 18)) & _M
        for k in range(4):
            h = ((h ^ _round(0, v[k])) * _P1 + _P4) & _M
    else:
        h = (seed + _P5) & _M
    h = (h + n) & _M
    while i + 8 <= n:
        h = (_rotl(h ^ _round(0, int.from_bytes(data[i:i + 8], "little")), 27) * _P1 + _P4) & _M
        i += 8
    if i + 4 <= n:
        h = (_rotl(h ^ ((int.from_bytes(data[i:i + 4], "little") * _P1) & _M), 23) * _P2 + _P3) & _M
        i += 4
    while i < n:
        h = (_rotl(h ^ ((data[i] * _P5) & _M), 11) * _P1) & _M
        i += 1
    h ^= h >> 33
    h = (h * _P2) & _M
    h ^= h >> 29
    h = (h * _P3) & _M
    h ^= h >> 32
    return h


def unity_mesh_file_id(mesh_name: str) -> int:
    """fileID Unity's ModelImporter (fileIdsGeneration 2) gives the Mesh sub-asset called `mesh_name`."""
    h = xxh64(("Type:Mesh->" + mesh_name + "0").encode("utf-8"))
    return h - (1 << 64) if h >= (1 << 63) else h


# ---- container ------------------------------------------------------------------------------------------------------------

_SCALAR = {"Y": "<h", "C": "<?", "I": "<i", "F": "<f", "D": "<d", "L": "<q"}
_ARRAY = {"f": "<f4", "d": "<f8", "l": "<i8", "i": "<i4", "b": "u1"}


def parse_fbx(path: str):
    """-> (version, [node, ...]) with node = (name, [property, ...], [child node, ...])."""
    data = open(path, "rb").read()
    if data[:21] != b"Kaydara FBX Binary  \x00":
        raise ValueError(f"{path}: not a binary FBX file (ASCII FBX is not supported)")
    version = struct.unpack_from("<I", data, 23)[0]
    wide = version >= 7500
    head = struct.Struct("<QQQ" if wide else "<III")

    def read_node(off: int):
        if off + head.size + 1 > len(data):
            raise ValueError(f"{path}: truncated node record")
        end, nprops, _plen = head.unpack_from(data, off)
        off += head.size
        nlen = data[off]
        name = data[off + 1:off + 1 + nlen].decode("ascii", "replace")
        off += 1 + nlen
        if end == 0:
            return None, off
        if end > len(data):
            raise ValueError(f"{path}: node '{name}' ends beyond the file")
        props = []
        for _ in range(nprops):
            t = chr(data[off]); off += 1
            if t in _SCALAR:
                props.append(struct.unpack_from(_SCALAR[t], data, off)[0]); off += struct.calcsize(_SCALAR[t])
            elif t in _ARRAY:
                n, enc, clen = struct.unpack_from("<III", data, off); off += 12
                raw = data[off:off + clen]; off += clen
                if enc == 1:
                    raw = zlib.decompress(raw)
                elif enc != 0:
                    raise ValueError(f"{path}: unknown array encoding {enc}")
                props.append(np.frombuffer(raw, dtype=_ARRAY[t], count=n))
            elif t in "SR":
                n = struct.unpack_from("<I", data, off)[0]; off += 4
                props.append(bytes(data[off:off + n]) if t == "R" else data[off:off + n].decode("utf-8", "replace")); off += n
            else:
                raise ValueError(f"{path}: unknown property type '{t}' in node '{name}'")
        children = []
        while off < end:
            child, off = read_node(off)
            if child is None:
                break
            children.append(child)
        return (name, props, children), end

    nodes, off = [], 27
    while off + head.size + 1 <= len(data):
        node, off = read_node(off)
        if node is None:
            break
        nodes.append(node)
    return version, nodes


def _child(node, name):
    return next((c for c in node[2] if c[0] == name), None)


def _clean(name: str) -> str:
    return name.split("\x00", 1)[0]


# ---- geometry ------------------------------------------------------------------------------------------------------------

def _layer_normals(geom, corner_vertex: np.ndarray):
    """Per-corner file normals (k, 3) float64, or None."""
    ln = _child(geom, "LayerElementNormal")
    if ln is None or _child(ln, "Normals") is None:
        return None
    mapping = _child(ln, "MappingInformationType")[1][0]
    ref = _child(ln, "ReferenceInformationType")[1][0]
    normals = np.asarray(_child(ln, "Normals")[1][0], dtype=np.float64).reshape(-1, 3)
    if mapping == "ByPolygonVertex":
        sel = np.arange(corner_vertex.size)
    elif mapping in ("ByVertice", "ByVertex", "ByControlPoint"):
        sel = corner_vertex
    else:
        return None
    if ref == "IndexToDirect":
        idx = np.asarray(_child(ln, "NormalsIndex")[1][0], dtype=np.int64)
        sel = idx[sel]
    elif ref != "Direct":
        return None
    return normals[sel]


def _calculated_normals(pos: np.ndarray, tri_corner: np.ndarray, smooth_angle_deg: float) -> np.ndarray:
    """Per-corner normals recomputed from the triangulated geometry (see the module docstring).  pos: (k, 3) corner positions,
    tri_corner: (t, 3) corner indices."""
    p = pos[tri_corner]                                                    # (t, 3, 3)
    e0, e1 = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
    fn = np.cross(e0, e1)
    area2 = np.linalg.norm(fn, axis=1)
    unit = fn / np.maximum(area2, 1e-300)[:, None]
    # interior angle at each corner
    ang = np.empty((len(p), 3))
    for c in range(3):
        a, b = p[:, (c + 1) % 3] - p[:, c], p[:, (c + 2) % 3] - p[:, c]
        cosv = np.einsum("ij,ij->i", a, b) / np.maximum(np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1), 1e-300)
        ang[:, c] = np.arccos(np.clip(cosv, -1.0, 1.0))
    # weld corners by position
    _, weld = np.unique(np.round(pos, 7), axis=0, return_inverse=True)
    weld = weld.reshape(-1)
    face_of = np.repeat(np.arange(len(p)), 3)
    corner_flat = tri_corner.reshape(-1)
    w_flat = (area2[:, None] * ang).reshape(-1)
    order = np.argsort(weld[corner_flat], kind="stable")
    groups = np.split(order, np.flatnonzero(np.diff(weld[corner_flat][order])) + 1)
    cos_limit = np.cos(np.radians(smooth_angle_deg)) - 1e-9
    out = np.zeros((len(pos), 3))
    for g in groups:
        f = face_of[g]
        n = unit[f]                                                       # face normals around this position
        close = (n @ n.T) >= cos_limit                                    # within the smoothing angle of each corner's own face
        acc = (close * w_flat[g][None, :]) @ n
        out[corner_flat[g]] = acc
    length = np.linalg.norm(out, axis=1)
    bad = length < 1e-30
    out[~bad] /= length[~bad, None]
    if bad.any():                                                         # degenerate faces: fall back to the face normal
        fallback = np.zeros((len(pos), 3)); fallback[corner_flat] = unit[face_of]
        out[bad] = fallback[bad]
    return out


def _geometry_mesh(geom, normal_import_mode: int, smooth_angle_deg: float) -> scenes.MeshDesc:
    verts = np.asarray(_child(geom, "Vertices")[1][0], dtype=np.float64).reshape(-1, 3)
    pvi = np.asarray(_child(geom, "PolygonVertexIndex")[1][0], dtype=np.int64)
    if pvi.size == 0:
        raise ValueError("FBX geometry without polygons")
    ends = pvi < 0
    corner_vertex = np.where(ends, ~pvi, pvi)
    if corner_vertex.max() >= len(verts):
        raise ValueError("FBX polygon index out of range")
    stop = np.flatnonzero(ends)
    start = np.concatenate([[0], stop[:-1] + 1])
    size = stop - start + 1
    if (size < 3).any():
        raise ValueError("FBX polygon with fewer than three corners")
    # fan triangulation: (first, i, i+1) for i = 1 .. size-2
    fans = size - 2
    poly = np.repeat(np.arange(len(size)), fans)
    k = np.arange(fans.sum()) - np.repeat(np.cumsum(fans) - fans, fans)
    first = start[poly]
    tri_corner = np.stack([first, first + k + 1, first + k + 2], axis=1)
    pos = verts[corner_vertex]                                            # per corner
    normals = _layer_normals(geom, corner_vertex) if normal_import_mode == 0 else None
    if normals is None:
        normals = _calculated_normals(pos, tri_corner, smooth_angle_deg)
    # Unity handedness: negate X, reverse the winding
    pos = pos * np.array([-1.0, 1.0, 1.0]); normals = normals * np.array([-1.0, 1.0, 1.0])
    tri_corner = tri_corner[:, ::-1]
    used = tri_corner.reshape(-1)
    return scenes.MeshDesc(np.ascontiguousarray(pos[used], dtype=np.float32), np.arange(used.size, dtype=np.int32),
                           np.ascontiguousarray(normals[used], dtype=np.float32))


def _meta_settings(fbx_path: str):
    """normalImportMode / normalSmoothAngle of the asset's .meta (defaults: import normals, 60 degrees)."""
    mode, angle = 0, 60.0
    try:
        import re
        text = open(fbx_path + ".meta").read()
        m = re.search(r"^\s*normalImportMode:\s*(\d+)", text, flags=re.M)
        if m:
            mode = int(m.group(1))
        m = re.search(r"^\s*normalSmoothAngle:\s*([0-9.]+)", text, flags=re.M)
        if m:
            angle = float(m.group(1))
    except OSError:
        pass
    return mode, angle


def load_fbx_meshes(path: str, normal_import_mode: int | None = None, smooth_angle_deg: float | None = None) -> dict:
    """{Unity fileID: (mesh name, MeshDesc)} for every geometry of the file, named after the node it is attached to (the name
    Unity gives the imported Mesh; the geometry's own name when it hangs on no node)."""
    mode, angle = _meta_settings(path)
    if normal_import_mode is not None:
        mode = normal_import_mode
    if smooth_angle_deg is not None:
        angle = smooth_angle_deg
    _version, nodes = parse_fbx(path)
    objects = next((n for n in nodes if n[0] == "Objects"), None)
    if objects is None:
        raise ValueError(f"{path}: no Objects section")
    geoms = {o[1][0]: o for o in objects[2] if o[0] == "Geometry" and len(o[1]) >= 3 and o[1][2] == "Mesh"}
    models = {o[1][0]: _clean(o[1][1]) for o in objects[2] if o[0] == "Model"}
    owner = {}
    conns = next((n for n in nodes if n[0] == "Connections"), None)
    for c in (conns[2] if conns else []):
        if len(c[1]) >= 3 and c[1][0] == "OO" and c[1][1] in geoms and c[1][2] in models:
            owner.setdefault(c[1][1], models[c[1][2]])
    out = {}
    for gid, g in geoms.items():
        name = owner.get(gid, _clean(g[1][1]))
        out[unity_mesh_file_id(name)] = (name, _geometry_mesh(g, mode, angle))
    return out
