"""CPU tests of the upload-time repack (csrc/rt_repack.cuh: planScene) through its device-free hook rtxPlanPairs: the
breadth-first NodePair layout must describe exactly the tree the host uploaded (same children, same order, same leaf ranges),
for every shared-memory budget, for shared meshes, and malformed trees must be refused."""
import ctypes as C

import numpy as np
import pytest

from conftest import CUDA_LIB
import ray_tracing_b200 as rt
from ray_tracing_b200 import capi, scenes

PAIR_DTYPE = np.dtype([("aMin", "<f4", 3), ("aMax", "<f4", 3), ("aStart", "<i4"), ("aCount", "<i4"),
                       ("bMin", "<f4", 3), ("bMax", "<f4", 3), ("bStart", "<i4"), ("bCount", "<i4")])
assert PAIR_DTYPE.itemsize == 64


def _plan(nodes, models, tri_count, budget, treelet_depth=None):
    L = C.CDLL(CUDA_LIB)                                   # host code only: no CUDA call is made
    pairs = np.zeros(max(len(nodes), 1), dtype=PAIR_DTYPE)
    roots = np.zeros(2 * len(models), dtype=np.int32)
    smem = C.c_int()
    msg = C.create_string_buffer(256)
    if treelet_depth is None:
        n = L.rtxPlanPairs(nodes.ctypes.data_as(C.c_void_p), len(nodes), models.ctypes.data_as(C.c_void_p), len(models), tri_count, budget,
                           pairs.ctypes.data_as(C.c_void_p), len(pairs), C.byref(smem), roots.ctypes.data_as(C.c_void_p), msg, 256)
    else:
        n = L.rtxPlanPairsOrdered(nodes.ctypes.data_as(C.c_void_p), len(nodes), models.ctypes.data_as(C.c_void_p), len(models), tri_count, budget, treelet_depth,
                                  pairs.ctypes.data_as(C.c_void_p), len(pairs), C.byref(smem), roots.ctypes.data_as(C.c_void_p), msg, 256)
    return n, pairs[:max(n, 0)], roots.reshape(-1, 2), smem.value, msg.value.decode()


def _scene_buffers():
    knot = scenes.knot_mesh(nu=80, nv=8)
    room = scenes.room_mesh()
    t0, n0, _ = rt.build_bvh(knot.vertices, knot.indices, knot.normals)
    t1, n1, _ = rt.build_bvh(room.vertices, room.indices, room.normals)
    nodes = np.concatenate([n0, n1])
    models = np.zeros(3, dtype=capi.MODEL_DTYPE)                                   # two instances of the knot + the room
    models["nodeOffset"] = [0, 0, len(n0)]
    models["triOffset"] = [0, 0, len(t0)]
    return nodes, models, len(t0) + len(t1), (len(n0), len(t0))


def _check_subtree(nodes, node_offset, tri_offset, node_index, pairs, ref, seen):
    """The device reference (start, count) of `node_index` must describe the same subtree as the uploaded node."""
    nd = nodes[node_offset + node_index] if node_index >= 0 else None
    start, count = ref
    if nd["triangleCount"] > 0:
        assert count == nd["triangleCount"] and start == tri_offset + nd["startIndex"]
        return 1
    assert count == 0
    p = pairs[start]
    seen.add(int(start))
    a, b = nodes[node_offset + nd["startIndex"]], nodes[node_offset + nd["startIndex"] + 1]
    assert np.array_equal(p["aMin"], a["boundsMin"]) and np.array_equal(p["aMax"], a["boundsMax"])       # child order A, B is kept
    assert np.array_equal(p["bMin"], b["boundsMin"]) and np.array_equal(p["bMax"], b["boundsMax"])
    return (_check_subtree(nodes, node_offset, tri_offset, int(nd["startIndex"]), pairs, (p["aStart"], p["aCount"]), seen)
            + _check_subtree(nodes, node_offset, tri_offset, int(nd["startIndex"]) + 1, pairs, (p["bStart"], p["bCount"]), seen))


@pytest.mark.parametrize("budget", [0, 1, 5, 64, 100000])
def test_pair_layout_describes_the_uploaded_trees(budget):
    import sys
    sys.setrecursionlimit(10000)
    nodes, models, tri_count, (n_knot, t_knot) = _scene_buffers()
    n, pairs, roots, smem, msg = _plan(nodes, models, tri_count, budget)
    inner = int((nodes["triangleCount"] <= 0).sum())
    assert n == inner and msg == ""                                                # one pair record per inner node
    assert smem == min(budget, 3072, inner)                                        # hot records are the front of the array
    assert np.array_equal(roots[0], roots[1])                                      # shared mesh -> shared records
    seen = set()
    leaves = _check_subtree(nodes, 0, 0, 0, pairs, tuple(roots[0]), seen)
    leaves += _check_subtree(nodes, n_knot, t_knot, 0, pairs, tuple(roots[2]), seen)
    assert leaves == int((nodes["triangleCount"] > 0).sum()) and len(seen) == inner   # every record reached exactly once
    if 0 < smem < inner:
        # breadth-first: the staged records are the TOP of the trees — every parent of a staged record is staged too
        parent = {}
        for i, p in enumerate(pairs):
            for s, c in ((p["aStart"], p["aCount"]), (p["bStart"], p["bCount"])):
                if c == 0:
                    parent[int(s)] = i
        assert all(parent[i] < smem for i in range(smem) if i in parent)


def test_malformed_trees_are_refused():
    nodes, models, tri_count, _ = _scene_buffers()
    bad = nodes.copy()
    inner = np.flatnonzero(bad["triangleCount"] <= 0)
    bad["startIndex"][inner[3]] = len(bad) + 7                                     # child index outside the buffer
    n, _, _, _, msg = _plan(bad, models, tri_count, 0)
    assert n == capi.RT_E_STATE and "out of range" in msg
    bad = nodes.copy()
    bad["startIndex"][inner[5]] = 0                                                # a child pointing back at the root: cycle
    n, _, _, _, msg = _plan(bad, models, tri_count, 0)
    assert n == capi.RT_E_STATE and msg != ""
    bad = nodes.copy()
    leaf = np.flatnonzero(bad["triangleCount"] > 0)[0]
    bad["startIndex"][leaf] = tri_count                                            # leaf range beyond the Triangles buffer
    n, _, _, _, msg = _plan(bad, models, tri_count, 0)
    assert n == capi.RT_E_STATE and "triangle range" in msg
    m2 = models.copy()
    m2["nodeOffset"][1] = len(nodes) + 1
    n, _, _, _, msg = _plan(nodes, m2, tri_count, 0)
    assert n == capi.RT_E_STATE


def test_single_leaf_mesh_has_no_pairs():
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    tris, nodes, _ = rt.build_bvh(v, np.array([0, 1, 2], dtype=np.int32), np.tile(np.float32([[0, 0, 1]]), (3, 1)))
    models = np.zeros(1, dtype=capi.MODEL_DTYPE)
    n, pairs, roots, smem, msg = _plan(nodes, models, 1, 16)
    assert n == 0 and smem == 0 and tuple(roots[0]) == (0, 1)                      # root is the leaf: (first triangle, count)


@pytest.mark.parametrize("depth,budget", [(1, 0), (2, 0), (3, 0), (5, 0), (32, 0), (1, 40), (3, 700)])
def test_treelet_orders_describe_the_same_trees(depth, budget):
    """The "pairOrder" layouts (treelets of `depth` levels in depth-first order) are permutations of the breadth-first one: the same
    trees hang off the same roots, every record is reached exactly once, the staged (hot) front is unchanged."""
    import sys
    sys.setrecursionlimit(10000)
    nodes, models, tri_count, (n_knot, t_knot) = _scene_buffers()
    n0, bfs, roots0, smem0, _ = _plan(nodes, models, tri_count, budget)
    n, pairs, roots, smem, msg = _plan(nodes, models, tri_count, budget, depth)
    assert n == n0 and msg == "" and smem == smem0
    seen = set()
    leaves = _check_subtree(nodes, 0, 0, 0, pairs, tuple(roots[0]), seen)
    leaves += _check_subtree(nodes, n_knot, t_knot, 0, pairs, tuple(roots[2]), seen)
    assert leaves == int((nodes["triangleCount"] > 0).sum()) and len(seen) == n
    for f in ("aMin", "aMax", "bMin", "bMax"):                                        # the hot front holds the same nodes (child ids inside may differ)
        assert np.array_equal(pairs[f][:smem], bfs[f][:smem])
    if depth >= 32:
        # one treelet spans the whole tree: breadth-first again
        assert np.array_equal(pairs.view(np.uint8), bfs.view(np.uint8))
    if depth == 1 and budget == 0:
        # plain pre-order: the record of an inner child A directly follows its parent's record
        inner_a = pairs["aCount"] == 0
        assert np.array_equal(pairs["aStart"][inner_a], np.flatnonzero(inner_a) + 1)
    if depth == 3 and budget == 0:
        # a descent of two steps from a treelet root stays within the treelet's 7 records
        root = int(roots[0][0])
        for s1 in (pairs[root]["aStart"], pairs[root]["bStart"]):
            assert root < s1 <= root + 2
            for s2, c2 in ((pairs[s1]["aStart"], pairs[s1]["aCount"]), (pairs[s1]["bStart"], pairs[s1]["bCount"])):
                assert c2 > 0 or root + 2 < s2 <= root + 6


# ---- TLAS over the models' world boxes (planTlas through the device-free hook rtxPlanTlas) -------------------------------------------

def _plan_tlas(boxes):
    L = C.CDLL(CUDA_LIB)                                   # host code only: no CUDA call is made
    n = len(boxes)
    pairs = np.zeros(max(n, 1), dtype=PAIR_DTYPE)
    leaves = np.full(n, -1, dtype=np.int32)
    root = np.zeros(2, dtype=np.int32)
    b = np.ascontiguousarray(boxes, dtype=np.float32)
    k = L.rtxPlanTlas(b.ctypes.data_as(C.c_void_p), n, pairs.ctypes.data_as(C.c_void_p), len(pairs), leaves.ctypes.data_as(C.c_void_p), root.ctypes.data_as(C.c_void_p))
    return k, pairs[:max(k, 0)], leaves, root


def _tlas_walk(pairs, leaves, boxes, ref, lo, hi, depth, seen, depths):
    """Every box of the tree must enclose the boxes of all the models below it; returns the models below `ref`."""
    start, count = ref
    if count > 0:
        assert count <= 4
        below = [int(m) for m in leaves[start:start + count]]
        depths.append(depth)
    else:
        assert start not in seen
        seen.add(int(start))
        p = pairs[start]
        below = (_tlas_walk(pairs, leaves, boxes, (p["aStart"], p["aCount"]), p["aMin"], p["aMax"], depth + 1, seen, depths)
                 + _tlas_walk(pairs, leaves, boxes, (p["bStart"], p["bCount"]), p["bMin"], p["bMax"], depth + 1, seen, depths))
    if lo is not None:
        for m in below:
            assert np.all(lo <= boxes[m, :3]) and np.all(hi >= boxes[m, 3:]), "a TLAS box does not enclose a model below it"
    return below


@pytest.mark.parametrize("n", [1, 3, 4, 5, 64, 65, 1000, 4096])
def test_tlas_plan_encloses_every_model_exactly_once(n):
    rng = np.random.RandomState(n)
    centre = rng.uniform(-50, 50, (n, 3))
    half = np.exp(rng.uniform(np.log(0.01), np.log(8.0), (n, 3)))
    boxes = np.concatenate([centre - half, centre + half], axis=1).astype(np.float32)
    if n >= 64:
        boxes[7] = [-np.inf] * 3 + [np.inf] * 3              # a model without trustworthy world bounds (buildModels)
        boxes[11, :] = boxes[12, :]                          # coincident boxes
    k, pairs, leaves, root = _plan_tlas(boxes)
    assert k >= 0
    assert sorted(leaves.tolist()) == list(range(n)), "every model is in exactly one leaf"
    seen, depths = set(), []
    below = _tlas_walk(pairs, leaves, boxes, (int(root[0]), int(root[1])), None, None, 0, seen, depths)
    assert sorted(below) == list(range(n)) and len(seen) == k
    assert max(depths) < 24, "the device walk keeps a stack of 24 entries (RT_TLAS_STACK)"
    if n <= 4:
        assert k == 0 and root[1] == n                       # the root is a leaf


def test_tlas_plan_refuses_more_models_than_the_device_mask_holds():
    boxes = np.zeros((4097, 6), dtype=np.float32)
    assert _plan_tlas(boxes)[0] < 0
