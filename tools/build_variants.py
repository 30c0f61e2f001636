#!/usr/bin/env python
"""Builds the compile-time A/B variants of librt_b200.so HERE (nvcc cross-compiles without a GPU) into ray_tracing_b200/variants/,
so that a gpurun call spends its minutes measuring, not compiling.  The .so files are git-ignored but travel with the snapshot.

    python tools/build_variants.py [name ...]        (no names: all)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ray_tracing_b200 import build   # noqa: E402

VARIANTS = {
    "r1": ("RT_DEFAULTS_R1",),
    "pushpred": ("RT_PUSH_PREDICATED",),
    "stream": ("RT_STREAM_IMAGES",),
    "cold": ("RT_POOL_COLD_GLOBAL",),
    "cold_ring16": ("RT_POOL_COLD_GLOBAL", "RT_SMEM_STACK=16"),
    "mb7": ("RT_WAVE_MINBLOCKS=7",),
    "mb8": ("RT_WAVE_MINBLOCKS=8",),
    "sphleaf1": ("RT_SPHERE_LEAF=1", "RT_SPHERE_SAH_DEPTH=20"),
    "sphleaf2": ("RT_SPHERE_LEAF=2", "RT_SPHERE_SAH_DEPTH=18"),
    "sphleaf8": ("RT_SPHERE_LEAF=8",),
    "sphsah20": ("RT_SPHERE_SAH_DEPTH=20",),                                   # the round-1 default kernels
    # round 2, third call: on top of the promoted defaults
    "pw28_m32": ("RT_POOL_WARPS=28",),                            # run with poolSlots 32 (28 warps x 64 slots + the stack ring exceed 227 KB)
    "stack4": ("RT_SMEM_STACK=4",),
    "stack16": ("RT_SMEM_STACK=16",),
    "vote274": ("RT_VOTE_WI=2", "RT_VOTE_WL=7", "RT_VOTE_WN=4"),
    "leaf3": ("RT_LEAF_REPEAT=3",),
    "tri64": ("RT_TRI_PAD64",),
    "vote132": ("RT_VOTE_WL=3", "RT_VOTE_WN=2"),
    "vote142": ("RT_VOTE_WL=4", "RT_VOTE_WN=2"),
    "vote274": ("RT_VOTE_WI=2", "RT_VOTE_WL=7", "RT_VOTE_WN=4"),
    "vote163": ("RT_VOTE_WL=6", "RT_VOTE_WN=3"),
    "vote132_pfcur": ("RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_PREFETCH_CUR"),
    "vote132_pfcur_treelet": ("RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_PREFETCH_CUR", "RT_TREELET_PREFETCH"),
    "treelet": ("RT_TREELET_PREFETCH",),
    "treelet_stacktop": ("RT_TREELET_PREFETCH", "RT_STACK_TOP_REG"),
    "stacktop": ("RT_STACK_TOP_REG",),
    "tri_na": ("RT_TRI_LOAD_POLICY=1",),
    "tri_ef": ("RT_TRI_LOAD_POLICY=2",),
    "pf": ("RT_PREFETCH_NEXT_PAIR",),
    "pfcur": ("RT_PREFETCH_CUR",),
    "pfcur_ir1": ("RT_PREFETCH_CUR", "RT_INNER_REPEAT=1"),
    "pfcur_treelet": ("RT_PREFETCH_CUR", "RT_TREELET_PREFETCH"),
    "smemstack4": ("RT_SMEM_STACK=4",),
    "smemstack8": ("RT_SMEM_STACK=8",),
    "ldg256": ("RT_LDG256",),
    "ldg256_tri64": ("RT_LDG256", "RT_TRI_PAD64"),
    "ldg256_tri64_na": ("RT_LDG256", "RT_TRI_PAD64", "RT_TRI_LOAD_POLICY=1"),
    "ldg256_vote132": ("RT_LDG256", "RT_VOTE_WL=3", "RT_VOTE_WN=2"),
    "ldg256_tri64_vote132_pfcur": ("RT_LDG256", "RT_TRI_PAD64", "RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_PREFETCH_CUR"),
    "sphsah": ("RT_SPHERE_SAH_DEPTH=14",),
    "leaf2": ("RT_LEAF_REPEAT=2",),
    "ir1": ("RT_INNER_REPEAT=1",),
    "ir3": ("RT_INNER_REPEAT=3",),
    "rayinv": ("RT_CACHE_RAYINV",),
    "pw20": ("RT_POOL_WARPS=20",),
    "pw16": ("RT_POOL_WARPS=16",),
    "skipsqrt": ("RT_SPHERE_SKIP_SQRT",),
    "zerodefocus": ("RT_SKIP_ZERO_DEFOCUS",),
    "glassool": ("RT_GLASS_OUT_OF_LINE",),
    "cornell_all": ("RT_SKIP_ZERO_DEFOCUS", "RT_GLASS_OUT_OF_LINE", "RT_SPHERE_SKIP_SQRT"),
    "zerodefocus_skipsqrt": ("RT_SKIP_ZERO_DEFOCUS", "RT_SPHERE_SKIP_SQRT"),
    "mb5": ("RT_WAVE_MINBLOCKS=5",),
    # round 2, second call: combinations of the first call's winners on soup4k (profiles/r02_a_sweep_first_call.jsonl)
    "c1": ("RT_LDG256", "RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_SMEM_STACK=8"),
    "c2": ("RT_LDG256", "RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_SMEM_STACK=8", "RT_LEAF_REPEAT=2"),
    "c3": ("RT_LDG256", "RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_SMEM_STACK=8", "RT_LEAF_REPEAT=2", "RT_SPHERE_SAH_DEPTH=14"),
    "c4": ("RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_SMEM_STACK=8", "RT_LEAF_REPEAT=2", "RT_SPHERE_SAH_DEPTH=14"),
    "c5": ("RT_LDG256", "RT_VOTE_WI=2", "RT_VOTE_WL=7", "RT_VOTE_WN=4", "RT_SMEM_STACK=8", "RT_LEAF_REPEAT=2", "RT_SPHERE_SAH_DEPTH=14"),
    "c6": ("RT_LDG256", "RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_SMEM_STACK=16", "RT_LEAF_REPEAT=2", "RT_SPHERE_SAH_DEPTH=14"),
    "c7": ("RT_LDG256", "RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_STACK_TOP_REG", "RT_LEAF_REPEAT=2", "RT_SPHERE_SAH_DEPTH=14"),
    "smemstack16": ("RT_SMEM_STACK=16",),
    "c8": ("RT_LDG256", "RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_SMEM_STACK=8", "RT_BRANCHLESS_POP", "RT_LEAF_REPEAT=2", "RT_SPHERE_SAH_DEPTH=14"),
    "pw28": ("RT_POOL_WARPS=28",),
    "pw32": ("RT_POOL_WARPS=32",),
    "pw28_c": ("RT_POOL_WARPS=28", "RT_LDG256", "RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_LEAF_REPEAT=2", "RT_SPHERE_SAH_DEPTH=14"),
    "pw32_c": ("RT_POOL_WARPS=32", "RT_LDG256", "RT_VOTE_WL=3", "RT_VOTE_WN=2", "RT_LEAF_REPEAT=2", "RT_SPHERE_SAH_DEPTH=14"),
    "all_mesh": ("RT_TREELET_PREFETCH", "RT_STACK_TOP_REG", "RT_TRI_LOAD_POLICY=1", "RT_LEAF_REPEAT=2"),
}
OUT_DIR = os.path.join(build.PKG_DIR, "variants")


def path(name: str) -> str:
    return os.path.join(OUT_DIR, f"librt_b200_{name}.so")


if __name__ == "__main__":
    os.makedirs(OUT_DIR, exist_ok=True)
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        build.build_cuda(force=True, defines=VARIANTS[n], out=path(n))
    print("built", len(names), "variants in", OUT_DIR)
