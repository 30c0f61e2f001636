"""The randomised searches of tools/ against the real kernels (librt_b200.so on the B200): random scenes, random sessions of host calls and
random meshes for rtBuildBVH, compared bit for bit with the oracle / the host builder.  Sorts last, so that `pytest -x -m gpu` reaches it
after every hand-made case.  On the CPU the same searches run against the SIMT interpreter build (tests/test_fuzz_smoke.py)."""
import os
import subprocess
import sys

import pytest

from conftest import CUDA_LIB, REPO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tool,args", [("simt_fuzz.py", ["--cases", "250", "--seed", "301", "--far", "0.03", "--spheres", "0.06", "--odd", "0.1"]),
                                       ("simt_fuzz_session.py", ["--cases", "50", "--seed", "302"]),
                                       ("simt_fuzz_bvh.py", ["--cases", "150", "--seed", "303"])])
def test_randomised_search_on_the_gpu_finds_nothing(tool, args):
    r = subprocess.run([sys.executable, os.path.join(REPO, "tools", tool), "--lib", CUDA_LIB] + args, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert r.stdout.strip().endswith("0 findings")
