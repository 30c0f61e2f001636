"""GPU parity of options added after the last GPU session of round 1 (they are bit-exact on the SIMT interpreter build,
tests/test_simt_kernels.py; this file sorts last so that `pytest -x -m gpu` reaches it after everything already validated on
the B200)."""
import pytest

from conftest import CUDA_LIB, ORACLE_LIB, assert_bit_equal, render
from ray_tracing_b200 import scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", [1, 3])
def test_pair_record_order_is_layout_only_on_gpu(order):
    sc = scenes.knot_room(160, 90, max_bounces=5, rays_per_pixel=2, nu=200, nv=12, glass=True)
    fo, ao, so = render(ORACLE_LIB, sc, frames=1, want_stats=True)
    for kernel, smem in ((2, 0), (2, 500), (1, 0)):
        fg, ag, sg = render(CUDA_LIB, sc, frames=1, options={"kernel": kernel, "pairOrder": order, "smemNodes": smem, "countStats": 1}, want_stats=True)
        assert_bit_equal(fg, fo, f"pairOrder={order} kernel={kernel} smemNodes={smem}")
        assert all(sg[k] == so[k] for k in ("rays", "boxTests", "triTests"))


@pytest.mark.parametrize("kernel", [1, 2])
def test_grid_fit_is_schedule_only_on_gpu(kernel):
    """A frame small enough that the B200 gets about two pixels per persistent lane, so the fitted grid differs from the full one."""
    sc = scenes.cornell_spheres(640, 360, 6, 4) if kernel == 1 else scenes.knot_room(640, 360, max_bounces=5, rays_per_pixel=2, nu=120, nv=10)
    fo, ao = render(CUDA_LIB, sc, frames=2, options={"kernel": kernel})
    fg, ag = render(CUDA_LIB, sc, frames=2, options={"kernel": kernel, "gridFit": 1})
    assert_bit_equal(ag, ao, f"gridFit kernel={kernel}")
