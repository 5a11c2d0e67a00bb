"""Experiment switches of the device path that are OFF by default: they must give the product's results bit for bit before
anybody times them.  Sorted last on purpose (pytest -x)."""
import numpy as np
import pytest

import oracle
import subzero_b200 as sz

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("inflate,seed", [(0.02, 31), (0.1, 32), (0.0005, 33)])
def test_class_c_split_in_two_kernels_equals_the_fused_kernel(inflate, seed):
    """sz_set_option("convex_split", 1): clip #1 in a sweep kernel, the polygon through a global buffer, the force law in a
    second kernel -- rows, pairs, polygons and per-floe outputs identical to the fused class C kernel and to the oracle"""
    prm, soa = sz.voronoi_field(6000, seed=seed, inflate=inflate)
    prm.want_clip_polys = 1
    with sz.ContactContext(0) as ctx:
        ctx.step(prm, soa, allow_pair_errors=True)
        off0, rows0 = ctx.rows()
        out0 = ctx.floe_outputs()
        c0 = ctx.narrow_class_ms()["C"][1]
        ctx.set_option("convex_split", 1)
        before = sz.abi.lib().sz_launch_count()
        ctx.step(prm, soa, allow_pair_errors=True)
        assert sz.abi.lib().sz_launch_count() > before
        off1, rows1 = ctx.rows()
        out1 = ctx.floe_outputs()
        assert ctx.narrow_class_ms()["C"][1] == c0 and (c0 > 1000 or inflate < 0.001)
        assert np.array_equal(off0, off1) and np.array_equal(rows0, rows1)
        for k in out0:
            assert np.array_equal(out0[k], out1[k]), k
        rep = oracle.compare_steps(ctx, oracle.OracleStep(prm, soa, broad_mode=1), rtol=1e-9)
        assert rep["rows_bit_exact"]
        ctx.set_option("convex_split", 0)
