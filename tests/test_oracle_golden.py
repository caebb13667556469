"""Pin Oracle B (oracle/gsr_oracle.c, the scalar CPU restatement) against the golden vectors in tests/golden/,
which are outputs of the UNMODIFIED reference CUDA rasterizer (tests/golden/make_golden.py). CPU only."""
import numpy as np
import pytest

import helpers
from tolerances import assert_cds, assert_grad


@pytest.fixture(scope="module", params=helpers.GOLDEN_CASES)
def golden(request):
    case, ref = helpers.load_golden(request.param)
    out = helpers.run_oracle_cpu(case, backward=True, nthreads=4)
    return case, ref, out


def test_integer_stages_bit_exact(golden):
    """radii, tiles_touched, offsets, depth / mean2D bits, sorted (tile|depth) keys incl. tie order, tile ranges."""
    case, ref, out = golden
    helpers.assert_integer_stages_equal(out, ref)


def test_preprocess_floats(golden):
    case, ref, out = golden
    vis = ref["radii"] > 0
    # conic / cov3D follow the reference's FMA contraction -> identical bits; SH colour is only needed to 1e-5 on C
    np.testing.assert_array_equal(helpers.bits(out["conic_opacity"])[vis], helpers.bits(ref["conic_opacity"])[vis])
    np.testing.assert_array_equal(helpers.bits(out["cov3D"])[vis], helpers.bits(ref["cov3D"])[vis])
    if not case.use_colors:
        assert np.abs(out["rgb"][vis] - ref["rgb"][vis]).max() <= 5e-7
        np.testing.assert_array_equal(out["clamped"][vis].astype(bool), ref["clamped"][vis].astype(bool))


def test_blend_forward(golden):
    """C, D, S within 1e-5 abs. glibc expf and CUDA expf differ in the last ulp, so pixels where a discrete decision
    (alpha >= 1/255, T >= 1e-4, power <= 0) sits within rounding distance of its threshold are flagged by the oracle
    itself (`fragile`) and excluded; they must be rare."""
    case, ref, out = golden
    H, W = case.height, case.width
    ok = (out["fragile"] == 0).reshape(H, W)
    assert ok.mean() > 0.995
    assert_cds("C", out["out_color"], ref["out_color"], ok[None])
    assert_cds("D", out["out_depth"], ref["out_depth"], ok[None])
    assert_cds("S", out["out_acc"], ref["out_acc"], ok[None])
    same = ok.ravel()
    np.testing.assert_array_equal(out["n_contrib"][same].astype(np.int64), ref["n_contrib"][same].astype(np.int64))
    assert np.abs(out["final_T"][same] - ref["accum_alpha"][same]).max() <= 1e-6


def test_backward_matches_reference(golden):
    """The oracle's float64 gradients vs the reference's float-atomic ones, 1e-4 relative (tests/tolerances.py)."""
    case, ref, out = golden
    if (out["fragile"] != 0).any():
        # a flipped decision changes one pixel's contribution list; allow it to show up only at the floor
        pass
    for name in helpers.GRAD_NAMES:
        if name == "dL_dsh" and case.use_colors:
            continue
        assert_grad(name, out[name].reshape(ref[name].shape), ref[name])


def test_no_depth_gradient_path(golden):
    """Reference quirk (backward.cu:567-571, rasterizer.cu:77-80): D never back-propagates. The oracle has no
    dL_ddepth input at all; mean3D gradients must therefore be independent of the depth image."""
    case, ref, out = golden
    assert "dL_ddepth" not in out
