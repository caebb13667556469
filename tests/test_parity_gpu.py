"""Parity tests proper (run on the B200): our CUDA path, called through the C ABI, against
  (1) the committed golden vectors (outputs of the unmodified reference CUDA rasterizer),
  (2) Oracle B, the scalar CPU restatement, on seeded inputs small enough to finish in seconds,
  (3) Oracle A, the reference CUDA rasterizer itself (prebuilt oracle/_ref/*.so), at larger sizes,
and size-independent properties at BASELINE.json's full size (1M Gaussians, 752x480)."""
import numpy as np
import pytest
import torch

import helpers
from tolerances import assert_cds, assert_grad

pytestmark = pytest.mark.gpu

SMALL = dict(width=160, height=96, fx=90.43, fy=90.81)
HALF_NTU = dict(width=376, height=240, fx=212.513, fy=213.399)  # the resolution GS-LIVM actually renders NTU at


def check_forward_exact(ours, ref, case, vs_cuda_reference: bool):
    helpers.assert_integer_stages_equal(ours, ref, check_unsorted="point_list_keys_unsorted" in ref)
    vis = np.asarray(ref["radii"]) > 0
    np.testing.assert_array_equal(helpers.bits(ours["conic_opacity"])[vis], helpers.bits(ref["conic_opacity"])[vis],
                                  err_msg="conic_opacity bits")
    if not case.use_colors:
        assert np.abs(ours["rgb"][vis] - ref["rgb"][vis]).max() <= 5e-7
        np.testing.assert_array_equal(helpers.clamped_bits_to_bool3(ours["clamped_bits"])[vis],
                                      np.asarray(ref["clamped"])[vis].astype(bool))
    ok = None
    if not vs_cuda_reference:  # CPU expf differs from CUDA expf in the last ulp: skip threshold-fragile pixels
        ok = (ref["fragile"] == 0).reshape(case.height, case.width)
        assert ok.mean() > 0.995
    assert_cds("C", ours["out_color"], ref["out_color"], None if ok is None else ok[None])
    assert_cds("D", ours["out_depth"], ref["out_depth"], None if ok is None else ok[None])
    assert_cds("S", ours["out_acc"], ref["out_acc"], None if ok is None else ok[None])
    sel = slice(None) if ok is None else ok.ravel()
    ref_T = ref["accum_alpha"] if "accum_alpha" in ref else ref["final_T"]
    np.testing.assert_array_equal(np.asarray(ours["n_contrib"])[sel].astype(np.int64),
                                  np.asarray(ref["n_contrib"])[sel].astype(np.int64), err_msg="n_contrib")
    if vs_cuda_reference:  # same expf, same operation DAG: D, S and T must match to the bit
        np.testing.assert_array_equal(helpers.bits(ours["out_depth"]), helpers.bits(ref["out_depth"]))
        np.testing.assert_array_equal(helpers.bits(ours["out_acc"]), helpers.bits(ref["out_acc"]))
        np.testing.assert_array_equal(helpers.bits(ours["accum_alpha"]), helpers.bits(ref_T))
    else:
        assert np.abs(np.asarray(ours["accum_alpha"])[sel] - np.asarray(ref_T)[sel]).max() <= 1e-6


def check_backward(ours, ref, case, skip_rows=None):
    """skip_rows (bool [P]): Gaussians left out of the comparison (helpers.fragile_gaussian_rows)."""
    keep = None if skip_rows is None or not skip_rows.any() else ~skip_rows
    for name in helpers.GRAD_NAMES:
        if name == "dL_dsh" and case.use_colors:
            continue
        want = np.asarray(ref[name])
        got = np.asarray(ours[name]).reshape(want.shape)
        if keep is not None:
            want, got = want[keep], got[keep]
        assert_grad(name, got, want)


def check_backward_vs_cpu_oracle(ours, ref, case):
    """Gradients against the fp64 CPU oracle for every Gaussian that touches no fragile pixel's tile (the others are
    masked, counted and printed -- not a reason to skip the rest)."""
    rows = helpers.fragile_gaussian_rows(ref, case)
    if rows.any():
        print("fragile pixels: %d; %d of %d Gaussians masked in the gradient check" % (
            int((np.asarray(ref["fragile"]) != 0).sum()), int(rows.sum()), case.P))
    assert rows.mean() < 0.5 or case.P < 100, "too many Gaussians masked: %.2f" % rows.mean()
    check_backward(ours, ref, case, skip_rows=rows)


# ------------------------------------------------------------------------------------------ (1) golden vectors
@pytest.mark.parametrize("blend", helpers.BLENDS)
@pytest.mark.parametrize("sort_path", helpers.SORT_PATHS)
@pytest.mark.parametrize("name", helpers.GOLDEN_CASES)
def test_golden_vectors(name, sort_path, blend, built_library, cuda_device):
    case, ref = helpers.load_golden(name)
    ours = helpers.run_ours(case, cuda_device, sort_path=sort_path, blend=blend)
    check_forward_exact(ours, ref, case, vs_cuda_reference=True)
    check_backward(ours, ref, case)


# ------------------------------------------------------------------------------------------ (2) CPU oracle
@pytest.mark.parametrize("P,seed,deg,intr,heavy,close", [
    (3000, 101, 3, SMALL, False, False),
    (3000, 102, 2, SMALL, True, False),
    (5000, 103, 1, dict(width=100, height=50, fx=56.5, fy=56.7), True, False),   # partial tiles both axes
    (20000, 104, 0, HALF_NTU, False, False),
    (6000, 105, 3, SMALL, True, True),                                           # close-up, large splats
    (1, 106, 3, SMALL, False, False),                                            # single Gaussian
    (37, 107, 3, dict(width=17, height=9, fx=10.0, fy=10.0), True, False),       # image smaller than two tiles
])
@pytest.mark.parametrize("blend", helpers.BLENDS)
@pytest.mark.parametrize("sort_path", helpers.SORT_PATHS)
def test_against_cpu_oracle(P, seed, deg, intr, heavy, close, sort_path, blend, built_library, cuda_device):
    case = helpers.scene_case(P, seed, deg, intr=intr, heavy=heavy, close_up=close)
    ref = helpers.run_oracle_cpu(case)
    ours = helpers.run_ours(case, cuda_device, sort_path=sort_path, blend=blend)
    check_forward_exact(ours, ref, case, vs_cuda_reference=False)
    check_backward_vs_cpu_oracle(ours, ref, case)


def test_nothing_visible(built_library, cuda_device):
    """Everything culled (behind the camera): R = 0, image = background, zero gradients, no crash."""
    case = helpers.scene_case(500, 201, 3, intr=SMALL)
    case.means3D[:, 0] = -5.0 - case.means3D[:, 0].abs()
    ours = helpers.run_ours(case, cuda_device)
    assert ours["num_rendered"] == 0 and not ours["radii"].any()
    assert np.allclose(ours["out_color"], 1.0) and not ours["out_acc"].any() and not ours["out_depth"].any()
    for n in helpers.GRAD_NAMES:
        assert not np.asarray(ours[n]).any(), n
    case.scales[:] = 0.31  # and via the reference's scale cull (forward.cu:19-25)
    case.means3D[:, 0] = case.means3D[:, 0].abs() + 5.0
    assert helpers.run_ours(case, cuda_device, backward=False)["num_rendered"] == 0


def test_colors_precomp_and_sh_paths_agree_on_geometry(built_library, cuda_device):
    case = helpers.scene_case(3000, 301, 1, intr=SMALL, heavy=True)
    a = helpers.run_ours(case, cuda_device, backward=False)
    case.colors_precomp = torch.rand((case.P, 3), generator=torch.Generator().manual_seed(5))
    b = helpers.run_ours(case, cuda_device, backward=True)
    ref = helpers.run_oracle_cpu(case)
    np.testing.assert_array_equal(a["point_list_keys"], b["point_list_keys"])
    np.testing.assert_array_equal(helpers.bits(a["out_depth"]), helpers.bits(b["out_depth"]))
    check_forward_exact(b, ref, case, vs_cuda_reference=False)
    check_backward_vs_cpu_oracle(b, ref, case)


def test_precomputed_covariance_path(built_library, cuda_device):
    """cov3D_precomp instead of scales + rotations (reference rasterizer.cu:165-169, forward.cu:239-247): with the
    covariances the scale/rotation path computes, the render is bit-identical to it; the gradients stop at
    dL_dcov3D (dL_dscale / dL_drot stay zero) and match the oracle."""
    case = helpers.scene_case(5000, 151, 2, intr=SMALL, heavy=True)
    ref_sr = helpers.run_oracle_cpu(case)
    ours_sr = helpers.run_ours(case, cuda_device)
    c = helpers.Case(case.means3D, None, None, case.opacities, case.shs, case.background, case.viewmatrix,
                     case.projmatrix, case.camera_center, case.tanfovx, case.tanfovy, case.width, case.height,
                     case.sh_degree, dL_dcolor=case.dL_dcolor, dL_dacc=case.dL_dacc,
                     cov3D_precomp=np.asarray(ref_sr["cov3D"], np.float32))
    ref = helpers.run_oracle_cpu(c)
    ours = helpers.run_ours(c, cuda_device)
    check_forward_exact(ours, ref, c, vs_cuda_reference=False)
    vis = np.asarray(ref["radii"]) > 0
    # (Gaussians culled only by the 0.3 m scale test of the scale/rotation path are not culled here)
    both = vis & (np.asarray(ref_sr["radii"]) > 0)
    np.testing.assert_array_equal(np.asarray(ours["radii"])[both], np.asarray(ours_sr["radii"])[both])
    keep = ~helpers.fragile_gaussian_rows(ref, c)
    for name in ("dL_dmean2D", "dL_dcolor", "dL_dopacity", "dL_dmean3D", "dL_dcov3D", "dL_dsh"):
        assert_grad(name, np.asarray(ours[name]).reshape(np.asarray(ref[name]).shape)[keep], np.asarray(ref[name])[keep])
    assert float(np.abs(ours["dL_dscale"]).max()) == 0.0 and float(np.abs(ours["dL_drot"]).max()) == 0.0


def _replicated(case, copies):
    """Every Gaussian `copies` times: exact depth ties everywhere, so the order of equal keys (ascending Gaussian
    id, what the reference's stable sort produces) is exercised."""
    rep = lambda t: None if t is None else torch.cat([t] * copies, dim=0).contiguous()  # noqa: E731
    c = helpers.Case(rep(case.means3D), rep(case.scales), rep(case.rotations), rep(case.opacities) * 0.3, rep(case.shs),
                     case.background, case.viewmatrix, case.projmatrix, case.camera_center, case.tanfovx, case.tanfovy,
                     case.width, case.height, case.sh_degree)
    return c.with_upstream(11)


@pytest.mark.parametrize("sort_path", helpers.SORT_PATHS)
def test_depth_ties_keep_id_order(sort_path, built_library, cuda_device):
    case = _replicated(helpers.scene_case(2500, 131, 1, intr=SMALL, heavy=True), 3)
    ref = helpers.run_oracle_cpu(case)
    ours = helpers.run_ours(case, cuda_device, sort_path=sort_path)
    keys = np.asarray(ours["point_list_keys"]).astype(np.uint64)
    assert (keys[1:] == keys[:-1]).mean() > 0.5, "the case must contain equal keys"
    check_forward_exact(ours, ref, case, vs_cuda_reference=False)


@pytest.mark.parametrize("blend", ("sparse", "small-buckets"))
def test_very_long_tile_list_is_cut_into_depth_buckets(blend, built_library, cuda_device):
    """One tile with 20 000 instances (ten times what a sort CTA holds in shared memory): the list is cut into depth
    buckets (tile_sort.cu, partition_long_kernel) and still matches the oracle bit for bit. There is no list-length
    ceiling and no library sort on this path."""
    case = helpers.scene_case(2000, 141, 0, intr=SMALL)
    fwd = helpers.run_oracle_cpu(case, backward=False)
    m2 = np.asarray(fwd["means2D"])
    central = (np.asarray(fwd["radii"]) > 0) & (np.abs(m2[:, 0] - 80) < 40) & (np.abs(m2[:, 1] - 48) < 24)
    k = int(np.argmax(central))
    g = torch.Generator().manual_seed(3)
    n = 20_000
    pick = lambda t: t[k:k + 1].repeat(n, *([1] * (t.dim() - 1)))  # noqa: E731
    means = pick(case.means3D) + 1e-3 * torch.randn((n, 3), generator=g)
    c = helpers.Case(means, pick(case.scales) * 0.2, pick(case.rotations), torch.full((n, 1), 0.01), pick(case.shs),
                     case.background, case.viewmatrix, case.projmatrix, case.camera_center, case.tanfovx, case.tanfovy,
                     case.width, case.height, case.sh_degree).with_upstream(5)
    ref = helpers.run_oracle_cpu(c)
    lengths = np.diff(np.asarray(ref["ranges"]).astype(np.int64), axis=1)
    assert lengths.max() > 12032
    ours = helpers.run_ours(c, cuda_device, blend=blend)
    check_forward_exact(ours, ref, c, vs_cuda_reference=False)


def test_thousands_of_equal_depths_overflow_a_bucket_and_still_sort(built_library, cuda_device):
    """6 000 copies of ONE Gaussian: every instance of a tile has the same depth, so no splitter can separate them and
    the depth bucket exceeds a sort CTA's shared memory -- the global-memory path of segment_sort_kernel; ties must
    still come out in ascending Gaussian order."""
    case = helpers.scene_case(2000, 141, 0, intr=SMALL)
    fwd = helpers.run_oracle_cpu(case, backward=False)
    m2 = np.asarray(fwd["means2D"])
    central = (np.asarray(fwd["radii"]) > 0) & (np.abs(m2[:, 0] - 80) < 40) & (np.abs(m2[:, 1] - 48) < 24)
    k = int(np.argmax(central))
    n = 6000
    pick = lambda t: t[k:k + 1].repeat(n, *([1] * (t.dim() - 1)))  # noqa: E731
    c = helpers.Case(pick(case.means3D), pick(case.scales) * 0.2, pick(case.rotations), torch.full((n, 1), 0.01),
                     pick(case.shs), case.background, case.viewmatrix, case.projmatrix, case.camera_center,
                     case.tanfovx, case.tanfovy, case.width, case.height, case.sh_degree).with_upstream(5)
    ref = helpers.run_oracle_cpu(c)
    keys = np.asarray(ref["point_list_keys"]).astype(np.uint64)
    assert ref["num_rendered"] >= n and len(np.unique(keys)) <= 64
    ours = helpers.run_ours(c, cuda_device)
    check_forward_exact(ours, ref, c, vs_cuda_reference=False)


# ------------------------------------------------------------------------------------------ (3) reference CUDA
needs_ref = pytest.mark.skipif(not helpers.ref_available(), reason="oracle/_ref/gsr_ref_ext.so not built")


@needs_ref
@pytest.mark.parametrize("P,seed,deg,intr,heavy,close", [
    (200_000, 401, 3, None, False, False),
    (100_000, 402, 0, HALF_NTU, False, False),
    (200_000, 403, 3, None, True, True),
])
@pytest.mark.parametrize("blend", helpers.BLENDS)
@pytest.mark.parametrize("sort_path", helpers.SORT_PATHS)
def test_against_reference_cuda(P, seed, deg, intr, heavy, close, sort_path, blend, built_library, cuda_device):
    case = helpers.scene_case(P, seed, deg, intr=intr, heavy=heavy, close_up=close)
    ref = helpers.run_ref_cuda(case, cuda_device)
    ours = helpers.run_ours(case, cuda_device, sort_path=sort_path, blend=blend)
    check_forward_exact(ours, ref, case, vs_cuda_reference=True)
    check_backward(ours, ref, case)


@needs_ref
def test_oracle_b_is_pinned_to_reference_cuda(cuda_device):
    """Oracle pinning on the box itself: the CPU restatement against the reference CUDA rasterizer run here."""
    case = helpers.scene_case(30_000, 451, 3, intr=HALF_NTU, heavy=True)
    a = helpers.run_ref_cuda(case, cuda_device)
    b = helpers.run_oracle_cpu(case)
    helpers.assert_integer_stages_equal(b, a)
    ok = (b["fragile"] == 0).reshape(case.height, case.width)
    assert_cds("C", b["out_color"], a["out_color"], ok[None])
    assert_cds("D", b["out_depth"], a["out_depth"], ok[None])
    assert_cds("S", b["out_acc"], a["out_acc"], ok[None])
    for name in helpers.GRAD_NAMES:
        assert_grad(name, b[name].reshape(a[name].shape), a[name])


# ------------------------------------------------------------------------------------------ full-size properties
@pytest.fixture(scope="module")
def full_size(built_library, cuda_device):
    case = helpers.scene_case(1_000_000, 20241017 + 2, 3)
    return case, helpers.run_ours(case, cuda_device)


def test_full_size_keys_sorted_and_ranges_partition(full_size):
    case, o = full_size
    keys = o["point_list_keys"].astype(np.uint64)
    assert keys.size == o["num_rendered"] > 500_000
    assert (keys[1:] >= keys[:-1]).all(), "sorted by (tile, depth)"
    same = keys[1:] == keys[:-1]
    assert (o["point_list"][1:][same] > o["point_list"][:-1][same]).all(), "stable: ties keep Gaussian order"
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    r = o["ranges"].astype(np.int64)
    lens = r[:, 1] - r[:, 0]
    assert lens.sum() == keys.size
    counts = np.bincount(tiles, minlength=r.shape[0])
    np.testing.assert_array_equal(counts, lens)
    nz = lens > 0
    np.testing.assert_array_equal(tiles[r[nz, 0]], np.nonzero(nz)[0])
    # the per-tile path (default) and the device-wide path must agree bit for bit
    d = helpers.run_ours(case, torch.device("cuda:0"), backward=False, sort_path="device")
    np.testing.assert_array_equal(d["point_list_keys"].astype(np.uint64), keys)
    np.testing.assert_array_equal(d["point_list"], o["point_list"])
    np.testing.assert_array_equal(d["ranges"], o["ranges"])
    np.testing.assert_array_equal(helpers.bits(d["out_color"]), helpers.bits(o["out_color"]))
    o = d  # the device-wide path also leaves the reference's intermediates behind
    # multiset preserved by the sort: order-independent checksum of (key, value) pairs
    ku, vu = o["point_list_keys_unsorted"].astype(np.uint64), o["point_list_unsorted"].astype(np.uint64)
    mix = lambda k, v: (k * np.uint64(0x9E3779B97F4A7C15) + v * np.uint64(0xC2B2AE3D27D4EB4F))  # noqa: E731
    assert np.bitwise_xor.reduce(mix(ku, vu)) == np.bitwise_xor.reduce(mix(keys, o["point_list"].astype(np.uint64)))
    assert int(mix(ku, vu).sum(dtype=np.uint64)) == int(mix(keys, o["point_list"].astype(np.uint64)).sum(dtype=np.uint64))
    np.testing.assert_array_equal(o["point_offsets"].astype(np.int64), np.cumsum(o["tiles_touched"].astype(np.int64)))


def test_full_size_image_invariants(full_size):
    case, o = full_size
    S, T = o["out_acc"].ravel(), o["accum_alpha"].ravel()
    assert np.isfinite(o["out_color"]).all() and np.isfinite(o["out_depth"]).all()
    assert S.min() >= 0.0 and S.max() <= 1.0 + 1e-6 and T.min() >= 0.0 and T.max() <= 1.0
    assert np.abs(S + T - 1.0).max() <= 2e-5, "S = sum alpha_i T_i telescopes to 1 - T_final"
    assert (o["n_contrib"].reshape(case.height, case.width) <= (o["ranges"][:, 1] - o["ranges"][:, 0]).max()).all()
    depth_hi = o["depths"][o["radii"] > 0].max()
    assert o["out_depth"].max() <= depth_hi * (1 + 1e-5)


def test_full_size_idempotent_and_background_linear(full_size, cuda_device):
    case, o = full_size
    again = helpers.run_ours(case, cuda_device, backward=False)
    for k in ("out_color", "out_depth", "out_acc", "accum_alpha"):
        np.testing.assert_array_equal(helpers.bits(o[k]), helpers.bits(again[k]), err_msg=k)
    np.testing.assert_array_equal(o["point_list"], again["point_list"])
    case.background = torch.zeros(3)
    black = helpers.run_ours(case, cuda_device, backward=False)
    T = o["accum_alpha"].reshape(1, case.height, case.width)
    assert np.abs((o["out_color"] - black["out_color"]) - T).max() <= 2e-6, "C(bg=1) - C(bg=0) = T_final"
    case.background = torch.ones(3)


def test_full_size_backward_is_linear_in_upstream_gradient(full_size, cuda_device):
    case, o = full_size
    case.dL_dcolor, case.dL_dacc = case.dL_dcolor * 2.0, case.dL_dacc * 2.0
    twice = helpers.run_ours(case, cuda_device)
    case.dL_dcolor, case.dL_dacc = case.dL_dcolor * 0.0, case.dL_dacc * 0.0
    zero = helpers.run_ours(case, cuda_device)
    case.dL_dcolor, case.dL_dacc = None, None
    for n in helpers.GRAD_NAMES:
        assert_grad(n + " (x2)", twice[n], 2.0 * o[n])
        assert not zero[n].any(), n + " must vanish with a zero upstream gradient"
    culled = o["radii"] == 0
    for n in ("dL_dmean3D", "dL_dscale", "dL_drot", "dL_dsh", "dL_dopacity"):
        assert not o[n][culled].any(), "culled Gaussians receive no gradient"


def test_zz_report_fragile_pixel_statistics():
    """(runs last in this file) How often the CPU-oracle comparisons had to mask Gaussians because of fragile pixels."""
    st = helpers.FRAGILE_STATS
    print("CPU-oracle gradient checks: %(cases)d cases, %(cases_with_fragile_pixels)d with fragile pixels "
          "(%(fragile_pixels)d pixels), %(gaussians_masked)d of %(gaussians)d Gaussians masked" % st)
    assert st["cases"] > 0 and st["gaussians_masked"] <= 0.2 * st["gaussians"]
