"""The numerical contract of BASELINE.json's north_star, in one place.

* sort keys, point lists, tile ranges, radii, tiles_touched, depth / mean2D bit patterns: BIT-EXACT.
* C (colour), D (depth), S (silhouette): <= 1e-5 absolute.
* gradients: <= 1e-4 relative, in the numpy.allclose sense
      |ours - ref| <= GRAD_RTOL * |ref| + GRAD_RTOL * max|ref|
  i.e. 1e-4 relative per element, with an absolute floor of 1e-4 of the tensor's scale. Gradients are sums
  of float atomics whose order differs between any two runs: the reference differs from ITSELF by up to
  ~1e-5 of the tensor's scale between two launches (measured, tools/dev_compare.py), so elements that are
  tiny through cancellation cannot be resolved more finely by any implementation.
"""
import numpy as np

CDS_ATOL = 1e-5
GRAD_RTOL = 1e-4


def assert_cds(name, ours, ref, mask=None):
    ours, ref = np.asarray(ours, np.float64), np.asarray(ref, np.float64)
    err = np.abs(ours - ref)
    if mask is not None:
        err = np.where(mask, err, 0.0)
    worst = float(err.max()) if err.size else 0.0
    assert worst <= CDS_ATOL, f"{name}: max abs diff {worst:.3e} > {CDS_ATOL}"


def grad_error(ours, ref):
    """max over elements of |ours-ref| / (|ref| + max|ref|); assert_grad requires <= GRAD_RTOL."""
    ours, ref = np.asarray(ours, np.float64), np.asarray(ref, np.float64)
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    if scale == 0.0:
        return float(np.abs(ours).max()) if ours.size else 0.0
    return float((np.abs(ours - ref) / (np.abs(ref) + scale)).max())


def assert_grad(name, ours, ref):
    assert np.isfinite(np.asarray(ours)).all(), f"{name}: non-finite gradient"
    assert np.asarray(ours).shape == np.asarray(ref).shape, f"{name}: shape {np.shape(ours)} vs {np.shape(ref)}"
    e = grad_error(ours, ref)
    assert e <= GRAD_RTOL, f"{name}: relative error {e:.3e} > {GRAD_RTOL}"
