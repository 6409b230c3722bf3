"""CPU: host-side logic of the B200 path that needs no GPU -- the symbolic Laplacian terms
 against the oracle's materialised Laplacians, and the refinement planner's fusion
decisions against a recording fake engine."""

import numpy as np
import pytest
import torch

import spectralcluster_b200 as scb
from spectralcluster_b200 import _native as nat
from spectralcluster_b200 import device as dev
from spectralcluster_b200 import laplacian as lap
from oracle import spectral_oracle as orc

RN = scb.RefinementName


@pytest.mark.parametrize("kind,ltype", [
    (None, None), ("unnormalized", scb.LaplacianType.Unnormalized),
    ("randomwalk", scb.LaplacianType.RandomWalk), ("graphcut", scb.LaplacianType.GraphCut)])
@pytest.mark.parametrize("row_normalised", [False, True])
def test_operator_terms_reproduce_the_reference_matrix(kind, ltype, row_normalised):
  """diag(delta) + sign*diag(left) S diag(right) == what the reference hands to np.linalg.eig, and
  its symmetrised form has the same spectrum; v = E u / |E u| are its eigenvectors."""
  rng = np.random.default_rng(0)
  x = rng.standard_normal((60, 8))
  s = orc.diffuse(orc.symmetrize(orc.row_threshold(orc.affinity(x), 0.9, 0.01)))
  r = 1.0 / s.max(axis=1) if row_normalised else None
  w_ref = s if r is None else orc.row_normalize(s)
  m_ref = w_ref if kind is None else orc.laplacian(w_ref, kind)
  rowsum = torch.from_numpy(s.sum(axis=1))
  rt = None if r is None else torch.from_numpy(r)
  delta, left, right, sign, which = lap.terms_from_row_sums(rowsum, rt, ltype)
  one = np.ones(60)
  d = np.zeros(60) if delta is None else delta.numpy()
  le = one if left is None else left.numpy()
  ri = one if right is None else right.numpy()
  m = np.diag(d) + sign * le[:, None] * s * ri[None, :]
  np.testing.assert_allclose(m, m_ref, rtol=1e-12, atol=1e-12)
  assert which == (nat.EIG_LARGEST if kind is None else nat.EIG_SMALLEST)
  c = np.sqrt(le * ri)
  t = np.diag(d) + sign * c[:, None] * s * c[None, :]
  np.testing.assert_allclose(t, t.T, atol=1e-12)
  wt, u = np.linalg.eigh(t)
  wm = np.sort(np.linalg.eigvals(m).real)
  np.testing.assert_allclose(wt, wm, rtol=1e-8, atol=1e-9)
  e = np.sqrt(le / ri)
  v = e[:, None] * u
  v /= np.linalg.norm(v, axis=0)
  np.testing.assert_allclose(m @ v, v * wt[None, :], atol=1e-8 * max(1.0, np.abs(wt).max()))


class FakeEngine:
  """Records the device calls run_refinement() makes; returns opaque tokens."""

  def __init__(self, n, tensor_cores=True):
    self.calls, self.n, self.tc = [], n, tensor_cores

  def gemm_engine(self, n):
    return nat.GEMM_TCGEN05 if self.tc else nat.GEMM_SIMT

  def _rec(self, name, *info):
    self.calls.append((name,) + info)
    return name

  def crop_values(self, a, n): return self._rec("crop_values")
  def crop_diagonal(self, a, n): return self._rec("crop_diagonal")
  def gaussian_blur(self, a, n, sigma, diag=None): return self._rec("gaussian_blur", sigma)
  def row_threshold(self, a, n, *args): return self._rec("row_threshold", args[0])
  def symmetrize(self, a, n, kind): return self._rec("symmetrize", kind)
  def blur_rowmax(self, a, n, sigma, diag, zero_diag): return self._rec("blur_rowmax", sigma, diag)

  gemm_precision = nat.GEMM_SPLIT3

  def upper_pass_ok(self, sigma):
    return dev.Engine.upper_pass_ok(sigma)

  def blur_upper_rowmax(self, a, n, sigma, diag, zero_diag):
    self._rec("blur_upper", sigma, diag)
    return "B", "m"

  def threshold_symmetrize_upper(self, b, n, m, p, mult, binarize, keep, sym, want_f32, want_planes,
                                 want_lo=True):
    self._rec("thrsym_upper", b, m, sym, want_f32, want_planes)
    return ("y" if want_f32 else None, "hi" if want_planes else None, "lo" if want_planes else None)

  def blur_threshold_symmetrize(self, a, n, sigma, diag, m, p, mult, binarize, keep, sym, want_f32,
                                want_planes):
    self._rec("blur_thrsym", sigma, diag, sym, want_f32, want_planes)
    return ("y" if want_f32 else None, "hi" if want_planes else None, "lo" if want_planes else None)

  def diffuse(self, n, y=None, hi=None, lo=None, want_stats=False, precision=None):
    self._rec("diffuse", y, hi, want_stats)
    if want_stats:    # (rowmax fp32, rowsum fp64) from the GEMM epilogue
      return "S", torch.ones(self.n, dtype=torch.float32) * 2.0, torch.ones(self.n, dtype=torch.float64) * 3.0
    return "S", None, None
  def row_normalize(self, a, n): return self._rec("row_normalize")

  def row_stats(self, a, n, want_max=True, want_sum=True):
    self._rec("row_stats", want_max, want_sum)
    return (torch.ones(self.n, dtype=torch.float64) * 2.0, None)


def names(eng):
  return [c[0] for c in eng.calls]


def test_planner_fuses_the_icassp_sequence():
  eng = FakeEngine(1000)
  opt = scb.RefinementOptions(refinement_sequence=list(scb.ICASSP2018_REFINEMENT_SEQUENCE))
  out = dev.run_refinement(eng, "A", 1000, opt, crop_vector="cropvec")
  # sigma = 1: the symmetric blur pair (upper tiles blurred once, result mirrored);
  # RowWiseNormalize's row maxima (and the degree) come out of the Diffuse epilogue: no row_stats
  assert names(eng) == ["blur_upper", "thrsym_upper", "diffuse"]
  assert eng.calls[0][1:] == (1.0, "cropvec")            # crop vector from the affinity epilogue
  assert eng.calls[1][1:] == ("B", "m", nat.SYMMETRIZE_MAX, False, True)   # planes only: Diffuse follows
  assert eng.calls[2][1:] == (None, "hi", True)          # tensor-core Diffuse on the planes
  assert out.symmetric and torch.allclose(out.row_scale, torch.full((1000,), 0.5, dtype=torch.float64))
  assert torch.allclose(out.rowsum, torch.full((1000,), 3.0, dtype=torch.float64))


def test_planner_variants():
  # no crop vector supplied -> crop values computed; sigma 0 -> fused chain without blur
  eng = FakeEngine(500)
  opt = scb.RefinementOptions(gaussian_blur_sigma=0, symmetrize_type=scb.SymmetrizeType.Average,
                              refinement_sequence=[RN.CropDiagonal, RN.RowWiseThreshold,
                                                   RN.Symmetrize])
  out = dev.run_refinement(eng, "A", 500, opt)
  assert names(eng) == ["crop_values", "blur_rowmax", "blur_thrsym"]
  assert eng.calls[2][1] == 0.0 and eng.calls[2][3:] == (nat.SYMMETRIZE_AVERAGE, True, False)
  assert out.symmetric and out.row_scale is None
  # small matrices use the SIMT GEMM: fp32 Y is kept instead of planes
  eng = FakeEngine(20, tensor_cores=False)
  opt = scb.RefinementOptions(refinement_sequence=list(scb.ICASSP2018_REFINEMENT_SEQUENCE))
  dev.run_refinement(eng, "A", 20, opt, crop_vector="c")
  assert eng.calls[1][4:] == (True, False) and eng.calls[2][1:] == ("y", None, False)
  assert names(eng)[-1] == "row_stats"                   # SIMT GEMM: separate reduction pass
  # a blur radius other than 4 keeps the two-pass form (blur recomputed in both passes)
  eng = FakeEngine(1000)
  opt = scb.RefinementOptions(gaussian_blur_sigma=2, refinement_sequence=list(scb.ICASSP2018_REFINEMENT_SEQUENCE))
  dev.run_refinement(eng, "A", 1000, opt, crop_vector="c")
  assert names(eng)[:2] == ["blur_rowmax", "blur_thrsym"]
  # Percentile thresholding and threshold-without-symmetrize are not fusable
  eng = FakeEngine(300)
  opt = scb.RefinementOptions(thresholding_type=scb.ThresholdType.Percentile,
                              refinement_sequence=[RN.RowWiseThreshold, RN.Symmetrize])
  out = dev.run_refinement(eng, "A", 300, opt)
  assert names(eng) == ["row_threshold", "symmetrize"] and out.symmetric
  eng = FakeEngine(300)
  out = dev.run_refinement(eng, "A", 300, scb.RefinementOptions(
      refinement_sequence=[RN.GaussianBlur, RN.RowWiseThreshold]))
  assert names(eng) == ["gaussian_blur", "row_threshold"] and not out.symmetric
  # RowWiseNormalize in the middle is materialised and breaks symmetry until Diffuse restores it
  eng = FakeEngine(300)
  out = dev.run_refinement(eng, "A", 300, scb.RefinementOptions(
      refinement_sequence=[RN.RowWiseNormalize, RN.Diffuse]))
  assert names(eng) == ["row_normalize", "diffuse"] and out.symmetric
  with pytest.raises(ValueError):
    dev.run_refinement(FakeEngine(10), "A", 10, scb.RefinementOptions(refinement_sequence=["x"]))
