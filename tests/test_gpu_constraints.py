"""GPU: constrained clustering against the unmodified reference
(tests/golden/make_golden_callers.py) and the reference's literal tests
(tests/constraint_test.py, tests/spectral_clusterer_test.py:243-328): AffinityIntegration as one
element-wise pass, ConstraintPropagation with its matrix inverse by Newton-Schulz on the GEMM
engines, the Turn-to-Diarize preset end to end."""

import os

import numpy as np
import pytest

import spectralcluster_b200 as scb
from spectralcluster_b200 import constraint as ct

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "callers")
ordered = scb.utils.enforce_ordered_labels
SIX = np.array([[1.0, 0.0], [1.1, 0.1], [0.0, 1.0], [0.1, 1.0], [0.9, -0.1], [0.0, 1.2]])


def test_constraint_operators_match_reference():
  z = np.load(os.path.join(GOLDEN, "constraints.npz"))
  a, q = z["a"], z["q"]
  np.testing.assert_array_equal(ct.ConstraintMatrix(list(z["scores"]), threshold=1).compute_diagonals(), q)
  f32 = 2.0 ** -23
  np.testing.assert_allclose(ct.AffinityIntegration(ct.IntegrationType.Max).adjust_affinity(a, q),
                             z["integ_max"], rtol=0, atol=2 * f32)
  np.testing.assert_allclose(ct.AffinityIntegration(ct.IntegrationType.Average).adjust_affinity(a, q),
                             z["integ_avg"], rtol=0, atol=2 * f32)
  for alpha in (0.4, 0.6):       # ~14 chained fp32-accurate products
    got = ct.ConstraintPropagation(alpha).adjust_affinity(a, q)
    np.testing.assert_allclose(got, z["prop_%d" % int(alpha * 10)], rtol=0, atol=2e-6)
  got = ct.ConstraintPropagation(0.6).adjust_affinity(z["a3"], z["q3"])     # constraint_test.py:25-32
  np.testing.assert_allclose(got, z["prop3"], rtol=0, atol=1e-6)
  np.testing.assert_allclose(got, [[1, 0.97, 0], [1.03, 1, 0], [0, 0, 1]], atol=0.01)
  with pytest.raises(ValueError):
    ct.AffinityIntegration().adjust_affinity(a, q[:10, :10])
  with pytest.raises(ValueError):
    ct.ConstraintMatrix([0, -1.0])


def turn_options():
  return scb.RefinementOptions(
      p_percentile=0.95, thresholding_type=scb.ThresholdType.Percentile,
      thresholding_with_binarization=True, thresholding_preserve_diagonal=True,
      symmetrize_type=scb.SymmetrizeType.Average,
      refinement_sequence=[scb.RefinementName.RowWiseThreshold, scb.RefinementName.Symmetrize])


def test_reference_constrained_predict_known_answers():     # spectral_clusterer_test.py:243-328
  q = np.array([[1, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0], [0, 0, 1, 1, 1, 1], [0, 0, 1, 1, 1, 1],
                [0, 0, 1, 1, 1, 1], [0, 0, 1, 1, 1, 1]])
  c = scb.SpectralClusterer(
      max_clusters=2, refinement_options=turn_options(),
      constraint_options=scb.ConstraintOptions(
          constraint_name=scb.ConstraintName.AffinityIntegration, apply_before_refinement=False,
          integration_type=scb.IntegrationType.Max),
      laplacian_type=scb.LaplacianType.GraphCut, row_wise_renorm=True)
  np.testing.assert_array_equal(ordered(c.predict(SIX, q)), [0, 0, 1, 1, 1, 1])
  q = np.array([[1, 1, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0], [0, 0, 0, 1, 0, 0],
                [0, 0, 0, 0, 1, -1], [0, 0, 0, 0, -1, 1]])
  c = scb.SpectralClusterer(
      max_clusters=2, refinement_options=turn_options(),
      constraint_options=scb.ConstraintOptions(
          constraint_name=scb.ConstraintName.ConstraintPropagation, apply_before_refinement=True,
          constraint_propagation_alpha=0.6),
      laplacian_type=scb.LaplacianType.GraphCut, row_wise_renorm=True)
  np.testing.assert_array_equal(ordered(c.predict(SIX, q)), [0, 0, 1, 1, 0, 1])


def test_turn_to_diarize_preset_matches_reference():
  z = np.load(os.path.join(GOLDEN, "constraints.npz"))
  c = scb.SpectralClusterer(
      min_clusters=2, max_clusters=7,
      refinement_options=scb.RefinementOptions(
          thresholding_soft_multiplier=0.01, thresholding_type=scb.ThresholdType.Percentile,
          thresholding_with_binarization=True, thresholding_preserve_diagonal=True,
          symmetrize_type=scb.SymmetrizeType.Average,
          refinement_sequence=list(scb.configs.TURNTODIARIZE_REFINEMENT_SEQUENCE)),
      constraint_options=scb.ConstraintOptions(
          constraint_name=scb.ConstraintName.ConstraintPropagation, apply_before_refinement=True,
          constraint_propagation_alpha=0.4),
      autotune=scb.AutoTune(p_percentile_min=0.40, p_percentile_max=0.95, init_search_step=0.05,
                            search_level=1),
      laplacian_type=scb.LaplacianType.GraphCut, row_wise_renorm=True, custom_dist="cosine")
  labels = c.predict(z["x"], z["q"])
  np.testing.assert_array_equal(ordered(labels), ordered(z["t2d_labels"]))
  assert c.refinement_options.p_percentile == float(z["t2d_p"])
  with pytest.raises(RuntimeError):
    scb.SpectralClusterer(max_spectral_size=100).predict(z["x"], z["q"])
