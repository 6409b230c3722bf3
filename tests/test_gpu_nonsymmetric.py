"""GPU: the general eigen path -- refinement sequences that leave a genuinely
non-symmetric matrix, np.linalg.eig + .real in the reference (utils.py:59-61), Krylov-Schur here --
against vectors produced by the unmodified reference (tests/golden/make_golden_callers.py) and
the reference's own auto-tune tests (tests/spectral_clusterer_test.py:156-241,
tests/autotune_test.py:40-81)."""

import os

import numpy as np
import pytest

import spectralcluster_b200 as scb

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "callers")
RN = scb.RefinementName
ordered = scb.utils.enforce_ordered_labels


def autotune_clusterer(p_min, p_max, step, max_clusters):
  return scb.SpectralClusterer(
      max_clusters=max_clusters,
      refinement_options=scb.RefinementOptions(
          thresholding_type=scb.ThresholdType.Percentile,
          refinement_sequence=[RN.RowWiseThreshold]),
      autotune=scb.AutoTune(p_percentile_min=p_min, p_percentile_max=p_max,
                            init_search_step=step, search_level=1),
      laplacian_type=scb.LaplacianType.GraphCut, row_wise_renorm=True)


def test_reference_autotune_tests_threshold_only_sequence():
  z = np.load(os.path.join(GOLDEN, "nonsymmetric.npz"))
  c = autotune_clusterer(0.60, 0.95, 0.05, 2)           # spectral_clusterer_test.py:156-184
  labels = c.predict(z["six"])
  assert c.last_details["solver"] == "krylov-schur"
  np.testing.assert_array_equal(ordered(labels), [0, 0, 1, 1, 0, 1])
  np.testing.assert_array_equal(ordered(labels), ordered(z["six_labels"]))
  c = autotune_clusterer(0.9, 0.95, 0.03, 4)            # :215-241, noise seeded
  labels = c.predict(z["k1000"])
  np.testing.assert_array_equal(ordered(labels), [0] * 400 + [1] * 300 + [2] * 200 + [3] * 100)
  np.testing.assert_array_equal(ordered(labels), ordered(z["k1000_labels"]))
  assert c.refinement_options.p_percentile == float(z["k1000_p"])     # quirk A.4-2: last searched


def test_2by2_autotune_goes_to_the_fallback_clusterer():   # spectral_clusterer_test.py:186-213
  c = scb.SpectralClusterer(
      max_clusters=2,
      refinement_options=scb.RefinementOptions(thresholding_type=scb.ThresholdType.Percentile,
                                               refinement_sequence=[RN.RowWiseThreshold]),
      autotune=scb.AutoTune(p_percentile_min=0.60, p_percentile_max=0.95, init_search_step=0.05,
                            search_level=1, proxy=scb.AutoTuneProxy.PercentileOverNME),
      fallback_options=scb.FallbackOptions(spectral_min_embeddings=3),
      laplacian_type=scb.LaplacianType.GraphCut, row_wise_renorm=True)
  np.testing.assert_array_equal(ordered(c.predict(np.array([[1.0, 0.0], [0.0, 1.0]]))), [0, 1])


@pytest.mark.parametrize("tag,seq,lap", [
    ("thr_graphcut", [RN.RowWiseThreshold], scb.LaplacianType.GraphCut),
    ("blur_thr_none", [RN.GaussianBlur, RN.RowWiseThreshold], None),
    ("thr_rw", [RN.CropDiagonal, RN.RowWiseThreshold], scb.LaplacianType.RandomWalk)])
def test_nonsymmetric_pipelines_match_reference(tag, seq, lap):
  z = np.load(os.path.join(GOLDEN, "nonsymmetric.npz"))
  c = scb.SpectralClusterer(
      min_clusters=2, max_clusters=7, laplacian_type=lap,
      refinement_options=scb.RefinementOptions(
          gaussian_blur_sigma=1, p_percentile=0.9, thresholding_soft_multiplier=0.01,
          refinement_sequence=list(seq)))
  labels = c.predict(z["syn"])
  assert c.last_details["solver"] == "krylov-schur"
  w, want = c.last_details["eigenvalues"], z[tag + "_w"]
  np.testing.assert_allclose(w, want, rtol=1e-5, atol=1e-6 * np.abs(want).max())
  assert c.last_details["n_clusters_raw"] == int(z[tag + "_k"])
  np.testing.assert_allclose(c.last_details["max_gap"], float(z[tag + "_gap"]), rtol=1e-3)
  np.testing.assert_array_equal(ordered(labels), ordered(z[tag + "_labels"]))


def test_tune_on_host_affinity_returns_every_eigenvector():   # autotune_test.py:40-81
  six = np.load(os.path.join(GOLDEN, "nonsymmetric.npz"))["six"]
  c = autotune_clusterer(0.60, 0.95, 0.05, 2)
  affinity = scb.utils.compute_affinity_matrix(six)

  def p_to_ratio(p):
    c.refinement_options.p_percentile = p
    vectors, k, gap = c._compute_eigenvectors_ncluster(affinity)
    return (1 - p) / gap, vectors, k

  vectors, k, p = c.autotune.tune(p_to_ratio)
  assert vectors.shape[0] == 6 and vectors.shape[1] >= 3 and k == 2 and p == 0.6
