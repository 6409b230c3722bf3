"""GPU: predict()'s caller branches on the device path against reference-generated vectors
(tests/golden/make_golden_callers.py) and the reference's own literal tests
(tests/spectral_clusterer_test.py:330-500): single-cluster decisions from device reductions over
the resident affinity (fallback_clusterer.py:127-187), max_spectral_size pre-clustering
(spectral_clusterer.py:170-199)."""

import os

import numpy as np
import pytest

import spectralcluster_b200 as scb
from spectralcluster_b200 import fallback_clusterer as fb
from spectralcluster_b200 import spectral_clusterer as sc_mod

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "callers")
ordered = scb.utils.enforce_ordered_labels


def load(name):
  return np.load(os.path.join(GOLDEN, name + ".npz"))


def test_device_affinity_statistics_match_numpy(engine):
  z = load("single_cluster")
  for row, a in enumerate((z["one"], z["many"])):
    n = a.shape[0]
    da = sc_mod.DeviceAffinity(engine.upload_matrix(a), n)
    got = fb.affinity_statistics(da)
    a32 = a.astype(np.float32).astype(np.float64)
    assert got["min"] == a32.min() and got["neighbor_min"] == np.diag(a32, k=1).min()
    np.testing.assert_allclose(got["std"], np.std(a32), rtol=1e-12)
    for col, (name, thr) in enumerate(zip(z["conditions"], z["thresholds"])):
      opt = fb.FallbackOptions(single_cluster_condition=getattr(fb.SingleClusterCondition, str(name)),
                               single_cluster_affinity_threshold=float(thr))
      want = bool(z["verdicts"][row, col])
      assert fb.check_single_cluster(opt, None, da) == want


@pytest.mark.parametrize("tag,name,thr", [("all", "AllAffinity", 0.6), ("nbr", "NeighborAffinity", 0.6),
                                          ("std", "AffinityStd", 0.05), ("bic", "AffinityGmmBic", 0.0)])
def test_min_clusters_1_matches_reference(tag, name, thr):
  z = load("predict_callers")
  for dn in ("x1", "x4"):
    c = scb.SpectralClusterer(
        min_clusters=1, max_clusters=7, laplacian_type=scb.LaplacianType.GraphCut,
        fallback_options=scb.FallbackOptions(
            single_cluster_condition=getattr(scb.SingleClusterCondition, name),
            single_cluster_affinity_threshold=thr),
        refinement_options=scb.RefinementOptions(
            gaussian_blur_sigma=1, p_percentile=0.95,
            refinement_sequence=list(scb.ICASSP2018_REFINEMENT_SEQUENCE)))
    got = c.predict(z[dn])
    np.testing.assert_array_equal(ordered(got), ordered(z["min1_%s_%s" % (tag, dn)]))


def test_max_spectral_size_matches_reference():
  z = load("predict_callers")
  c = scb.SpectralClusterer(
      min_clusters=2, max_clusters=7, max_spectral_size=300,
      refinement_options=scb.RefinementOptions(
          gaussian_blur_sigma=0, p_percentile=0.95,
          refinement_sequence=list(scb.ICASSP2018_REFINEMENT_SEQUENCE)))
  got = c.predict(z["big"])
  assert got.dtype == z["big_labels"].dtype == np.float64
  np.testing.assert_array_equal(ordered(got), ordered(z["big_labels"]))


SINGLE = np.array([[1.0, 0.0], [1.1, 0.1], [1.0, 0.0], [1.1, 0.0], [0.9, -0.1], [1.0, 0.2]])
OUTLIER_LAST = np.array([[1.0, 0.0], [1.1, 0.1], [1.0, 0.0], [1.1, 0.0], [0.9, -0.1], [1.0, 0.5]])
OUTLIER_MID = np.array([[1.0, 0.0], [1.1, 0.1], [1.0, 0.0], [1.0, 0.5], [1.1, 0.0], [0.9, -0.1]])


def test_reference_single_cluster_known_answers():       # spectral_clusterer_test.py:330-500
  c = scb.SpectralClusterer(min_clusters=1, refinement_options=scb.RefinementOptions(
      gaussian_blur_sigma=0, p_percentile=0.95,
      refinement_sequence=list(scb.ICASSP2018_REFINEMENT_SEQUENCE)))
  np.testing.assert_array_equal(ordered(c.predict(SINGLE)), [0] * 6)
  SC = scb.SingleClusterCondition
  table = [
      (OUTLIER_LAST, dict(single_cluster_condition=SC.AllAffinity, single_cluster_affinity_threshold=0.93), [0, 0, 0, 0, 0, 1]),
      (OUTLIER_LAST, dict(single_cluster_condition=SC.AllAffinity, single_cluster_affinity_threshold=0.91), [0] * 6),
      (OUTLIER_MID, dict(single_cluster_condition=SC.NeighborAffinity, single_cluster_affinity_threshold=0.96), [0, 0, 0, 1, 0, 0]),
      (OUTLIER_MID, dict(single_cluster_condition=SC.NeighborAffinity, single_cluster_affinity_threshold=0.94), [0] * 6),
      (OUTLIER_MID, dict(single_cluster_condition=SC.AffinityStd, single_cluster_affinity_threshold=0.02), [0, 0, 0, 1, 0, 0]),
      (OUTLIER_MID, dict(single_cluster_condition=SC.AffinityStd, single_cluster_affinity_threshold=0.03), [0] * 6),
      (OUTLIER_MID, dict(single_cluster_condition=SC.FallbackClusterer,
                         fallback_clusterer_type=scb.FallbackClustererType.Naive, naive_threshold=0.95), [0, 0, 0, 1, 0, 0]),
      (OUTLIER_MID, dict(single_cluster_condition=SC.FallbackClusterer,
                         fallback_clusterer_type=scb.FallbackClustererType.Naive, naive_threshold=0.9), [0] * 6),
  ]
  for x, opts, want in table:
    c = scb.SpectralClusterer(min_clusters=1, laplacian_type=scb.LaplacianType.GraphCut,
                              refinement_options=None, fallback_options=scb.FallbackOptions(**opts))
    np.testing.assert_array_equal(ordered(c.predict(x)), want)
