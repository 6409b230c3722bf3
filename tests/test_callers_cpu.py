"""CPU: the host-side callers around the hot path against vectors produced by the unmodified
reference (tests/golden/make_golden_callers.py): naive / fallback clusterers
(naive_clusterer.py, fallback_clusterer.py:95-124), get_cluster_centroids / chain_labels
(utils.py:159-206), check_single_cluster on host matrices (fallback_clusterer.py:127-187), and
predict()'s tiny-input fallback branch (spectral_clusterer.py:230-234)."""

import os

import numpy as np
import pytest

import spectralcluster_b200 as scb
from spectralcluster_b200 import fallback_clusterer as fb
from spectralcluster_b200 import naive_clusterer as nc
from spectralcluster_b200 import utils

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "callers")


def load(name):
  return np.load(os.path.join(GOLDEN, name + ".npz"))


def test_naive_clusterer_matches_reference_labels():
  z = load("naive")
  for (thr, adapt), want in zip(z["thresholds"], z["labels"]):
    got = nc.NaiveClusterer(thr, None if adapt < 0 else adapt).predict(z["x"])
    np.testing.assert_array_equal(got, want)
  with pytest.raises(ValueError):
    nc.NaiveClusterer(0.8, 0.5)
  c = nc.NaiveClusterer(0.5)
  assert c.predict_next(np.array([1.0, 0.0])) == 0 and len(c.centroids) == 1
  c.reset()
  assert c.centroids == []


def test_naive_6by2_known_answer():          # tests/fallback_clusterer_test.py:19-36
  m = np.array([[1.0, 0.0], [1.1, 0.1], [0.0, 1.0], [0.1, 1.0], [0.9, -0.1], [0.0, 1.2]])
  for kind, kw in ((fb.FallbackClustererType.Naive, dict(naive_threshold=0.5)),
                   (fb.FallbackClustererType.Agglomerative, dict(agglomerative_threshold=0.5))):
    labels = fb.FallbackClusterer(fb.FallbackOptions(fallback_clusterer_type=kind, **kw)).predict(m)
    np.testing.assert_array_equal(utils.enforce_ordered_labels(labels), [0, 0, 1, 1, 0, 1])


def test_fallback_clusterer_matches_reference():
  z = load("fallback")
  got = fb.FallbackClusterer(fb.FallbackOptions(
      fallback_clusterer_type=fb.FallbackClustererType.Naive, naive_threshold=0.6)).predict(z["x"])
  np.testing.assert_array_equal(got, z["naive"])
  got = fb.FallbackClusterer(fb.FallbackOptions(
      fallback_clusterer_type=fb.FallbackClustererType.Agglomerative,
      agglomerative_threshold=0.4)).predict(z["x"])
  np.testing.assert_array_equal(got, z["agglomerative"])


def test_centroids_and_chain_labels():
  z = load("utils")
  np.testing.assert_allclose(utils.get_cluster_centroids(z["x"], z["pre"]), z["centroids"],
                             rtol=1e-13, atol=1e-15)
  chained = utils.chain_labels(z["pre"], z["main"])
  np.testing.assert_array_equal(chained, z["chained"])
  assert chained.dtype == z["chained"].dtype == np.float64       # quirk A.4-5
  main = z["main"]
  assert utils.chain_labels(None, main) is main
  with pytest.raises(ValueError):
    utils.chain_labels(z["pre"], z["main"][:-1])


def test_check_single_cluster_host_matrices():
  z = load("single_cluster")
  for row, a in enumerate((z["one"], z["many"])):
    for col, (name, thr) in enumerate(zip(z["conditions"], z["thresholds"])):
      opt = fb.FallbackOptions(single_cluster_condition=getattr(fb.SingleClusterCondition, str(name)),
                               single_cluster_affinity_threshold=float(thr))
      assert fb.check_single_cluster(opt, None, a) == bool(z["verdicts"][row, col]), (row, name, thr)
  # literal cases of tests/fallback_clusterer_test.py:61-76
  opt = fb.FallbackOptions(single_cluster_condition=fb.SingleClusterCondition.AffinityGmmBic)
  assert fb.check_single_cluster(opt, None, np.array([[1, 0.999, 1.001], [0.999, 1, 1], [1.001, 1, 1]]))
  assert not fb.check_single_cluster(opt, None, np.array([[1, 2, 2], [2, 1, 1], [2, 1, 1.0]]))
  with pytest.raises(TypeError):
    fb.check_single_cluster(fb.FallbackOptions(single_cluster_condition="x"), None, z["one"])


def test_predict_uses_the_fallback_clusterer_below_spectral_min_embeddings():
  z = load("predict_callers")
  c = scb.SpectralClusterer(
      fallback_options=scb.FallbackOptions(spectral_min_embeddings=20, naive_threshold=0.6))
  np.testing.assert_array_equal(c.predict(z["tiny"]), z["tiny_labels"])   # no GPU involved


def test_max_spectral_size_argument_errors():     # spectral_clusterer.py:239-246
  x = np.random.default_rng(0).standard_normal((50, 4))
  with pytest.raises(RuntimeError):
    scb.SpectralClusterer(max_spectral_size=10).predict(x, np.zeros((50, 50)))
  for kw in (dict(max_spectral_size=1), dict(max_spectral_size=5, max_clusters=5),
             dict(max_spectral_size=5, min_clusters=7)):
    with pytest.raises(ValueError):
      scb.SpectralClusterer(**kw).predict(x)


def test_match_labels_known_answers():        # tests/multi_stage_clusterer_test.py:13-80
  from spectralcluster_b200 import multi_stage_clusterer as ms
  cases = [([1, 0], [0], [0, 1]),
           ([0, 1, 2, 3, 4, 5], [0, 0, 0, 1, 2], [0, 3, 4, 1, 2, 5]),
           ([0, 0, 0, 1, 1, 1, 2, 2], [0, 0, 1, 2, 2, 3, 4], [0, 0, 0, 2, 2, 2, 4, 4]),
           ([1, 1, 1, 0, 0, 1], [0, 0, 0, 1, 1], [0, 0, 0, 1, 1, 0]),
           ([1, 1, 1, 0, 0, 2], [0, 0, 0, 1, 1], [0, 0, 0, 1, 1, 2]),
           ([0, 1, 1, 0, 0, 2], [0, 0, 0, 1, 1], [1, 0, 0, 1, 1, 2]),
           ([0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5], [0, 0, 3, 3, 1, 1, 4, 4, 5, 5, 2],
            [0, 0, 3, 3, 1, 1, 4, 4, 5, 5, 2, 2])]
  for current, previous, want in cases:
    np.testing.assert_array_equal(ms.match_labels(np.array(current), np.array(previous)), want)
  with pytest.raises(ValueError):
    ms.match_labels(np.array([0, 1]), np.array([0, 1]))
  with pytest.raises(ValueError):
    ms.MultiStageClusterer(scb.SpectralClusterer(max_spectral_size=50))


def test_naive_centroid_surface_matches_reference_semantics():
  """NaiveCentroid (naive_clusterer.py:5-22) and the aliases configs / spectral_clusterer export."""
  from spectralcluster_b200 import configs, naive_clusterer, spectral_clusterer
  c = naive_clusterer.NaiveCentroid(np.array([1.0, 0.0]))
  c.merge(np.array([0.0, 1.0]))
  assert c.count == 2 and np.allclose(c.embedding, [0.5, 0.5])
  assert abs(c.cosine(np.array([1.0, 1.0])) - 1.0) < 1e-12
  nc = naive_clusterer.NaiveClusterer(0.5)
  labels = nc.predict(np.array([[1, 0], [0.9, 0.1], [0, 1.0], [0.1, 0.9]]))
  assert labels.tolist() == [0, 0, 1, 1]
  assert [k.count for k in nc.centroids] == [2, 2]
  assert np.allclose(nc.centroids[0].embedding, [0.95, 0.05])
  for name in ("AutoTune", "ConstraintName", "ConstraintOptions", "LaplacianType", "RefinementName",
               "RefinementOptions", "ThresholdType", "SymmetrizeType", "SpectralClusterer"):
    assert hasattr(configs, name), name
  for name in ("ConstraintName", "ConstraintOptions", "AutoTune", "FallbackOptions", "LaplacianType"):
    assert hasattr(spectral_clusterer, name), name
