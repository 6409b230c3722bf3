"""CPU: the literal known-answer vectors of the reference's own unit tests, replayed on the
oracle (reference test file:line cited per test)."""

import numpy as np

from oracle import spectral_oracle as orc


def test_affinity_4by2():                    # tests/utils_test.py:10-15
  m = np.array([[3, 4], [-4, 3], [6, 8], [-3, -4]])
  want = np.array([[1, 0.5, 1, 0], [0.5, 1, 0.5, 0.5], [1, 0.5, 1, 0], [0, 0.5, 0, 1]])
  np.testing.assert_equal(want, orc.affinity(m))


def test_sorted_eig_order():                 # tests/utils_test.py:21-37
  a = orc.affinity(np.array([[1, 2], [3, 4], [1, 3]]))
  w, v = orc.sorted_eig(a)
  assert w.shape == (3,) and v.shape == (3, 3) and w[0] > w[1] > w[2]
  w, v = orc.sorted_eig(a, descend=False)
  assert w[0] < w[1] < w[2]


def test_number_of_clusters():               # tests/utils_test.py:43-67
  k, gap = orc.number_of_clusters(np.array([1.0, 0.9, 0.8, 0.2, 0.1]))
  assert k == 3 and abs(gap - 4.0) < 0.01
  w = np.array([1.0, 0.9, 0.8, 0.7, 0.6, 0.5])
  k, gap = orc.number_of_clusters(w)
  assert k == 5 and abs(gap - 1.2) < 0.01
  k, gap = orc.number_of_clusters(w, max_clusters=2)
  assert k == 2 and abs(gap - 1.125) < 0.01
  k, gap = orc.number_of_clusters(np.array([1.0, 0.9, 0.8, 0.2, 0.1]), max_clusters=3,
                                  descend=False)
  assert k == 2 and abs(gap - 0.88) < 0.01


def test_ordered_labels():                   # tests/utils_test.py:73-77
  np.testing.assert_equal(np.array([0, 0, 1, 2, 3, 3, 1]),
                          orc.ordered(np.array([2, 2, 1, 0, 3, 3, 1])))


M3 = np.array([[0.5, 2.0, 3.0], [3.0, 4.0, 5.0], [4.0, 2.0, 1.0]])


def test_crop_diagonal():                    # tests/refinement_test.py:12-16
  got = orc.crop_diagonal(np.array([[1, 2, 3], [3, 4, 5], [4, 2, 1]]))
  np.testing.assert_equal(np.array([[3, 2, 3], [3, 5, 5], [4, 2, 4]]), got)


def test_gaussian_blur():                    # tests/refinement_test.py:22-27
  got = orc.gaussian_blur(np.array([[1.0, 2.0, 3.0], [3.0, 4.0, 5.0], [4.0, 2.0, 1.0]]), 1)
  want = np.array([[2.12, 2.61, 3.10], [2.76, 2.90, 3.06], [3.16, 2.78, 2.46]])
  np.testing.assert_allclose(want, got, atol=0.01)


def test_row_threshold_variants():           # tests/refinement_test.py:33-70
  got = orc.row_threshold(M3, 0.5, 0.01, "percentile")
  assert np.allclose([[0.005, 2.0, 3.0], [0.03, 4.0, 5.0], [4.0, 2.0, 0.01]], got, atol=0.001)
  got = orc.row_threshold(M3, 0.5, 0.01, "rowmax")
  np.testing.assert_allclose([[0.005, 2.0, 3.0], [3.0, 4.0, 5.0], [4.0, 2.0, 0.01]], got,
                             atol=0.001)
  got = orc.row_threshold(M3, 0.5, 0.01, "rowmax", binarize=True)
  np.testing.assert_allclose([[0.005, 1.0, 1.0], [1.0, 1.0, 1.0], [1.0, 1.0, 0.01]], got,
                             atol=0.001)
  got = orc.row_threshold(M3, 0.5, 0.01, "rowmax", binarize=True, preserve_diagonal=True)
  np.testing.assert_allclose(np.ones((3, 3)), got, atol=0.001)


def test_symmetrize():                       # tests/refinement_test.py:76-87
  m = np.array([[1, 2, 3], [3, 4, 5], [4, 2, 1]])
  np.testing.assert_equal(np.array([[1, 3, 4], [3, 4, 5], [4, 5, 1]]), orc.symmetrize(m))
  np.testing.assert_equal(np.array([[1, 2.5, 3.5], [2.5, 4, 3.5], [3.5, 3.5, 1]]),
                          orc.symmetrize(m, "average"))


def test_diffuse_and_rownorm():              # tests/refinement_test.py:93-108
  np.testing.assert_equal(np.array([[5, 11], [11, 25]]), orc.diffuse(np.array([[1, 2], [3, 4]])))
  want = np.array([[0.167, 0.667, 1.0], [0.6, 0.8, 1.0], [1.0, 0.5, 0.25]])
  np.testing.assert_allclose(want, orc.row_normalize(M3), atol=0.001)


def test_laplacians():                       # tests/laplacian_test.py:13-45
  a = orc.affinity(np.array([[3, 4], [-4, 3], [6, 8], [-3, -4]]))
  np.testing.assert_equal(a, orc.laplacian(a, "affinity"))
  want = np.array([[1.5, -0.5, -1, 0], [-0.5, 1.5, -0.5, -0.5], [-1, -0.5, 1.5, 0],
                   [0, -0.5, 0, 0.5]])
  np.testing.assert_equal(want, orc.laplacian(a, "unnormalized"))
  want = np.array([[0.6, -0.2, -0.4, 0], [-0.2, 0.6, -0.2, -0.26], [-0.4, -0.2, 0.6, 0],
                   [0, -0.26, 0, 0.33]])
  np.testing.assert_allclose(want, orc.laplacian(a, "graphcut"), atol=0.01)
  want = np.array([[0.6, -0.2, -0.4, 0], [-0.2, 0.6, -0.2, -0.2], [-0.4, -0.2, 0.6, 0],
                   [0, -0.33, 0, 0.33]])
  np.testing.assert_allclose(want, orc.laplacian(a, "randomwalk"), atol=0.01)


SIX = np.array([[1.0, 0.0], [1.1, 0.1], [0.0, 1.0], [0.1, 1.0], [0.9, -0.1], [0.0, 1.2]])


def test_kmeans_6by2():                      # tests/custom_distance_kmeans_test.py:14-46
  for metric in ("cosine", "euclidean"):
    got = orc.ordered(orc.run_kmeans(SIX, 2, metric, 300))
    np.testing.assert_equal(np.array([0, 0, 1, 1, 0, 1]), got)


def test_predict_6by2_icassp():              # tests/spectral_clusterer_test.py:33-51
  opt = orc.options(sequence=orc.ICASSP2018, sigma=0, p=0.95)
  np.testing.assert_equal(np.array([0, 0, 1, 1, 0, 1]), orc.ordered(orc.predict(SIX, opt)))


def test_predict_6by2_graphcut():            # tests/spectral_clusterer_test.py:112-132
  opt = orc.options(max_clusters=2, laplacian="graphcut", row_wise_renorm=True)
  np.testing.assert_equal(np.array([0, 0, 1, 1, 0, 1]), orc.ordered(orc.predict(SIX, opt)))


def test_autotune_range():                   # tests/autotune_test.py:18-38
  np.testing.assert_allclose(orc.autotune_range(0.60, 0.66, 0.01),
                             [0.60, 0.61, 0.62, 0.63, 0.64, 0.65, 0.66], atol=0.01)
