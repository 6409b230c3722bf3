"""GPU: the full-spectrum dense solver's parallel route: every eigenvalue by
Sturm-count bisection, the requested eigenvectors by inverse iteration with Gram-Schmidt inside
clusters, the eigengap deciding between the two phases how many vectors are needed
(max_clusters=None scans the whole spectrum, utils.py:100-102).  Checked against LAPACK (fp64)."""

import time

import numpy as np
import pytest

import spectralcluster_b200 as scb
from spectralcluster_b200 import _native as nat
from spectralcluster_b200 import device as dev
from oracle import spectral_oracle as orc
from test_gpu_parity_fullsize import report

pytestmark = pytest.mark.gpu


def sym(n, seed):
  a = np.random.default_rng(seed).standard_normal((n, n))
  return ((a + a.T) / 2).astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("n,which", [(300, nat.EIG_LARGEST), (1000, nat.EIG_SMALLEST), (2500, nat.EIG_LARGEST)])
def test_bisection_and_inverse_iteration(engine, n, which):
  a = sym(n, n)
  w, v, _ = engine.eigh(engine.upload_matrix(a), n, None, None, None, 1.0, which, n, 12, True)
  v = v.cpu().numpy()
  ref = np.linalg.eigvalsh(a)
  ref = ref[::-1] if which == nat.EIG_LARGEST else ref
  scale = np.abs(ref).max()
  np.testing.assert_allclose(w, ref, rtol=0, atol=1e-11 * scale)
  assert np.abs(a @ v - v * w[None, :12]).max() <= 1e-9 * scale
  np.testing.assert_allclose(v.T @ v, np.eye(12), atol=1e-9)


def test_repeated_eigenvalues_dense_route(engine):
  """Five identical diagonal blocks: every eigenvalue five times; the five copies of each wanted
  vector must come out orthonormal with small residuals."""
  rng = np.random.default_rng(5)
  b = rng.standard_normal((120, 120))
  b = (b + b.T) / 2
  a = np.kron(np.eye(5), b).astype(np.float32).astype(np.float64)
  n = a.shape[0]
  w, v, _ = engine.eigh(engine.upload_matrix(a), n, None, None, None, 1.0, nat.EIG_LARGEST, n, 15, True)
  v = v.cpu().numpy()
  ref = np.linalg.eigvalsh(a)[::-1]
  np.testing.assert_allclose(w, ref, rtol=0, atol=1e-11 * np.abs(ref).max())
  assert np.abs(a @ v - v * w[None, :15]).max() <= 1e-8 * np.abs(ref).max()
  np.testing.assert_allclose(v.T @ v, np.eye(15), atol=1e-8)


def test_full_spectrum_n4096_time(engine):
  n = 4096
  a = sym(n, 1)
  s = engine.upload_matrix(a)
  engine.eigh(s, n, None, None, None, 1.0, nat.EIG_LARGEST, n, 8, True)      # warm
  t = dev.torch()
  t.cuda.synchronize()
  t0 = time.perf_counter()
  w, v, _ = engine.eigh(s, n, None, None, None, 1.0, nat.EIG_LARGEST, n, 8, True)
  t.cuda.synchronize()
  sec = time.perf_counter() - t0
  ref = np.linalg.eigvalsh(a)[::-1]
  err = float(np.abs(w - ref).max() / np.abs(ref).max())
  report("dense_full_spectrum_n4096", dict(seconds=sec, max_rel_err=err))
  assert err <= 1e-11
  assert sec <= 2.0, sec


def test_max_clusters_none_scans_the_full_spectrum_n16384():
  n = 16384
  x, truth = orc.synthetic_dvectors(n, 256, 5, seed=2, return_labels=True)
  c = scb.SpectralClusterer(min_clusters=2, max_clusters=None,
                            laplacian_type=scb.LaplacianType.GraphCut,
                            refinement_options=scb.RefinementOptions(
                                gaussian_blur_sigma=1, p_percentile=0.95,
                                refinement_sequence=list(scb.ICASSP2018_REFINEMENT_SEQUENCE)))
  t0 = time.perf_counter()
  labels = c.predict(x.astype(np.float32))
  sec = time.perf_counter() - t0
  report("max_clusters_none_n16384", dict(seconds=sec, n_clusters=int(c.last_details["n_clusters"])))
  assert c.last_details["solver"] == "dense" and len(c.last_details["eigenvalues"]) == n
  assert c.last_details["n_clusters"] == 5
  assert np.array_equal(scb.utils.enforce_ordered_labels(labels), orc.ordered(truth))
