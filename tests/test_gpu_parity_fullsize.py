"""GPU: parity pinned against float64 at the sizes BASELINE.json quotes.

  * Diffuse (refinement.py:232-234) at K = 16,384 and K = 65,536: 512 sampled rows of S = Y Y^T
    against the float64 product of the same fp32 Y (torch on the GPU, used as a checker only),
    rtol 3e-6 -- the two-level accumulation must hold its error at the headline K;
  * configs[1] (ICASSP, no Laplacian) and configs[2] (GraphCut) at N = 8,192 and 16,384: the oracle's
    refinement stage by stage in float64 on the host, LAPACK eigh of the symmetrised matrix
    (SURVEY.md A.2), eigengap and k-means -> eigenvalues within 1e-5 relative (+1e-6 lambda_max
    floor for the Laplacian's lambda_0 ~ 0) AND identical labels;
  * eigenvalues of exactly block-diagonal affinities (true multiplicity k) against LAPACK;
  * configs[2] at its REAL size, N = 65,536 (and configs[4]'s N = 32,768 without a Laplacian):
    the whole chain in float64 torch on the GPU (tests/fp64_checker.py, a checker pinned to the
    oracle at N = 2,400 by test_fp64_checker_matches_oracle), ARPACK eigenvalues of the
    symmetrised operator -> eigenvalues within 1e-5 relative AND identical labels.
"""

import numpy as np
import pytest
import scipy.linalg

import spectralcluster_b200 as scb
from spectralcluster_b200 import _native as nat
from spectralcluster_b200 import device as dev
from oracle import spectral_oracle as orc

import fp64_checker

pytestmark = pytest.mark.gpu

EPS = 1e-10


def report(name, values):
  """Measured numbers go to gpurun_out/ (when it exists) so that they can be committed."""
  import json
  import os
  out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
  if os.path.isdir(out):
    with open(os.path.join(out, "parity_measurements.jsonl"), "a") as f:
      f.write(json.dumps(dict(name=name, **values)) + "\n")


def icassp_options(seq=None):
  return scb.RefinementOptions(
      gaussian_blur_sigma=1, p_percentile=0.95, thresholding_soft_multiplier=0.01,
      refinement_sequence=list(seq if seq is not None else scb.ICASSP2018_REFINEMENT_SEQUENCE))


def fused_y(engine, n, speakers, seed):
  """fp32 Y = Symmetrize(Threshold(Blur(Crop(A)))) of the synthetic workload, on the device."""
  t = dev.torch()
  x = t.from_numpy(orc.synthetic_dvectors(n, 256, speakers, seed=seed).astype(np.float32)).to(engine.device)
  a, crop = engine.affinity(x, want_crop_vector=True)
  refined = dev.run_refinement(engine, a, n, icassp_options(scb.ICASSP2018_REFINEMENT_SEQUENCE[:4]),
                               crop_vector=crop)
  return refined.s


@pytest.mark.parametrize("n", [16384, 65536])
def test_diffuse_rows_vs_fp64_at_headline_k(engine, n):
  t = dev.torch()
  y = fused_y(engine, n, 6, seed=1)
  s, _, _ = engine.diffuse(n, y=y)                      # operator-level API: split3
  rows = t.from_numpy(np.sort(np.random.default_rng(n).choice(n, 512, replace=False))).to(engine.device)
  yr = y[rows, :n].double()
  want = t.empty((512, n), dtype=t.float64, device=engine.device)
  for c0 in range(0, n, 8192):                          # float64 checker, chunked to bound memory
    want[:, c0:c0 + 8192] = yr @ y[c0:c0 + 8192, :n].double().T
  got = s[rows, :n].double()
  rel = (got - want).abs() / want.abs().clamp_min(1e-300)
  worst, rms = rel.max().item(), rel.pow(2).mean().sqrt().item()
  beyond = (rel > 3e-6).double().mean().item()
  report("diffuse_fp64_rows_n%d" % n, dict(max_rel=worst, rms_rel=rms, frac_beyond_3e6=beyond,
                                           chains=n // 128))
  # fp32 round-to-nearest accumulation of n/128 chains: a random walk of ~3e-8 sqrt(n/128) per
  # element (measured 9.2e-7 rms, 4.2e-6 max at n = 65,536; 6.7e-7 / 1.6e-6 at n = 16,384)
  assert rms <= 1e-6, rms
  assert worst <= (3e-6 if n <= 16384 else 5e-6), worst
  assert beyond <= (1e-5 if n <= 16384 else 1e-3), beyond      # measured 3.8e-4 at n = 65,536
  # the mirrored half is the same numbers: S[rows, :] == S[:, rows]^T up to the diagonal tiles
  sym = ((s[rows, :n] - s[:n, rows].T).abs() / want.abs().clamp_min(1e-300).float()).max().item()
  assert sym <= 5e-6, sym


def host_reference(x, laplacian, max_clusters, n_values):
  """Oracle refinement in float64 + LAPACK eigh of the symmetrised matrix."""
  opt = orc.options(min_clusters=2, max_clusters=max_clusters, sequence=orc.ICASSP2018[:-1])
  s = orc.refine(orc.affinity(x), opt)                  # Crop..Diffuse, float64
  n = s.shape[0]
  r = 1.0 / s.max(axis=1)                               # RowWiseNormalize: W = diag(r) S
  if laplacian is None:
    left, right = r, np.ones(n)
    c = np.sqrt(left * right)
    tmat = s * c[:, None] * c[None, :]
    w, u = scipy.linalg.eigh(tmat, subset_by_index=[n - n_values, n - 1])
    w, u = w[::-1], u[:, ::-1]
  else:                                                 # GraphCut (laplacian.py:54-58)
    d = r * s.sum(axis=1)
    inv = 1.0 / (np.sqrt(d) + EPS)
    delta, left, right = inv * d * inv, inv * r, inv
    c = np.sqrt(left * right)
    tmat = np.diag(delta) - s * c[:, None] * c[None, :]
    w, u = scipy.linalg.eigh(tmat, subset_by_index=[0, n_values - 1])
  v = u * np.sqrt(left / right)[:, None]
  v /= np.linalg.norm(v, axis=0, keepdims=True)
  if laplacian is None:
    k, _ = orc.number_of_clusters(w, max_clusters, 1e-2, "ratio", descend=True)
  else:
    k, _ = orc.number_of_clusters(w, max_clusters, eigengap="ratio", descend=False)
  k = max(k, 2)
  return w, k, orc.run_kmeans(v[:, :k], k)


@pytest.mark.parametrize("n,laplacian,max_clusters,speakers", [
    (8192, None, 7, 4), (8192, "graphcut", 10, 6),
    (16384, "graphcut", 10, 6)])       # 16,384: the size from which Diffuse issues a single MMA
def test_configs_vs_float64_eigh(n, laplacian, max_clusters, speakers):
  x, truth = orc.synthetic_dvectors(n, 256, speakers, seed=0, return_labels=True)
  w_ref, k_ref, labels_ref = host_reference(x, laplacian, max_clusters, max_clusters + 1)
  c = scb.SpectralClusterer(
      min_clusters=2, max_clusters=max_clusters, refinement_options=icassp_options(),
      laplacian_type=scb.LaplacianType.GraphCut if laplacian else None)
  labels = c.predict(x)
  w = c.last_details["eigenvalues"]
  report("config_n%d_%s" % (n, laplacian or "nolaplacian"),
         dict(eigenvalues=w.tolist(), reference=w_ref.tolist(),
              tolerance_units=float(np.max(np.abs(w - w_ref) / (1e-5 * np.abs(w_ref) + 1e-6 * np.abs(w_ref).max()))),
              max_rel=float(np.max(np.abs(w - w_ref) / np.abs(w_ref).clip(1e-3 * np.abs(w_ref).max())))))
  assert c.last_details["n_clusters"] == k_ref == speakers
  np.testing.assert_allclose(w, w_ref, rtol=1e-5, atol=1e-6 * np.abs(w_ref).max())
  assert np.array_equal(scb.utils.enforce_ordered_labels(labels), orc.ordered(labels_ref))
  assert np.array_equal(scb.utils.enforce_ordered_labels(labels), orc.ordered(truth))


def block_diagonal_affinity(blocks, size, seed):
  """`blocks` IDENTICAL dense blocks on the diagonal: every eigenvalue has multiplicity `blocks`
  (the Laplacian of a graph with identical connected components)."""
  rng = np.random.default_rng(seed)
  b = orc.affinity(rng.standard_normal((size, 16)))
  a = np.zeros((blocks * size, blocks * size))
  for i in range(blocks):
    a[i * size:(i + 1) * size, i * size:(i + 1) * size] = b
  return a


@pytest.mark.parametrize("laplacian", [None, scb.LaplacianType.GraphCut])
@pytest.mark.parametrize("blocks", [3, 6])
def test_extremal_solver_finds_every_copy_of_a_repeated_eigenvalue(laplacian, blocks):
  size = 400
  a = block_diagonal_affinity(blocks, size, seed=blocks)
  n = a.shape[0]
  c = scb.SpectralClusterer(min_clusters=2, max_clusters=10, laplacian_type=laplacian,
                            refinement_options=scb.RefinementOptions(refinement_sequence=[]),
                            affinity_function=lambda x: a)
  labels = c.predict(np.zeros((n, 2)))
  assert c.last_details["solver"] != "dense"            # the extremal (Krylov) path is under test
  w = c.last_details["eigenvalues"]
  if laplacian is None:
    want = np.sort(np.linalg.eigvalsh(a))[::-1][:11]
  else:
    want = np.sort(np.linalg.eigvalsh(orc.laplacian(a, "graphcut")))[:11]
  np.testing.assert_allclose(w, want, rtol=1e-5, atol=1e-6 * np.abs(want).max())
  # lambda_0 of the Laplacian (top eigenvalue of the affinity) repeats `blocks` times -> k
  assert c.last_details["n_clusters_raw"] == (blocks if laplacian is not None else c.last_details["n_clusters_raw"])
  if laplacian is not None:
    truth = np.repeat(np.arange(blocks), size)
    assert np.array_equal(scb.utils.enforce_ordered_labels(labels), truth)


def checker_reference(x, laplacian, max_clusters, device):
  """eigenvalues, k and labels of the float64 GPU checker (tests/fp64_checker.py)."""
  s = fp64_checker.refine_through_diffuse(x, device)
  terms = fp64_checker.operator_terms(s, laplacian)
  w, v = fp64_checker.extremal_eigh(s, terms, max_clusters + 1)
  del s
  if laplacian is None:
    k, _ = orc.number_of_clusters(w, max_clusters, 1e-2, "ratio", descend=True)
  else:
    k, _ = orc.number_of_clusters(w, max_clusters, eigengap="ratio", descend=False)
  k = max(k, 2)
  return w, k, orc.run_kmeans(v[:, :k], k)


@pytest.mark.parametrize("laplacian,max_clusters,speakers", [(None, 7, 4), ("graphcut", 10, 6)])
def test_fp64_checker_matches_oracle(engine, laplacian, max_clusters, speakers):
  """Pins the checker: at N = 2,400 it must reproduce the (reference-pinned) oracle."""
  n = 2400
  x = orc.synthetic_dvectors(n, 256, speakers, seed=3)
  opt = orc.options(min_clusters=2, max_clusters=max_clusters, sequence=orc.ICASSP2018,
                    laplacian=laplacian)
  labels_ref, det = orc.predict(x, opt, return_details=True)      # np.linalg.eig, as the reference
  w_ref = np.asarray(det["eigenvalues"])[:max_clusters + 1]
  w, k, labels = checker_reference(x, laplacian, max_clusters, engine.device)
  np.testing.assert_allclose(w, w_ref, rtol=1e-9, atol=1e-11 * np.abs(w_ref).max())
  assert k == det["n_clusters"]
  assert np.array_equal(orc.ordered(labels), orc.ordered(labels_ref))


@pytest.mark.parametrize("n,laplacian,max_clusters,speakers", [
    (32768, None, 7, 6),            # configs[4]'s size, ICASSP preset eigen-decomposition
    (65536, "graphcut", 10, 6)])    # configs[2]: the configuration the metric is quoted on
def test_configs_at_real_size_vs_float64_checker(engine, n, laplacian, max_clusters, speakers):
  t = dev.torch()
  x, truth = orc.synthetic_dvectors(n, 256, speakers, seed=0, return_labels=True)
  c = scb.SpectralClusterer(
      min_clusters=2, max_clusters=max_clusters, refinement_options=icassp_options(),
      laplacian_type=scb.LaplacianType.GraphCut if laplacian else None)
  labels = c.predict(x)
  w = np.array(c.last_details["eigenvalues"])
  k_found = c.last_details["n_clusters"]
  t.cuda.synchronize()
  t.cuda.empty_cache()
  w_ref, k_ref, labels_ref = checker_reference(x, laplacian, max_clusters, engine.device)
  t.cuda.empty_cache()
  units = float(np.max(np.abs(w - w_ref) / (1e-5 * np.abs(w_ref) + 1e-6 * np.abs(w_ref).max())))
  report("config_n%d_%s_fp64_gpu_checker" % (n, laplacian or "nolaplacian"),
         dict(eigenvalues=w.tolist(), reference=w_ref.tolist(), tolerance_units=units,
              diffuse_mma_per_product={dev.nat.GEMM_SPLIT3: 3, dev.nat.GEMM_SPLIT2: 2,
                                       dev.nat.GEMM_SINGLE: 1}[engine.diffuse_precision_for(n)]))
  assert k_found == k_ref == speakers
  np.testing.assert_allclose(w, w_ref, rtol=1e-5, atol=1e-6 * np.abs(w_ref).max())
  assert np.array_equal(scb.utils.enforce_ordered_labels(labels), orc.ordered(labels_ref))
  assert np.array_equal(scb.utils.enforce_ordered_labels(labels), orc.ordered(truth))
