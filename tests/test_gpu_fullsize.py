"""GPU: BASELINE.json's full-size configurations, checked through size-independent properties
(the CPU reference cannot run them: eig is ~N^3 and N=65,536 needs >137 GB of float64):

  * the synthetic speaker-turn generator has a known ground truth: predicted labels must equal it
    up to a permutation (the same property the oracle satisfies at every size it can run);
  * eigenpairs returned by the Lanczos solver satisfy M v = lambda v on the implicit operator
    (residual checked with an independent fp64 matvec in torch);
  * Y = sym(thr(blur(A))) and S = Y Y^T are symmetric; the mirrored Diffuse tiles equal the
    computed ones bit for bit.
"""

import numpy as np
import pytest

import spectralcluster_b200 as scb
from spectralcluster_b200 import _native as nat
from spectralcluster_b200 import device as dev
from spectralcluster_b200 import laplacian as lap_lib
from oracle import spectral_oracle as orc

pytestmark = pytest.mark.gpu


def same_partition(a, b):
  return np.array_equal(scb.utils.enforce_ordered_labels(np.asarray(a)),
                        scb.utils.enforce_ordered_labels(np.asarray(b)))


def icassp_options():
  return scb.RefinementOptions(gaussian_blur_sigma=1, p_percentile=0.95,
                               thresholding_soft_multiplier=0.01,
                               refinement_sequence=list(scb.ICASSP2018_REFINEMENT_SEQUENCE))


def test_config2_n16384_icassp_no_laplacian():
  x, truth = orc.synthetic_dvectors(16384, 256, 4, seed=0, return_labels=True)
  c = scb.SpectralClusterer(min_clusters=2, max_clusters=7, refinement_options=icassp_options())
  labels = c.predict(x)
  assert c.last_details["solver"] == "lanczos" and c.last_details["n_clusters"] == 4
  assert same_partition(labels, truth)
  w = c.last_details["eigenvalues"]
  assert np.all(np.diff(w) <= 1e-9 * w[0]) and w[3] / w[4] > 10     # descending, clear gap at k=4


def test_config3_n65536_graphcut_eigengap():
  x, truth = orc.synthetic_dvectors(65536, 256, 6, seed=0, return_labels=True)
  c = scb.SpectralClusterer(min_clusters=2, max_clusters=10,
                            laplacian_type=scb.LaplacianType.GraphCut,
                            refinement_options=icassp_options())
  labels = c.predict(x.astype(np.float32))
  assert c.last_details["n_clusters"] == 6
  assert same_partition(labels, truth)
  w = c.last_details["eigenvalues"]
  assert abs(w[0]) < 1e-6 and np.all(np.diff(w) >= -1e-9)             # lambda_0 ~ 0, ascending


def test_config5_autotune_n32768_sweep():
  x, truth = orc.synthetic_dvectors(32768, 256, 6, seed=0, return_labels=True)
  c = scb.SpectralClusterer(
      min_clusters=2, max_clusters=10, refinement_options=icassp_options(),
      autotune=scb.AutoTune(p_percentile_min=0.60, p_percentile_max=0.95, init_search_step=0.045,
                            search_level=1))
  assert len(c.autotune.get_percentile_range()) == 8
  labels = c.predict(x.astype(np.float32))
  assert same_partition(labels, truth)
  assert 0.6 <= c.last_details["best_p_percentile"] <= 0.95


def test_symmetry_and_lanczos_residual_n16384(engine):
  t = dev.torch()
  n = 16384
  x = t.from_numpy(orc.synthetic_dvectors(n, 256, 5, seed=7).astype(np.float32)).to(engine.device)
  a, crop = engine.affinity(x, want_crop_vector=True)
  opt = icassp_options()
  refined = dev.run_refinement(engine, a, n, opt, crop_vector=crop)
  s = refined.s[:, :n]
  # mirrored tiles are copies; tiles crossing the diagonal are computed twice -> equal to rounding
  rel = ((s - s.T).abs().max() / s.abs().max()).item()
  assert rel <= 2e-6
  delta, left, right, sign, which = lap_lib.operator_terms(engine, refined, scb.LaplacianType.GraphCut)
  w, v, stats = engine.eigh(refined.s, n, delta, left, right, sign, which, 11, 8, dense=False)
  # independent fp64 check of M v = lambda v with M = diag(delta) - diag(left) S diag(right)
  s64 = s.double()
  for col in range(8):
    vec = v[:, col]
    mv = delta * vec - left * (s64 @ (right * vec))
    resid = (mv - w[col] * vec).abs().max().item()
    assert resid <= 1e-7, (col, resid)
  assert stats[2] == 11
