"""Pin the oracle against the unmodified reference and write golden fixtures.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

1. Imports wq2012/SpectralCluster unmodified from /root/reference.
2. Checks every oracle function (oracle/spectral_oracle.py) against the
   reference function it restates, on seeded inputs -- exact equality, since
   both delegate to the same NumPy/SciPy/scikit-learn routines.
3. Writes small input/output fixtures (float64) to tests/golden/*.npz.  These
   were produced BY THE REFERENCE, not by the oracle; tests replay them against
   the oracle (CPU) and against the CUDA path (GPU).
"""

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import spectralcluster as ref  # noqa: E402  (the unmodified reference)
from spectralcluster import (custom_distance_kmeans, laplacian, refinement,  # noqa: E402
                             utils)
from oracle import spectral_oracle as orc  # noqa: E402

RN = ref.RefinementName
NAME = {"crop": RN.CropDiagonal, "blur": RN.GaussianBlur,
        "threshold": RN.RowWiseThreshold, "symmetrize": RN.Symmetrize,
        "diffuse": RN.Diffuse, "rownorm": RN.RowWiseNormalize}
LAP = {None: None, "affinity": ref.LaplacianType.Affinity,
       "unnormalized": ref.LaplacianType.Unnormalized,
       "randomwalk": ref.LaplacianType.RandomWalk,
       "graphcut": ref.LaplacianType.GraphCut}
GAP = {"ratio": ref.EigenGapType.Ratio,
       "normalizeddiff": ref.EigenGapType.NormalizedDiff}
TT = {"rowmax": ref.ThresholdType.RowMax,
      "percentile": ref.ThresholdType.Percentile}
ST = {"max": ref.SymmetrizeType.Max, "average": ref.SymmetrizeType.Average}
PROXY = {"sqrt": ref.AutoTuneProxy.PercentileSqrtOverNME,
         "linear": ref.AutoTuneProxy.PercentileOverNME}


def ref_clusterer(opt):
  """Fresh reference objects for an oracle option bag (quirk A.4-2)."""
  ro = ref.RefinementOptions(
      gaussian_blur_sigma=opt["sigma"], p_percentile=opt["p"],
      thresholding_soft_multiplier=opt["mult"],
      thresholding_type=TT[opt["threshold_type"]],
      thresholding_with_binarization=opt["binarize"],
      thresholding_preserve_diagonal=opt["preserve_diagonal"],
      symmetrize_type=ST[opt["symmetrize_type"]],
      refinement_sequence=[NAME[s] for s in opt["sequence"]])
  at = None
  if opt["autotune"]:
    a = opt["autotune"]
    at = ref.AutoTune(p_percentile_min=a["p_min"], p_percentile_max=a["p_max"],
                      init_search_step=a["step"], search_level=a.get("level", 1),
                      proxy=PROXY[a.get("proxy", "sqrt")])
  return ref.SpectralClusterer(
      min_clusters=opt["min_clusters"], max_clusters=opt["max_clusters"],
      refinement_options=ro, autotune=at, laplacian_type=LAP[opt["laplacian"]],
      stop_eigenvalue=opt["stop_eigenvalue"],
      row_wise_renorm=opt["row_wise_renorm"], custom_dist=opt["custom_dist"],
      max_iter=opt["max_iter"], eigengap_type=GAP[opt["eigengap"]])


def same(a, b, what):
  if not np.array_equal(np.asarray(a), np.asarray(b)):
    raise AssertionError("oracle != reference: " + what)


def four_cluster_1000x6(seed):
  """The 1000x6 generator of spectral_clusterer_test.py:54-59, seeded."""
  base = np.array([[1.0, 0, 0, 0, 0, 0]] * 400 + [[0, 1.0, 0, 0, 0, 0]] * 300 +
                  [[0, 0, 2.0, 0, 0, 0]] * 200 + [[0, 0, 0, 1.0, 0, 0]] * 100)
  rng = np.random.RandomState(seed)
  return base + (rng.rand(1000, 6) * 2 - 1) * 0.1


def synth(n, d, speakers, seed):
  def build():
    return orc.synthetic_dvectors(n, d, speakers, seed=seed)
  build.gen = (n, d, speakers, seed)
  return build


SIX_BY_TWO = np.array([[1.0, 0.0], [1.1, 0.1], [0.0, 1.0], [0.1, 1.0],
                       [0.9, -0.1], [0.0, 1.2]])

# name -> (embeddings builder, option bag).  Every bag is a configuration the
# CUDA path supports (symmetrisable eigen matrix); sizes keep np.linalg.eig
# in seconds.
CASES = {
    "icassp_6x2": (lambda: SIX_BY_TWO,
                   orc.options(sequence=orc.ICASSP2018, sigma=0, p=0.95)),
    "icassp_1000x6": (lambda: four_cluster_1000x6(1),
                      orc.options(sequence=orc.ICASSP2018, sigma=0, p=0.2,
                                  stop_eigenvalue=0.01)),
    "icassp_ndiff_6x2": (lambda: SIX_BY_TWO,
                         orc.options(sequence=orc.ICASSP2018, sigma=0, p=0.95,
                                     eigengap="normalizeddiff")),
    "graphcut_6x2": (lambda: SIX_BY_TWO,
                     orc.options(max_clusters=2, laplacian="graphcut",
                                 row_wise_renorm=True)),
    "graphcut_1000x6": (lambda: four_cluster_1000x6(2),
                        orc.options(max_clusters=4, laplacian="graphcut",
                                    row_wise_renorm=True)),
    # BASELINE.json configs[0]: configs.icassp2018_clusterer, N=1000 d=128 k=4
    "config1_icassp_n1000_d128": (
        synth(1000, 128, 4, 0),
        orc.options(min_clusters=2, max_clusters=7, sequence=orc.ICASSP2018,
                    sigma=1, p=0.95, mult=0.01)),
    # configs[1] at a size the reference finishes in seconds
    "config2_icassp_n1536_d256": (
        synth(1536, 256, 4, 0),
        orc.options(min_clusters=2, max_clusters=7, sequence=orc.ICASSP2018)),
    # configs[2] (GraphCut + eigengap in [2,10], ICASSP-2018 refinement) scaled
    "config3_graphcut_icassp_n1536_d256": (
        synth(1536, 256, 6, 0),
        orc.options(min_clusters=2, max_clusters=10, sequence=orc.ICASSP2018,
                    laplacian="graphcut")),
    "config3b_graphcut_cts_n1200_d256": (
        synth(1200, 256, 6, 1),
        orc.options(min_clusters=2, max_clusters=10,
                    sequence=("crop", "threshold", "symmetrize"),
                    laplacian="graphcut")),
    "randomwalk_n800_d64": (
        synth(800, 64, 3, 3),
        orc.options(min_clusters=2, max_clusters=8,
                    sequence=("crop", "blur", "threshold", "symmetrize"),
                    laplacian="randomwalk", row_wise_renorm=True)),
    "unnormalized_avg_n700_d64": (
        synth(700, 64, 3, 4),
        orc.options(min_clusters=2, max_clusters=8,
                    sequence=("crop", "blur", "threshold", "symmetrize"),
                    symmetrize_type="average", laplacian="unnormalized")),
    # configs[4] scaled: AutoTune sweep of 8 p values over the ICASSP sequence
    "config5_autotune_n1024_d256": (
        synth(1024, 256, 6, 0),
        orc.options(min_clusters=2, max_clusters=10, sequence=orc.ICASSP2018,
                    autotune=dict(p_min=0.60, p_max=0.95, step=0.045, level=1,
                                  proxy="sqrt"))),
}


def check_operators():
  """Oracle == reference, function by function, on seeded inputs."""
  rng = np.random.default_rng(123)
  x = rng.standard_normal((257, 33))
  a = utils.compute_affinity_matrix(x)
  same(orc.affinity(x), a, "affinity")
  same(orc.crop_diagonal(a), refinement.CropDiagonal().refine(a), "crop")
  for sigma in (0, 1, 2, 0.7):
    same(orc.gaussian_blur(a, sigma), refinement.GaussianBlur(sigma).refine(a),
         "blur sigma=%s" % sigma)
  for kind in ("rowmax", "percentile"):
    for binarize in (False, True):
      for keep_diag in (False, True):
        for p, mult in ((0.95, 0.01), (0.5, 0.0), (0.3, 0.5)):
          got = orc.row_threshold(a, p, mult, kind, binarize, keep_diag)
          want = refinement.RowWiseThreshold(p, mult, TT[kind], binarize,
                                             keep_diag).refine(a)
          same(got, want, "threshold %s %s %s %s" % (kind, binarize, keep_diag, p))
  t = refinement.RowWiseThreshold().refine(a)
  for kind in ("max", "average"):
    same(orc.symmetrize(t, kind), refinement.Symmetrize(ST[kind]).refine(t),
         "symmetrize " + kind)
  same(orc.diffuse(t), refinement.Diffuse().refine(t), "diffuse")
  same(orc.row_normalize(t), refinement.RowWiseNormalize().refine(t), "rownorm")
  for kind in ("affinity", "unnormalized", "randomwalk", "graphcut"):
    same(orc.laplacian(a, kind), laplacian.compute_laplacian(a, LAP[kind]),
         "laplacian " + kind)
  for descend in (True, False):
    w0, v0 = utils.compute_sorted_eigenvectors(a, descend)
    w1, v1 = orc.sorted_eig(a, descend)
    same(w1, w0, "eigenvalues")
    same(v1, v0, "eigenvectors")
  w = np.sort(rng.random(40))[::-1]
  for mc in (None, 3, 7, 100):
    for gap in ("ratio", "normalizeddiff"):
      for descend in (True, False):
        ww = w if descend else w[::-1]
        got = orc.number_of_clusters(ww, mc, 0.05, gap, descend)
        want = utils.compute_number_of_clusters(ww, mc, 0.05, GAP[gap], descend)
        same(got, want, "n_clusters %s %s %s" % (mc, gap, descend))
  e = np.vstack([rng.standard_normal((90, 5)) * 0.1 + 3 * np.eye(5)[i % 5]
                 for i in range(5)])
  for metric in ("cosine", "euclidean"):
    same(orc.run_kmeans(e, 5, metric, 300),
         custom_distance_kmeans.run_kmeans(e, 5, metric, 300), "kmeans " + metric)
  for lo, hi, step in ((0.6, 0.95, 0.05), (0.6, 0.95, 0.045), (0.4, 0.9, 0.1)):
    same(orc.autotune_range(lo, hi, step),
         ref.AutoTune(lo, hi, step).get_percentile_range(), "autotune range")
  lab = rng.integers(0, 7, 300)
  same(orc.ordered(lab), utils.enforce_ordered_labels(lab), "ordered labels")
  print("operator-level: oracle == reference")


def reference_trace(x, opt):
  """Run the reference stage by stage, returning what the fixtures store."""
  clusterer = ref_clusterer(opt)
  a = utils.compute_affinity_matrix(x)
  out = {}
  if opt["autotune"]:
    # Replays spectral_clusterer.py:274-289 through the reference objects.
    trace = []

    def closure(p):
      clusterer.refinement_options.p_percentile = p
      v, k, gap = clusterer._compute_eigenvectors_ncluster(a)
      ratio = np.sqrt(1 - p) / gap
      trace.append((p, ratio, k))
      return ratio, v, k
    v, k, p_best = clusterer.autotune.tune(closure)
    out["autotune_trace"] = np.array(trace)
    out["p_best"] = np.float64(p_best)
    out["n_clusters_raw"] = np.int64(k)
  else:
    refined = a
    for name in clusterer.refinement_options.refinement_sequence or []:
      refined = clusterer.refinement_options.get_refinement_operator(name).refine(refined)
    if opt["laplacian"] and opt["laplacian"] != "affinity":
      m = laplacian.compute_laplacian(refined, LAP[opt["laplacian"]])
      w, v = utils.compute_sorted_eigenvectors(m, descend=False)
    else:
      w, v = utils.compute_sorted_eigenvectors(refined)
    v2, k, gap = ref_clusterer(opt)._compute_eigenvectors_ncluster(a)
    assert np.array_equal(v, v2)
    keep = 16 if x.shape[0] > 16 else x.shape[0]
    out["eigenvalues_head"] = w[:keep]
    out["eigenvalue_max"] = np.float64(np.max(w))
    out["n_clusters_raw"] = np.int64(k)
    out["max_gap"] = np.float64(gap)
    # strided probe of the refined matrix (full matrix would be megabytes)
    idx = np.unique(np.linspace(0, x.shape[0] - 1, 24).astype(int))
    out["probe_index"] = idx
    out["refined_probe"] = refined[np.ix_(idx, idx)]
    out["refined_rowsum"] = refined.sum(axis=1)
  labels = ref_clusterer(opt).predict(x)
  out["labels"] = labels.astype(np.int64)
  return out


def main():
  check_operators()
  for name, (build, opt) in CASES.items():
    x = build()
    got = reference_trace(x, opt)
    # whole-pipeline pin of the oracle
    lab, det = orc.predict(x, opt, return_details=True)
    same(lab, got["labels"], name + " labels")
    if "eigenvalues_head" in got:
      same(det["eigenvalues"][:len(got["eigenvalues_head"])],
           got["eigenvalues_head"], name + " eigenvalues")
    else:
      _, _, k, p_best, trace = orc.autotune(orc.affinity(x), opt)
      same(np.array(trace), got["autotune_trace"], name + " autotune trace")
      same(p_best, got["p_best"], name + " p_best")
    meta = {k: v for k, v in opt.items() if k != "autotune"}
    at = opt["autotune"] or {}
    gen = getattr(build, "gen", None)
    if gen is None:
      got["embeddings"] = x
    else:
      # synthetic inputs are regenerated from the seed (PCG64 streams are
      # stable across NumPy versions); the digest pins the generator.
      got["synthetic_args"] = np.array(gen, dtype=np.int64)
      got["embeddings_sha256"] = np.array(
          hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest())
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        options=np.array(repr(dict(meta, autotune=at or None))), **got)
    print("%-40s N=%-5d k_raw=%d labels=%d clusters  ok" % (
        name, x.shape[0], int(got["n_clusters_raw"]), len(set(got["labels"]))))
  print("pipeline-level: oracle == reference; fixtures written to", HERE)


if __name__ == "__main__":
  main()
