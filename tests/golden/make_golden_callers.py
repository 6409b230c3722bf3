"""Golden vectors for the callers either side of the hot path:
naive / fallback clusterers, get_cluster_centroids / chain_labels, check_single_cluster,
predict() with min_clusters=1, spectral_min_embeddings and max_spectral_size.

Run in the build container only (imports the UNMODIFIED reference from /root/reference):

    python tests/golden/make_golden_callers.py

Outputs tests/golden/callers/*.npz -- every array in them was produced by the reference.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import spectralcluster as ref  # noqa: E402
from spectralcluster import fallback_clusterer as fb, naive_clusterer as nc, utils  # noqa: E402
from oracle import spectral_oracle as orc  # noqa: E402

OUT = os.path.join(HERE, "callers")


def blobs(n, d, k, seed, spread=0.1):
  rng = np.random.default_rng(seed)
  centres = np.eye(d)[:k] + 0.05 * rng.standard_normal((k, d))
  lab = rng.integers(0, k, n)
  return centres[lab] + spread * rng.standard_normal((n, d))


def main():
  os.makedirs(OUT, exist_ok=True)
  # ---- naive clusterer (naive_clusterer.py:58-105)
  x = blobs(240, 8, 4, seed=11)
  cases = [(0.5, -1.0), (0.7, 0.9), (0.9, 0.97)]
  np.savez(os.path.join(OUT, "naive.npz"), x=x, thresholds=np.array(cases),
           labels=np.stack([nc.NaiveClusterer(t, None if a < 0 else a).predict(x)
                            for t, a in cases]))
  # ---- fallback clusterer, both types (fallback_clusterer.py:95-124)
  x = blobs(120, 6, 3, seed=5)
  np.savez(
      os.path.join(OUT, "fallback.npz"), x=x,
      naive=fb.FallbackClusterer(fb.FallbackOptions(
          fallback_clusterer_type=fb.FallbackClustererType.Naive, naive_threshold=0.6)).predict(x),
      agglomerative=fb.FallbackClusterer(fb.FallbackOptions(
          fallback_clusterer_type=fb.FallbackClustererType.Agglomerative,
          agglomerative_threshold=0.4)).predict(x))
  # ---- utils.get_cluster_centroids / chain_labels (utils.py:159-206)
  rng = np.random.default_rng(2)
  x = rng.standard_normal((300, 7))
  pre = rng.permutation(np.arange(300) % 13)
  main_labels = rng.integers(0, 4, 13)
  np.savez(os.path.join(OUT, "utils.npz"), x=x, pre=pre, main=main_labels,
           centroids=utils.get_cluster_centroids(x, pre),
           chained=utils.chain_labels(pre, main_labels))
  # ---- check_single_cluster on affinities (fallback_clusterer.py:127-187)
  one = orc.affinity(blobs(90, 16, 1, seed=3, spread=0.05))
  many = orc.affinity(blobs(90, 16, 3, seed=4, spread=0.05))
  conds = [("AllAffinity", 0.75), ("AllAffinity", 0.999), ("NeighborAffinity", 0.75),
           ("NeighborAffinity", 0.999), ("AffinityStd", 0.05), ("AffinityStd", 0.001),
           ("AffinityGmmBic", 0.0)]
  verdicts = []
  for a in (one, many):
    for name, thr in conds:
      opt = fb.FallbackOptions(single_cluster_condition=getattr(fb.SingleClusterCondition, name),
                               single_cluster_affinity_threshold=thr)
      verdicts.append(bool(fb.check_single_cluster(opt, None, a)))
  np.savez(os.path.join(OUT, "single_cluster.npz"), one=one, many=many,
           conditions=np.array([c for c, _ in conds]), thresholds=np.array([t for _, t in conds]),
           verdicts=np.array(verdicts).reshape(2, len(conds)))
  # ---- predict(): min_clusters=1 with each condition; tiny-input fallback; max_spectral_size
  x1 = blobs(400, 32, 1, seed=7, spread=0.05)
  x4 = orc.synthetic_dvectors(400, 32, 4, seed=7)
  out = dict(x1=x1, x4=x4)
  for tag, name, thr in (("all", "AllAffinity", 0.6), ("nbr", "NeighborAffinity", 0.6),
                         ("std", "AffinityStd", 0.05), ("bic", "AffinityGmmBic", 0.0)):
    for dn, data in (("x1", x1), ("x4", x4)):
      c = ref.SpectralClusterer(
          min_clusters=1, max_clusters=7, laplacian_type=ref.LaplacianType.GraphCut,
          fallback_options=fb.FallbackOptions(
              single_cluster_condition=getattr(fb.SingleClusterCondition, name),
              single_cluster_affinity_threshold=thr),
          refinement_options=ref.RefinementOptions(
              gaussian_blur_sigma=1, p_percentile=0.95,
              refinement_sequence=ref.ICASSP2018_REFINEMENT_SEQUENCE))
      out["min1_%s_%s" % (tag, dn)] = c.predict(data)
  tiny = blobs(12, 6, 2, seed=9)
  out["tiny"] = tiny
  out["tiny_labels"] = ref.SpectralClusterer(
      fallback_options=fb.FallbackOptions(spectral_min_embeddings=20, naive_threshold=0.6)).predict(tiny)
  big = orc.synthetic_dvectors(900, 32, 4, seed=1)
  out["big"] = big
  out["big_labels"] = ref.SpectralClusterer(
      min_clusters=2, max_clusters=7, max_spectral_size=300,
      refinement_options=ref.RefinementOptions(
          gaussian_blur_sigma=0, p_percentile=0.95,
          refinement_sequence=ref.ICASSP2018_REFINEMENT_SEQUENCE)).predict(big)
  np.savez(os.path.join(OUT, "predict_callers.npz"), **out)
  print("wrote", sorted(os.listdir(OUT)))





def nonsymmetric():
  """sequences that leave a genuinely non-symmetric matrix (np.linalg.eig +
  .real in the reference, utils.py:59-61): the reference's own auto-tune tests
  (tests/spectral_clusterer_test.py:156-241) plus two synthetic cases with eigenvalues."""
  RN = ref.RefinementName
  six = np.array([[1.0, 0.0], [1.1, 0.1], [0.0, 1.0], [0.1, 1.0], [0.9, -0.1], [0.0, 1.2]])
  out = {}

  def autotune_clusterer(p_min, p_max, step, max_clusters):
    return ref.SpectralClusterer(
        max_clusters=max_clusters,
        refinement_options=ref.RefinementOptions(
            thresholding_type=ref.ThresholdType.Percentile,
            refinement_sequence=[RN.RowWiseThreshold]),
        autotune=ref.AutoTune(p_percentile_min=p_min, p_percentile_max=p_max,
                              init_search_step=step, search_level=1),
        laplacian_type=ref.LaplacianType.GraphCut, row_wise_renorm=True)

  c = autotune_clusterer(0.60, 0.95, 0.05, 2)
  out["six"] = six
  out["six_labels"] = c.predict(six)
  base = np.array([[1.0, 0, 0, 0, 0, 0]] * 400 + [[0, 1.0, 0, 0, 0, 0]] * 300 +
                  [[0, 0, 2.0, 0, 0, 0]] * 200 + [[0, 0, 0, 1.0, 0, 0]] * 100)
  x = base + (np.random.RandomState(3).rand(1000, 6) * 2 - 1) * 0.1
  c = autotune_clusterer(0.9, 0.95, 0.03, 4)
  out["k1000"] = x
  out["k1000_labels"] = c.predict(x)
  out["k1000_p"] = np.array(c.refinement_options.p_percentile)
  # no autotune: eigenvalues and labels of two non-symmetric pipelines
  x = orc.synthetic_dvectors(800, 64, 4, seed=5)
  out["syn"] = x
  for tag, seq, lap in (("thr_graphcut", [RN.RowWiseThreshold], ref.LaplacianType.GraphCut),
                        ("blur_thr_none", [RN.GaussianBlur, RN.RowWiseThreshold], None),
                        ("thr_rw", [RN.CropDiagonal, RN.RowWiseThreshold], ref.LaplacianType.RandomWalk)):
    c = ref.SpectralClusterer(
        min_clusters=2, max_clusters=7, laplacian_type=lap,
        refinement_options=ref.RefinementOptions(
            gaussian_blur_sigma=1, p_percentile=0.9, thresholding_soft_multiplier=0.01,
            refinement_sequence=seq))
    a = utils.compute_affinity_matrix(x)
    vec, k, gap = c._compute_eigenvectors_ncluster(a)
    for op in seq:
      a = c.refinement_options.get_refinement_operator(op).refine(a)
    if lap is None:
      w, _ = utils.compute_sorted_eigenvectors(a)
    else:
      from spectralcluster import laplacian as lp
      w, _ = utils.compute_sorted_eigenvectors(lp.compute_laplacian(a, lap), descend=False)
    out[tag + "_w"] = np.real(w[:8])
    out[tag + "_k"] = np.array(k)
    out[tag + "_gap"] = np.array(gap)
    out[tag + "_labels"] = c.predict(x)
  np.savez(os.path.join(OUT, "nonsymmetric.npz"), **out)
  print("nonsymmetric:", {k: (v.shape if v.ndim else v.item()) for k, v in out.items()
                          if k.endswith(("_k", "_p", "_gap"))})




def constraints():
  """constraint operators (constraint.py:95-164) and constrained predict()
  (tests/spectral_clusterer_test.py:243-328, configs.turntodiarize_clusterer)."""
  from spectralcluster import constraint as rc
  rng = np.random.default_rng(21)
  out = {}
  x = orc.synthetic_dvectors(300, 32, 3, seed=2)
  a = utils.compute_affinity_matrix(x)
  scores = np.where(rng.random(300) < 0.1, rng.random(300) * 3 + 0.5, 0.0)
  q = rc.ConstraintMatrix(list(scores), threshold=1).compute_diagonals()
  out.update(a=a, q=q, scores=scores)
  out["integ_max"] = rc.AffinityIntegration(rc.IntegrationType.Max).adjust_affinity(a, q)
  out["integ_avg"] = rc.AffinityIntegration(rc.IntegrationType.Average).adjust_affinity(a, q)
  for alpha in (0.4, 0.6):
    out["prop_%d" % int(alpha * 10)] = rc.ConstraintPropagation(alpha).adjust_affinity(a, q)
  # the non-symmetric 3x3 case of tests/constraint_test.py:25-32
  a3 = np.array([[1, 0.25, 0], [0.31, 1, 0], [0, 0, 1]])
  q3 = np.array([[1, 1, 0], [1, 1, 0], [0, 0, 0]])
  out.update(a3=a3, q3=q3, prop3=rc.ConstraintPropagation(0.6).adjust_affinity(a3, q3))
  # constrained predict(): Turn-to-Diarize preset on synthetic turns (fresh objects, quirk A.4-2)
  from spectralcluster import configs as rcfg
  c = ref.SpectralClusterer(
      min_clusters=2, max_clusters=7,
      refinement_options=ref.RefinementOptions(
          thresholding_soft_multiplier=0.01, thresholding_type=ref.ThresholdType.Percentile,
          thresholding_with_binarization=True, thresholding_preserve_diagonal=True,
          symmetrize_type=ref.SymmetrizeType.Average,
          refinement_sequence=[ref.RefinementName.RowWiseThreshold, ref.RefinementName.Symmetrize]),
      constraint_options=rc.ConstraintOptions(
          constraint_name=rc.ConstraintName.ConstraintPropagation, apply_before_refinement=True,
          constraint_propagation_alpha=0.4),
      autotune=ref.AutoTune(p_percentile_min=0.40, p_percentile_max=0.95, init_search_step=0.05,
                            search_level=1),
      laplacian_type=ref.LaplacianType.GraphCut, row_wise_renorm=True, custom_dist="cosine")
  out["x"] = x
  out["t2d_labels"] = c.predict(x, q)
  out["t2d_p"] = np.array(c.refinement_options.p_percentile)
  np.savez(os.path.join(OUT, "constraints.npz"), **out)
  print("constraints: t2d clusters", len(set(out["t2d_labels"].tolist())), "p", out["t2d_p"])


def multistage():
  """MultiStageClusterer.streaming_predict (multi_stage_clusterer.py:125-180)."""
  from spectralcluster import multi_stage_clusterer as rm
  x = orc.synthetic_dvectors(130, 24, 3, seed=9, turn=(8, 30))
  out = {"x": x}
  checkpoints = [5, 12, 30, 31, 45, 59, 60, 61, 90, 130]
  for name in ("NoDeflicker", "OrderBased", "Hungarian"):
    main = ref.SpectralClusterer(
        min_clusters=1, max_clusters=5,
        refinement_options=ref.RefinementOptions(
            gaussian_blur_sigma=0, p_percentile=0.9,
            refinement_sequence=ref.ICASSP2018_REFINEMENT_SEQUENCE))
    ms = rm.MultiStageClusterer(main, fallback_threshold=0.5, L=8, U1=30, U2=60,
                                deflicker=getattr(rm.Deflicker, name))
    for i in range(130):
      labels = ms.streaming_predict(x[i])
      if i + 1 in checkpoints:
        out["%s_%d" % (name, i + 1)] = np.asarray(labels)
  out["checkpoints"] = np.array(checkpoints)
  np.savez(os.path.join(OUT, "multistage.npz"), **out)
  print("multistage:", {k: len(set(np.asarray(v).tolist())) for k, v in out.items() if k.endswith("_130")})


if __name__ == "__main__":
  main()
  nonsymmetric()
  constraints()
  multistage()
