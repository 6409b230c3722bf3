"""SpectralClusterer -- drop-in mirror of the reference orchestrator
(/root/reference/spectralcluster/spectral_clusterer.py: __init__ :29-106,
_compute_eigenvectors_ncluster :108-168, predict :201-314) whose arithmetic runs on a B200.

predict(): one H2D copy of the [N, d] embeddings, every N x N intermediate stays in HBM
(affinity GEMM -> fused refinement -> Diffuse GEMM -> row statistics -> symmetric eigensolve ->
k-means), one D2H copy of N labels.  The callers around the path follow the reference: the
fallback clusterer for tiny inputs and the max_spectral_size AHC pre-clustering are host code
(scikit-learn, as in the reference), the single-cluster checks are device reductions over the
resident affinity.
"""

from __future__ import annotations

import typing

import numpy as np

from . import _native as nat
from . import autotune as autotune_lib
from . import constraint as constraint_lib
from . import custom_distance_kmeans
from . import device as dev
from . import fallback_clusterer
from . import laplacian as laplacian_lib
from . import refinement
from . import utils

AutoTune = autotune_lib.AutoTune
AutoTuneProxy = autotune_lib.AutoTuneProxy
ConstraintName = constraint_lib.ConstraintName
ConstraintOptions = constraint_lib.ConstraintOptions
FallbackOptions = fallback_clusterer.FallbackOptions
LaplacianType = laplacian_lib.LaplacianType
RefinementName = refinement.RefinementName
RefinementOptions = refinement.RefinementOptions
EigenGapType = utils.EigenGapType


class DeviceAffinity:
  """An affinity matrix resident in HBM (what predict() passes between its stages)."""

  def __init__(self, matrix, n: int, crop_vector=None, symmetric: bool = True):
    self.matrix = matrix
    self.n = n
    self.crop_vector = crop_vector
    self.symmetric = symmetric


class SpectralClusterer:
  """Spectral clustering of embeddings; same constructor and attributes as the reference."""

  def __init__(self,
               min_clusters: typing.Optional[int] = None,
               max_clusters: typing.Optional[int] = None,
               refinement_options: typing.Optional[RefinementOptions] = None,
               autotune: typing.Optional[AutoTune] = None,
               fallback_options: typing.Optional[FallbackOptions] = None,
               laplacian_type: typing.Optional[LaplacianType] = None,
               stop_eigenvalue: float = 1e-2,
               row_wise_renorm: bool = False,
               custom_dist: typing.Union[str, typing.Callable] = "cosine",
               max_iter: int = 300,
               constraint_options=None,
               eigengap_type: EigenGapType = EigenGapType.Ratio,
               max_spectral_size: typing.Optional[int] = None,
               affinity_function: typing.Callable = utils.compute_affinity_matrix,
               post_eigen_cluster_function: typing.Callable = (
                   custom_distance_kmeans.run_kmeans)):
    self.min_clusters = min_clusters
    self.max_clusters = max_clusters
    self.refinement_options = refinement_options or RefinementOptions()
    self.autotune = autotune
    self.fallback_options = fallback_options or FallbackOptions()
    self.laplacian_type = laplacian_type
    self.row_wise_renorm = row_wise_renorm
    self.stop_eigenvalue = stop_eigenvalue
    self.custom_dist = custom_dist
    self.max_iter = max_iter
    self.constraint_options = constraint_options
    self.eigengap_type = eigengap_type
    self.max_spectral_size = max_spectral_size
    self.affinity_function = affinity_function
    self.post_eigen_cluster_function = post_eigen_cluster_function
    # Diagnostics of the last predict(): eigenvalues used by the eigengap, cluster count, solver.
    self.last_details: typing.Dict[str, typing.Any] = {}
    # Set collect_timings to get, after every predict(), last_timings = {C-ABI entry point: ms of
    # device time (CUDA events)} -- the per-stage breakdown bench.py prints.
    self.collect_timings = False
    self.last_timings: typing.Dict[str, float] = {}
    # Set to a torch.distributed process group (or True for the default group) to spread the
    # AutoTune grid over the ranks, one p_percentile evaluation at a time per GPU
    # (BASELINE.json configs[4]); every rank must call predict() with the same embeddings.
    self.autotune_group = None

  # ------------------------------------------------------------------ eigen stage
  def _constrain(self, eng, affinity: "DeviceAffinity", constraint_matrix) -> "DeviceAffinity":
    """constraint_operator.adjust_affinity on the resident affinity."""
    q_host = np.asarray(constraint_matrix, dtype=np.float64)
    op = self.constraint_options.constraint_operator
    op.check_input(np.empty((affinity.n, affinity.n), dtype=np.bool_), q_host)
    q = eng.upload_matrix(q_host)
    out = op.adjust_on_device(eng, affinity.matrix, q, affinity.n)
    sym = affinity.symmetric and bool(np.array_equal(q_host, q_host.T))
    return DeviceAffinity(out, affinity.n, None, sym)

  def _eigen_on_device(self, eng, affinity: DeviceAffinity, constraint_matrix=None):
    """Refinement + Laplacian terms + eigensolve; returns (w host, V device [n, nv], k, gap)."""
    n = affinity.n
    refined = dev.run_refinement(eng, affinity.matrix, n, self.refinement_options,
                                 crop_vector=affinity.crop_vector,
                                 a_symmetric=affinity.symmetric,
                                 diffuse_precision=eng.diffuse_precision_for(n))
    if (self.constraint_options and not self.constraint_options.apply_before_refinement and
        constraint_matrix is not None):
      # constraint on the refined affinity (spectral_clusterer.py:137-142)
      s = refined.s
      if refined.row_scale is not None:
        s = eng.row_normalize(s, n)
      adjusted = self._constrain(eng, DeviceAffinity(s, n, None, refined.symmetric and
                                                      refined.row_scale is None), constraint_matrix)
      refined = dev.Refined(adjusted.matrix, n, adjusted.symmetric)
    delta, left, right, sign, which = laplacian_lib.operator_terms(eng, refined,
                                                                   self.laplacian_type)
    descend = which == nat.EIG_LARGEST
    limit = n
    if self.max_clusters and self.max_clusters + 1 < limit:
      limit = self.max_clusters + 1
    n_vectors = min(n, max(limit, self.min_clusters or 0))
    if not refined.symmetric:
      # not symmetric nor diagonally similar to a symmetric matrix (e.g. RowWiseThreshold without
      # Symmetrize): np.linalg.eig + .real semantics (utils.py:59-61) by Krylov-Schur
      from . import arnoldi
      w, v, stats = arnoldi.eig_extremal_device(eng, refined.s, n, delta, left, right, sign,
                                                not descend, limit, n_vectors)
      if descend:
        k, gap = utils.compute_number_of_clusters(
            w, max_clusters=self.max_clusters, stop_eigenvalue=self.stop_eigenvalue,
            eigengap_type=self.eigengap_type, descend=True)
      else:
        if self.eigengap_type == EigenGapType.NormalizedDiff and limit < n:
          top, _, _ = arnoldi.eig_extremal_device(eng, refined.s, n, delta, left, right, sign,
                                                  False, 1, 0)
          w = np.concatenate([w, top])
        k, gap = utils.compute_number_of_clusters(
            w, max_clusters=self.max_clusters, eigengap_type=self.eigengap_type, descend=False)
      self.last_details = dict(eigenvalues=np.array(w[:limit]), n_clusters_raw=k, max_gap=gap,
                               solver="krylov-schur", lanczos_stats=stats)
      return w, v, k, gap
    # The block Lanczos solver needs n >= 4 max(2*limit+32, 64) and reads S a handful of times; the
    # dense solver (Householder tridiagonalisation, 4/3 n^3) is kept for small matrices and for
    # max_clusters=None / > 31 (every eigenvalue, or more than the extremal solver's 32, is needed).
    basis = max(2 * limit + 32, 64)
    dense = (n <= eng.dense_eig_max or n < 4 * basis or not self.max_clusters or limit > 32)

    def count_clusters(values):
      if descend:
        return utils.compute_number_of_clusters(
            values, max_clusters=self.max_clusters, stop_eigenvalue=self.stop_eigenvalue,
            eigengap_type=self.eigengap_type, descend=True)
      return utils.compute_number_of_clusters(
          values, max_clusters=self.max_clusters, eigengap_type=self.eigengap_type, descend=False)

    if dense and limit == n and n > 256:
      # every eigenvalue is needed (max_clusters=None scans the whole spectrum, utils.py:100-102)
      # but only the columns predict() will select: the eigengap runs between the two phases
      picked = {}

      def picker(values):
        picked["k"], picked["gap"] = count_clusters(values)
        return max(picked["k"], self.min_clusters or 0, 1)

      w, v = eng.eigh_dense_pick(refined.s, n, delta, left, right, sign, which, picker)
      k, gap = picked["k"], picked["gap"]
      self.last_details = dict(eigenvalues=np.array(w[:limit]), n_clusters_raw=k, max_gap=gap,
                               solver="dense", lanczos_stats=None)
      return w, v, k, gap
    if dense:
      w, v, stats = eng.eigh(refined.s, n, delta, left, right, sign, which, n, n_vectors, True)
    else:
      w, v, stats = eng.eigh(refined.s, n, delta, left, right, sign, which, limit, n_vectors,
                             False)
      if not descend and self.eigengap_type == EigenGapType.NormalizedDiff:
        # np.max(eigenvalues) of the full spectrum (utils.py:109): one more extremal solve
        top, _, _ = eng.eigh(refined.s, n, delta, left, right, sign, nat.EIG_LARGEST, 1, 0, False)
        w = np.concatenate([w, top])
    if descend:
      k, gap = utils.compute_number_of_clusters(
          w, max_clusters=self.max_clusters, stop_eigenvalue=self.stop_eigenvalue,
          eigengap_type=self.eigengap_type, descend=True)
    else:
      k, gap = utils.compute_number_of_clusters(
          w, max_clusters=self.max_clusters, eigengap_type=self.eigengap_type, descend=False)
    self.last_details = dict(eigenvalues=np.array(w[:limit]), n_clusters_raw=k, max_gap=gap,
                             solver="dense" if dense else "lanczos",
                             lanczos_stats=None if dense else stats.tolist())
    return w, v, k, gap

  def _compute_eigenvectors_ncluster(self, affinity, constraint_matrix=None):
    """(eigenvectors, n_clusters, max eigengap) for an affinity matrix.

    `affinity` is a host ndarray (reference signature) or a DeviceAffinity.  Eigenvectors come
    back as a host ndarray [n, n_vec] in the first case, a device fp64 tensor in the second;
    n_vec covers every column predict() can select (all n when max_clusters is None)."""
    eng = dev.Engine.get()
    if isinstance(affinity, DeviceAffinity):
      _, v, k, gap = self._eigen_on_device(eng, affinity, constraint_matrix)
      return v, k, gap
    a = np.asarray(affinity)
    if a.ndim != 2:
      raise ValueError("affinity must be 2-dimensional")
    if a.shape[0] != a.shape[1]:
      raise ValueError("affinity must be a square matrix")
    sym = bool(np.allclose(a, a.T, rtol=1e-6, atol=1e-9))
    da = DeviceAffinity(eng.upload_matrix(a), a.shape[0], None, sym)
    _, v, k, gap = self._eigen_on_device(eng, da, constraint_matrix)
    return v.to("cpu").numpy(), k, gap

  def _reduce_size_and_predict(self, embeddings: np.ndarray) -> np.ndarray:
    """spectral_clusterer.py:170-199: complete-linkage cosine AHC down to `max_spectral_size`
    clusters (scikit-learn, host -- the reference's own lossy answer to large N; the device path
    does not need it, BASELINE configs[3] shards the exact problem instead), spectral clustering
    of the cluster centroids on the device, labels chained back."""
    from sklearn.cluster import AgglomerativeClustering
    pre = AgglomerativeClustering(n_clusters=self.max_spectral_size, metric="cosine",
                                  linkage="complete").fit_predict(embeddings)
    centroids = utils.get_cluster_centroids(embeddings, pre)
    return utils.chain_labels(pre, self.predict(centroids))

  # ------------------------------------------------------------------ predict
  def predict(self, embeddings: np.ndarray, constraint_matrix=None) -> np.ndarray:
    """Cluster the rows of `embeddings` ([n_samples, n_features] ndarray) -> int64 labels."""
    if self.collect_timings and dev.torch().cuda.is_available():
      eng = dev.Engine.get()
      if eng.profile is None:
        eng.start_profile()
        try:
          return self._predict(embeddings, constraint_matrix)
        finally:
          self.last_timings = eng.stop_profile()
    return self._predict(embeddings, constraint_matrix)

  def _predict(self, embeddings: np.ndarray, constraint_matrix=None) -> np.ndarray:
    num_embeddings = embeddings.shape[0]
    if not isinstance(embeddings, np.ndarray):
      raise TypeError("embeddings must be a numpy array")
    if len(embeddings.shape) != 2:
      raise ValueError("embeddings must be 2-dimensional")
    if num_embeddings < self.fallback_options.spectral_min_embeddings:
      # too few embeddings for a spectrum (spectral_clusterer.py:230-234): host fallback clusterer
      return fallback_clusterer.FallbackClusterer(self.fallback_options).predict(embeddings)
    if self.max_spectral_size is not None and num_embeddings > self.max_spectral_size:
      if constraint_matrix is not None:
        raise RuntimeError("Cannot handle constraint_matrix when max_spectral_size is set")
      if (self.max_spectral_size < 2 or
          (self.max_clusters and self.max_spectral_size <= self.max_clusters) or
          (self.min_clusters and self.max_spectral_size <= self.min_clusters)):
        raise ValueError("max_spectral_size should be a relatively big number")
      return self._reduce_size_and_predict(embeddings)
    eng = dev.Engine.get()
    t = dev.torch()
    sequence = list(self.refinement_options.refinement_sequence or [])
    if self.affinity_function is utils.compute_affinity_matrix:
      x = np.ascontiguousarray(embeddings)
      if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
      x_dev = t.from_numpy(x).to(eng.device, non_blocking=True)
      crop_first = bool(sequence) and sequence[0] == RefinementName.CropDiagonal
      a, crop = eng.affinity(x_dev, want_crop_vector=crop_first)
      affinity = DeviceAffinity(a, num_embeddings, crop, True)
    else:
      host = np.asarray(self.affinity_function(embeddings))
      affinity = DeviceAffinity(eng.upload_matrix(host), num_embeddings, None,
                                bool(np.allclose(host, host.T, rtol=1e-6, atol=1e-9)))

    if self.min_clusters == 1:
      # single-vs-multi cluster decision on the resident affinity (spectral_clusterer.py:253-256)
      if fallback_clusterer.check_single_cluster(self.fallback_options, embeddings, affinity):
        return np.array([0] * num_embeddings)

    if (self.constraint_options and self.constraint_options.apply_before_refinement and
        constraint_matrix is not None):
      # constraint on the raw affinity (spectral_clusterer.py:258-264), on the device
      affinity = self._constrain(eng, affinity, constraint_matrix)

    if self.autotune:
      if RefinementName.RowWiseThreshold not in sequence:
        raise ValueError("AutoTune is only effective when the refinement sequence"
                         "contains RowWiseThreshold")
      proxy = self.autotune.proxy

      def p_percentile_to_ratio(p_percentile: float):
        self.refinement_options.p_percentile = p_percentile   # shared state, as the reference
        vectors, k, gap = self._compute_eigenvectors_ncluster(affinity, constraint_matrix)
        if proxy == AutoTuneProxy.PercentileSqrtOverNME:
          return np.sqrt(1 - p_percentile) / gap, vectors, k
        if proxy == AutoTuneProxy.PercentileOverNME:
          return (1 - p_percentile) / gap, vectors, k
        raise ValueError("Unsupported value of AutoTuneProxy")

      if self.autotune_group is not None:
        return self._predict_parallel_autotune(eng, affinity, p_percentile_to_ratio)
      eigenvectors, n_clusters, best_p = self.autotune.tune(p_percentile_to_ratio)
      self.last_details["best_p_percentile"] = best_p
    else:
      eigenvectors, n_clusters, _ = self._compute_eigenvectors_ncluster(affinity, constraint_matrix)
    del affinity

    return self._cluster_embeddings(eng, eigenvectors, n_clusters)

  def _cluster_embeddings(self, eng, eigenvectors, n_clusters):
    """spectral_clusterer.py:295-313: clamp k, slice, optional row renorm, k-means."""
    if self.min_clusters is not None:
      n_clusters = max(n_clusters, self.min_clusters)
    self.last_details["n_clusters"] = n_clusters

    spectral = eigenvectors[:, :n_clusters].contiguous()
    if self.row_wise_renorm and n_clusters > 0:
      eng.row_renorm(spectral)
    if self.post_eigen_cluster_function is custom_distance_kmeans.run_kmeans:
      return custom_distance_kmeans.run_kmeans(
          spectral_embeddings=spectral, n_clusters=n_clusters, custom_dist=self.custom_dist,
          max_iter=self.max_iter)
    return self.post_eigen_cluster_function(
        spectral_embeddings=spectral.to("cpu").numpy(), n_clusters=n_clusters,
        custom_dist=self.custom_dist, max_iter=self.max_iter)

  def _predict_parallel_autotune(self, eng, affinity, p_percentile_to_ratio):
    """One p_percentile per rank: rank r evaluates grid[r::world] on its own copy of the base
    affinity; the (ratio, k) pairs are all-gathered, every rank picks the same winner, the rank
    that evaluated it runs k-means and broadcasts the labels."""
    import torch.distributed as dist
    from . import sharded
    t = dev.torch()
    if self.autotune.search_level != 1:
      raise NotImplementedError("parallel AutoTune supports search_level == 1")
    group = None if self.autotune_group is True else self.autotune_group
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    grid = self.autotune.get_percentile_range()
    kept = {}

    def evaluate(p):
      ratio, vectors, k = p_percentile_to_ratio(p)
      if not kept or ratio < kept["ratio"]:
        kept.update(ratio=ratio, vectors=vectors, k=k, p=p)
      return ratio, k

    best, best_p, ratio, k, owner = sharded.parallel_autotune(
        evaluate, grid, dist=dist, group=group, world=world, rank=rank)
    if len(grid) > 1 and self.autotune.search_step >= autotune_lib.MIN_SEARCH_STEP:
      self.autotune.narrow(grid, best)          # the reference stores the narrowed range (A.4-2)
    self.last_details["best_p_percentile"] = best_p
    n = affinity.n
    on_gpu = dist.get_backend(group) == "nccl"
    labels = t.empty((n,), dtype=t.int64, device=eng.device if on_gpu else "cpu")
    if rank == owner:
      assert kept["p"] == best_p
      labels.copy_(t.from_numpy(self._cluster_embeddings(eng, kept["vectors"], kept["k"])))
    dist.broadcast(labels, src=dist.get_global_rank(group, owner) if group is not None else owner,
                   group=group)
    return labels.cpu().numpy()
