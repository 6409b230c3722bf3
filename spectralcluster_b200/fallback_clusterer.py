"""Fallback clusterer and the single-vs-multi cluster decision around the hot path
(/root/reference/spectralcluster/fallback_clusterer.py: options :23-92, FallbackClusterer
:95-124, check_single_cluster :127-187).

What runs where:
  * AllAffinity / NeighborAffinity / AffinityStd are reductions over the N x N affinity: one
    streaming pass on the device (`sc_affinity_stats`), four doubles come back;
  * AffinityGmmBic fits scikit-learn GMMs to the N^2/2 upper-triangular affinities: host, on a
    copy of the device affinity ("keep on host");
  * FallbackClusterer (scikit-learn average-linkage AHC, or the sequential naive clusterer) only
    ever sees inputs below `spectral_min_embeddings` or acts as a probe: host.
"""

from __future__ import annotations

import dataclasses
import enum
import typing

import numpy as np

from . import naive_clusterer


class SingleClusterCondition(enum.Enum):
  AffinityGmmBic = enum.auto()      # BIC of a 1- vs 2-component GMM on the affinities
  AllAffinity = enum.auto()         # every affinity above the threshold
  NeighborAffinity = enum.auto()    # every affinity[i, i+1] above the threshold
  AffinityStd = enum.auto()         # standard deviation of the affinities below the threshold
  FallbackClusterer = enum.auto()   # the fallback clusterer finds one cluster


class FallbackClustererType(enum.Enum):
  Agglomerative = enum.auto()
  Naive = enum.auto()


@dataclasses.dataclass
class FallbackOptions:
  spectral_min_embeddings: int = 1
  single_cluster_condition: SingleClusterCondition = SingleClusterCondition.AffinityGmmBic
  single_cluster_affinity_threshold: float = 0.75
  single_cluster_affinity_diagonal_offset: int = 1
  fallback_clusterer_type: FallbackClustererType = FallbackClustererType.Naive
  agglomerative_threshold: float = 0.5
  naive_threshold: float = 0.5
  naive_adaptation_threshold: typing.Optional[float] = None


class FallbackClusterer:
  """fallback_clusterer.py:95-124."""

  def __init__(self, options: FallbackOptions):
    self.options = options
    kind = options.fallback_clusterer_type
    if kind == FallbackClustererType.Agglomerative:
      from sklearn.cluster import AgglomerativeClustering
      self.clusterer = AgglomerativeClustering(
          n_clusters=None, metric="cosine", linkage="average",
          distance_threshold=options.agglomerative_threshold)
    elif kind == FallbackClustererType.Naive:
      self.clusterer = naive_clusterer.NaiveClusterer(
          threshold=options.naive_threshold,
          adaptation_threshold=options.naive_adaptation_threshold)
    # any other value: the reference builds a ValueError without raising it (:121, SURVEY.md
    # A.4-4) and fails later with AttributeError in predict(); same here.

  def predict(self, embeddings: np.ndarray) -> np.ndarray:
    return self.clusterer.fit_predict(embeddings)


def affinity_statistics(affinity) -> typing.Dict[str, float]:
  """{min, neighbor_min, std} of a host ndarray or a device-resident affinity."""
  if isinstance(affinity, np.ndarray):
    return dict(min=float(affinity.min()),
                neighbor_min=float(np.diag(affinity, k=1).min()) if affinity.shape[0] > 1 else np.inf,
                std=float(np.std(affinity)))
  import ctypes
  from . import device as dev
  eng = dev.Engine.get()
  out = np.zeros(4, dtype=np.float64)
  n = affinity.n
  eng.call("sc_affinity_stats", dev._ptr(affinity.matrix), n, affinity.matrix.stride(0),
           out.ctypes.data_as(ctypes.c_void_p), eng.stream)
  count = float(n) * float(n)
  mean = out[1] / count
  var = max(out[2] / count - mean * mean, 0.0)
  return dict(min=float(out[0]), neighbor_min=float(out[3]), std=float(np.sqrt(var)))


def check_single_cluster(fallback_options: FallbackOptions,
                         embeddings: typing.Optional[np.ndarray], affinity) -> bool:
  """True when the data form a single cluster (only consulted when min_clusters == 1).

  `affinity` is a host ndarray (reference signature) or the DeviceAffinity predict() holds."""
  cond = fallback_options.single_cluster_condition
  thr = fallback_options.single_cluster_affinity_threshold
  if cond == SingleClusterCondition.AllAffinity:
    return affinity_statistics(affinity)["min"] > thr
  if cond == SingleClusterCondition.NeighborAffinity:
    return affinity_statistics(affinity)["neighbor_min"] > thr
  if cond == SingleClusterCondition.AffinityStd:
    return affinity_statistics(affinity)["std"] < thr
  if cond == SingleClusterCondition.AffinityGmmBic:
    from sklearn.mixture import GaussianMixture
    if not isinstance(affinity, np.ndarray):
      from . import device as dev
      affinity = dev.Engine.get().download_matrix(affinity.matrix, affinity.n)
    n = affinity.shape[0]
    offset = fallback_options.single_cluster_affinity_diagonal_offset
    if offset >= n - 1:
      raise ValueError("single_cluster_affinity_diagonal_offset must be significantly "
                       "smaller than affinity matrix dimension")
    values = affinity[np.triu_indices(n, offset)][:, None]
    bic = [GaussianMixture(n_components=c).fit(values).bic(values) for c in (1, 2)]
    return bool(bic[0] < bic[1])
  if cond == SingleClusterCondition.FallbackClusterer:
    labels = FallbackClusterer(fallback_options).predict(embeddings)
    return bool(np.unique(labels).size == 1)
  raise TypeError("Unsupported single_cluster_condition")
