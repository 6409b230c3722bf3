"""Option bag of the reference's fallback / single-cluster logic
(/root/reference/spectralcluster/fallback_clusterer.py:23-92).  Only the options travel: the
fallback clusterers themselves (scikit-learn AHC / GMM, the sequential naive clusterer) are
outside the B200 hot path, so SpectralClusterer raises
NotImplementedError where the reference would branch into them."""

import dataclasses
import enum
import typing


class SingleClusterCondition(enum.Enum):
  AffinityGmmBic = enum.auto()
  AllAffinity = enum.auto()
  NeighborAffinity = enum.auto()
  AffinityStd = enum.auto()
  FallbackClusterer = enum.auto()


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
