"""K-means with a custom distance on the spectral embeddings -- device-backed mirror of
/root/reference/spectralcluster/custom_distance_kmeans.py (run_kmeans :13-52, CustomKMeans
:55-141).  Seeding, the single scikit-learn Lloyd step and the custom-distance loop all run in
CUDA (csrc/kmeans.cu); only the MT19937 draws of RandomState(0) are taken on the host."""

from __future__ import annotations

import typing

import numpy as np

from . import device as dev

_METRICS = {"cosine": 0, "euclidean": 1}


def run_kmeans(spectral_embeddings, n_clusters: int,
               custom_dist: typing.Union[str, typing.Callable], max_iter: int) -> np.ndarray:
  """Cluster the rows of `spectral_embeddings` (host ndarray or device fp64 tensor) into
  `n_clusters` groups; returns int64 labels on the host."""
  if not isinstance(custom_dist, str) or custom_dist not in _METRICS:
    raise NotImplementedError(
        "custom_dist=%r: the B200 path implements 'cosine' and 'euclidean' "
        "(arbitrary scipy metrics/callables are out of scope, SURVEY.md section 2 row 5)"
        % (custom_dist,))
  eng = dev.Engine.get()
  t = dev.torch()
  if isinstance(spectral_embeddings, np.ndarray):
    e = t.from_numpy(np.ascontiguousarray(spectral_embeddings, dtype=np.float64)).to(eng.device)
  else:
    e = spectral_embeddings.contiguous()
  labels, _ = eng.kmeans(e, int(n_clusters), _METRICS[custom_dist], int(max_iter))
  return labels
