"""Online "naive" clusterer -- the caller-side fallback of the hot path
(/root/reference/spectralcluster/naive_clusterer.py:5-105; used by FallbackClusterer,
fallback_clusterer.py:114-117).

Strictly sequential by definition (the label of frame t depends on the centroids after frames
0..t-1), and only ever run on inputs too small for spectral clustering
(`spectral_min_embeddings`) or as a single-vs-multi cluster probe: host code.  State is kept as
one [clusters, d] matrix of running means so that each step is a single matrix-vector product.
"""

from __future__ import annotations

import typing

import numpy as np


class NaiveCentroid:
  """One cluster of the naive algorithm as the reference exposes it (naive_clusterer.py:5-22):
  the running mean `embedding` of its `count` members."""

  def __init__(self, embedding: np.ndarray, count: int = 1):
    self.embedding = embedding
    self.count = count

  def merge(self, embedding: np.ndarray):
    total = self.count + 1
    self.embedding = (self.embedding * self.count + embedding) / total
    self.count = total

  def cosine(self, embedding: np.ndarray) -> float:
    norms = np.linalg.norm(self.embedding) * np.linalg.norm(embedding)
    return float(np.dot(self.embedding, embedding) / norms)


class NaiveClusterer:
  """Assign each embedding to the most similar running centroid, or open a new cluster."""

  def __init__(self, threshold: float, adaptation_threshold: typing.Optional[float] = None):
    if adaptation_threshold is not None and adaptation_threshold < threshold:
      raise ValueError("adaptation_threshold cannot be smaller than threshold")
    self.threshold = threshold
    self.adaptation_threshold = threshold if adaptation_threshold is None else adaptation_threshold
    self.reset()

  def reset(self):
    self._means = None                     # [clusters, d] running means
    self._counts = []                      # members merged into each mean

  @property
  def centroids(self) -> typing.List[NaiveCentroid]:
    """Snapshot of the clusters as NaiveCentroid objects (the reference keeps such a list)."""
    if self._means is None:
      return []
    return [NaiveCentroid(row.copy(), count) for row, count in zip(self._means, self._counts)]

  def _open(self, embedding) -> int:
    row = np.asarray(embedding, dtype=np.float64)[None, :]
    self._means = row.copy() if self._means is None else np.concatenate([self._means, row])
    self._counts.append(1)
    return len(self._counts) - 1

  def predict_next(self, embedding: np.ndarray) -> int:
    """Label of one new embedding (naive_clusterer.py:58-89)."""
    if self._means is None:
      return self._open(embedding)
    e = np.asarray(embedding, dtype=np.float64)
    cos = (self._means @ e) / (np.linalg.norm(self._means, axis=1) * np.linalg.norm(e))
    if cos.max() < self.threshold:
      return self._open(embedding)
    label = int(cos.argmax())
    if cos[label] > self.adaptation_threshold:
      c = self._counts[label]
      self._means[label] = (self._means[label] * c + e) / (c + 1)    # :14-18
      self._counts[label] = c + 1
    return label

  def predict(self, embeddings: np.ndarray) -> np.ndarray:
    return np.array([self.predict_next(e) for e in embeddings])

  def fit_predict(self, embeddings: np.ndarray) -> np.ndarray:
    return self.predict(embeddings)
