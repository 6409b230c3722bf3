"""Laplacian formation -- device-backed mirror of /root/reference/spectralcluster/laplacian.py.

`compute_laplacian` (reference :24-60) materialises the Laplacian for callers that want the
matrix; `SpectralClusterer` never does: it hands the eigensolver the symbolic form
diag(delta) - diag(left) S diag(right) built by `operator_terms` (O(N) vectors).
"""

from __future__ import annotations

import enum

import numpy as np

from . import _native as nat
from . import device as dev

EPS = 1e-10


class LaplacianType(enum.Enum):
  Affinity = enum.auto()       # W itself
  Unnormalized = enum.auto()   # L = D - W
  RandomWalk = enum.auto()     # D^-1 L
  GraphCut = enum.auto()       # D^-1/2 L D^-1/2


_NATIVE = {
    LaplacianType.Affinity: nat.LAPLACIAN_AFFINITY,
    LaplacianType.Unnormalized: nat.LAPLACIAN_UNNORMALIZED,
    LaplacianType.RandomWalk: nat.LAPLACIAN_RANDOMWALK,
    LaplacianType.GraphCut: nat.LAPLACIAN_GRAPHCUT,
}


def compute_laplacian(affinity: np.ndarray,
                      laplacian_type: LaplacianType = LaplacianType.GraphCut,
                      eps: float = EPS) -> np.ndarray:
  """Laplacian of a host affinity matrix, computed on the device (fp32 storage)."""
  if not isinstance(laplacian_type, LaplacianType):
    raise TypeError("laplacian_type must be a LaplacianType")
  n = affinity.shape[0]
  eng = dev.Engine.get()
  w = eng.upload_matrix(affinity)
  if laplacian_type == LaplacianType.Affinity:
    return eng.download_matrix(w, n)
  return eng.download_matrix(eng.laplacian(w, n, _NATIVE[laplacian_type], eps), n)


def operator_terms(eng, refined, laplacian_type, eps: float = EPS):
  """(delta, left, right, sign, which) such that the matrix the reference decomposes is
  diag(delta) + sign * diag(left) S diag(right), with S = refined.s symmetric.

  No Laplacian / Affinity : M = R S                      (largest eigenvalues, utils.py:146)
  Unnormalized            : M = D - R S                  (smallest, utils.py:160)
  RandomWalk              : M = D~ D - (D~ R) S, D~ = 1/(d+eps)      (laplacian.py:51-53)
  GraphCut                : M = D^ D D^ - (D^ R) S D^, D^ = 1/(sqrt(d)+eps)  (laplacian.py:56-58)
  where R = refined.row_scale (or I) and d = R * rowsum(S) (laplacian.py:41).
  """
  r = refined.row_scale
  if laplacian_type is None or laplacian_type == LaplacianType.Affinity:
    return None, r, None, 1.0, nat.EIG_LARGEST
  if not isinstance(laplacian_type, LaplacianType):
    raise TypeError("laplacian_type must be a LaplacianType")
  rowsum = getattr(refined, "rowsum", None)
  if rowsum is None:
    _, rowsum = eng.row_stats(refined.s, refined.n, want_max=False, want_sum=True)
  return terms_from_row_sums(rowsum, r, laplacian_type, eps)


def terms_from_row_sums(rowsum, r, laplacian_type, eps: float = EPS):
  """operator_terms given rowsum(S) (device fp64 vector) and the row scaling r (or None)."""
  t = dev.torch()
  if laplacian_type is None or laplacian_type == LaplacianType.Affinity:
    return None, r, None, 1.0, nat.EIG_LARGEST
  if not isinstance(laplacian_type, LaplacianType):
    raise TypeError("laplacian_type must be a LaplacianType")
  d = rowsum if r is None else rowsum * r
  if laplacian_type == LaplacianType.Unnormalized:
    return d, r, None, -1.0, nat.EIG_SMALLEST
  if laplacian_type == LaplacianType.RandomWalk:
    inv = 1.0 / (d + eps)
    return inv * d, (inv if r is None else inv * r), None, -1.0, nat.EIG_SMALLEST
  if laplacian_type == LaplacianType.GraphCut:
    inv = 1.0 / (t.sqrt(d) + eps)
    return inv * d * inv, (inv if r is None else inv * r), inv, -1.0, nat.EIG_SMALLEST
  raise ValueError("Unsupported laplacian_type.")
