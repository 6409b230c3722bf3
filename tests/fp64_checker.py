"""float64 restatement of the ICASSP-2018 hot path in torch ON THE GPU -- a CHECKER for the sizes
the host oracle cannot reach (BASELINE configs[2] at N = 65,536 needs four live N x N float64
arrays on the host and ~19 h of LAPACK).  Test infrastructure only: nothing in the product
imports it.  It is pinned against the oracle (oracle/spectral_oracle.py, itself pinned against
the unmodified reference) at a size both can run: tests/test_gpu_parity_fullsize.py::
test_fp64_checker_matches_oracle.

Follows, stage by stage (reference file:line under /root/reference/spectralcluster/):
  utils.py:20-41            cosine affinity (X^ X^T + 1) / 2
  refinement.py:145-151     CropDiagonal
  refinement.py:160-162     GaussianBlur == scipy.ndimage.gaussian_filter (axis 0 then axis 1,
                            mode='reflect', truncate=4.0)
  refinement.py:168-210     RowWiseThreshold (RowMax)
  refinement.py:219-226     Symmetrize (Max)
  refinement.py:232-234     Diffuse  A A^T
  refinement.py:240-245     RowWiseNormalize   (kept as the row scaling r)
  laplacian.py:24-60        GraphCut Laplacian (symmetrised form)
  utils.py:44-71            eigenvalues / eigenvectors: ARPACK (scipy eigsh) on the symmetrised
                            operator, matvec in float64 on the GPU
"""

import numpy as np
import scipy.sparse.linalg as spla
import torch

EPS = 1e-10
CHUNK = 4096


def gaussian_weights(sigma, truncate=4.0):
  radius = int(truncate * sigma + 0.5)
  x = np.arange(-radius, radius + 1, dtype=np.float64)
  w = np.exp(-0.5 / (sigma * sigma) * x * x)
  return w / w.sum(), radius


def reflect_index(idx, n):
  """scipy 'reflect' (half-sample symmetric): d c b a | a b c d | d c b a."""
  idx = np.asarray(idx)
  period = 2 * n
  idx = np.mod(idx, period)
  return np.where(idx >= n, period - 1 - idx, idx)


def refine_through_diffuse(x, device, sigma=1.0, p=0.95, mult=0.01):
  """float64 S = Diffuse(Symmetrize(Threshold(Blur(Crop(affinity(x)))))) as a torch tensor on
  `device`.  Peak memory: three N x N float64 matrices."""
  t = torch
  xd = t.from_numpy(np.asarray(x, dtype=np.float64)).to(device)
  n = xd.shape[0]
  xn = xd / t.linalg.norm(xd, dim=1, keepdim=True)
  a = t.empty((n, n), dtype=t.float64, device=device)
  for r0 in range(0, n, CHUNK):
    a[r0:r0 + CHUNK] = (xn[r0:r0 + CHUNK] @ xn.T + 1.0) * 0.5
  # CropDiagonal: diag <- 0, then diag <- row maximum
  a.fill_diagonal_(0.0)
  a.diagonal().copy_(a.max(dim=1).values)
  # GaussianBlur, axis 0 then axis 1
  w, radius = gaussian_weights(sigma)
  wt = [float(v) for v in w]
  b = t.empty_like(a)
  for r0 in range(0, n, CHUNK):
    r1 = min(n, r0 + CHUNK)
    idx = t.from_numpy(reflect_index(np.arange(r0 - radius, r1 + radius), n)).to(device)
    ext = a[idx]                                            # (rows + 2R) x n
    acc = wt[radius] * ext[radius:radius + (r1 - r0)]
    for k in range(1, radius + 1):
      acc += wt[radius + k] * (ext[radius + k:radius + k + (r1 - r0)] + ext[radius - k:radius - k + (r1 - r0)])
    b[r0:r1] = acc
  cidx = t.from_numpy(reflect_index(np.arange(-radius, n + radius), n)).to(device)
  for r0 in range(0, n, CHUNK):
    r1 = min(n, r0 + CHUNK)
    ext = b[r0:r1][:, cidx]                                 # rows x (n + 2R)
    acc = wt[radius] * ext[:, radius:radius + n]
    for k in range(1, radius + 1):
      acc += wt[radius + k] * (ext[:, radius + k:radius + k + n] + ext[:, radius - k:radius - k + n])
    a[r0:r1] = acc                                          # a now holds the blurred matrix
  del ext, acc
  # RowWiseThreshold (RowMax), in place
  for r0 in range(0, n, CHUNK):
    blk = a[r0:r0 + CHUNK]
    m = blk.max(dim=1, keepdim=True).values
    blk.copy_(t.where(blk < m * p, blk * mult, blk))
  # Symmetrize (Max) -> b
  for r0 in range(0, n, CHUNK):
    r1 = min(n, r0 + CHUNK)
    b[r0:r1] = t.maximum(a[r0:r1], a[:, r0:r1].T)
  # Diffuse -> a
  for r0 in range(0, n, CHUNK):
    r1 = min(n, r0 + CHUNK)
    t.matmul(b[r0:r1], b.T, out=a[r0:r1])
  del b
  return a


def operator_terms(s, laplacian):
  """(delta, left, right, sign, descending): the matrix the reference hands to np.linalg.eig is
  diag(delta) + sign * diag(left) S diag(right) (RowWiseNormalize, then laplacian.py:24-60)."""
  r = 1.0 / s.max(dim=1).values                            # RowWiseNormalize: W = diag(r) S
  if laplacian is None:
    return None, r, torch.ones_like(r), 1.0, True
  if laplacian != "graphcut":
    raise ValueError(laplacian)
  d = r * s.sum(dim=1)
  inv = 1.0 / (torch.sqrt(d) + EPS)
  return inv * d * inv, inv * r, inv, -1.0, False


def extremal_eigh(s, terms, n_values, tol=1e-13):
  """Eigenvalues (sorted as the reference sorts them) and unit-norm eigenvectors of the REFERENCE
  (non-symmetric) matrix, through its symmetrised form T = diag(delta) + sign * c S c, c =
  sqrt(left * right); v = E u / |E u|, E = sqrt(left / right)."""
  delta, left, right, sign, descending = terms
  n = s.shape[0]
  c = torch.sqrt(left * right)

  def matvec(v):
    xv = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64).reshape(-1)).to(s.device)
    y = sign * c * (s @ (c * xv))
    if delta is not None:
      y = y + delta * xv
    return y.cpu().numpy()

  op = spla.LinearOperator((n, n), matvec=matvec, dtype=np.float64)
  rng = np.random.default_rng(0)
  w, u = spla.eigsh(op, k=n_values, which="LA" if descending else "SA", tol=tol,
                    ncv=max(4 * n_values, 48), v0=rng.standard_normal(n), maxiter=20000)
  order = np.argsort(-w) if descending else np.argsort(w)
  w, u = w[order], u[:, order]
  e = torch.sqrt(left / right).cpu().numpy()
  v = u * e[:, None]
  v /= np.linalg.norm(v, axis=0, keepdims=True)
  return w, v
