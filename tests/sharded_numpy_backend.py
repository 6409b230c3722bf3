"""NumPy stand-in for spectralcluster_b200.sharded.DeviceBackend: the same block interface,
computed with the oracle's arithmetic on torch CPU float64 buffers, so that the sharded
orchestration (row plan, halo recompute, collectives and their order) runs under gloo."""

import numpy as np
import torch


def reflect(i, n):
  i = np.mod(i, 2 * n)
  return np.where(i < n, i, 2 * n - 1 - i)


def weights(sigma):
  r = int(4.0 * sigma + 0.5)
  x = np.arange(-r, r + 1)
  w = np.exp(-0.5 * x * x / (sigma * sigma))
  return w / w.sum(), r


class NumpyBackend:

  def normalize(self, x):
    x = np.asarray(x, dtype=np.float64)
    return x / np.linalg.norm(x, axis=1)[:, None]

  def affinity_block(self, xn, n, row_begin, row_count, want_crop):
    a = (xn[row_begin:row_begin + row_count] @ xn.T + 1.0) / 2.0
    crop = None
    if want_crop:
      crop = np.zeros(n)
      for r in range(row_count):
        row = a[r].copy()
        row[row_begin + r] = 0.0
        crop[row_begin + r] = row.max()
    return a, crop

  def new_row_vector(self, length):
    return torch.zeros(length, dtype=torch.float64)

  def new_planes(self, rows, n):
    return (torch.zeros((rows, n), dtype=torch.float64),)

  def new_block(self, rows, n):
    return np.zeros((rows, n))

  def _blurred_rows(self, a_ext, n, plan, crop, sigma):
    """Blurred values of the owned rows from the halo'd block (global reflect indexing)."""
    c = a_ext.copy()
    if crop is not None:
      for gr in range(plan.halo_begin, plan.halo_end):
        c[gr - plan.halo_begin, gr] = crop[gr]
    if sigma <= 1e-15:
      return c[plan.row_begin - plan.halo_begin:plan.row_end - plan.halo_begin]
    w, r = weights(sigma)
    rows = np.arange(plan.row_begin, plan.row_end)
    vert = np.zeros((len(rows), n))
    for k in range(-r, r + 1):
      src = reflect(rows + k, n) - plan.halo_begin
      vert += w[k + r] * c[src]
    out = np.zeros_like(vert)
    cols = np.arange(n)
    for k in range(-r, r + 1):
      out += w[k + r] * vert[:, reflect(cols + k, n)]
    return out

  def blur_rowmax_block(self, a_ext, n, plan, crop, sigma, zero_diag, m_full):
    b = self._blurred_rows(a_ext, n, plan, crop, sigma).copy()
    if zero_diag:
      for i in range(plan.row_begin, plan.row_end):
        b[i - plan.row_begin, i] = 0.0
    m_full[plan.row_begin:plan.row_end] = torch.from_numpy(b.max(axis=1))

  def thrsym_block(self, a_ext, n, plan, crop, sigma, m_full, opt, sym_max, y_full):
    b = self._blurred_rows(a_ext, n, plan, crop, sigma)
    m = m_full.numpy()[:n]
    p, mult = opt.p_percentile, opt.thresholding_soft_multiplier

    def rule(cut):
      keep = 1.0 if opt.thresholding_with_binarization else b
      return np.where(b < cut, b * mult, keep)
    t1 = rule(m[plan.row_begin:plan.row_end, None] * p)
    t2 = rule(m[None, :] * p)
    y = np.maximum(t1, t2) if sym_max else 0.5 * (t1 + t2)
    if opt.thresholding_preserve_diagonal:
      for i in range(plan.row_begin, plan.row_end):
        y[i - plan.row_begin, i] = 1.0
    y_full[0][plan.row_begin:plan.row_end] = torch.from_numpy(y)

  def gemm_block(self, y_full, a_row, a_rows, b_row, b_rows, n, s_block, s_row):
    y = y_full[0].numpy()
    s_block[s_row:s_row + a_rows, b_row:b_row + b_rows] = (
        y[a_row:a_row + a_rows] @ y[b_row:b_row + b_rows].T)

  def reserve_comm_sms(self, on):
    pass

  def transport(self, dist, group):
    return dist.get_backend(group)          # "gloo": the send/recv schedule

  def mark(self):
    return None

  def new_dense(self, rows, cols):
    return torch.zeros((rows, cols), dtype=torch.float64)

  def transposed_block(self, s_block, row_begin, rows, col_begin, cols):
    return torch.from_numpy(np.ascontiguousarray(
        s_block[row_begin:row_begin + rows, col_begin:col_begin + cols].T))

  def place_block(self, s_block, row_begin, col_begin, dense):
    r, c = dense.shape
    s_block[row_begin:row_begin + r, col_begin:col_begin + c] = dense.numpy()

  def row_stats_block(self, s_block, rows, n):
    return s_block.max(axis=1), s_block.sum(axis=1)
