"""Row-sharded refinement across the GPUs of one box (BASELINE.json configs[3]).

One process per GPU (torch.distributed: NCCL on GPUs, gloo in the CPU tests).  Rank g owns the
row block [row_begin, row_end) of every N x N matrix:

  1. embeddings are replicated; every rank normalises all N rows (N*d work, no traffic);
  2. affinity rows of the block PLUS the blur halo (R rows each side) are computed locally --
     the halo is recomputed from the embeddings, not exchanged; CropDiagonal is row-local;
  3. blur statistics pass -> row maxima of the owned rows -> ALL-GATHER of N floats
     (the fused threshold/symmetrize rule needs m_i and m_j);
  4. blur + threshold + symmetrize pass -> owned rows of Y as split fp16 planes;
  5. Diffuse: S_block = Y_block * Y^T needs every row of Y: the peers' row blocks arrive by
     asynchronous BROADCASTs (one per peer, issued up front on the communicator's stream) while
     the tensor cores work through S_block[:, cols of peer p] = Y_block * Y_p^T in the order
     the blocks land -- the own diagonal block first, which needs no traffic at all;
  6. row maxima / sums of S_block (RowWiseNormalize, Laplacian degree) are row-local.

The orchestration below is backend-agnostic: `DeviceBackend` runs the CUDA kernels through the C
ABI; the CPU tests plug in a NumPy backend and run the same code under gloo with 2 ranks.
"""

from __future__ import annotations

import math
import typing

import numpy as np


def _round_up(x: int, m: int) -> int:
  return (x + m - 1) // m * m


class ShardPlan:
  """Row partition of an n x n problem over `world` ranks, with the blur halo of `radius`."""

  def __init__(self, n: int, world: int, rank: int, radius: int, align: int = 128):
    if world < 1 or not 0 <= rank < world:
      raise ValueError("bad world/rank")
    self.n, self.world, self.rank, self.radius = n, world, rank, radius
    self.block = _round_up(int(math.ceil(n / world)), align)
    if (world - 1) * self.block >= n:
      raise ValueError("n=%d is too small to give each of %d ranks a row block of %d"
                       % (n, world, self.block))
    self.row_begin, self.row_end = self.rows_of(rank)
    self.halo_begin = max(0, self.row_begin - radius)
    self.halo_end = min(n, self.row_end + radius)

  def rows_of(self, rank: int) -> typing.Tuple[int, int]:
    lo = rank * self.block
    return lo, min(lo + self.block, self.n)

  @property
  def rows(self) -> int:
    return self.row_end - self.row_begin

  @property
  def halo_rows(self) -> int:
    return self.halo_end - self.halo_begin

  def peer_order(self) -> typing.List[int]:
    """Own block first (no traffic), then the peers in broadcast (= arrival) order."""
    return [self.rank] + [p for p in range(self.world) if p != self.rank]


def blur_radius(sigma: float) -> int:
  return int(4.0 * sigma + 0.5) if sigma > 1e-15 else 0


class ShardedRefiner:
  """[CropDiagonal] [GaussianBlur] RowWiseThreshold(RowMax) Symmetrize Diffuse on row blocks.

  `backend` supplies the arithmetic, `dist` is torch.distributed (or None for world == 1).
  run() returns the owned rows of S = Y Y^T together with their row maxima / sums.
  """

  def __init__(self, backend, options, dist=None, group=None):
    from . import refinement as rf
    self.backend, self.options, self.dist, self.group = backend, options, dist, group
    names = list(options.refinement_sequence or [])
    RN = rf.RefinementName
    self.has_crop = bool(names) and names[0] == RN.CropDiagonal
    rest = names[1:] if self.has_crop else names
    self.has_blur = bool(rest) and rest[0] == RN.GaussianBlur
    rest = rest[1:] if self.has_blur else rest
    tail = [RN.RowWiseThreshold, RN.Symmetrize, RN.Diffuse]
    if rest[:3] != tail or rest[3:] not in ([], [RN.RowWiseNormalize]):
      raise NotImplementedError(
          "the sharded pipeline covers [CropDiagonal] [GaussianBlur] RowWiseThreshold "
          "Symmetrize Diffuse [RowWiseNormalize] (the ICASSP-2018 family)")
    if options.thresholding_type != rf.ThresholdType.RowMax:
      raise NotImplementedError("sharded pipeline: RowMax thresholding only")
    self.sigma = float(options.gaussian_blur_sigma) if self.has_blur else 0.0
    self.sym_max = options.symmetrize_type == rf.SymmetrizeType.Max

  def run(self, embeddings, world: int = 1, rank: int = 0):
    be, opt = self.backend, self.options
    n = int(embeddings.shape[0])
    plan = ShardPlan(n, world, rank, blur_radius(self.sigma))
    planes = be.normalize(embeddings)
    # affinity of the owned rows + halo; CropDiagonal values for exactly those rows
    a_ext, crop = be.affinity_block(planes, n, plan.halo_begin, plan.halo_rows, self.has_crop)
    # blur statistics -> row maxima of the owned rows, gathered into a full-length vector
    m_full = be.new_row_vector(world * plan.block)
    be.blur_rowmax_block(a_ext, n, plan, crop, self.sigma,
                         bool(opt.thresholding_preserve_diagonal), m_full)
    if world > 1:
      self._all_gather_blocks(m_full, plan)
    # owned rows of Y straight into the full-size planes (the peers' blocks land beside them)
    y_full = be.new_planes(world * plan.block, n)
    be.thrsym_block(a_ext, n, plan, crop, self.sigma, m_full, opt, self.sym_max, y_full)
    del a_ext
    works = []
    if world > 1:
      for p in range(world):                       # same order on every rank
        lo = p * plan.block
        for plane in y_full:
          works.append((p, self.dist.broadcast(plane[lo:lo + plan.block], src=p,
                                               group=self.group, async_op=True)))
    s_block = be.new_block(plan.rows, n)
    for p in plan.peer_order():
      if p != rank:
        for q, w in works:
          if q == p:
            w.wait()                               # stream-ordered on CUDA, blocking on gloo
      lo, hi = plan.rows_of(p)
      be.gemm_block(y_full, plan.row_begin, plan.rows, lo, hi - lo, n, s_block)
    for _, w in works:
      w.wait()
    rowmax, rowsum = be.row_stats_block(s_block, plan.rows, n)
    return dict(plan=plan, s_block=s_block, rowmax=rowmax, rowsum=rowsum, y_planes=y_full)

  def _all_gather_blocks(self, full, plan):
    """full[p*block:(p+1)*block] <- rank p's slice (equal, padded blocks)."""
    mine = full[plan.rank * plan.block:(plan.rank + 1) * plan.block].clone()
    self.dist.all_gather_into_tensor(full, mine, group=self.group)


class DeviceBackend:
  """The CUDA kernels behind the C ABI (tcgen05 GEMMs, fused blur passes)."""

  def __init__(self, engine):
    from . import _native as nat
    from . import device as dev
    self.eng, self.nat, self.dev = engine, nat, dev
    self.t = dev.torch()

  def _p(self, t):
    return self.dev._ptr(t)

  def normalize(self, x):
    eng, t = self.eng, self.t
    n, d = int(x.shape[0]), int(x.shape[1])
    hi, lo = eng.planes(n, d)
    eng.call("sc_normalize_rows", self._p(x), 1 if x.dtype == t.float64 else 0, n, d,
             x.stride(0), None, 0, self._p(hi), self._p(lo), hi.stride(0), eng.stream)
    return hi, lo, d

  def affinity_block(self, planes, n, row_begin, row_count, want_crop):
    eng, t = self.eng, self.t
    hi, lo, d = planes
    a = t.empty((row_count, self.dev.round_up(n, 64)), dtype=t.float32, device=eng.device)
    crop = t.zeros((n,), dtype=t.float32, device=eng.device) if want_crop else None
    eng.call("sc_affinity_cosine_block", eng.gemm_precision, self._p(hi), self._p(lo),
             hi.stride(0), n, d, row_begin, row_count, self._p(a), a.stride(0),
             None if crop is None else self.dev.ctypes.c_void_p(crop.data_ptr() + 4 * row_begin),
             eng.stream)
    return a, crop

  def new_row_vector(self, length):
    return self.t.zeros((length,), dtype=self.t.float32, device=self.eng.device)

  def new_planes(self, rows, n):
    t, ld = self.t, self.dev.round_up(n, 64)
    return (t.empty((rows, ld), dtype=t.float16, device=self.eng.device),
            t.empty((rows, ld), dtype=t.float16, device=self.eng.device))

  def new_block(self, rows, n):
    return self.t.empty((rows, self.dev.round_up(n, 64)), dtype=self.t.float32,
                        device=self.eng.device)

  def blur_rowmax_block(self, a_ext, n, plan, crop, sigma, zero_diag, m_full):
    eng = self.eng
    eng.call("sc_gaussian_blur_rowmax_block", self._p(a_ext), n, a_ext.stride(0),
             plan.halo_begin, plan.halo_rows, plan.row_begin, plan.row_end, self._p(crop),
             float(sigma), int(zero_diag), self._p(m_full), eng.stream)

  def thrsym_block(self, a_ext, n, plan, crop, sigma, m_full, opt, sym_max, y_full):
    eng, nat = self.eng, self.nat
    hi, lo = y_full
    off = plan.row_begin * hi.stride(0) * 2
    c = self.dev.ctypes.c_void_p
    eng.call("sc_blur_threshold_symmetrize_block", self._p(a_ext), n, a_ext.stride(0),
             plan.halo_begin, plan.halo_rows, plan.row_begin, plan.row_end, self._p(crop),
             float(sigma), self._p(m_full), float(opt.p_percentile),
             float(opt.thresholding_soft_multiplier), int(bool(opt.thresholding_with_binarization)),
             int(bool(opt.thresholding_preserve_diagonal)),
             nat.SYMMETRIZE_MAX if sym_max else nat.SYMMETRIZE_AVERAGE, None, 0,
             c(hi.data_ptr() + off), c(lo.data_ptr() + off), hi.stride(0), eng.stream)

  def gemm_block(self, y_full, a_row, a_rows, b_row, b_rows, n, s_block):
    eng = self.eng
    hi, lo = y_full
    ld = hi.stride(0)
    c = self.dev.ctypes.c_void_p
    eng.call("sc_gemm_nt_planes", eng.gemm_precision,
             c(hi.data_ptr() + 2 * a_row * ld), c(lo.data_ptr() + 2 * a_row * ld), ld, a_rows,
             c(hi.data_ptr() + 2 * b_row * ld), c(lo.data_ptr() + 2 * b_row * ld), ld, b_rows, n,
             c(s_block.data_ptr() + 4 * b_row), s_block.stride(0), eng.stream)

  def row_stats_block(self, s_block, rows, n):
    eng, t = self.eng, self.t
    mx = t.empty((rows,), dtype=t.float64, device=eng.device)
    sm = t.empty((rows,), dtype=t.float64, device=eng.device)
    eng.call("sc_row_stats_block", self._p(s_block), rows, n, s_block.stride(0), self._p(mx),
             self._p(sm), eng.stream)
    return mx, sm


def parallel_autotune(evaluate: typing.Callable[[float], typing.Tuple[float, int]],
                      grid: typing.Sequence[float], dist=None, group=None, world: int = 1,
                      rank: int = 0):
  """AutoTune with one p_percentile per rank (BASELINE.json configs[4]).

  Every rank holds the base affinity; rank r evaluates grid[r::world] with `evaluate(p) ->
  (ratio, n_clusters)`; the (ratio, k) pairs are all-gathered and every rank picks the same
  winner: the smallest ratio, the lowest grid index on ties (autotune.py:106-111 keeps the
  first strict minimum of an ascending scan).  Returns (best_index, best_p, ratio, k, owner).
  """
  mine = [(idx, evaluate(grid[idx])) for idx in range(rank, len(grid), world)]
  table = np.full((len(grid), 2), np.inf)
  for idx, (ratio, k) in mine:
    table[idx] = (ratio, k)
  if world > 1:
    import torch as t
    local = t.from_numpy(table)
    device = None
    if dist.get_backend(group) == "nccl":
      device = t.device("cuda", t.cuda.current_device())
      local = local.to(device)
    gathered = [t.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local, group=group)
    table = t.stack(gathered).min(dim=0).values.cpu().numpy()   # unevaluated entries are +inf
  best = int(np.argmin(table[:, 0]))          # first minimum == lowest index on ties
  return best, grid[best], float(table[best, 0]), int(table[best, 1]), best % world
