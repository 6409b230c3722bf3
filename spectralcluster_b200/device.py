"""Device side of the hot path: buffers (torch tensors used only as HBM allocations), thin
wrappers over the C ABI, and the refinement planner that fuses the reference's operator
sequence (spectral_clusterer.py:131-135) into kernel launches.

Layout in HBM: every N x N matrix is fp32 row-major with a padded leading dimension
(`ld` = N rounded up to 64 elements, so rows start on 256-byte boundaries and 128-bit
loads/TMA boxes are aligned); the GEMM operands additionally exist as two fp16 planes
(hi, lo) of the same shape, i.e. the same 4 bytes per element.  Row statistics, scaling
vectors, eigenvectors and the k-means state are fp64 vectors / N x k arrays.
"""

from __future__ import annotations

import ctypes
import os
import typing

import numpy as np

from . import _native as nat

_torch = None

# MMAs per product of the Diffuse GEMM inside predict().  "auto" picks by N from the measured error
# of each scheme (profiles/r02_diffuse_precision.md): the rounding of the fp16 operand planes is a
# zero-mean perturbation per product that averages down with sqrt(K) = sqrt(N).  In units of the
# parity tolerance (|dw| <= 1e-5 |w| + 1e-6 max|w|) the worst eigenvalue error over configs[1],[2]
# and seeds 0-4 was 0.85 (single) / 0.38 (split2) / 0.014 (split3) at N = 2,400 and 0.33 / 0.15 at
# N = 16,384 -- labels identical in every run.  So: one MMA from N = 16,384 (<= 1/3 of the
# tolerance, falling), two from N = 4,096, three below.
DEFAULT_DIFFUSE_PRECISION = "auto"
AUTO_SINGLE_FROM, AUTO_SPLIT2_FROM = 16384, 4096


def torch():
  global _torch
  if _torch is None:
    import torch as _t
    _torch = _t
  return _torch


def _ptr(t) -> ctypes.c_void_p:
  return ctypes.c_void_p(0 if t is None else t.data_ptr())


def round_up(x: int, m: int) -> int:
  return (x + m - 1) // m * m


class Engine:
  """One per CUDA device: the sc_context plus allocation / launch helpers."""

  _instances: typing.Dict[int, "Engine"] = {}

  # matrices smaller than this use the SIMT fp64-accumulate GEMM (a 128x256 tcgen05 tile would
  # be mostly padding); everything else goes through the tensor-core kernel.
  simt_below = 64
  # force the dense (full-spectrum) eigensolver up to this N even when Lanczos would do (tests)
  dense_eig_max = 0

  def __init__(self, device: int = 0):
    t = torch()
    if not t.cuda.is_available():
      raise RuntimeError(
          "spectralcluster_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    self.device_index = device
    self.device = t.device("cuda", device)
    self.ctx = nat.context(device)
    self.gemm_precision = nat.GEMM_SPLIT3       # affinity: its entries feed a threshold
    # Diffuse inside predict() (K = N): MMAs per product, SCB_DIFFUSE_PRECISION = split3 | split2
    # | single.  The operator-level API (refinement.Diffuse) always uses split3.
    self.diffuse_precision = {"auto": None, "split3": nat.GEMM_SPLIT3, "split2": nat.GEMM_SPLIT2,
                              "single": nat.GEMM_SINGLE}[
                                  os.environ.get("SCB_DIFFUSE_PRECISION", DEFAULT_DIFFUSE_PRECISION)]
    self.profile = None

  @classmethod
  def get(cls, device: typing.Optional[int] = None) -> "Engine":
    if device is None:
      device = torch().cuda.current_device() if torch().cuda.is_available() else 0
    eng = cls._instances.get(device)
    if eng is None:
      eng = cls(device)
      cls._instances[device] = eng
    return eng

  def diffuse_precision_for(self, n: int) -> int:
    """The sc_gemm_precision predict() uses for an n x n Diffuse (None = decide by n)."""
    if self.diffuse_precision is not None:
      return self.diffuse_precision
    if n >= AUTO_SINGLE_FROM:
      return nat.GEMM_SINGLE
    return nat.GEMM_SPLIT2 if n >= AUTO_SPLIT2_FROM else nat.GEMM_SPLIT3

  # ---- buffers
  @property
  def stream(self) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch().cuda.current_stream(self.device).cuda_stream)

  def matrix(self, n: int, dtype=None):
    t = torch()
    return t.empty((n, round_up(n, 64)), dtype=dtype or t.float32, device=self.device)

  def planes(self, n: int, cols: typing.Optional[int] = None):
    t = torch()
    ld = round_up(cols if cols is not None else n, 64)
    return (t.empty((n, ld), dtype=t.float16, device=self.device),
            t.empty((n, ld), dtype=t.float16, device=self.device))

  def vector(self, n: int, dtype=None, zero=False):
    t = torch()
    f = t.zeros if zero else t.empty
    return f((n,), dtype=dtype or t.float64, device=self.device)

  def upload_matrix(self, a: np.ndarray):
    """Host [n, n] array -> device fp32 matrix with padded rows."""
    t = torch()
    n = a.shape[0]
    m = self.matrix(n)
    src = t.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    m[:, :n].copy_(src, non_blocking=False)
    return m

  def download_matrix(self, m, n: int) -> np.ndarray:
    return m[:, :n].to("cpu").numpy().astype(np.float64)

  # C-ABI entry point -> the reference function whose work it does (NVTX range names)
  REFERENCE_NAMES = {
      "sc_normalize_rows": "utils.compute_affinity_matrix", "sc_affinity_cosine": "utils.compute_affinity_matrix",
      "sc_crop_diagonal": "refinement.CropDiagonal", "sc_crop_diagonal_values": "refinement.CropDiagonal",
      "sc_gaussian_blur": "refinement.GaussianBlur", "sc_gaussian_blur_rowmax": "refinement.GaussianBlur",
      "sc_blur_upper_rowmax": "refinement.GaussianBlur",
      "sc_blur_threshold_symmetrize": "refinement.RowWiseThreshold+Symmetrize",
      "sc_threshold_symmetrize_upper": "refinement.RowWiseThreshold+Symmetrize",
      "sc_row_threshold": "refinement.RowWiseThreshold", "sc_symmetrize": "refinement.Symmetrize",
      "sc_diffuse": "refinement.Diffuse", "sc_gemm_nt_planes": "refinement.Diffuse",
      "sc_row_normalize": "refinement.RowWiseNormalize", "sc_row_stats": "refinement.RowWiseNormalize",
      "sc_laplacian": "laplacian.compute_laplacian",
      "sc_eigh_dense": "utils.compute_sorted_eigenvectors",
      "sc_eigh_extremal": "utils.compute_sorted_eigenvectors",
      "sc_eigh_extremal_sharded": "utils.compute_sorted_eigenvectors",
      "sc_kmeans": "custom_distance_kmeans.run_kmeans", "sc_row_renorm": "spectral_clusterer.row_wise_renorm",
      "sc_affinity_stats": "fallback_clusterer.check_single_cluster",
      "sc_constraint_combine": "constraint.adjust_affinity", "sc_scale_shift": "constraint.adjust_affinity",
  }
  nvtx = os.environ.get("SCB_NVTX") == "1"

  def call(self, name, *args, exc=nat.NativeError):
    """One C-ABI call on the current stream; with `profile` on, bracketed by CUDA events; with
    SCB_NVTX=1 inside an NVTX range named after the reference function it replaces."""
    if self.nvtx:
      torch().cuda.nvtx.range_push("%s [%s]" % (self.REFERENCE_NAMES.get(name, name), name))
      try:
        return self._call(name, *args, exc=exc)
      finally:
        torch().cuda.nvtx.range_pop()
    return self._call(name, *args, exc=exc)

  def _call(self, name, *args, exc=nat.NativeError):
    if self.profile is None:
      nat.call(name, self.ctx, *args, exc=exc)
      return
    t = torch()
    start, stop = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    start.record()
    nat.call(name, self.ctx, *args, exc=exc)
    stop.record()
    self.profile.append((name, start, stop))

  def start_profile(self):
    self.profile = []

  def stop_profile(self) -> typing.Dict[str, float]:
    """Milliseconds per C-ABI entry point since start_profile() (device time, CUDA events)."""
    torch().cuda.synchronize(self.device)
    out: typing.Dict[str, float] = {}
    for name, start, stop in self.profile or []:
      out[name] = out.get(name, 0.0) + start.elapsed_time(stop)
    self.profile = None
    return out

  # ---- operators (device in / device out)
  def gemm_engine(self, n: int) -> int:
    if os.environ.get("SCB_FORCE_SIMT") == "1":    # debugging aid: validation GEMM everywhere
      return nat.GEMM_SIMT
    return nat.GEMM_SIMT if n < self.simt_below else nat.GEMM_TCGEN05

  def affinity(self, x_dev, want_crop_vector: bool):
    """utils.compute_affinity_matrix on a device [n, d] fp32/fp64 tensor."""
    t = torch()
    n, d = int(x_dev.shape[0]), int(x_dev.shape[1])
    engine = self.gemm_engine(n)
    a = self.matrix(n)
    crop = self.vector(n, t.float32, zero=True) if want_crop_vector else None
    is64 = 1 if x_dev.dtype == t.float64 else 0
    if engine == nat.GEMM_SIMT:
      xn = t.empty((n, round_up(d, 4)), dtype=t.float32, device=self.device)
      self.call("sc_normalize_rows", _ptr(x_dev), is64, n, d, x_dev.stride(0), _ptr(xn),
                xn.stride(0), None, None, 0, self.stream)
      self.call("sc_affinity_cosine", engine, self.gemm_precision, _ptr(xn), xn.stride(0), None,
                None, 0, n, d, _ptr(a), a.stride(0), _ptr(crop), self.stream)
    else:
      hi, lo = self.planes(n, d)
      self.call("sc_normalize_rows", _ptr(x_dev), is64, n, d, x_dev.stride(0), None, 0,
                _ptr(hi), _ptr(lo), hi.stride(0), self.stream)
      self.call("sc_affinity_cosine", engine, self.gemm_precision, None, 0, _ptr(hi), _ptr(lo),
                hi.stride(0), n, d, _ptr(a), a.stride(0), _ptr(crop), self.stream)
    return a, crop

  def crop_diagonal(self, a, n):
    out = self.matrix(n)
    self.call("sc_crop_diagonal", _ptr(a), n, a.stride(0), _ptr(out), out.stride(0), self.stream)
    return out

  def crop_values(self, a, n):
    d = self.vector(n, torch().float32)
    self.call("sc_crop_diagonal_values", _ptr(a), n, a.stride(0), _ptr(d), self.stream)
    return d

  def gaussian_blur(self, a, n, sigma, diag=None):
    out = self.matrix(n)
    self.call("sc_gaussian_blur", _ptr(a), n, a.stride(0), _ptr(diag), float(sigma), _ptr(out),
              out.stride(0), None, self.stream)
    return out

  def row_threshold(self, a, n, type_, p, mult, binarize, preserve_diag):
    out = self.matrix(n)
    self.call("sc_row_threshold", _ptr(a), n, a.stride(0), int(type_), float(p), float(mult),
              int(bool(binarize)), int(bool(preserve_diag)), _ptr(out), out.stride(0),
              self.stream, exc=ValueError)
    return out

  def symmetrize(self, a, n, type_):
    out = self.matrix(n)
    self.call("sc_symmetrize", _ptr(a), n, a.stride(0), int(type_), _ptr(out), out.stride(0),
              self.stream, exc=ValueError)
    return out

  def blur_rowmax(self, a, n, sigma, diag, zero_diag):
    t = torch()
    m = self.vector(n, t.float32, zero=True)
    self.call("sc_gaussian_blur_rowmax", _ptr(a), n, a.stride(0), _ptr(diag), float(sigma),
              int(bool(zero_diag)), _ptr(m), self.stream)
    return m

  def blur_threshold_symmetrize(self, a, n, sigma, diag, rowmax, p, mult, binarize,
                                preserve_diag, sym_type, want_f32, want_planes):
    y = self.matrix(n) if want_f32 else None
    hi, lo = self.planes(n) if want_planes else (None, None)
    self.call("sc_blur_threshold_symmetrize", _ptr(a), n, a.stride(0), _ptr(diag), float(sigma),
              _ptr(rowmax), float(p), float(mult), int(bool(binarize)), int(bool(preserve_diag)),
              int(sym_type), _ptr(y), 0 if y is None else y.stride(0), _ptr(hi), _ptr(lo),
              0 if hi is None else hi.stride(0), self.stream)
    return y, hi, lo

  def blur_upper_rowmax(self, a, n, sigma, diag, zero_diag):
    """Symmetric pass 1: (b with its upper tiles = blur(a), m = row maxima of blur(a))."""
    t = torch()
    b = self.matrix(n)
    m = self.vector(n, t.float32)
    self.call("sc_blur_upper_rowmax", _ptr(a), n, a.stride(0), _ptr(diag), float(sigma),
              int(bool(zero_diag)), _ptr(b), b.stride(0), _ptr(m), self.stream)
    return b, m

  def threshold_symmetrize_upper(self, b, n, rowmax, p, mult, binarize, preserve_diag, sym_type,
                                 want_f32, want_planes, want_lo=True):
    t = torch()
    y = self.matrix(n) if want_f32 else None
    hi = lo = None
    if want_planes and want_lo:
      hi, lo = self.planes(n)
    elif want_planes:      # a single-MMA Diffuse reads only the hi plane: 2 B/element less to write
      hi = t.empty((n, round_up(n, 64)), dtype=t.float16, device=self.device)
    self.call("sc_threshold_symmetrize_upper", _ptr(b), n, b.stride(0), _ptr(rowmax), float(p),
              float(mult), int(bool(binarize)), int(bool(preserve_diag)), int(sym_type), _ptr(y),
              0 if y is None else y.stride(0), _ptr(hi), _ptr(lo),
              0 if hi is None else hi.stride(0), self.stream)
    return y, hi, lo

  @staticmethod
  def upper_pass_ok(sigma: float) -> bool:
    """The symmetric blur pair covers radius-4 filters (sigma ~ 1, every BASELINE config)."""
    return sigma > 1e-15 and int(4.0 * sigma + 0.5) == 4 and os.environ.get("SCB_BLUR_TWO_PASS") != "1"

  def split_planes(self, a, n):
    hi, lo = self.planes(n)
    self.call("sc_split_planes", _ptr(a), n, a.stride(0), _ptr(hi), _ptr(lo), hi.stride(0),
              self.stream)
    return hi, lo

  def diffuse(self, n, y=None, hi=None, lo=None, want_stats=False, precision=None):
    """S = Y Y^T.  With want_stats (tcgen05 engine) also returns (rowmax fp32, rowsum fp64) of S
    from the GEMM epilogue; otherwise (None, None).  `precision` defaults to the fp32-accurate
    three-MMA split (the operator-level API); predict() passes `diffuse_precision`."""
    t = torch()
    precision = self.gemm_precision if precision is None else precision
    engine = self.gemm_engine(n)
    s = self.matrix(n)
    if engine == nat.GEMM_SIMT:
      assert y is not None
      self.call("sc_diffuse", engine, precision, _ptr(y), y.stride(0), None, None, 0,
                n, _ptr(s), s.stride(0), None, None, self.stream)
      return s, None, None
    if hi is None:
      hi, lo = self.split_planes(y, n)
    mx = self.vector(n, t.float32) if want_stats else None
    sm = self.vector(n, t.float64) if want_stats else None
    self.call("sc_diffuse", engine, precision, None, 0, _ptr(hi), _ptr(lo),
              hi.stride(0), n, _ptr(s), s.stride(0), _ptr(mx), _ptr(sm), self.stream)
    return s, mx, sm

  def row_stats(self, a, n, want_max=True, want_sum=True):
    mx = self.vector(n) if want_max else None
    sm = self.vector(n) if want_sum else None
    self.call("sc_row_stats", _ptr(a), n, a.stride(0), _ptr(mx), _ptr(sm), self.stream)
    return mx, sm

  def row_normalize(self, a, n):
    out = self.matrix(n)
    self.call("sc_row_normalize", _ptr(a), n, a.stride(0), _ptr(out), out.stride(0), self.stream)
    return out

  def laplacian(self, w, n, type_, eps=1e-10):
    out = self.matrix(n)
    self.call("sc_laplacian", _ptr(w), n, w.stride(0), int(type_), float(eps), _ptr(out),
              out.stride(0), self.stream, exc=ValueError)
    return out

  def eigh(self, s, n, delta, left, right, sign, which, n_values, n_vectors, dense, tol=1e-9,
           max_matvecs=0):
    """Sorted extremal eigenvalues (host fp64) and eigenvectors (device fp64 [n, n_vectors])."""
    t = torch()
    w = np.empty(max(n_values, 1), dtype=np.float64)
    v = t.empty((n, max(n_vectors, 1)), dtype=t.float64, device=self.device)
    wp = w.ctypes.data_as(ctypes.c_void_p)
    stats = np.zeros(4, dtype=np.int64)
    if dense:
      self.call("sc_eigh_dense", _ptr(s), n, s.stride(0), _ptr(delta), _ptr(left), _ptr(right),
                float(sign), int(which), n_values, n_vectors, wp, _ptr(v), None, None, self.stream)
    else:
      self.call("sc_eigh_extremal", _ptr(s), n, s.stride(0), _ptr(delta), _ptr(left),
                _ptr(right), float(sign), int(which), n_values, n_vectors, float(tol),
                int(max_matvecs), wp, _ptr(v), stats.ctypes.data_as(ctypes.c_void_p),
                self.stream)
    return w[:n_values], v[:, :n_vectors], stats

  def eigh_dense_pick(self, s, n, delta, left, right, sign, which, picker):
    """Full spectrum by the dense solver; `picker(w) -> count` (host) decides from the sorted
    eigenvalues how many eigenvectors are computed.  Returns (w[n], v device [n, count])."""
    t = torch()
    w = np.empty(n, dtype=np.float64)
    box = {}

    def pick(_user, w_ptr, n_values, out_ptr):
      try:
        count = int(picker(np.ctypeslib.as_array(w_ptr, shape=(int(n_values),)).copy()))
        count = max(0, min(count, int(n_values)))
        box["v"] = t.empty((n, max(count, 1)), dtype=t.float64, device=self.device)
        out_ptr[0] = box["v"].data_ptr()
        box["count"] = count
        return count
      except Exception as e:                         # never unwind through the C frame
        box["error"] = e
        return -1

    callback = nat.PICK_FN(pick)
    try:
      self.call("sc_eigh_dense", _ptr(s), n, s.stride(0), _ptr(delta), _ptr(left), _ptr(right),
                float(sign), int(which), n, 0, w.ctypes.data_as(ctypes.c_void_p), None, callback,
                None, self.stream)
    except nat.NativeError:
      if "error" in box:
        raise box["error"]
      raise
    return w, box["v"][:, :box["count"]]

  def row_renorm(self, e):
    self.call("sc_row_renorm", _ptr(e), int(e.shape[0]), int(e.shape[1]), self.stream)

  def kmeans(self, e, k, metric, max_iter, tol=0.001):
    """run_kmeans on a contiguous device fp64 [n, k_dim] array -> host int64 labels."""
    n, kd = int(e.shape[0]), int(e.shape[1])
    # scikit-learn's draws: RandomState(0).choice(n, p=uniform) then uniform(size=trials) per
    # additional centre (sklearn/cluster/_kmeans.py:231, :249).  Host RNG logic only.
    rs = np.random.RandomState(0)
    first = int(rs.choice(n, p=np.full(n, 1.0 / n)))
    trials = 2 + int(np.log(k))
    u = np.ascontiguousarray(
        [rs.uniform(size=trials) for _ in range(k - 1)], dtype=np.float64).reshape(-1)
    labels = np.empty(n, dtype=np.int64)
    iters = np.zeros(1, dtype=np.int64)
    self.call("sc_kmeans", _ptr(e), n, kd, int(k), first,
              u.ctypes.data_as(ctypes.c_void_p) if u.size else None, trials, int(metric),
              int(max_iter), float(tol), labels.ctypes.data_as(ctypes.c_void_p),
              iters.ctypes.data_as(ctypes.c_void_p), self.stream, exc=ValueError)
    return labels, int(iters[0])


class Refined:
  """Result of the refinement sequence in the structured form the eigensolver wants:
  the matrix is diag(row_scale) * S with S fp32 on the device (`row_scale` None = ones).
  `symmetric` tells whether S is symmetric (up to rounding)."""

  def __init__(self, s, n, symmetric, row_scale=None, rowsum=None):
    self.s = s
    self.n = n
    self.symmetric = symmetric
    self.row_scale = row_scale
    self.rowsum = rowsum      # fp64 row sums of S when the Diffuse epilogue produced them


def run_refinement(eng: Engine, a, n: int, options, crop_vector=None, a_symmetric=True,
                   diffuse_precision=None) -> Refined:
  """Apply options.refinement_sequence to the device affinity `a`.

  Fusions (all exact restatements):
    [CropDiagonal] GaussianBlur? RowWiseThreshold(RowMax) Symmetrize  on a symmetric input
        -> crop vector + blur statistics pass + blur/threshold/symmetrize pass (12 B/element),
           emitting fp16 planes directly when Diffuse follows;
    RowWiseNormalize as the last operator -> kept as a row scaling for the eigensolver.
  Anything else runs operator by operator.
  """
  from . import refinement as rf
  names = list(options.refinement_sequence or [])
  RN = rf.RefinementName
  cur = a
  sym = a_symmetric
  planes = None          # (hi, lo) of `cur` when available
  cur_is_planes_only = False
  row_scale = None
  stats = None           # (rowmax fp32, rowsum fp64) of `cur` straight from the Diffuse epilogue
  i = 0
  while i < len(names):
    name = names[i]
    if not isinstance(name, RN):
      raise ValueError("Unknown refinement operation: {}".format(name))
    # ---- fused chain
    j = i
    has_crop = names[j] == RN.CropDiagonal
    if has_crop:
      j += 1
    has_blur = j < len(names) and names[j] == RN.GaussianBlur
    if has_blur:
      j += 1
    fusable = (sym and row_scale is None and not cur_is_planes_only and
               j + 1 < len(names) and names[j] == RN.RowWiseThreshold and
               names[j + 1] == RN.Symmetrize and
               options.thresholding_type == rf.ThresholdType.RowMax and
               isinstance(options.symmetrize_type, rf.SymmetrizeType))
    if fusable:
      sigma = float(options.gaussian_blur_sigma) if has_blur else 0.0
      diag = None
      if has_crop:
        if crop_vector is not None and i == 0:
          diag = crop_vector
        else:
          diag = eng.crop_values(cur, n)
      zero_diag = bool(options.thresholding_preserve_diagonal)
      next_is_diffuse = j + 2 < len(names) and names[j + 2] == RN.Diffuse
      use_tc = eng.gemm_engine(n) == nat.GEMM_TCGEN05
      want_planes = next_is_diffuse and use_tc
      want_f32 = not want_planes
      sym_type = (nat.SYMMETRIZE_MAX if options.symmetrize_type == rf.SymmetrizeType.Max
                  else nat.SYMMETRIZE_AVERAGE)
      if eng.upper_pass_ok(sigma):
        # blur(A) is symmetric: blur the upper tiles once, keep them, mirror the result
        b, m = eng.blur_upper_rowmax(cur, n, sigma, diag, zero_diag)
        want_lo = (diffuse_precision if diffuse_precision is not None
                   else eng.gemm_precision) != nat.GEMM_SINGLE
        y, hi, lo = eng.threshold_symmetrize_upper(
            b, n, m, options.p_percentile, options.thresholding_soft_multiplier,
            options.thresholding_with_binarization, zero_diag, sym_type, want_f32, want_planes,
            want_lo)
        del b
      else:
        m = eng.blur_rowmax(cur, n, sigma, diag, zero_diag)
        y, hi, lo = eng.blur_threshold_symmetrize(
            cur, n, sigma, diag, m, options.p_percentile, options.thresholding_soft_multiplier,
            options.thresholding_with_binarization, zero_diag, sym_type, want_f32, want_planes)
      cur = y
      planes = (hi, lo) if want_planes else None
      cur_is_planes_only = want_planes
      sym = True
      i = j + 2
      continue
    # ---- single operators
    if name == RN.Diffuse:
      if row_scale is not None:
        cur, row_scale = _materialise_scale(eng, cur, n, row_scale), None
      # the row reductions of a closing RowWiseNormalize / of the Laplacian degree come out of the
      # GEMM epilogue when Diffuse is the last matrix-valued operator of the sequence
      tail = names[i + 1:]
      fuse_stats = (tail in ([], [RN.RowWiseNormalize]) and
                    eng.gemm_engine(n) == nat.GEMM_TCGEN05)
      if planes is not None:
        cur, mx, sm = eng.diffuse(n, hi=planes[0], lo=planes[1], want_stats=fuse_stats,
                                  precision=diffuse_precision)
      else:
        cur, mx, sm = eng.diffuse(n, y=cur, want_stats=fuse_stats, precision=diffuse_precision)
      stats = (mx, sm) if mx is not None else None
      planes, cur_is_planes_only, sym = None, False, True
      i += 1
      continue
    if cur_is_planes_only:
      raise AssertionError("internal: planes-only matrix consumed by a non-GEMM operator")
    if name == RN.RowWiseNormalize and i == len(names) - 1 and row_scale is None:
      if stats is not None:
        mx = stats[0].double()
      else:
        mx, _ = eng.row_stats(cur, n, want_max=True, want_sum=False)
      row_scale = 1.0 / mx
      i += 1
      continue
    stats = None
    if row_scale is not None:
      cur, row_scale = _materialise_scale(eng, cur, n, row_scale), None
      sym = False
    if name == RN.CropDiagonal:
      cur = eng.crop_diagonal(cur, n)
    elif name == RN.GaussianBlur:
      cur = eng.gaussian_blur(cur, n, options.gaussian_blur_sigma)
    elif name == RN.RowWiseThreshold:
      if not isinstance(options.thresholding_type, rf.ThresholdType):
        raise TypeError("thresholding_type must be a ThresholdType")
      cur = eng.row_threshold(
          cur, n, nat.THRESHOLD_ROWMAX if options.thresholding_type == rf.ThresholdType.RowMax
          else nat.THRESHOLD_PERCENTILE, options.p_percentile,
          options.thresholding_soft_multiplier, options.thresholding_with_binarization,
          options.thresholding_preserve_diagonal)
      sym = False
    elif name == RN.Symmetrize:
      if options.symmetrize_type == rf.SymmetrizeType.Max:
        cur = eng.symmetrize(cur, n, nat.SYMMETRIZE_MAX)
      elif options.symmetrize_type == rf.SymmetrizeType.Average:
        cur = eng.symmetrize(cur, n, nat.SYMMETRIZE_AVERAGE)
      else:
        raise ValueError("Unsupported symmetrize_type.")
      sym = True
    elif name == RN.RowWiseNormalize:
      cur = eng.row_normalize(cur, n)
      sym = False
    else:
      raise ValueError("Unknown refinement operation: {}".format(name))
    i += 1
  return Refined(cur, n, sym, row_scale, stats[1] if stats is not None else None)


def _materialise_scale(eng: Engine, s, n: int, row_scale):
  # diag(row_scale) S : only reached by sequences that keep refining after RowWiseNormalize
  return eng.row_normalize(s, n)
