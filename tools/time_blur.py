#!/usr/bin/env python
"""Times the two passes of the fused refinement chain (CUDA events, L2 flushed by the matrix
size) for one N: python tools/time_blur.py --n 65536"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectralcluster_b200 import device as dev
ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=65536); ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
eng = dev.Engine.get(0); n = a.n
A = eng.matrix(n); A.uniform_(0.3, 1.0)
crop = torch.rand(n, device=eng.device, dtype=torch.float32)
for name, fn in (("stats", lambda: eng.blur_rowmax(A, n, 1.0, crop, False)),):
  m = fn()
def thr():
  return eng.blur_threshold_symmetrize(A, n, 1.0, crop, m, 0.95, 0.01, False, False, 0, False, True)
res = {}
for name, fn in (("stats_4B", lambda: eng.blur_rowmax(A, n, 1.0, crop, False)), ("thrsym_8B", thr)):
  for _ in range(2): fn()
  torch.cuda.synchronize()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(a.iters): fn()
  e.record(); torch.cuda.synchronize()
  ms = s.elapsed_time(e) / a.iters
  bytes_ = n * n * (4 if "4B" in name else 8)
  res[name] = (ms, bytes_ / ms / 1e6)
print("N=%d tiles_per_cta=%s " % (n, os.environ.get("SCB_BLUR_TILES_PER_CTA", "default")) +
      "  ".join("%s %.2f ms %.0f GB/s" % (k, v[0], v[1]) for k, v in res.items()))
