#!/usr/bin/env python
"""One device-resident predict() step between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
      --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --n 16384
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from spectralcluster_b200 import synthetic  # noqa: E402
from spectralcluster_b200 import device as dev  # noqa: E402
from spectralcluster_b200 import spectral_clusterer as sc_mod  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--d", type=int, default=256)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--stop-after", default="")      # e.g. "diffuse": skip eigensolve + k-means
args = ap.parse_args()

eng = dev.Engine.get(0)
x = torch.from_numpy(synthetic.speaker_turn_dvectors(args.n, args.d, 6, seed=0).astype(np.float32)).to(eng.device)
clusterer = bench.make_clusterer()


def step():
  a, crop = eng.affinity(x, want_crop_vector=True)
  aff = sc_mod.DeviceAffinity(a, args.n, crop, True)
  if args.stop_after == "diffuse":
    dev.run_refinement(eng, a, args.n, clusterer.refinement_options, crop_vector=crop)
    return
  v, k, _ = clusterer._compute_eigenvectors_ncluster(aff)
  k = max(k, clusterer.min_clusters)
  eng.kmeans(v[:, :k].contiguous(), k, 0, clusterer.max_iter)


for _ in range(args.warm):
  step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step at N=%d" % args.n)
