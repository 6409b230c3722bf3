"""GPU: the row-sharded pipeline (sharded.py + the *_block C-ABI entry points) against the
unsharded device pipeline.  world=1 runs in-process; world=2 runs two ranks -- over NCCL on two
GPUs when the box has them, else both ranks on GPU 0 with gloo moving the CUDA tensors (same
orchestration and kernels, only the transport differs)."""

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def icassp_options(scb, sigma=1):
  return scb.RefinementOptions(
      gaussian_blur_sigma=sigma, p_percentile=0.95, thresholding_soft_multiplier=0.01,
      refinement_sequence=list(scb.ICASSP2018_REFINEMENT_SEQUENCE))


def unsharded_reference(eng, x_dev, n, opt):
  """S = Diffuse(...) and its row statistics from the single-GPU fused path."""
  import spectralcluster_b200 as scb
  from spectralcluster_b200 import device as dev
  a, crop = eng.affinity(x_dev, want_crop_vector=True)
  pre = scb.RefinementOptions(**{**opt.__dict__, "refinement_sequence":
                                 list(opt.refinement_sequence)[:-1]})   # stop before RowNormalize
  refined = dev.run_refinement(eng, a, n, pre, crop_vector=crop)
  return refined.s


@pytest.mark.parametrize("n", [1000, 2304])
def test_block_kernels_world1_match_unsharded(engine, n):
  import torch
  import spectralcluster_b200 as scb
  from spectralcluster_b200 import sharded
  from oracle import spectral_oracle as orc
  x = torch.from_numpy(orc.synthetic_dvectors(n, 64, 4, seed=3).astype(np.float32)).to(engine.device)
  opt = icassp_options(scb)
  res = sharded.ShardedRefiner(sharded.DeviceBackend(engine), opt).run(x, 1, 0)
  want = unsharded_reference(engine, x, n, opt)
  got = res["s_block"][:, :n].cpu().numpy()
  np.testing.assert_allclose(got, want[:, :n].cpu().numpy(), rtol=2e-6, atol=0)
  np.testing.assert_allclose(res["rowmax"].cpu().numpy(), got.max(axis=1), rtol=1e-6)


def _worker(rank, world, port, n, use_nccl, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  import torch
  import torch.distributed as dist
  torch.cuda.set_device(rank if use_nccl else 0)
  dist.init_process_group("nccl" if use_nccl else "gloo", rank=rank, world_size=world)
  import spectralcluster_b200 as scb
  from spectralcluster_b200 import device as dev
  from spectralcluster_b200 import sharded
  from oracle import spectral_oracle as orc
  eng = dev.Engine.get(rank if use_nccl else 0)
  x = torch.from_numpy(orc.synthetic_dvectors(n, 64, 4, seed=3).astype(np.float32)).to(eng.device)
  opt = icassp_options(scb)
  res = sharded.ShardedRefiner(sharded.DeviceBackend(eng), opt, dist=dist).run(x, world, rank)
  plan = res["plan"]
  want = unsharded_reference(eng, x, n, opt)[plan.row_begin:plan.row_end, :n]
  got = res["s_block"][:, :n]
  err = float(((got - want).abs() / want.abs().clamp_min(1e-30)).max())
  flags = torch.tensor([err], device=eng.device if use_nccl else "cpu")
  dist.all_reduce(flags, op=dist.ReduceOp.MAX)
  if rank == 0:
    out.put(float(flags[0]))
  dist.destroy_process_group()


def test_two_ranks_match_unsharded():
  import torch
  import torch.multiprocessing as mp
  use_nccl = torch.cuda.device_count() >= 2
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  ctx = mp.get_context("spawn")
  out = ctx.Queue()
  mp.spawn(_worker, args=(2, port, 2304, use_nccl, out), nprocs=2, join=True)
  assert out.get(timeout=30) <= 2e-6


def _autotune_worker(rank, world, port, use_nccl, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  import torch
  import torch.distributed as dist
  torch.cuda.set_device(rank if use_nccl else 0)
  dist.init_process_group("nccl" if use_nccl else "gloo", rank=rank, world_size=world)
  sys.path.insert(0, os.path.join(ROOT, "tests"))
  import spectralcluster_b200 as scb
  from conftest import load_golden
  from test_gpu_predict import make_clusterer
  case = load_golden("config5_autotune_n1024_d256")
  c = make_clusterer(case["options"])
  c.autotune_group = True
  labels = c.predict(case["embeddings"])
  ok = (np.array_equal(scb.utils.enforce_ordered_labels(labels),
                       scb.utils.enforce_ordered_labels(case["labels"])) and
        c.last_details["best_p_percentile"] == float(case["p_best"]))
  flags = torch.tensor([1 if ok else 0], device="cuda" if use_nccl else "cpu")
  dist.all_reduce(flags, op=dist.ReduceOp.MIN)
  if rank == 0:
    out.put(int(flags[0]))
  dist.destroy_process_group()


def test_parallel_autotune_two_ranks_matches_reference_fixture():
  """BASELINE.json configs[4] scaled: 8 p values over 2 ranks -> the reference's p and labels."""
  import torch
  import torch.multiprocessing as mp
  use_nccl = torch.cuda.device_count() >= 2
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  ctx = mp.get_context("spawn")
  out = ctx.Queue()
  mp.spawn(_autotune_worker, args=(2, port, use_nccl, out), nprocs=2, join=True)
  assert out.get(timeout=30) == 1


def _predict_worker(rank, world, port, use_nccl, transport, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SCB_SHARDED_TRANSPORT=transport)
  import torch
  import torch.distributed as dist
  torch.cuda.set_device(rank if use_nccl else 0)
  dist.init_process_group("nccl" if use_nccl else "gloo", rank=rank, world_size=world)
  import spectralcluster_b200 as scb
  from spectralcluster_b200 import sharded
  from oracle import spectral_oracle as orc
  n = 3072
  x, truth = orc.synthetic_dvectors(n, 128, 5, seed=4, return_labels=True)
  ok = True
  for lap in (None, scb.LaplacianType.GraphCut):
    c = scb.SpectralClusterer(min_clusters=2, max_clusters=9, laplacian_type=lap,
                              refinement_options=icassp_options(scb))
    got = sharded.predict_sharded(c, x, dist=dist)
    single = scb.SpectralClusterer(min_clusters=2, max_clusters=9, laplacian_type=lap,
                                   refinement_options=icassp_options(scb))
    ref = single.predict(x)
    ok = ok and np.array_equal(scb.utils.enforce_ordered_labels(got),
                               scb.utils.enforce_ordered_labels(ref))
    ok = ok and np.array_equal(scb.utils.enforce_ordered_labels(got),
                               scb.utils.enforce_ordered_labels(truth))
    w1, w2 = c.last_details["eigenvalues"], single.last_details["eigenvalues"]
    ok = ok and np.allclose(w1, w2[:len(w1)], rtol=1e-6, atol=1e-7 * np.abs(w2).max())
    # two GPUs over NCCL: the NVLink peer-memory schedule unless send/recv was asked for
    want = ("peer" if transport == "peer" else "nccl") if use_nccl else "gloo"
    ok = ok and c.last_details["transport"] == want
  flags = torch.tensor([1 if ok else 0], device="cuda" if use_nccl else "cpu")
  dist.all_reduce(flags, op=dist.ReduceOp.MIN)
  if rank == 0:
    out.put(int(flags[0]))
  dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["peer", "sendrecv"])
def test_predict_sharded_two_ranks_matches_single_gpu(transport):
  import torch
  import torch.multiprocessing as mp
  use_nccl = torch.cuda.device_count() >= 2
  if transport == "sendrecv" and not use_nccl:
    pytest.skip("one GPU: the gloo run of the other parametrisation already is send/recv")
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  ctx = mp.get_context("spawn")
  out = ctx.Queue()
  mp.spawn(_predict_worker, args=(2, port, use_nccl, transport, out), nprocs=2, join=True)
  assert out.get(timeout=30) == 1
