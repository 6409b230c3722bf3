"""CPU, world_size 2, gloo: the row-sharded orchestration (spectralcluster_b200/sharded.py) with
the NumPy backend reproduces the oracle's unsharded refinement; one-p-per-rank AutoTune picks the
oracle's winner."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _refine_worker(rank, world, port, n, sigma, crop, sym, keep_diag, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  import spectralcluster_b200 as scb
  from spectralcluster_b200 import sharded
  from sharded_numpy_backend import NumpyBackend
  from oracle import spectral_oracle as orc
  RN = scb.RefinementName
  seq = ([RN.CropDiagonal] if crop else []) + ([RN.GaussianBlur] if sigma else []) + [
      RN.RowWiseThreshold, RN.Symmetrize, RN.Diffuse, RN.RowWiseNormalize]
  opt = scb.RefinementOptions(
      gaussian_blur_sigma=sigma or 1, p_percentile=0.9, thresholding_preserve_diagonal=keep_diag,
      symmetrize_type=scb.SymmetrizeType.Max if sym == "max" else scb.SymmetrizeType.Average,
      refinement_sequence=seq)
  x = orc.synthetic_dvectors(n, 16, 3, seed=5)
  res = sharded.ShardedRefiner(NumpyBackend(), opt, dist=dist).run(x, world, rank)
  plan = res["plan"]
  names = (["crop"] if crop else []) + (["blur"] if sigma else []) + ["threshold", "symmetrize",
                                                                      "diffuse"]
  want = orc.refine(orc.affinity(x), orc.options(
      sequence=tuple(names), sigma=sigma or 1, p=0.9, preserve_diagonal=keep_diag,
      symmetrize_type=sym))
  blk = want[plan.row_begin:plan.row_end]
  ok = (np.allclose(res["s_block"], blk, rtol=1e-11, atol=1e-11) and
        np.allclose(res["rowmax"], blk.max(axis=1), rtol=1e-11) and
        np.allclose(res["rowsum"], blk.sum(axis=1), rtol=1e-11))
  flags = torch.tensor([1 if ok else 0])
  dist.all_reduce(flags, op=dist.ReduceOp.MIN)
  if rank == 0:
    out.put(int(flags[0]))
  dist.destroy_process_group()


@pytest.mark.parametrize("n,sigma,crop,sym,keep_diag", [
    (300, 1, True, "max", False), (257, 1, False, "average", False), (300, 0, True, "max", True),
    (384, 2, True, "max", False)])
def test_sharded_refinement_two_ranks_gloo(n, sigma, crop, sym, keep_diag):
  ctx = mp.get_context("spawn")
  out = ctx.Queue()
  mp.spawn(_refine_worker, args=(2, free_port(), n, sigma, crop, sym, keep_diag, out), nprocs=2,
           join=True)
  assert out.get(timeout=10) == 1


@pytest.mark.parametrize("world", [3, 4])
def test_sharded_refinement_mirrored_blocks_gloo(world):
  """3 and 4 ranks exercise the transposed-block exchange (odd and even block schedules)."""
  ctx = mp.get_context("spawn")
  out = ctx.Queue()
  mp.spawn(_refine_worker, args=(world, free_port(), 900, 1, True, "max", False, out),
           nprocs=world, join=True)
  assert out.get(timeout=10) == 1


def test_shard_plan():
  from spectralcluster_b200 import sharded
  p = sharded.ShardPlan(131072, 8, 3, 4)
  assert (p.block, p.row_begin, p.row_end, p.halo_begin, p.halo_end) == (16384, 49152, 65536,
                                                                         49148, 65540)
  p = sharded.ShardPlan(1000, 2, 1, 4)
  assert (p.block, p.row_begin, p.row_end, p.halo_begin, p.halo_end) == (512, 512, 1000, 508, 1000)
  assert p.compute_jobs() == [(0, (0, 488), (256, 512))] and p.mirror_jobs() == [(0, (0, 256), (0, 488))]
  covered = []
  for r in range(4):
    q = sharded.ShardPlan(70000, 4, r, 4)
    covered += list(range(q.row_begin, q.row_end))
    assert q.rows % 32 == 0 or q.row_end == 70000
  assert covered == list(range(70000))
  with pytest.raises(ValueError):
    sharded.ShardPlan(200, 4, 0, 4)
  # block schedule: every off-diagonal element of S is produced exactly once (computed or mirrored)
  for world, n in ((1, 256), (2, 1000), (3, 900), (4, 1300), (5, 2100), (8, 8 * 384 - 50)):
    plans = [sharded.ShardPlan(n, world, r, 4) for r in range(world)]
    cover = np.zeros((n, n), dtype=np.int32)
    for g, pl in enumerate(plans):
      cover[pl.row_begin:pl.row_end, pl.row_begin:pl.row_end] += 1
      for p2, (r0, r1), (c0, c1) in pl.compute_jobs():
        lo = plans[p2].row_begin
        cover[pl.row_begin + r0:pl.row_begin + r1, lo + c0:lo + c1] += 1
      for q, (r0, r1), (c0, c1) in pl.mirror_jobs():
        lo = plans[q].row_begin
        cover[pl.row_begin + c0:pl.row_begin + c1, lo + r0:lo + r1] += 1
      assert [q for q, _ in pl.y_requests()] == [q for q, _, _ in pl.mirror_jobs()]
    assert cover.min() == 1 and cover.max() == 1, (world, n)
  for world in (2, 3, 4, 8):                       # balance: ~G/2 block products per rank
    n = world * 4096
    plans = [sharded.ShardPlan(n, world, r, 4) for r in range(world)]
    work = [pl.rows ** 2 / 2 + sum((r1 - r0) * (c1 - c0) for _, (r0, r1), (c0, c1) in pl.compute_jobs())
            for pl in plans]
    assert max(work) <= 1.02 * (n * n / 2 / world), (world, work)


def _autotune_worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from spectralcluster_b200 import sharded
  from oracle import spectral_oracle as orc
  x = orc.synthetic_dvectors(400, 32, 4, seed=1)
  a = orc.affinity(x)
  opt = orc.options(min_clusters=2, max_clusters=8, sequence=orc.ICASSP2018,
                    autotune=dict(p_min=0.6, p_max=0.95, step=0.045, level=1, proxy="sqrt"))
  grid = orc.autotune_range(0.6, 0.95, 0.045)
  calls = []

  def evaluate(p):
    calls.append(p)
    _, _, k, gap = orc.eigenvectors_ncluster(a, dict(opt, p=p))
    return np.sqrt(1 - p) / gap, k
  best, p, ratio, k, owner = sharded.parallel_autotune(evaluate, grid, dist=dist, world=world,
                                                       rank=rank)
  _, _, k_ref, p_ref, trace = orc.autotune(a, opt)
  ok = (p == p_ref and k == k_ref and calls == grid[rank::world] and owner == best % world)
  flags = torch.tensor([1 if ok else 0])
  dist.all_reduce(flags, op=dist.ReduceOp.MIN)
  if rank == 0:
    out.put(int(flags[0]))
  dist.destroy_process_group()


def test_parallel_autotune_two_ranks_gloo():
  ctx = mp.get_context("spawn")
  out = ctx.Queue()
  mp.spawn(_autotune_worker, args=(2, free_port(), out), nprocs=2, join=True)
  assert out.get(timeout=10) == 1


def _subgroup_worker(rank, world, port, n, out):
  """Three processes, the refinement sharded over the sub-group {1, 2}: the plan speaks group-local
  ranks, P2POp wants global ranks."""
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  group = dist.new_group([1, 2])
  ok = True
  if rank in (1, 2):
    import spectralcluster_b200 as scb
    from spectralcluster_b200 import sharded
    from sharded_numpy_backend import NumpyBackend
    from oracle import spectral_oracle as orc
    RN = scb.RefinementName
    opt = scb.RefinementOptions(gaussian_blur_sigma=1, p_percentile=0.9, refinement_sequence=[
        RN.CropDiagonal, RN.GaussianBlur, RN.RowWiseThreshold, RN.Symmetrize, RN.Diffuse])
    x = orc.synthetic_dvectors(n, 16, 3, seed=5)
    res = sharded.ShardedRefiner(NumpyBackend(), opt, dist=dist, group=group).run(
        x, 2, dist.get_rank(group))
    plan = res["plan"]
    want = orc.refine(orc.affinity(x), orc.options(
        sequence=("crop", "blur", "threshold", "symmetrize", "diffuse"), sigma=1, p=0.9))
    ok = np.allclose(res["s_block"], want[plan.row_begin:plan.row_end], rtol=1e-11, atol=1e-11)
  flags = torch.tensor([1 if ok else 0])
  dist.all_reduce(flags, op=dist.ReduceOp.MIN)
  if rank == 0:
    out.put(int(flags[0]))
  dist.destroy_process_group()


def test_sharded_refinement_on_a_subgroup():
  ctx = mp.get_context("spawn")
  out = ctx.Queue()
  mp.spawn(_subgroup_worker, args=(3, free_port(), 700, out), nprocs=3, join=True)
  assert out.get(timeout=30) == 1
