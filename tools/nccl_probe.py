#!/usr/bin/env python
"""Transport probe for the sharded pipeline (run under torchrun, 2+ GPUs): NCCL broadcast /
all-gather bandwidth, and copy-engine peer copies through CUDA IPC handles."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
nbytes = 4 << 30
buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
def timed(fn, iters=3):
  fn(); torch.cuda.synchronize(); dist.barrier()
  t0 = time.perf_counter()
  for _ in range(iters): fn()
  torch.cuda.synchronize(); dist.barrier()
  return (time.perf_counter() - t0) / iters
t = timed(lambda: dist.broadcast(buf, src=0))
if rank == 0: print("nccl broadcast 4 GiB: %.1f ms  %.1f GB/s" % (t * 1e3, nbytes / t / 1e9), flush=True)
full = torch.empty(world * (1 << 30), dtype=torch.uint8, device=dev)
t = timed(lambda: dist.all_gather_into_tensor(full, buf[:1 << 30]))
if rank == 0: print("nccl all_gather 1 GiB/rank: %.1f ms  recv %.1f GB/s" % (t * 1e3, (world - 1) * (1 << 30) / t / 1e9), flush=True)
# CUDA IPC + copy engine
try:
  handle = buf.untyped_storage()._share_cuda_()
  handles = [None] * world
  dist.all_gather_object(handles, handle)
  peer = (rank + 1) % world
  st = torch.UntypedStorage._new_shared_cuda(*handles[peer])
  src = torch.empty(0, dtype=torch.uint8, device=torch.device("cuda", st.device.index)).set_(st, 0, (nbytes,), (1,))
  dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
  t = timed(lambda: dst.copy_(src, non_blocking=True))
  print("rank %d: IPC peer copy (pull from rank %d, device %s) 4 GiB: %.1f ms  %.1f GB/s" % (rank, peer, src.device, t * 1e3, nbytes / t / 1e9), flush=True)
except Exception as e:
  print("rank %d: IPC path failed: %r" % (rank, e), flush=True)
dist.destroy_process_group()
