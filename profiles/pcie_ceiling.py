#!/usr/bin/env python
"""The host-buffer ceiling of a box: N ranks (one per GPU) copy the packed step-output block of a 1 M-env shard
(68,158,208 B: obs 60 + reward 4 + flags 1 B per env, 256-byte aligned parts) device -> pinned host with plain
cudaMemcpyAsync, all at the same time -- no kernel, no Python per-step work.  This is the upper bound of
BatchedDroneEnv.step_host()'s e2e rate at N GPUs; bench.py reports e2e.frac_of_pcie_ceiling against the same
measurement taken inside the bench run.

    python profiles/pcie_ceiling.py                         # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        profiles/pcie_ceiling.py                              # N GPUs concurrently
Prints one JSON line on rank 0."""
import json
import os
import time

import torch
import torch.distributed as dist

BYTES = 68158208
ws = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if ws > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29511")
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev)
d = torch.empty(BYTES, dtype=torch.uint8, device=dev).fill_(1)
h = torch.empty(BYTES, dtype=torch.uint8, pin_memory=True)
h.fill_(0)                                                  # first touch by this rank


def timed(fn, reps):
    fn(3)
    if ws > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(reps); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if ws > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        lo = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        return float(t.item()), float(lo.item())
    return ms, ms


def d2h(k):
    for _ in range(k):
        h.copy_(d, non_blocking=True)


def h2d(k):
    for _ in range(k):
        d.copy_(h, non_blocking=True)


def d2h_sync(k):                                            # one copy, wait for it: what a step loop sees
    for _ in range(k):
        h.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()


ms_d2h, ms_d2h_min = timed(d2h, 60)
ms_h2d, _ = timed(h2d, 60)
t0 = time.perf_counter(); ms_sync, _ = timed(d2h_sync, 60)
if rank == 0:
    g = lambda ms: BYTES / (ms * 1e-3) / 1e9
    print(json.dumps({"n_gpus": ws, "bytes": BYTES, "d2h_ms_max_over_ranks": ms_d2h, "d2h_gbs_per_gpu": g(ms_d2h), "d2h_gbs_fastest_rank": g(ms_d2h_min),
                      "d2h_gbs_aggregate": ws * g(ms_d2h), "h2d_gbs_per_gpu": g(ms_h2d), "d2h_with_sync_per_copy_gbs_per_gpu": g(ms_sync),
                      "env_steps_per_s_ceiling": ws * (1 << 20) / (ms_sync * 1e-3), "cpus": os.cpu_count()}), flush=True)
if ws > 1:
    dist.barrier(); dist.destroy_process_group()
