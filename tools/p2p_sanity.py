#!/usr/bin/env python
"""Sanity + timing of the sharded step (dint_shard_submit_many) under torchrun or a single process (world 1)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from dint_b200 import wire
from dint_b200.shard import ShardedEngine
import trace_gen as T
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl")
n = 1 << 20
K = 8
reqs = [torch.from_numpy(T.fasst_random(n, 24_000_000, seed=100 * rank + i, weights=(0.6, 0.15, 0.05, 0.2))).cuda().view(torch.uint8).reshape(-1) for i in range(K)]
res = {}
modes = os.environ.get("SANITY_MODES", "p2p,slabs").split(",")
for mode in modes:
    se = ShardedEngine(wire.FASST, chunk=n + n // 2, use_slabs=True, use_p2p=(mode == "p2p"), p2p_max_n=n, strict=False)
    outs = se.submit_many(reqs)
    torch.cuda.synchronize(); dist.barrier()
    if mode == "p2p":
        fl = se.check_p2p()
        if fl != (0, 0):
            print("rank", rank, "p2p flags after the first sequence:", fl, "-- giving up", flush=True)
            sys.exit(1)
    for _ in range(2):
        se.submit_many(reqs)
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(5):
        se.submit_many(reqs)
    t_cpu = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * K)
    flags = se.check_p2p() if mode == "p2p" else se.check_overflow()
    res[mode] = [o.cpu() for o in outs]
    if rank == 0:
        print(f"{mode}: {us:.1f} us per 1M-request batch per rank ({n / us / 1e3:.2f} G req/s per GPU), host enqueue {t_cpu * 1e6 / (5 * K):.1f} us/batch, flags {flags}", flush=True)
    se.close()
if len(modes) == 2:
    same = all(torch.equal(a, b) for a, b in zip(res["p2p"], res["slabs"]))
    if rank == 0:
        print("p2p == slabs replies:", same)
dist.destroy_process_group()
