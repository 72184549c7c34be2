#!/usr/bin/env python
"""Per-phase timing of the sharded (multi-GPU) request path; run under torchrun.  Development aid."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from dint_b200 import wire
from dint_b200.shard import ShardedEngine
import trace_gen as T

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
n = 1 << 20
req = torch.from_numpy(T.fasst_random(n, 24_000_000, seed=rank, weights=(0.6, 0.15, 0.05, 0.2))).cuda()
se = ShardedEngine(wire.FASST, chunk=n + n // 2, use_slabs=True, strict=False)
eng = se.engine
W = world
mean = (n + W - 1) // W
cap = (int(mean * se.slab_slack) + int(8 * mean ** 0.5) + 64 + 15) // 16 * 16

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

acc = {}
for it in range(25):
    t = [ev()]
    t.append(ev())
    slabs, state, cap = se._dispatch_local(req, n, None); t.append(ev())
    recv = torch.empty_like(slabs); dist.all_to_all_single(recv, slabs); t.append(ev())
    out_local = torch.empty_like(recv); eng.submit_tensor(recv, out_local); t.append(ev())
    back = torch.empty_like(slabs); dist.all_to_all_single(back, out_local); t.append(ev())
    out = torch.empty(n * 9, dtype=torch.uint8, device="cuda")
    eng.route_combine(eng.slab_ptrs(back.data_ptr(), W, cap * 9), state, n, W, cap, out); t.append(ev())
    torch.cuda.synchronize()
    if it >= 5:
        for name, a, b in zip(["-", "dispatch", "a2a_out", "engine", "a2a_back", "combine"], t[:-1], t[1:]):
            acc[name] = acc.get(name, 0.0) + a.elapsed_time(b)
        acc["total"] = acc.get("total", 0.0) + t[0].elapsed_time(t[-1])
# whole calls back to back (no per-phase events, no sync in between)
torch.cuda.synchronize(); dist.barrier()
e0 = ev()
for it in range(20):
    se.submit_tensor(req)
e1 = ev(); torch.cuda.synchronize()
# fused P2P path
sp = ShardedEngine(wire.FASST, chunk=n + n // 2, use_p2p=True, p2p_max_n=n, strict=False)
for it in range(5):
    o2 = sp.submit_tensor(req)
torch.cuda.synchronize(); dist.barrier()
p0 = ev()
for it in range(20):
    o2 = sp.submit_tensor(req)
p1 = ev(); torch.cuda.synchronize()
p2p_us = p0.elapsed_time(p1) / 20 * 1e3
p2p_flags = sp.check_p2p()
if rank == 0:
    print("p2p back-to-back submit_tensor: %.1f us per call" % p2p_us, "flags", p2p_flags)
    print("per-phase us (1M requests per rank, world=%d):" % world, {k: round(v / 20 * 1e3, 1) for k, v in acc.items()})
    print("back-to-back submit_tensor: %.1f us per call" % (e0.elapsed_time(e1) / 20 * 1e3), "overflow", se.check_overflow())
dist.destroy_process_group()
