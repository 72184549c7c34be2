#!/usr/bin/env python
"""Where does the N=2 step lose time?  Rank 0 times each kernel of the sharded step ALONE (CUDA events, 30
back-to-back launches, no flag polling anywhere), with its buffers in plain device memory, in local symmetric
(peer-mapped) memory, and in rank 1's memory over NVLink; rank 1 only maps its buffer and waits.

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/step_probe.py

DESIGN.md section 9, "open question" of round 1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
from dint_b200 import Engine, wire
from dint_b200.engine import DintPeerPtrs
import trace_gen as T

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl")
W, n, msg = 2, 1 << 20, 9
cap = (int(n / W * 1.02) + 8 * int((n / W) ** 0.5) + 64 + 15) // 16 * 16
region = (W * cap * msg + 255) // 256 * 256
sym = symm_mem.empty(2 * region, dtype=torch.uint8, device=dev)          # [inbox | reply buffer]
hdl = symm_mem.rendezvous(sym, group=dist.group.WORLD.group_name)
sym.zero_(); torch.cuda.synchronize(); hdl.barrier()
ptrs = [int(p) for p in hdl.buffer_ptrs]

def timed(fn, it=30):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / it

if rank == 0:
    eng = Engine(wire.FASST, n_shards=W, shard_id=0, chunk=n + n // 2)
    req = torch.from_numpy(T.fasst_random(n, 24_000_000, seed=1, weights=(0.6, 0.15, 0.05, 0.2))).to(dev).view(torch.uint8).reshape(-1)
    flags = torch.zeros(2, dtype=torch.int32, device=dev)
    state = eng.route_state(n, dev)
    plain = torch.empty(W * cap * msg, dtype=torch.uint8, device=dev)
    out = torch.empty(n * msg, dtype=torch.uint8, device=dev)
    slab = cap * msg
    targets = {
        "plain device memory": DintPeerPtrs.of([plain.data_ptr(), plain.data_ptr() + slab]),
        "local symmetric memory": DintPeerPtrs.of([ptrs[0], ptrs[0] + slab]),
        "shard 1's slab in rank 1's memory (NVLink)": DintPeerPtrs.of([ptrs[0], ptrs[1] + slab]),
        "both slabs in rank 1's memory (NVLink)": DintPeerPtrs.of([ptrs[1], ptrs[1] + slab]),
    }
    for name, p in targets.items():
        d = timed(lambda: eng.route_dispatch(req, n, W, 0, cap, p, flags, state=state))
        c = timed(lambda: eng.route_combine(p, state, n, W, cap, out))
        print(f"dispatch -> {name}: {d:.1f} us   combine <- same: {c:.1f} us   roundtrip ok={bool(torch.equal(out, req))}", flush=True)
    # the engine on one batch of W * cap records (what a rank runs per step), input / output in the three kinds of memory
    eng.route_dispatch(req, n, W, 0, cap, targets["plain device memory"], flags, state=state)
    sym[: W * cap * msg].copy_(plain)
    nb = W * cap * msg
    resp_plain = torch.empty_like(plain)
    print(f"engine, plain in/out: {timed(lambda: eng.submit_tensor(plain, resp_plain)):.1f} us", flush=True)
    print(f"engine, local symmetric in/out: {timed(lambda: eng.submit_tensor(sym[:nb], sym[region:region + nb])):.1f} us", flush=True)
    remote = hdl.get_buffer(1, (2 * region,), torch.uint8)
    print(f"engine, plain in, replies stored into rank 1's memory (NVLink): {timed(lambda: eng.submit_tensor(plain, remote[region:region + nb])):.1f} us", flush=True)
    print(f"engine, local symmetric in, replies stored into rank 1's memory: {timed(lambda: eng.submit_tensor(sym[:nb], remote[region:region + nb])):.1f} us", flush=True)
    print("flags", flags.tolist())
    eng.close()
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
