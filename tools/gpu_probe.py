#!/usr/bin/env python
"""Quick device-resident throughput probe (development aid; bench.py is the contract)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from dint_b200 import Engine, wire
import trace_gen as T

def probe(kind, req, reps=5, **cfg):
    with Engine(kind, **cfg) as eng:
        if cfg.get("populate"):
            pass
        d = torch.from_numpy(req).cuda()
        out = torch.empty_like(d)
        n = req.size // eng.msg
        for _ in range(2):
            eng.submit_tensor(d, out)
        torch.cuda.synchronize()
        # wall time without per-kernel event profiling
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.submit_tensor(d, out)
        e1.record(); torch.cuda.synchronize()
        ms_noprof = e0.elapsed_time(e1) / reps
        eng.reset_stats(); eng.profile(True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.submit_tensor(d, out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        eng.sync()
        kt = eng.kernel_times(); st = eng.stats()
        print(f"{wire.KIND_NAMES[kind]:10s} n={n} chunk={eng.cfg.chunk} {ms:8.3f} ms/pass (unprofiled {ms_noprof:.3f} ms = {n/ms_noprof/1e3:.0f} Mreq/s)  {n/ms/1e3:9.1f} Mreq/s  conflicted={st['conflicted']/reps:.0f} max_run={st['max_run']}")
        for k, (l, t) in kt.items():
            print(f"      {k:12s} launches={l:5d} avg={t/l*1e3:9.1f} us total={t:8.3f} ms")

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
    for chunk in (1 << 18, 1 << 20, 1 << 22):
        probe(wire.FASST, T.fasst_random(n, 24_000_000, seed=1, weights=(0.6, 0.15, 0.05, 0.2)), chunk=chunk)
    probe(wire.FASST, T.fasst_random(n, 4800, seed=1), chunk=1 << 20)
    probe(wire.LOCK2PL, T.lock2pl_random(n, 24_000_000, seed=2), chunk=1 << 20)
    probe(wire.LOG, T.log_random(n // 4, seed=3), chunk=1 << 20)
    t = time.time()
    req = T.store_random(n, 200000, seed=4, p_set=0.0, p_miss=0.0)
    probe(wire.STORE, req, chunk=1 << 20, subs_populate=200000, populate=True)
    print("store populate+probe wall", time.time() - t)
