#!/usr/bin/env python
"""Tiny driver for ncu: one engine kind, a few device-resident passes.  usage: prof_run.py KIND [n] [chunk]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from dint_b200 import Engine, wire
import trace_gen as T

kind = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 21
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
if kind == "fasst":
    k, req, cfg = wire.FASST, T.fasst_random(n, 24_000_000, seed=1, weights=(0.6, 0.15, 0.05, 0.2)), {}
elif kind == "fasst_hot":
    k, req, cfg = wire.FASST, T.fasst_random(n, 4800, seed=1), {}
elif kind == "store":
    k, req, cfg = wire.STORE, T.store_random(n, 2_000_000, seed=4, p_set=0.0, p_miss=0.0), dict(populate=True)
elif kind == "lock2pl":
    k, req, cfg = wire.LOCK2PL, T.lock2pl_random(n, 24_000_000, seed=2), {}
else:
    raise SystemExit("kind?")
with Engine(k, chunk=chunk, **cfg) as eng:
    d = torch.from_numpy(req).cuda()
    out = torch.empty_like(d)
    for _ in range(3):
        eng.submit_tensor(d, out)
    eng.sync()
    torch.cuda.synchronize()
print("done")
