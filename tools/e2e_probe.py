#!/usr/bin/env python
"""Host-path (dint_submit) time per call over a sweep of the slice-schedule knobs: lock_fasst, pinned memory."""
import sys, os, time, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from dint_b200 import Engine, wire
import trace_gen as T
msg = wire.MSG_SIZE[wire.FASST]
sizes = [int(x) for x in os.environ.get("N", "1048576,4194304").split(",")]
nmax = max(sizes)
reqs = [T.fasst_random(nmax, 24_000_000, seed=10 + i, weights=(0.6, 0.15, 0.05, 0.2)) for i in range(2)]
pin_in = [torch.from_numpy(np.ascontiguousarray(r).view(np.uint8).reshape(-1)).pin_memory() for r in reqs]
pin_out = torch.empty(nmax * msg, dtype=torch.uint8).pin_memory()
grid = [dict(DINT_HOST_CHUNK=c, DINT_HOST_MIN_SLICE=m, DINT_HOST_RAMP_UP=u)
        for c, m, u in itertools.product([262144, 524288, 1048576], [32768, 65536, 131072], [0, 1])]
grid.append(dict(DINT_HOST_CHUNK=262144, DINT_HOST_MIN_SLICE=1 << 20, DINT_HOST_RAMP_UP=0))   # the flat schedule
if len(sys.argv) > 1:
    grid = [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[1:]]
for g in grid:
    for k, v in g.items(): os.environ[k] = str(v)
    eng = Engine(wire.FASST)
    line = " ".join(f"{k[10:]}={v}" for k, v in g.items())
    for n in sizes:
        ts = []
        for i in range(28):
            a = pin_in[i % 2].numpy()[: n * msg]; o = pin_out.numpy()[: n * msg]
            t0 = time.perf_counter(); eng.submit(a, out=o); ts.append(time.perf_counter() - t0)
        ts = sorted(ts[4:])
        line += f" | N={n}: min {ts[0]*1e6:.0f} med {ts[len(ts)//2]*1e6:.0f} us = {n/ts[len(ts)//2]/1e9:.2f} G/s"
    print(line, flush=True)
    del eng
