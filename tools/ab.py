#!/usr/bin/env python
"""A/B helper: fasst (uniform 24M ids) + store GET kernel times for the current env settings."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from dint_b200 import Engine, wire
import trace_gen as T
from gpu_probe import probe
n = 1 << 22
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("DINT_"))
print("==", tag or "defaults")
probe(wire.FASST, T.fasst_random(n, 24_000_000, seed=1, weights=(0.6, 0.15, 0.05, 0.2)), chunk=int(os.environ.get("CHUNK", 1 << 20)))
if "--store" in sys.argv:
    probe(wire.STORE, T.store_random(n, 2_000_000, seed=4, p_set=0.0, p_miss=0.0), chunk=1 << 20, populate=True)
if "--hot" in sys.argv:
    probe(wire.FASST, T.fasst_random(n, 4800, seed=1), chunk=1 << 20)
    probe(wire.LOCK2PL, T.lock2pl_random(n, 4800, seed=2), chunk=1 << 20)
