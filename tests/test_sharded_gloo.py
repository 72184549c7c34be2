"""World-size-2 gloo tests (CPU, no GPU): the sharded routing logic of dint_b200.shard -- owner
computation, stable dispatch, variable-count all-to-all, combine -- with the oracle standing in for each
shard's engine.  The routed result must equal ONE sequential server fed the rank-major concatenation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, n_per_rank, seed, ret):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    import trace_gen as T
    from dint_b200 import wire
    from dint_b200.shard import ShardedEngine
    try:
        if kind == wire.FASST:
            mk = lambda r: T.fasst_random(n_per_rank, 200, seed=seed + r)
            cfg = {}
        elif kind == wire.LOCK2PL:
            mk = lambda r: T.lock2pl_random(n_per_rank, 50, seed=seed + r)
            cfg = {}
        else:
            mk = lambda r: T.store_random(n_per_rank, 50, seed=seed + r)
            cfg = dict(subs_populate=50)
        ora = O.Oracle(kind, **cfg)                      # this shard's server (full tables; it only sees its keys)
        se = ShardedEngine(kind, local_submit=ora.process, **cfg)
        mine = mk(rank)
        got = se.submit(mine)
        # a second collective batch checks that shard state carries over
        mine2 = mk(rank + 100)
        got2 = se.submit(mine2)
        # expectation: ONE sequential server over the rank-major concatenation
        seq = O.Oracle(kind, **cfg)
        all1 = np.concatenate([mk(r) for r in range(world)])
        all2 = np.concatenate([mk(r + 100) for r in range(world)])
        want1 = seq.process(all1)
        want2 = seq.process(all2)
        msg = wire.MSG_SIZE[kind]
        lo, hi = rank * n_per_rank * msg, (rank + 1) * n_per_rank * msg
        ret[rank] = bool(np.array_equal(got, want1[lo:hi]) and np.array_equal(got2, want2[lo:hi]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", [0, 1, 3])
def test_sharded_routing_world2(kind):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), kind, 3000, 7, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_owner_host_twin_matches_reference_slot():
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_lib as O
    import trace_gen as T
    from dint_b200 import wire
    from dint_b200.engine import DintCfg
    from dint_b200.shard import owners_cpu
    cfg = DintCfg()
    cfg.lock_slots, cfg.subs_sizing, cfg.accts_sizing = 36000000, 2000000, 24000000
    req = T.fasst_random(2000, 10**6, seed=1)
    own = owners_cpu(wire.FASST, cfg, 8, 0, req)
    ora = O.Oracle(wire.FASST)
    rec = wire.as_records(wire.FASST, req)
    want = np.array([ora.lock_slot(0, int(l)) % 8 for l in rec["lid"]], dtype=np.uint8)
    assert np.array_equal(own, want)
