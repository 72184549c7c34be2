"""-m gpu, needs >= 2 GPUs: the real NCCL path of dint_b200.shard with one GPU engine per rank, against ONE
sequential oracle fed the rank-major concatenation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, n_per_rank, mode, ret):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    import oracle_lib as O
    import trace_gen as T
    from dint_b200 import wire
    from dint_b200.shard import ShardedEngine
    try:
        if kind == wire.FASST:
            mk, cfg = (lambda r: T.fasst_random(n_per_rank, 3000, seed=5 + r)), {}
        elif kind == wire.LOCK2PL:
            mk, cfg = (lambda r: T.lock2pl_random(n_per_rank, 500, seed=5 + r)), {}
        else:
            mk, cfg = (lambda r: T.store_random(n_per_rank, 400, seed=5 + r)), dict(subs_populate=400)
        se = ShardedEngine(kind, chunk=1 << 13, use_slabs=(mode == 'slabs'), use_p2p=mode.startswith('p2p'), p2p_max_n=n_per_rank, **cfg)
        if kind == wire.STORE:
            se.populate()
        got1 = se.submit(mk(rank))
        got2 = se.submit(mk(rank + 50))
        seq = O.Oracle(kind, **cfg)
        want1 = seq.process(np.concatenate([mk(r) for r in range(world)]))
        want2 = seq.process(np.concatenate([mk(r + 50) for r in range(world)]))
        msg = wire.MSG_SIZE[kind]
        lo, hi = rank * n_per_rank * msg, (rank + 1) * n_per_rank * msg
        ok = bool(np.array_equal(got1, want1[lo:hi]) and np.array_equal(got2, want2[lo:hi]))
        # a pipelined sequence of batches (dispatch of k+1 | engine of k | combine of k-1) = the same batches in turn
        K = 7
        dev = [torch.from_numpy(np.ascontiguousarray(mk(rank + 100 + 10 * i)).view(np.uint8).reshape(-1)).cuda() for i in range(K)]
        outs = se.submit_many(dev)
        torch.cuda.synchronize()
        for i in range(K):
            want = seq.process(np.concatenate([mk(r + 100 + 10 * i) for r in range(world)]))
            ok = ok and bool(np.array_equal(outs[i].cpu().numpy(), want[lo:hi]))
        if mode.startswith('p2p'):
            # batches of different sizes (per rank too) in one pipelined call, every batch's slabs sized to the batch
            sizes = [n_per_rank - 1000 * i - 7 * r for i in range(5) for r in [rank]]
            varied = [np.ascontiguousarray(mk(rank + 300 + 10 * i)).view(np.uint8).reshape(-1, msg)[:sizes[i]].reshape(-1) for i in range(5)]
            caps = [se.cap_for(n_per_rank - 1000 * i) for i in range(5)]
            outs = se.submit_many([torch.from_numpy(v).cuda() for v in varied], caps=caps)
            torch.cuda.synchronize()
            for i in range(5):
                parts = [np.ascontiguousarray(mk(r + 300 + 10 * i)).view(np.uint8).reshape(-1, msg)[:n_per_rank - 1000 * i - 7 * r].reshape(-1) for r in range(world)]
                want = seq.process(np.concatenate(parts))
                lo2 = sum(p.size for p in parts[:rank])
                ok = ok and bool(np.array_equal(outs[i].cpu().numpy(), want[lo2:lo2 + parts[rank].size]))
            ok = ok and se.check_p2p() == (0, 0)
        else:
            ok = ok and not se.check_overflow()
        ret[rank] = ok
        se.close()
    finally:
        dist.destroy_process_group()


MODES = ["exact", "slabs", "p2p"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("kind", [0, 1, 3])
def test_sharded_nccl_world2(kind, mode):
    """exact = variable-count NCCL all-to-all; slabs = fixed-capacity NCCL exchange with padding records;
    p2p = fused dispatch / combine through NVLink peer memory (no NCCL on the data path)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), kind, 20000, mode, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
