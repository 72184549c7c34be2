"""-m gpu parity tests: the CUDA path (through the C ABI of libdint_b200.so) against the oracle.

Bar: bit-exact response streams (integer/byte work) and bit-exact final server state.
"""
import os
import numpy as np
import pytest

import oracle_lib as O
import trace_gen as T
from dint_b200 import Engine, wire
from dint_b200.workloads import Workload, record_trace, REF, HOT

pytestmark = pytest.mark.gpu


def first_diff(a, b, msg):
    a = a.reshape(-1, msg)
    b = b.reshape(-1, msg)
    bad = np.nonzero((a != b).any(axis=1))[0]
    if bad.size == 0:
        return None
    i = int(bad[0])
    return f"{bad.size} of {a.shape[0]} replies differ; first at {i}: got {a[i].tolist()} want {b[i].tolist()}"


def check_stream(eng, req, want, chunked_device=True):
    got = eng.submit(req)
    d = first_diff(got, want, eng.msg)
    assert d is None, d


# ---------------------------------------------------------------- lock_fasst -------------------------
@pytest.mark.parametrize("chunk", [256, 4096, 1 << 16])
@pytest.mark.parametrize("n_keys", [8, 300, 100000])
def test_fasst_random(chunk, n_keys):
    req = T.fasst_random(40000, n_keys, seed=n_keys + chunk)
    ora = O.Oracle(wire.FASST)
    want = ora.process(req)
    with Engine(wire.FASST, chunk=chunk) as eng:
        check_stream(eng, req, want)
        for lid in range(min(n_keys, 50)):
            s = eng.lock_slot(0, lid)
            assert s == ora.lock_slot(0, lid)
            assert eng.lock_state(0, s) == ora.lock_state(0, s)
        st = eng.stats()
        assert st["requests"] == 40000 and st["errors"] == 0


@pytest.mark.parametrize("fam", [REF, HOT], ids=["REF", "HOT"])
def test_fasst_closed_loop(fam):
    ora = O.Oracle(wire.FASST)
    wl = Workload(wire.FASST, n_clients=2048, seed=20230, **fam)
    req, want = record_trace(wl, ora.process, 60)
    with Engine(wire.FASST, chunk=1 << 15) as eng:
        check_stream(eng, req, want)
    # the same engine code driven closed-loop must take the same decisions round by round
    wl2 = Workload(wire.FASST, n_clients=2048, seed=20230, **fam)
    with Engine(wire.FASST) as eng:
        req2, got2 = record_trace(wl2, eng.submit, 60)
    assert np.array_equal(req2, req) and np.array_equal(got2, want)
    assert wl2.stats() == wl.stats()


# ---------------------------------------------------------------- lock_2pl ---------------------------
@pytest.mark.parametrize("chunk", [256, 1 << 14])
@pytest.mark.parametrize("n_keys", [5, 2000])
def test_lock2pl_random(chunk, n_keys):
    req = T.lock2pl_random(30000, n_keys, seed=7 * n_keys + chunk)
    ora = O.Oracle(wire.LOCK2PL)
    want = ora.process(req)
    with Engine(wire.LOCK2PL, chunk=chunk) as eng:
        check_stream(eng, req, want)
        for lid in range(min(n_keys, 50)):
            s = eng.lock_slot(0, lid)
            assert eng.lock_state(0, s) == ora.lock_state(0, s)


@pytest.mark.parametrize("fam", [REF, HOT], ids=["REF", "HOT"])
def test_lock2pl_closed_loop(fam):
    ora = O.Oracle(wire.LOCK2PL)
    wl = Workload(wire.LOCK2PL, n_clients=2048, seed=1, **fam)
    req, want = record_trace(wl, ora.process, 60)
    with Engine(wire.LOCK2PL, chunk=1 << 16) as eng:
        check_stream(eng, req, want)


# ---------------------------------------------------------------- log_server -------------------------
@pytest.mark.parametrize("ring,chunk", [(1000000, 1 << 14), (1000, 4096), (7, 256)])
def test_log(ring, chunk):
    req = T.log_random(20000, seed=ring)
    ora = O.Oracle(wire.LOG, log_ring=ring)
    want = ora.process(req)
    with Engine(wire.LOG, log_ring=ring, chunk=chunk) as eng:
        check_stream(eng, req, want)
        ring_got, appended = eng.dump_log()
        assert appended == ora.log_appended() == 20000
        assert np.array_equal(ring_got, ora.log_ring())


# ---------------------------------------------------------------- store ------------------------------
@pytest.mark.parametrize("chunk", [256, 1 << 14])
@pytest.mark.parametrize("n_subs", [3, 2000])
def test_store_random(chunk, n_subs):
    req = T.store_random(30000, n_subs, seed=n_subs + chunk)
    ora = O.Oracle(wire.STORE, subs_populate=n_subs)
    want = ora.process(req)
    with Engine(wire.STORE, subs_populate=n_subs, chunk=chunk, populate=True) as eng:
        assert eng.kv_count(0) == ora.kv_count(0) == 12 * n_subs
        check_stream(eng, req, want)
        for s in range(min(n_subs, 20)):
            for sf in (1, 4):
                k = int(T.store_key(s, sf, 8))
                assert eng.kv_get(0, k) == ora.kv_get(0, k)


def test_store_closed_loop_contention():
    ora = O.Oracle(wire.STORE, subs_populate=20000)
    wl = Workload(wire.STORE, n_clients=4096, set_pct=50, store_subscribers=20000)
    req, want = record_trace(wl, ora.process, 20)
    with Engine(wire.STORE, subs_populate=20000, populate=True) as eng:
        check_stream(eng, req, want)


# ---------------------------------------------------------------- smallbank --------------------------
@pytest.mark.parametrize("chunk", [256, 1 << 14])
@pytest.mark.parametrize("n_accts", [4, 3000])
def test_smallbank_random(chunk, n_accts):
    req = T.smallbank_random(30000, n_accts, seed=n_accts + chunk)
    ora = O.Oracle(wire.SMALLBANK, accts_populate=n_accts)
    want = ora.process(req)
    with Engine(wire.SMALLBANK, accts_populate=n_accts, chunk=chunk, populate=True) as eng:
        check_stream(eng, req, want)
        for tb in (0, 1):
            for a in range(min(n_accts, 20)):
                assert eng.kv_get(tb, a)[0][:8] == ora.kv_get(tb, a)[0][:8]
                assert eng.kv_get(tb, a)[1] == ora.kv_get(tb, a)[1]
                s = eng.lock_slot(tb, a)
                assert s == ora.lock_slot(tb, a)
                assert eng.lock_state(tb, s) == ora.lock_state(tb, s)
        ring_got, appended = eng.dump_log()
        assert appended == ora.log_appended()
        assert np.array_equal(ring_got, ora.log_ring())


# ---------------------------------------------------------------- tatp -------------------------------
@pytest.mark.parametrize("chunk", [256, 1 << 14])
@pytest.mark.parametrize("n_subs", [2, 300])
def test_tatp_random(chunk, n_subs):
    ora = O.Oracle(wire.TATP, subs_populate=n_subs)
    req = T.tatp_random(20000, n_subs, seed=n_subs + chunk, oracle=ora)
    want = ora.process(req)
    with Engine(wire.TATP, subs_populate=n_subs, chunk=chunk, populate=True) as eng:
        check_stream(eng, req, want)
        for tb, key in T.tatp_key_universe(min(n_subs, 10)):
            assert eng.kv_get(tb, key) == ora.kv_get(tb, key), (tb, hex(key))
            s = eng.lock_slot(tb, key)
            assert s == ora.lock_slot(tb, key)
            assert eng.lock_state(tb, s)[0] == ora.lock_state(tb, s)[0]
        for tb in range(5):
            assert eng.kv_count(tb) == ora.kv_count(tb)
        ring_got, appended = eng.dump_log()
        assert appended == ora.log_appended()
        assert np.array_equal(ring_got, ora.log_ring())


# ---------------------------------------------------------------- closed-loop transaction drivers ----
@pytest.mark.parametrize("kind,n,clients", [(wire.TATP, 3000, 2000), (wire.SMALLBANK, 5000, 1500)])
def test_txn_drivers_closed_loop_three_shards(kind, n, clients):
    """TATP (7 txn types) / SmallBank (6) clients against three shard servers: GPU engines must take the
    same commit/abort decisions, reply for reply, as three oracle servers."""
    from dint_b200.txn_workloads import TxnWorkload, Cluster
    cfg = dict(subs_populate=n) if kind == wire.TATP else dict(accts_populate=n)
    msg = wire.MSG_SIZE[kind]

    def run(servers):
        wl = TxnWorkload(kind, n_clients=clients, n_shards=3, subscribers=n)
        cl = Cluster(servers, msg)
        trace = []
        for _ in range(80):
            rq, dst = wl.next()
            rs = cl.submit(rq, dst)
            wl.feed(rs)
            trace.append((rq.copy(), dst.copy(), rs.copy()))
        return trace, wl.stats()

    oras = [O.Oracle(kind, **cfg) for _ in range(3)]
    want, st_want = run([o.process for o in oras])
    engs = [Engine(kind, populate=True, chunk=4096, **cfg) for _ in range(3)]
    try:
        got, st_got = run([e.submit for e in engs])
        for r, ((q1, d1, s1), (q2, d2, s2)) in enumerate(zip(want, got)):
            assert np.array_equal(q1, q2) and np.array_equal(d1, d2), f"round {r}: clients diverged"
            assert first_diff(s2, s1, msg) is None, f"round {r}: {first_diff(s2, s1, msg)}"
        assert st_got == st_want and st_got["committed"] > 0
        for e, o in zip(engs, oras):
            for tb in range(5 if kind == wire.TATP else 2):
                assert e.kv_count(tb) == o.kv_count(tb)
            ring, appended = e.dump_log()
            assert appended == o.log_appended() and np.array_equal(ring, o.log_ring())
    finally:
        for e in engs:
            e.close()


def test_kv_tombstones_are_reclaimed_under_insert_delete_churn():
    """The reference's chained kvs frees entries on delete (store/udp/kvs.h:124-133).  Here deletes leave tombstones;
    a churn of ever-new call-forwarding rows through a 1024-entry table (20x its capacity in total) must neither fill
    the table nor change a single reply, and the engine must have rehashed the table on the way (kv_rebuilds)."""
    from dint_b200.wire import Tatp
    n_subs = 20
    cfg = dict(subs_populate=n_subs, kv_capacity_log2=[0, 0, 0, 0, 10])
    ora = O.Oracle(wire.TATP, subs_populate=n_subs)
    rng = np.random.default_rng(5)
    with Engine(wire.TATP, populate=True, chunk=4096, **cfg) as eng:
        next_sid = 1000
        for call in range(40):
            m = 512                                                # rows inserted, read and deleted again by this call
            keys = (np.arange(next_sid, next_sid + m, dtype=np.uint64) | (np.uint64(1) << np.uint64(32)) | (np.uint64(8) << np.uint64(40)))
            next_sid += m
            rec = np.zeros(4 * m, dtype=wire.MSG_DTYPE[wire.TATP])
            rec["table"] = Tatp.kCallForwarding
            rec["key"] = np.concatenate([keys, keys, keys, keys + np.uint64(1 << 20)])
            rec["type"] = np.concatenate([np.full(m, Tatp.kInsertBck), np.full(m, Tatp.kRead), np.full(m, Tatp.kDeleteBck),
                                          np.full(m, Tatp.kRead)]).astype(np.uint8)       # last quarter: keys that never existed
            rec["val"] = rng.integers(0, 256, size=(4 * m, 40))
            req = wire.as_bytes(rec)
            want = ora.process(req)
            got = eng.submit(req)
            assert first_diff(got, want, 55) is None, f"call {call}: {first_diff(got, want, 55)}"
        st = eng.stats()
        assert st["kv_rebuilds"] >= 1, st
        assert st["errors"] == 0
        assert eng.kv_count(4) == ora.kv_count(4)


# ---------------------------------------------------------------- device path / edge cases -----------
def test_device_path_and_empty():
    import torch
    req = T.fasst_random(100000, 5000, seed=3)
    want = O.Oracle(wire.FASST).process(req)
    with Engine(wire.FASST, chunk=1 << 13) as eng:
        assert eng.submit(np.zeros(0, dtype=np.uint8)).size == 0           # empty batch
        d_req = torch.from_numpy(req).cuda()
        d_out = eng.submit_tensor(d_req)
        eng.sync()
        torch.cuda.synchronize()
        assert first_diff(d_out.cpu().numpy(), want, 9) is None
        # ragged tail: n not a multiple of the 256-record tile, 16-byte TMA body + byte tail
    for n in (1, 2, 255, 257, 1000):
        r = req[: n * 9]
        w = O.Oracle(wire.FASST).process(r)
        with Engine(wire.FASST) as eng:
            assert first_diff(eng.submit(r), w, 9) is None


def test_invalid_requests_are_flagged_not_fatal():
    rec = np.zeros(6, dtype=wire.MSG_DTYPE[wire.FASST])
    rec["type"] = [0, 9, 1, 200, 3, 0]
    rec["lid"] = 77
    with Engine(wire.FASST) as eng:
        out = wire.as_records(wire.FASST, eng.submit(wire.as_bytes(rec), check=False))
        assert out["type"].tolist() == [4, 0xFF, 5, 0xFF, 8, 4]
        assert out["ver"][5] == 1
        assert eng.stats()["errors"] == 2


# ---------------------------------------------------------------- reference golden fixtures ----------
import golden_util as G   # noqa: E402


@pytest.mark.parametrize("name", G.names())
def test_engine_reproduces_reference_golden(name):
    kind, req, resp, cfg = G.load(name)
    with Engine(kind, populate=True, chunk=2048, **cfg) as eng:
        ours = eng.submit(req)
    assert G.mismatch(kind, ours, resp) is None, G.mismatch(kind, ours, resp)


def test_state_cloned_from_a_reference_server_answers_strictly_bit_exact():
    """sweep a live reference server with kRead, dint_load what it answered (instead of
    dint_populate) -- then even the value bytes the reference's populate_* leaves indeterminate are the reference's, and
    the comparison with the reference binary's replies needs NO tolerance.  The sweep and the replies are the golden
    fixture recorded from the unmodified tatp server (tools/make_golden.py)."""
    from dint_b200.wire import Tatp
    kind, req, resp, cfg = G.load("tatp_sweep_random")
    assert kind == wire.TATP
    rq = wire.as_records(kind, req)
    rs = wire.as_records(kind, resp)
    n_sweep = len(T.tatp_key_universe(cfg["subs_populate"]))
    hit = rs["type"][:n_sweep] == Tatp.kGrantRead
    with Engine(wire.TATP, subs_populate=cfg["subs_populate"], chunk=2048) as eng:      # tables sized as usual, NOT populated
        for tb in range(5):
            sel = hit & (rq["table"][:n_sweep] == tb)
            if sel.any():
                eng.load(tb, rq["key"][:n_sweep][sel], np.ascontiguousarray(rs["val"][:n_sweep][sel]))
        got = eng.submit(req)
        d = first_diff(got, resp, 55)
        assert d is None, d                                                              # strict: no indeterminate-byte mask


# ---------------------------------------------------------------- multi-GPU dispatch kernels (1 GPU) -----
def _route_case(kind, n, seed=3):
    cfg = {}
    if kind == wire.FASST:
        req = T.fasst_random(n, 10**6, seed=seed)
    elif kind == wire.LOCK2PL:
        req = T.lock2pl_random(n, 10**5, seed=seed)
    elif kind == wire.LOG:
        req = T.log_random(n, seed=seed)
    elif kind == wire.STORE:
        req = T.store_random(n, 500, seed=seed)
        cfg = dict(subs_populate=500)
    elif kind == wire.SMALLBANK:
        req = T.smallbank_random(n, 300, seed=seed)
        cfg = dict(accts_populate=300)
    else:
        req = T.tatp_random(n, 20, seed=seed)
        cfg = dict(subs_populate=20)
    return np.ascontiguousarray(req).view(np.uint8).reshape(-1), cfg


@pytest.mark.parametrize("kind,world,n,by_dst", [
    (wire.FASST, 2, 70001, False), (wire.FASST, 8, (1 << 20) + 13, False), (wire.FASST, 1, 5000, False),
    (wire.LOCK2PL, 3, 4097, False), (wire.LOG, 2, 1000, False), (wire.STORE, 4, 70001, False),
    (wire.TATP, 3, 30011, True), (wire.TATP, 5, 30011, False), (wire.SMALLBANK, 6, 12345, True), (wire.FASST, 4, 1, False),
])
def test_route_dispatch_combine(kind, world, n, by_dst):
    """k_route_dispatch / k_route_combine against numpy: slab o = the records owned by shard o in request order,
    then padding; the combine of the slabs themselves is the identity; a client-chosen shard >= world is
    undeliverable (0xFF reply)."""
    import torch
    from dint_b200.shard import owners_cpu
    req, cfg = _route_case(kind, n)
    msg = wire.MSG_SIZE[kind]
    rank = world - 1
    rng = np.random.default_rng(5)
    with Engine(kind, n_shards=1 if by_dst else world, shard_id=0 if by_dst else rank, **cfg) as eng:
        d = torch.from_numpy(req).cuda()
        if by_dst:
            want_owner = rng.integers(0, world, n).astype(np.uint8)
            want_owner[rng.integers(0, n, 7)] = 200                      # shards that do not exist
            owner_in = torch.from_numpy(want_owner).cuda()
        else:
            want_owner = owners_cpu(kind, eng.cfg, world, rank, req)
            owner_in = None
        counts = np.bincount(want_owner[want_owner < world], minlength=world)[:world]
        cap = (int(counts.max()) + 16 + 15) // 16 * 16
        slabs = torch.zeros(world * cap * msg, dtype=torch.uint8, device="cuda")
        flags = torch.zeros(2, dtype=torch.int32, device="cuda")
        ptrs = Engine.slab_ptrs(slabs.data_ptr(), world, cap * msg)
        state = eng.route_dispatch(d, n, world, rank, cap, ptrs, flags, owner_in=owner_in)
        got = slabs.cpu().numpy().reshape(world, cap, msg)
        rec = req.reshape(n, msg)
        own = state[0].cpu().numpy()[:n]
        assert np.array_equal(own, np.where(want_owner < world, want_owner, 255))
        for o in range(world):
            assert np.array_equal(got[o, :counts[o]], rec[want_owner == o]), o
            assert (got[o, counts[o]:] == 0xFE).all(), o
        assert flags.cpu().tolist() == [0, 0]
        out = torch.empty(n * msg, dtype=torch.uint8, device="cuda")
        eng.route_combine(ptrs, state, n, world, cap, out)
        back = out.cpu().numpy().reshape(n, msg)
        ok = want_owner < world
        assert np.array_equal(back[ok], rec[ok])
        assert (back[~ok] == 0xFF).all()


def test_route_dispatch_overflow_is_counted():
    import torch
    n, world = 50000, 2
    req, cfg = _route_case(wire.FASST, n)
    with Engine(wire.FASST, n_shards=world, shard_id=0) as eng:
        d = torch.from_numpy(req).cuda()
        cap = 16000                                                      # < n / 2
        slabs = torch.zeros(world * cap * 9, dtype=torch.uint8, device="cuda")
        flags = torch.zeros(2, dtype=torch.int32, device="cuda")
        ptrs = Engine.slab_ptrs(slabs.data_ptr(), world, cap * 9)
        state = eng.route_dispatch(d, n, world, 0, cap, ptrs, flags)
        own = state[0].cpu().numpy()[:n]
        counts = np.bincount(own, minlength=world)[:world]
        assert flags.cpu().tolist()[0] == int(np.maximum(counts - cap, 0).sum()) > 0
        rec = req.reshape(n, 9)
        got = slabs.cpu().numpy().reshape(world, cap, 9)
        for o in range(world):
            assert np.array_equal(got[o], rec[own == o][:cap])


@pytest.mark.parametrize("kind,world", [(wire.FASST, 2), (wire.FASST, 8), (wire.STORE, 3), (wire.TATP, 5)])
def test_route_owner_partition_unpermute(kind, world):
    """k_route_owner / k_route_count+scan+scatter / k_route_unpermute against numpy: owner = the slot one
    server would compute, modulo the shard count; the partition is stable; unpermute is its inverse."""
    import torch
    from dint_b200.shard import owners_cpu
    n = 70001
    cfg = {}
    if kind == wire.FASST:
        req = T.fasst_random(n, 10**6, seed=3)
    elif kind == wire.STORE:
        req = T.store_random(n, 500, seed=3)
        cfg = dict(subs_populate=500)
    else:
        req = T.tatp_random(n, 20, seed=3)
        cfg = dict(subs_populate=20)
    msg = wire.MSG_SIZE[kind]
    with Engine(kind, n_shards=world, shard_id=1, **cfg) as eng:
        d = torch.from_numpy(req).cuda()
        owner = eng.route_owner(d)
        want_owner = owners_cpu(kind, eng.cfg, world, 1, req)
        assert np.array_equal(owner.cpu().numpy(), want_owner)
        srt, perm, counts = eng.route_partition(d, owner, world)
        order = np.argsort(want_owner, kind="stable")
        assert np.array_equal(perm.cpu().numpy().astype(np.int64), order)
        assert np.array_equal(counts.cpu().numpy(), np.bincount(want_owner, minlength=world)[:world])
        assert np.array_equal(srt.cpu().numpy().reshape(-1, msg), req.reshape(-1, msg)[order])
        back = eng.route_unpermute(srt, perm)
        assert np.array_equal(back.cpu().numpy(), req)


# ---------------------------------------------------------------- the UDP front-end ---------------------
def _udp_cases():
    return [
        ("lock_2pl", wire.LOCK2PL, lambda: T.lock2pl_random(12000, 2000, seed=22), {}, ()),
        ("lock_fasst", wire.FASST, lambda: T.fasst_random(20000, 3000, seed=21), {}, ()),
        ("log_server", wire.LOG, lambda: T.log_random(6000, seed=23), {}, ()),
        ("store", wire.STORE, lambda: T.store_random(8000, 400, seed=24), dict(subs_populate=400), ("--populate", "400")),
        ("tatp", wire.TATP, lambda: T.tatp_random(8000, 30, seed=25), dict(subs_populate=30), ("--populate", "30")),
        ("smallbank", wire.SMALLBANK, lambda: T.smallbank_random(12000, 500, seed=26), dict(accts_populate=500), ("--populate", "500")),
        # the same server with the key space on several shards (here all resident on GPU 0): owner = slot % 2
        ("lock_fasst", wire.FASST, lambda: T.fasst_random(20000, 3000, seed=27), {}, ("--gpus", "2", "--devices", "0,0")),
    ]


@pytest.mark.parametrize("name,kind,make,cfg,extra", _udp_cases(), ids=[c[0] + ("+gpus" if "--gpus" in c[4] else "") for c in _udp_cases()])
def test_udp_front_end_serves_the_wire_protocol_bit_exact(name, kind, make, cfg, extra):
    """dint_udp_server behind a real socket, every server kind: one client socket, windows of 64 datagrams (loopback
    keeps their order), replies must equal ONE sequential reference server's."""
    import socket
    import subprocess
    import time
    from dint_b200 import _build
    msg = wire.MSG_SIZE[kind]
    req = make()
    want = O.Oracle(kind, **cfg).process(req).reshape(-1, msg)
    with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as s0:
        s0.bind(("127.0.0.1", 0))
        port = s0.getsockname()[1]
    srv = subprocess.Popen([_build.UDP_SERVER, name, "--port", str(port), "--bind", "127.0.0.1", *extra], stderr=subprocess.PIPE)
    try:
        os.set_blocking(srv.stderr.fileno(), False)
        banner, t0 = b"", time.time()
        while b"sockets, batches" not in banner:           # printed once the engine exists and the sockets are bound
            assert srv.poll() is None and time.time() - t0 < 120, banner
            time.sleep(0.1)
            banner += srv.stderr.read() or b""
        c = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        c.settimeout(5.0)
        c.connect(("127.0.0.1", port))
        rec = np.ascontiguousarray(req).view(np.uint8).reshape(-1, msg)
        got = np.empty_like(rec)
        for lo in range(0, len(rec), 64):
            hi = min(lo + 64, len(rec))
            for i in range(lo, hi):
                c.send(rec[i].tobytes())
            for i in range(lo, hi):
                got[i] = np.frombuffer(c.recv(256), dtype=np.uint8)
        assert first_diff(got, want, msg) is None, first_diff(got, want, msg)
    finally:
        srv.terminate()
        srv.wait(timeout=20)
