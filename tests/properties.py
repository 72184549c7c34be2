"""Size-independent properties of the hot path, written once against a minimal server interface and run twice:
on the CPU oracle at small sizes (tests/test_properties_cpu.py, which is how the property code itself is checked)
and on the CUDA engine at BASELINE.json's full table sizes (tests/test_gpu_properties.py), where an oracle replay of
the same volume would take too long for a unit test.

server: any object with `.submit(req_bytes) -> resp_bytes` (uint8 arrays of packed wire structs).
Every property is a consequence of the reference's handler semantics (file:line cited per property) for ONE
sequential server; none depends on the table size or on which keys collide in a slot.
"""
import numpy as np

from dint_b200 import wire
from dint_b200.wire import Fasst, Lock2pl, Store

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(h):
    h = h ^ (h >> np.uint64(23))
    h = h * np.uint64(0x2127599BF4325C37)
    return h ^ (h >> np.uint64(47))


def fasthash64_u32(x, seed=0xDEADBEEF):
    """Vectorised fasthash64(&x, 4, seed) of the reference (lock_2pl/udp/utils.h:20-57, tail-switch path)."""
    with np.errstate(over="ignore"):
        m = np.uint64(0x880355F21E6D1965)
        h = np.uint64(seed) ^ (np.uint64(4) * m)
        v = np.asarray(x, dtype=np.uint64)
        h = (h ^ _mix(v)) * m
        return _mix(h)


def fasst_records(types, lids, vers=0):
    rec = np.zeros(len(lids), dtype=wire.MSG_DTYPE[wire.FASST])
    rec["type"], rec["lid"], rec["ver"] = types, lids, vers
    return wire.as_bytes(rec)


def lock2pl_records(actions, lids, types):
    rec = np.zeros(len(lids), dtype=wire.MSG_DTYPE[wire.LOCK2PL])
    rec["action"], rec["lid"], rec["type"] = actions, lids, types
    return wire.as_bytes(rec)


# ------------------------------------------------------------------------------------------------------------
def fasst_acquire_abort_roundtrip(server, n, n_keys, seed):
    """lock_fasst/udp/server.cc:92-107: ACQUIRE is CAS(0->1), ABORT is CAS(1->0).  Acquiring a batch of ids, aborting
    exactly the granted ones and acquiring the same batch again must reproduce the first grant pattern (the table is
    back where it started), and every reject of the first pass is a slot some EARLIER request of the pass won."""
    rng = np.random.default_rng(seed)
    lids = rng.integers(0, n_keys, size=n).astype(np.uint32)
    acq = fasst_records(Fasst.kAcquireLock, lids)
    r1 = wire.as_records(wire.FASST, server.submit(acq))
    granted = r1["type"] == Fasst.kGrantLock
    assert set(np.unique(r1["type"])) <= {Fasst.kGrantLock, Fasst.kRejectLock}
    assert np.array_equal(r1["lid"], lids)                          # the reply is the request buffer, mutated
    r2 = wire.as_records(wire.FASST, server.submit(fasst_records(Fasst.kAbort, lids[granted])))
    assert (r2["type"] == Fasst.kAbortAck).all()
    r3 = wire.as_records(wire.FASST, server.submit(acq))
    assert np.array_equal(r3["type"], r1["type"])
    server.submit(fasst_records(Fasst.kAbort, lids[granted]))      # leave the table clean
    # exactly one grant per distinct slot, and it is the FIRST request on that slot (index order = arrival order)
    slot = fasthash64_u32(lids) % np.uint64(server.lock_slots)
    first = np.zeros(n, dtype=bool)
    first[np.unique(slot, return_index=True)[1]] = True
    assert np.array_equal(granted, first)
    return int(granted.sum())


def fasst_commit_checksum(server, n, n_keys, seed):
    """lock_fasst/udp/server.cc:86-90,109-114: COMMIT is ver_table[slot]++ and unlock, READ returns ver_table[slot].
    After k rounds of acquire + commit on the same ids every slot's version has advanced by k x (requests that were
    granted on it per round = 1): sum over distinct slots of (ver_after - ver_before) == number of CommitAcks."""
    rng = np.random.default_rng(seed)
    lids = rng.integers(0, n_keys, size=n).astype(np.uint32)
    rd = fasst_records(Fasst.kRead, lids)
    before = wire.as_records(wire.FASST, server.submit(rd))["ver"].astype(np.int64)
    acks = 0
    for _ in range(3):
        r = wire.as_records(wire.FASST, server.submit(fasst_records(Fasst.kAcquireLock, lids)))
        won = r["type"] == Fasst.kGrantLock
        c = wire.as_records(wire.FASST, server.submit(fasst_records(Fasst.kCommit, lids[won])))
        assert (c["type"] == Fasst.kCommitAck).all()
        acks += int(won.sum())
    rr = wire.as_records(wire.FASST, server.submit(rd))
    assert (rr["type"] == Fasst.kGrantRead).all()
    after = rr["ver"].astype(np.int64)
    slot = fasthash64_u32(lids) % np.uint64(server.lock_slots)
    _, idx = np.unique(slot, return_index=True)
    delta = (after - before) % (1 << 32)
    assert int(delta[idx].sum()) == acks
    # every request that shares a slot reads the same version
    order = np.argsort(slot, kind="stable")
    same = slot[order][1:] == slot[order][:-1]
    assert np.array_equal(after[order][1:][same], after[order][:-1][same])
    return acks


def lock2pl_counters_balance(server, n, n_keys, seed):
    """lock_2pl/udp/server.cc:83-119: shared grants need num_ex == 0, exclusive grants need both counters 0, releases
    decrement.  Releasing exactly what was granted brings every slot back to (0, 0): a second identical pass must
    produce the identical reply stream (idempotence of acquire-all / release-all)."""
    rng = np.random.default_rng(seed)
    lids = rng.integers(0, n_keys, size=n).astype(np.uint32)
    mode = (rng.random(n) < 0.2).astype(np.uint8)                   # 20 % exclusive, as in the reference traces
    acq = lock2pl_records(Lock2pl.kAcquireLock, lids, mode)
    passes = []
    for _ in range(2):
        r = wire.as_records(wire.LOCK2PL, server.submit(acq))
        got = r["action"] == Lock2pl.kGrantLock
        assert set(np.unique(r["action"])) <= {Lock2pl.kGrantLock, Lock2pl.kRejectLock}
        rel = wire.as_records(wire.LOCK2PL, server.submit(lock2pl_records(Lock2pl.kReleaseLock, lids[got], mode[got])))
        assert (rel["action"] == Lock2pl.kReleaseAck).all()
        passes.append(r["action"].copy())
    assert np.array_equal(passes[0], passes[1])
    # an exclusive grant is alone on its slot among the grants of the pass
    got = passes[0] == Lock2pl.kGrantLock
    slot = fasthash64_u32(lids) % np.uint64(server.lock_slots)
    gx = slot[got & (mode == 1)]
    assert len(np.unique(gx)) == len(gx)
    assert not np.isin(slot[got & (mode == 0)], gx).any()
    return int(got.sum())


def store_read_your_writes(server, keys, seed):
    """store/udp/kvs.h:37-75: kvs_set overwrites the 40-byte value and increments the version, kvs_get returns both.
    GET all, SET all with fresh values, GET all: every key reads back the LAST value written to it and its version
    advanced by the number of SETs it received (keys may repeat inside the batch)."""
    rng = np.random.default_rng(seed)
    n = len(keys)

    def recs(t, vals=None):
        rec = np.zeros(n, dtype=wire.MSG_DTYPE[wire.STORE])
        rec["type"], rec["key"] = t, keys
        if vals is not None:
            rec["val"] = vals
        return wire.as_bytes(rec)

    g0 = wire.as_records(wire.STORE, server.submit(recs(Store.kRead)))
    assert (g0["type"] == Store.kGrantRead).all(), "the property needs keys that exist"
    vals = rng.integers(0, 256, size=(n, 40), dtype=np.uint8)
    s = wire.as_records(wire.STORE, server.submit(recs(Store.kSet, vals)))
    assert (s["type"] == Store.kSetAck).all()
    g1 = wire.as_records(wire.STORE, server.submit(recs(Store.kRead)))
    uniq, inv, cnt = np.unique(keys, return_inverse=True, return_counts=True)
    last = np.zeros(len(uniq), dtype=np.int64)
    np.maximum.at(last, inv, np.arange(n))                           # the last index at which each key was written
    assert np.array_equal(g1["val"], vals[last[inv]])
    assert np.array_equal((g1["ver"].astype(np.int64) - g0["ver"].astype(np.int64)) % (1 << 32), cnt[inv])
    return n


def fasst_version_counts_commits(server, k, lid=12345):
    """lock_fasst/udp/server.cc:92-114: [ACQUIRE, COMMIT] x k on ONE id inside ONE batch.  Sequential semantics grant
    every acquire (the previous commit released the lock) and the version advances by exactly k -- across the 15-bit
    boundary of the experimental 16-bit version layout when k > 32767."""
    before = wire.as_records(wire.FASST, server.submit(fasst_records([Fasst.kRead], np.array([lid], dtype=np.uint32))))["ver"][0]
    types = np.tile(np.array([Fasst.kAcquireLock, Fasst.kCommit], dtype=np.uint8), k)
    r = wire.as_records(wire.FASST, server.submit(fasst_records(types, np.full(2 * k, lid, dtype=np.uint32))))
    assert (r["type"][0::2] == Fasst.kGrantLock).all() and (r["type"][1::2] == Fasst.kCommitAck).all()
    after = wire.as_records(wire.FASST, server.submit(fasst_records([Fasst.kRead], np.array([lid], dtype=np.uint32))))["ver"][0]
    assert (int(after) - int(before)) % (1 << 32) == k
    return int(after)
