"""The size-independent properties (tests/properties.py) on the CPU oracle: this is how the property code itself is
pinned before tests/test_gpu_properties.py applies it to the CUDA engine at BASELINE.json's full sizes."""
import numpy as np
import pytest

import oracle_lib as O
import properties as P
import trace_gen as T
from dint_b200 import wire


class OracleServer:
    def __init__(self, kind, lock_slots=36000000, **cfg):
        self.ora = O.Oracle(kind, lock_slots=lock_slots, **cfg)
        self.lock_slots = lock_slots

    def submit(self, req):
        return self.ora.process(req)


def test_vectorised_fasthash_matches_the_reference_hash():
    import struct
    rng = np.random.default_rng(3)
    xs = np.concatenate([[0, 1, 7, 9, 4799, 23999999, 0xFFFFFFFF], rng.integers(0, 2**32, size=200)]).astype(np.uint64)
    got = P.fasthash64_u32(xs)
    for x, h in zip(xs.tolist(), got.tolist()):
        assert O.fasthash64(struct.pack("<I", x)) == h, x


@pytest.mark.parametrize("lock_slots,n_keys", [(36000000, 24000000), (1009, 5000), (16, 16)])
def test_fasst_properties_on_the_oracle(lock_slots, n_keys):
    srv = OracleServer(wire.FASST, lock_slots=lock_slots)
    assert P.fasst_acquire_abort_roundtrip(srv, 20000, n_keys, seed=1) > 0
    assert P.fasst_commit_checksum(srv, 20000, n_keys, seed=2) > 0
    assert P.fasst_version_counts_commits(srv, 40000) >= 40000


@pytest.mark.parametrize("lock_slots,n_keys", [(36000000, 24000000), (1009, 5000)])
def test_lock2pl_properties_on_the_oracle(lock_slots, n_keys):
    srv = OracleServer(wire.LOCK2PL, lock_slots=lock_slots)
    assert P.lock2pl_counters_balance(srv, 20000, n_keys, seed=3) > 0


def test_store_properties_on_the_oracle():
    srv = OracleServer(wire.STORE, subs_populate=500)
    keys = wire.as_records(wire.STORE, T.store_random(8000, 500, seed=4, p_set=0.0, p_miss=0.0))["key"].copy()
    assert P.store_read_your_writes(srv, keys, seed=5) == 8000


def test_the_gpu_suites_exact_calls_hold_on_the_oracle_at_full_size():
    """tests/test_gpu_properties.py, call for call (2^21 requests, 36 M slots, 24 M ids, the 24 M-key store), on the oracle:
    whatever fails there on the GPU is then a parity bug of the engine, not a bug of the property."""
    N = 1 << 21
    srv = OracleServer(wire.FASST)
    assert P.fasst_acquire_abort_roundtrip(srv, N, 24_000_000, seed=1) > N // 2
    assert P.fasst_commit_checksum(srv, N, 24_000_000, seed=2) > N // 2
    assert P.fasst_acquire_abort_roundtrip(srv, N, 4800, seed=3) <= 4800
    assert P.fasst_commit_checksum(srv, N, 4800, seed=4) <= 3 * 4800
    srv = OracleServer(wire.LOCK2PL)
    assert P.lock2pl_counters_balance(srv, N, 24_000_000, seed=5) > N // 2
    assert P.lock2pl_counters_balance(srv, N, 4800, seed=6) > 0
    srv = OracleServer(wire.STORE, subs_populate=2_000_000)
    keys = wire.as_records(wire.STORE, T.store_random(N, 2_000_000, seed=7, p_set=0.0, p_miss=0.0))["key"].copy()
    assert P.store_read_your_writes(srv, keys, seed=8) == N
    hot = wire.as_records(wire.STORE, T.store_random(N, 50, seed=9, p_set=0.0, p_miss=0.0))["key"].copy()
    assert P.store_read_your_writes(srv, hot, seed=10) == N
