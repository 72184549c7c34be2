"""ctypes binding of libdint_wl.so -- the reference's closed-loop clients as round-based generators.

    wl = Workload(wire.FASST, n_clients=4096, seed=20230)
    for _ in range(rounds):
        req = wl.next()                 # uint8 [n_clients * msg], client order = trace order
        wl.feed(server.submit(req))     # any server: the GPU engine, the oracle, the reference binary
"""
import ctypes as C

import numpy as np

from . import _build
from .wire import MSG_SIZE


class WlCfg(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("n_clients", C.c_uint32), ("seed", C.c_uint64), ("n_keys", C.c_uint32),
                ("zipf_theta", C.c_double), ("read_pct", C.c_uint32), ("set_pct", C.c_uint32),
                ("store_subscribers", C.c_uint32), ("store_hot", C.c_uint32), ("reserved", C.c_uint32 * 5)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _build.build()
        L = C.CDLL(_build.WL_LIB)
        L.dint_wl_create.restype = C.c_void_p
        L.dint_wl_create.argtypes = [C.POINTER(WlCfg)]
        L.dint_wl_destroy.argtypes = [C.c_void_p]
        L.dint_wl_next.restype = C.c_uint64
        L.dint_wl_next.argtypes = [C.c_void_p, C.c_void_p]
        L.dint_wl_feed.argtypes = [C.c_void_p, C.c_void_p]
        L.dint_wl_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


# trace families of SURVEY.md section 8(d)
REF = dict(n_keys=24_000_000, zipf_theta=0.0)       # reference-faithful: uniform over 24 M ids
HOT = dict(n_keys=4800, zipf_theta=0.8)             # BASELINE.json's literal reading


class Workload:
    def __init__(self, kind, n_clients, seed=20230, n_keys=24_000_000, zipf_theta=0.0, read_pct=80,
                 set_pct=0, store_subscribers=2_000_000, store_hot=False):
        if kind > 3:
            raise ValueError("tatp / smallbank drivers live in dint_b200.txn_workloads")
        self.kind = kind
        self.msg = MSG_SIZE[kind]
        self.n_clients = n_clients
        cfg = WlCfg(kind=kind, n_clients=n_clients, seed=seed, n_keys=n_keys, zipf_theta=zipf_theta,
                    read_pct=read_pct, set_pct=set_pct, store_subscribers=store_subscribers,
                    store_hot=1 if store_hot else 0)
        self.h = lib().dint_wl_create(C.byref(cfg))
        if not self.h:
            raise RuntimeError("dint_wl_create failed")

    def close(self):
        if self.h:
            lib().dint_wl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def next(self, out=None):
        if out is None:
            out = np.empty(self.n_clients * self.msg, dtype=np.uint8)
        lib().dint_wl_next(self.h, out.ctypes.data)
        return out

    def feed(self, resp):
        resp = np.ascontiguousarray(resp).view(np.uint8).reshape(-1)
        assert resp.size == self.n_clients * self.msg
        lib().dint_wl_feed(self.h, resp.ctypes.data)

    def stats(self):
        out = (C.c_uint64 * 6)()
        lib().dint_wl_stats(self.h, out)
        keys = ["requests", "committed", "validation_aborts", "lock_rejects", "not_exist", "rounds"]
        return dict(zip(keys, [int(x) for x in out]))


def record_trace(wl, server_submit, rounds):
    """Drive `wl` for `rounds` rounds against `server_submit(req)->resp`; returns (requests, responses)."""
    reqs, resps = [], []
    for _ in range(rounds):
        r = wl.next()
        s = server_submit(r)
        wl.feed(s)
        reqs.append(r.copy())
        resps.append(np.array(s, dtype=np.uint8, copy=True).reshape(-1))
    return np.concatenate(reqs), np.concatenate(resps)
