"""ctypes binding of oracle/libdint_oracle.so and a runner for oracle/_ref -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORA_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORA_DIR, "_ref")

LOCK2PL, FASST, LOG, STORE, TATP, SMALLBANK = range(6)
KIND_NAMES = ["lock_2pl", "lock_fasst", "log_server", "store", "tatp", "smallbank"]
MSG_SIZE = [6, 9, 53, 53, 55, 23]
REF_BIN = ["lock_2pl_server", "lock_fasst_server", "log_server_server", "store_server",
           "tatp_server_shard", "smallbank_server_shard"]


class OracleCfg(C.Structure):
    _fields_ = [("lock_slots", C.c_uint32), ("log_ring", C.c_uint32), ("subs_sizing", C.c_uint32),
                ("subs_populate", C.c_uint32), ("accts_sizing", C.c_uint32), ("accts_populate", C.c_uint32)]


_lib = None


def build_oracle(ref=False):
    """(re)build the C restatement (and, with ref=True and /root/reference present, oracle/_ref)."""
    subprocess.run(["make", "-s", "-C", ORA_DIR, "oracle"] + (["ref"] if ref else []), check=True)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORA_DIR, "libdint_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.dint_oracle_create.restype = C.c_void_p
        L.dint_oracle_create.argtypes = [C.c_int, C.POINTER(OracleCfg)]
        L.dint_oracle_destroy.argtypes = [C.c_void_p]
        L.dint_oracle_populate.argtypes = [C.c_void_p]
        L.dint_oracle_process.restype = C.c_int64
        L.dint_oracle_process.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.dint_oracle_default_cfg.argtypes = [C.c_int, C.POINTER(OracleCfg)]
        L.dint_oracle_fasthash64.restype = C.c_uint64
        L.dint_oracle_fasthash64.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.dint_oracle_kv_get.restype = C.c_int
        L.dint_oracle_kv_get.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint32)]
        L.dint_oracle_kv_count.restype = C.c_uint64
        L.dint_oracle_kv_count.argtypes = [C.c_void_p, C.c_int]
        L.dint_oracle_lock_state.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_uint32)]
        L.dint_oracle_lock_slot.restype = C.c_uint32
        L.dint_oracle_lock_slot.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.dint_oracle_log_appended.restype = C.c_uint64
        L.dint_oracle_log_appended.argtypes = [C.c_void_p]
        L.dint_oracle_log_entry_size.restype = C.c_uint32
        L.dint_oracle_log_entry_size.argtypes = [C.c_void_p]
        L.dint_oracle_log_ring.restype = C.c_void_p
        L.dint_oracle_log_ring.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def default_cfg(kind, **over):
    cfg = OracleCfg()
    lib().dint_oracle_default_cfg(kind, C.byref(cfg))
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def fasthash64(data: bytes, seed=0xDEADBEEF):
    buf = C.create_string_buffer(data, len(data))
    return lib().dint_oracle_fasthash64(buf, len(data), seed)


class Oracle:
    """Sequential CPU restatement of one reference server (see oracle/dint_oracle.c)."""

    def __init__(self, kind, populate=True, **cfg_over):
        self.kind = kind
        self.msg = MSG_SIZE[kind]
        self.cfg = default_cfg(kind, **cfg_over)
        self.h = lib().dint_oracle_create(kind, C.byref(self.cfg))
        if not self.h:
            raise RuntimeError("oracle create failed")
        if populate:
            lib().dint_oracle_populate(self.h)

    def close(self):
        if self.h:
            lib().dint_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def process(self, req: np.ndarray) -> np.ndarray:
        """req: uint8 array of n*msg bytes (or [n, msg]); returns the response stream, same shape."""
        req = np.ascontiguousarray(req, dtype=np.uint8)
        n = req.size // self.msg
        out = np.empty_like(req)
        rc = lib().dint_oracle_process(self.h, req.ctypes.data, n, out.ctypes.data)
        if rc != 0:
            raise ValueError(f"reference would panic() at request {-rc - 1}")
        return out

    def kv_get(self, table, key):
        val = (C.c_uint8 * 40)()
        ver = C.c_uint32(0)
        rc = lib().dint_oracle_kv_get(self.h, table, key, val, C.byref(ver))
        return None if rc else (bytes(val), ver.value)

    def kv_count(self, table):
        return lib().dint_oracle_kv_count(self.h, table)

    def lock_slot(self, table, key):
        return lib().dint_oracle_lock_slot(self.h, table, key)

    def lock_state(self, table, slot):
        out = (C.c_uint32 * 2)()
        lib().dint_oracle_lock_state(self.h, table, slot, out)
        return out[0], out[1]

    def log_appended(self):
        return lib().dint_oracle_log_appended(self.h)

    def log_ring(self):
        es = lib().dint_oracle_log_entry_size(self.h)
        n = self.cfg.log_ring
        p = lib().dint_oracle_log_ring(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n * es,)).reshape(n, es).copy()


def ref_available(kind=None):
    if not os.path.exists(os.path.join(REF_DIR, "replay_shim.so")):
        return False
    kinds = range(6) if kind is None else [kind]
    return all(os.path.exists(os.path.join(REF_DIR, REF_BIN[k])) for k in kinds)


def run_ref(kind, req: np.ndarray, threads=1, repeat=1, want_out=True, spread=False, timeout=1800):
    """Run the UNMODIFIED reference server binary over a request trace; returns (responses|None, stats)."""
    req = np.ascontiguousarray(req, dtype=np.uint8)
    with tempfile.TemporaryDirectory() as td:
        tp, op, sp = (os.path.join(td, x) for x in ("trace.bin", "out.bin", "stats.json"))
        req.tofile(tp)
        env = dict(os.environ, LD_PRELOAD=os.path.join(REF_DIR, "replay_shim.so"), DINT_TRACE=tp,
                   DINT_STATS=sp, DINT_THREADS=str(threads), DINT_REPEAT=str(repeat))
        if want_out:
            env["DINT_OUT"] = op
        if spread:
            env["DINT_SHIM_SPREAD"] = "1"
        argv = [os.path.join(REF_DIR, REF_BIN[kind])]
        if kind in (TATP, SMALLBANK):
            argv.append("2")          # shard id 2: same tables, no CPU-monitor threads (server_shard.cc:308)
        argv.append(str(threads))
        r = subprocess.run(argv, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError(f"reference server exited {r.returncode}: {r.stderr[-500:]!r}")
        stats = json.load(open(sp))
        out = np.fromfile(op, dtype=np.uint8).reshape(req.shape) if want_out else None
    return out, stats


def run_ref_udp(kind, req: np.ndarray, server_threads=8, client_threads=8, window=32, seconds=5.0, port=None):
    """SURVEY 8(d) baseline B1, "the reference UDP server as shipped": the UNMODIFIED server binary with REAL
    sockets on loopback (only its bind address is rewritten, oracle/udp_shim.c), driven by the multi-socket
    replayer oracle/udp_blast.c.  Replies are counted, not compared.  Returns the replayer's JSON dict."""
    import signal
    import socket
    import time
    req = np.ascontiguousarray(req, dtype=np.uint8)
    if port is None:
        with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as s_:
            s_.bind(("127.0.0.1", 0))
            port = s_.getsockname()[1]
    with tempfile.TemporaryDirectory() as td:
        tp = os.path.join(td, "trace.bin")
        req.tofile(tp)
        env = dict(os.environ, LD_PRELOAD=os.path.join(REF_DIR, "udp_shim.so"), DINT_UDP_PORT=str(port))
        argv = [os.path.join(REF_DIR, REF_BIN[kind])]
        if kind in (TATP, SMALLBANK):
            argv.append("2")
        argv.append(str(server_threads))
        srv = subprocess.Popen(argv, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, start_new_session=True)
        try:
            time.sleep(1.0 if kind in (LOCK2PL, FASST, LOG) else 40.0)        # table population of the KV servers
            if srv.poll() is not None:
                raise RuntimeError(f"reference UDP server exited {srv.returncode} before serving")
            r = subprocess.run([os.path.join(REF_DIR, "udp_blast"), tp, str(MSG_SIZE[kind]), str(port), str(client_threads),
                                str(window), str(seconds)], capture_output=True, timeout=seconds + 60)
            if r.returncode != 0:
                raise RuntimeError(f"udp_blast exited {r.returncode}: {r.stderr[-300:]!r}")
            out = json.loads(r.stdout.decode().strip().splitlines()[-1])
        finally:
            try:
                os.killpg(srv.pid, signal.SIGKILL)       # exactly the process group we started
            except ProcessLookupError:
                pass
            srv.wait()
    out.update(server_threads=server_threads, port=port)
    return out
