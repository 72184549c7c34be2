"""The UDP front-end's own logic (dint_b200/csrc/udp_server.cc: recvmmsg batching, arrival order, reply addressing,
datagram-size filter, clean shutdown) without a GPU: the server binary runs against tests/stub/stub_abi.c, a stand-in
for libdint_b200.so that answers the C-ABI calls with the CPU oracle (LD_LIBRARY_PATH in front of the real library).
The engine behind the same calls is covered by the -m gpu tests."""
import os
import signal
import socket
import subprocess
import time

import numpy as np
import pytest

import oracle_lib as O
import trace_gen as T
from dint_b200 import _build, wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def stub_dir(tmp_path_factory):
    from dint_b200 import engine as E
    E.lib()                                               # builds the real library and the server binary
    O.lib()                                               # builds oracle/libdint_oracle.so
    d = tmp_path_factory.mktemp("stub")
    subprocess.run(["gcc", "-O1", "-std=gnu11", "-fPIC", "-shared", "-o", str(d / "libdint_b200.so"),
                    os.path.join(ROOT, "tests", "stub", "stub_abi.c"), os.path.join(ROOT, "oracle", "dint_oracle.c")], check=True)
    return str(d)


def _serve(stub_dir, kind_name, extra=()):
    with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as s0:
        s0.bind(("127.0.0.1", 0))
        port = s0.getsockname()[1]
    env = dict(os.environ, LD_LIBRARY_PATH=stub_dir, DINT_STUB_SMALL="300")
    srv = subprocess.Popen([_build.UDP_SERVER, kind_name, "--port", str(port), "--bind", "127.0.0.1", *extra], env=env,
                           stderr=subprocess.PIPE)
    time.sleep(0.5)
    return srv, port


def _stop(srv):
    srv.send_signal(signal.SIGTERM)
    try:
        return srv.communicate(timeout=10)[1].decode()
    except subprocess.TimeoutExpired:
        srv.kill()
        raise


@pytest.mark.parametrize("kind,make", [
    (wire.FASST, lambda: T.fasst_random(6000, 40, seed=31)),
    (wire.LOCK2PL, lambda: T.lock2pl_random(6000, 40, seed=32)),
    (wire.STORE, lambda: T.store_random(3000, 300, seed=33)),
])
def test_front_end_preserves_arrival_order_and_answers_the_sender(stub_dir, kind, make):
    req = make()
    msg = wire.MSG_SIZE[kind]
    cfg = dict(subs_populate=300) if kind == wire.STORE else {}
    want = O.Oracle(kind, **cfg).process(req).reshape(-1, msg)
    srv, port = _serve(stub_dir, wire.KIND_NAMES[kind], ("--batch", "97"))     # odd batch size: batches split mid-window
    try:
        assert srv.poll() is None, srv.stderr.read()
        rec = np.ascontiguousarray(req).view(np.uint8).reshape(-1, msg)
        got = np.zeros_like(rec)
        # two client sockets interleave in a fixed global order; every reply must come back on the socket that asked
        socks = [socket.socket(socket.AF_INET, socket.SOCK_DGRAM) for _ in range(2)]
        for s in socks:
            s.settimeout(5.0)
            s.connect(("127.0.0.1", port))
        for lo in range(0, len(rec), 48):
            hi = min(lo + 48, len(rec))
            s = socks[(lo // 48) % 2]                      # one window per socket at a time: the global order is fixed
            for i in range(lo, hi):
                s.send(rec[i].tobytes())
            if lo == 96:
                s.send(b"\x01\x02\x03")                   # a datagram of the wrong size is dropped, not served
            for i in range(lo, hi):
                got[i] = np.frombuffer(s.recv(256), dtype=np.uint8)
        assert np.array_equal(got, want)
    finally:
        log = _stop(srv)
    assert "datagrams in" in log and " 1 dropped" in log, log


def test_front_end_refuses_bad_arguments(stub_dir):
    env = dict(os.environ, LD_LIBRARY_PATH=stub_dir)
    assert subprocess.run([_build.UDP_SERVER, "lock_fasst", "--bind", "not.an.address"], env=env, capture_output=True).returncode == 2
    assert subprocess.run([_build.UDP_SERVER, "lock_fasst", "--frobnicate", "1"], env=env, capture_output=True).returncode == 2
