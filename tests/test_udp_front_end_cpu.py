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


def test_front_end_answers_the_caladan_control_handshake(stub_dir):
    """lock_2pl/caladan/client_caladan.cc:248-271: `net_req{int nports}` on the well-known port -> `net_resp{int nports;
    uint16_t ports[]}`; the data then flows on the fresh ports and is served exactly like the well-known one."""
    import struct
    req = T.fasst_random(3000, 40, seed=41)
    rec = np.ascontiguousarray(req).view(np.uint8).reshape(-1, 9)
    want = O.Oracle(wire.FASST).process(req).reshape(-1, 9)
    srv, port = _serve(stub_dir, "lock_fasst", ("--batch", "64", "--sockets", "2"))
    try:
        with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as ctl:
            ctl.settimeout(5.0)
            ctl.sendto(struct.pack("<i", 3), ("127.0.0.1", port))
            resp = ctl.recv(256)
        nports, = struct.unpack_from("<i", resp)
        ports = struct.unpack_from("<3H", resp, 4)
        assert nports == 3 and len(resp) == 4 + 2 * 3 and all(p not in (0, port) for p in ports) and len(set(ports)) == 3
        socks = [socket.socket(socket.AF_INET, socket.SOCK_DGRAM) for _ in ports]
        got = np.zeros_like(rec)
        for s, p in zip(socks, ports):
            s.settimeout(5.0)
            s.connect(("127.0.0.1", p))
        for lo in range(0, len(rec), 32):                  # one window at a time, data ports in turn: a fixed global order
            s = socks[(lo // 32) % 3]
            for i in range(lo, min(lo + 32, len(rec))):
                s.send(rec[i].tobytes())
            for i in range(lo, min(lo + 32, len(rec))):
                got[i] = np.frombuffer(s.recv(256), dtype=np.uint8)
        assert np.array_equal(got, want)
    finally:
        log = _stop(srv)
    assert " 1 control requests" in log, log


def test_front_end_answers_the_utilisation_channel(stub_dir):
    """tatp/udp/server_shard.cc:241-274: any datagram on the monitor port is answered with {double ucores; double kcores}."""
    import struct
    with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as s0:
        s0.bind(("127.0.0.1", 0))
        mon = s0.getsockname()[1]
    srv, port = _serve(stub_dir, "lock_fasst", ("--mon-port", str(mon)))
    try:
        with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as c:
            c.settimeout(5.0)
            c.sendto(struct.pack("<dd", 0.0, 0.0), ("127.0.0.1", mon))
            u, k = struct.unpack("<dd", c.recv(64))
        assert 0.0 <= u < 256.0 and 0.0 <= k < 256.0
    finally:
        _stop(srv)


def test_front_end_serves_one_port_per_shard_with_gpus(stub_dir):
    """--gpus 3, smallbank: shard i listens on port + i; the port a datagram arrives on is the shard the client chose
    (smallbank/caladan/client_udp_shard.cc: primary / backups / log have their own addresses)."""
    n, accts = 3000, 300
    req = T.smallbank_random(n, accts, seed=43)
    rec = np.ascontiguousarray(req).view(np.uint8).reshape(-1, 23)
    rng = np.random.default_rng(5)
    dst = rng.integers(0, 3, size=n)
    oras = [O.Oracle(wire.SMALLBANK, accts_populate=accts) for _ in range(3)]
    want = np.zeros_like(rec)
    for i in range(n):
        want[i] = oras[dst[i]].process(rec[i].copy()).reshape(-1)
    srv, port = _serve(stub_dir, "smallbank", ("--gpus", "3", "--sockets", "1", "--batch", "50"))
    try:
        socks = []
        for g in range(3):
            s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
            s.settimeout(5.0)
            s.connect(("127.0.0.1", port + g))
            socks.append(s)
        got = np.zeros_like(rec)
        for i in range(n):                                  # strictly one at a time: the global order is the index order
            socks[dst[i]].send(rec[i].tobytes())
            got[i] = np.frombuffer(socks[dst[i]].recv(256), dtype=np.uint8)
        assert np.array_equal(got, want)
    finally:
        _stop(srv)
