#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark (see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            # this repo's GPU engine
  python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU server (oracle/_ref)

Metric (BASELINE.json): committed txns/sec on lock_fasst.  Workload at N = 1: the reference's own
lock_fasst trace shape (lock_fasst/caladan/trace_init.sh: 24,000,000 lock ids, uniform, 5-10 ids per
transaction, write probability 0.2) driven closed-loop by 1,048,576 logical clients through the FaSST
protocol of lock_fasst/caladan/client.cc against a 36,000,000-slot lock table ("REF" in SURVEY.md 8(d)).
One STEP = one batch of 4 client rounds = 4,194,304 wire requests.  Every step replays a DIFFERENT
segment of one long recorded closed-loop trace (so inputs are never L2-resident from the previous
step) from a freshly reset server state, and the reply stream of every step is checked bit-for-bit
against the closed-loop recording.  BASELINE.json's literal "4800 keys, Zipf 0.8" reading ("HOT") and
the store GET path are measured too and reported under "extra".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIENTS = 1 << 20
ROUNDS_PER_STEP = 4
STEP_REQS = CLIENTS * ROUNDS_PER_STEP
# algorithmic bytes per request: wire in + wire out + state at the reference's field granularity
FASST_BYTES = {4: 22, 5: 26, 6: 22, 7: 22, 8: 30}          # by reply type
STORE_GET_BYTES = 186


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (NVML, ~1 ms period; the timed region
    is only tens of milliseconds long, too short for `nvidia-smi -lms`)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            while not self.stop_flag:
                try:
                    reasons = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    reasons = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), reasons))
                time.sleep(0.001)
        except Exception as ex:       # no NVML: report that rather than inventing numbers
            self.samples.append((None, repr(ex)))

    def stop(self):
        self.stop_flag = True
        self.join(timeout=2)
        sm = sorted(s[0] for s in self.samples if isinstance(s[0], int))
        bits = 0
        for s in self.samples:
            if isinstance(s[1], int):
                bits |= s[1]
        # NVML reason bits: 0x8 hw_slowdown, 0x40 hw_thermal_slowdown, 0x20 sw_thermal_slowdown, 0x4 sw_power_cap
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in names.items() if bits & b], "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
def record_closed_loop(eng, wl, n_steps, pinned):
    """Drive the client state machines against the engine; returns per-step request / reply arrays
    and per-step committed-transaction counts."""
    msg = eng.msg
    reqs = np.empty((n_steps, STEP_REQS * msg), dtype=np.uint8)
    resps = np.empty_like(reqs)
    committed = []
    rq, rs = pinned
    for s in range(n_steps):
        before = wl.stats()["committed"]
        for r in range(ROUNDS_PER_STEP):
            wl.next(rq.array[: CLIENTS * msg])
            eng.submit(rq.array[: CLIENTS * msg], out=rs.array[: CLIENTS * msg])
            wl.feed(rs.array[: CLIENTS * msg])
            reqs[s, r * CLIENTS * msg:(r + 1) * CLIENTS * msg] = rq.array[: CLIENTS * msg]
            resps[s, r * CLIENTS * msg:(r + 1) * CLIENTS * msg] = rs.array[: CLIENTS * msg]
        committed.append(wl.stats()["committed"] - before)
    return reqs, resps, committed


def run_fasst(args, torch, dist, rank, world, fam_name, fam, steps, warmup, do_e2e=True, do_cpu=True):
    from dint_b200 import Engine, PinnedBuffer, wire
    from dint_b200.workloads import Workload

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    n_steps = steps + warmup
    msg = 9
    pinned = (PinnedBuffer(STEP_REQS * msg), PinnedBuffer(STEP_REQS * msg))
    # ---- record the closed-loop trace against the GPU engine itself ----
    with Engine(wire.FASST, device=dev.index, chunk=args.chunk) as eng:
        wl = Workload(wire.FASST, n_clients=CLIENTS, seed=20230 + rank, **fam)
        reqs, resps, committed = record_closed_loop(eng, wl, n_steps, pinned)
        wl_stats = wl.stats()
    # ---- device-resident replay from a fresh state: the timed region ----
    out = {}
    with Engine(wire.FASST, device=dev.index, chunk=args.chunk) as eng:
        d_req = torch.from_numpy(reqs).to(dev)
        d_resp = torch.empty((STEP_REQS * msg,), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev)
        ok = True
        for s in range(warmup):
            eng.submit_tensor(d_req[s], d_resp)
        torch.cuda.synchronize(dev)
        ok &= bool((d_resp.cpu().numpy() == resps[warmup - 1]).all()) if warmup else True
        eng.reset_stats()
        eng.profile(Engine.PROF_APPLY)        # events around the dominant kernel only: all-kernel profiling costs ~20 %
        sampler = ClockSampler(dev.index)
        sampler.start()
        time.sleep(0.01)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in range(warmup, n_steps):
            eng.submit_tensor(d_req[s], d_resp)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        eng.profile(False)
        kt = eng.kernel_times()
        st = eng.stats()
        ok &= bool((d_resp.cpu().numpy() == resps[n_steps - 1]).all())
        out.update(ms=ms, kernel_times=kt, stats=st, clocks=clocks, parity_last_step=ok)
        # a second, fully profiled replay (not the timed one) for the per-kernel breakdown
        eng.reset_stats()
        eng.profile(True)
        for s in range(warmup, n_steps):
            eng.submit_tensor(d_req[s], d_resp)
        torch.cuda.synchronize(dev)
        eng.profile(False)
        out["all_kernel_times"] = eng.kernel_times()
        del d_req
    timed_committed = sum(committed[warmup:])
    timed_reqs = steps * STEP_REQS
    out.update(committed=timed_committed, requests=timed_reqs, wl_stats=wl_stats)
    # algorithmic bytes of the timed region, from the reply types
    types = np.concatenate([resps[s].reshape(-1, msg)[:, 0] for s in range(warmup, n_steps)])
    cnt = np.bincount(types, minlength=9)
    out["alg_bytes"] = int(sum(FASST_BYTES[t] * int(cnt[t]) for t in FASST_BYTES))
    out["reply_mix"] = {str(t): int(cnt[t]) for t in FASST_BYTES}
    # ---- end to end through the host-facing C ABI call (pinned host buffers, H2D + D2H inside) ----
    if do_e2e:
        with Engine(wire.FASST, device=dev.index, chunk=args.chunk) as eng:
            rq, rs = pinned
            for s in range(warmup):
                rq.array[:] = reqs[s]
                eng.submit(rq.array, out=rs.array)
            host_in = [PinnedBuffer(STEP_REQS * msg) for _ in range(min(steps, 4))]
            t_e2e, ok2 = 0.0, True
            for s in range(warmup, n_steps):
                b = host_in[(s - warmup) % len(host_in)]
                b.array[:] = reqs[s]                        # staging into pinned memory: not timed
                if dist is not None:
                    dist.barrier()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                eng.submit(b.array, out=rs.array)           # timed: H2D + kernels + D2H, returns when resp is complete
                t_e2e += time.perf_counter() - t0
                if s == n_steps - 1:
                    ok2 = bool((rs.array == resps[s]).all())
            out.update(e2e_s=t_e2e, e2e_parity=ok2)
    # ---- CPU baseline: the unmodified reference server, one handler thread, bounded sample ----
    if do_cpu and rank == 0:
        out["cpu_baseline"] = cpu_baseline(wire.FASST, reqs[0], wl_stats["committed"] / wl_stats["requests"], threads=1,
                                           target_s=12.0, check_against=resps[0])
    return out


def cpu_baseline(kind, sample_req, txn_per_req, threads, target_s, check_against=None):
    """Time oracle/_ref (the reference's own server.cc under the replay shim) on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    if not O.ref_available(kind):
        # the C restatement, timed in-process ("port")
        ora = O.Oracle(kind)
        t0 = time.perf_counter()
        ora.process(sample_req)
        dt = time.perf_counter() - t0
        n = sample_req.size // O.MSG_SIZE[kind]
        return {"value": n / dt * txn_per_req, "unit": "txn/s", "cores": 1, "kind": "port",
                "sample": f"{n} requests once through oracle/libdint_oracle.so", "req_per_s": n / dt}
    n = sample_req.size // O.MSG_SIZE[kind]
    # calibrate the repeat count on one pass
    out, st = O.run_ref(kind, sample_req, threads=1, repeat=1, want_out=check_against is not None)
    parity = None
    if check_against is not None:
        parity = bool(np.array_equal(out, check_against))
    rate1 = st["req_per_s"]
    repeat = max(1, int(target_s * rate1 * max(1, threads) * 0.7 / n))
    _, st = O.run_ref(kind, sample_req, threads=threads, repeat=repeat, want_out=False, spread=threads > 1)
    res = {"value": st["req_per_s"] * txn_per_req, "unit": "txn/s", "cores": threads, "kind": "reference",
           "sample": f"first step of the trace ({n} requests) x {repeat} passes through oracle/_ref "
                     f"`server {threads}` under the replay shim ({st['seconds']:.1f} s)",
           "req_per_s": st["req_per_s"], "gpu_replies_equal_reference": parity}
    res["udp_as_shipped"] = udp_as_shipped(O, kind, sample_req, txn_per_req)
    return res


def udp_as_shipped(O, kind, sample_req, txn_per_req, seconds=4.0):
    """the unmodified reference server with REAL sockets on loopback (`server 8`, the reference's
    thread count, exp/run_lock_fasst.sh), two syscalls per request as deployed; informational, never the value."""
    try:
        cores = os.cpu_count() or 8
        ct = max(8, min(32, cores // 4))
        r = O.run_ref_udp(kind, sample_req, server_threads=8, client_threads=ct, window=32, seconds=seconds)
        return {"req_per_s": r["req_per_s"], "txn_per_s": r["req_per_s"] * txn_per_req, "server_threads": 8,
                "client_threads": ct, "lost_datagrams": r["lost"], "seconds": r["seconds"],
                "note": "oracle/_ref server, bind address rewritten to 127.0.0.1, replies counted not compared"}
    except Exception as ex:
        return {"unavailable": repr(ex)[:200]}


def run_closed_loop_extra(args, torch, rank, kind_name, rounds=16, warm=24):
    """lock_2pl / log_server side measurement: the reference's closed-loop clients (workloads.cc) recorded against
    the GPU engine through the host path, then replayed device-resident and timed; replies must be bit-exact."""
    from dint_b200 import Engine, wire
    from dint_b200.workloads import Workload, REF
    kind = {"lock_2pl": wire.LOCK2PL, "log_server": wire.LOG}[kind_name]
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    fam = REF if kind == wire.LOCK2PL else {}
    wl = Workload(kind, n_clients=CLIENTS, seed=20230 + rank, **fam)
    if kind == wire.LOG:
        warm = 3                                          # no protocol state to warm up
    reqs, resps, committed = [], [], []
    with Engine(kind, device=dev.index, chunk=args.chunk) as eng:
        for _ in range(warm + rounds):
            before = wl.stats()["committed"]
            q = wl.next()
            a = eng.submit(q)
            wl.feed(a)
            reqs.append(q.copy()); resps.append(a.copy())
            committed.append(wl.stats()["committed"] - before)   # transactions whose last reply arrived this round
    st = wl.stats()
    with Engine(kind, device=dev.index, chunk=args.chunk) as eng:
        d_req = [torch.from_numpy(r).to(dev) for r in reqs]
        d_out = torch.empty_like(d_req[0])
        for r in range(warm):
            eng.submit_tensor(d_req[r], d_out)
        torch.cuda.synchronize(dev)
        eng.reset_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(warm, warm + rounds):
            eng.submit_tensor(d_req[r], d_out)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        ok = bool((d_out.cpu().numpy() == resps[-1]).all())
        est = eng.stats()
    n_req = rounds * CLIENTS
    res = {"workload": ("lock_2pl REF: 24,000,000 uniform lock ids, 5-10 ids/txn, p(exclusive)=0.2, closed-loop 2PL clients "
                        "(acquire in id order, release in reverse, retry after a reject)" if kind == wire.LOCK2PL else
                        "log_server: key uniform [0, 7,009,999], ver [0,127], 40 random bytes per append") +
                       f"; {CLIENTS} logical clients, {rounds} rounds timed (device-resident replay of the recorded closed loop)",
           "requests_per_s": n_req / (ms * 1e-3), "replies_bit_exact": ok, "gpu_launches": est["kernel_launches"],
           "conflicted_fraction": est["conflicted"] / max(1, est["requests"])}
    if kind == wire.LOCK2PL:
        res["txn_per_s"] = sum(committed[warm:]) / (ms * 1e-3)
        res["requests_per_txn"] = n_req / max(1, sum(committed[warm:]))
        res["lock_rejects"] = st["lock_rejects"]
    return res


def run_udp_front_end(seconds=4.0):
    """dint_udp_server (the reference's UDP server shape over the C ABI, dint_b200/csrc/udp_server.cc) with the GPU
    engine behind it, driven over loopback by the same multi-socket replayer that times the unmodified reference
    server for cpu_baseline.udp_as_shipped.  Runs in child processes with deadlines."""
    import signal
    import socket
    import tempfile
    from dint_b200 import _build, wire
    blast = os.path.join(ROOT, "oracle", "_ref", "udp_blast")
    if not (os.path.exists(_build.UDP_SERVER) and os.path.exists(blast)):
        return {"unavailable": "dint_udp_server or oracle/_ref/udp_blast not built"}
    n = 1 << 20
    rng = np.random.default_rng(20230)
    rec = np.zeros(n, dtype=wire.MSG_DTYPE[wire.FASST])
    rec["type"] = rng.choice(4, size=n, p=(0.6, 0.15, 0.05, 0.2))          # read / acquire / abort / commit mix of the REF trace
    rec["lid"] = rng.integers(0, 24_000_000, size=n)
    with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as s0:
        s0.bind(("127.0.0.1", 0))
        port = s0.getsockname()[1]
    with tempfile.TemporaryDirectory() as td:
        tp = os.path.join(td, "trace.bin")
        wire.as_bytes(rec).tofile(tp)
        srv = subprocess.Popen([_build.UDP_SERVER, "lock_fasst", "--bind", "127.0.0.1", "--port", str(port), "--sockets", "8"],
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, start_new_session=True)
        try:
            t0 = time.time()
            os.set_blocking(srv.stderr.fileno(), False)
            banner = b""
            while b"sockets, batches" not in banner:                         # printed once the engine exists and the sockets are bound
                if srv.poll() is not None or time.time() - t0 > 90:
                    return {"unavailable": "server did not come up: " + banner.decode(errors="replace")[-200:]}
                time.sleep(0.2)
                try:
                    banner += srv.stderr.read() or b""
                except (BlockingIOError, TypeError):
                    pass
            cores = os.cpu_count() or 8
            ct = max(8, min(32, cores // 4))
            r = subprocess.run([blast, tp, "9", str(port), str(ct), "64", str(seconds)], capture_output=True, timeout=seconds + 60)
            out = json.loads(r.stdout.decode().strip().splitlines()[-1])
        finally:
            try:
                os.killpg(srv.pid, signal.SIGTERM)          # exactly the process group we started
            except ProcessLookupError:
                pass
            try:
                srv.wait(timeout=20)
            except subprocess.TimeoutExpired:
                os.killpg(srv.pid, signal.SIGKILL)
                srv.wait()
    return {"req_per_s": out["req_per_s"], "lost_datagrams": out["lost"], "server_sockets": 8, "client_threads": out["client_threads"],
            "window": out["window"], "seconds": out["seconds"],
            "note": "loopback UDP, one datagram per request, recvmmsg/sendmmsg front-end + dint_submit; same replayer as "
                    "cpu_baseline.udp_as_shipped; replies counted, not compared (parity of this path: tests)"}


def run_store_get(args, torch, rank, steps, warmup):
    """The store lookup path: 100 % kRead, NURand keys over the reference's 24 M-key population."""
    from dint_b200 import Engine, wire
    from dint_b200.workloads import Workload
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    n = 1 << 22
    wl = Workload(wire.STORE, n_clients=n, seed=1 + rank)
    first = wl.next().copy()
    deferred = None
    if rank == 0:                                          # BASELINE.json configs[0]: the reference store server on the host CPU
        deferred = DeferredCpuBaseline(wire.STORE, first, 1.0, "the first step of the GET trace")
        deferred.start()
    with Engine(wire.STORE, device=dev.index, chunk=args.chunk, populate=True) as eng:
        bufs = [torch.from_numpy(first).to(dev)] + [torch.from_numpy(wl.next().copy()).to(dev) for _ in range(steps + warmup - 1)]   # open-loop
        d_out = torch.empty_like(bufs[0])
        for s in range(warmup):
            eng.submit_tensor(bufs[s], d_out)
        torch.cuda.synchronize(dev)
        eng.reset_stats()
        eng.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(warmup, warmup + steps):
            eng.submit_tensor(bufs[s], d_out)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        eng.profile(False)
        kt = eng.kernel_times()
        hits = int((d_out.view(-1, 53)[:, 0] == 3).sum().item())
    peak, how = peaks()
    l, t = kt["k_apply"]
    ach = STORE_GET_BYTES * n * steps / l / (t / l * 1e-3) / 1e9
    return {"workload": "store kRead, NURand keys, 24,000,000-key table (reference population), device-resident",
            "get_per_s": n * steps / (ms * 1e-3), "hit_fraction_last_step": hits / n,
            "roofline": {"kernel": "k_apply<store>", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "peak_source": how, "avg_launch_us": t / l * 1e3},
            "kernel_ms": {k: v[1] for k, v in kt.items()}, "_deferred_cpu": deferred}


class DeferredCpuBaseline(threading.Thread):
    """The unmodified reference shard server (oracle/_ref, replay shim, one handler thread) over a recorded request
    stream, in the background: its table population takes most of a minute of host time, the GPU work goes on."""

    def __init__(self, kind, req, txn_per_req, what):
        super().__init__(daemon=True)
        self.kind, self.req, self.txn_per_req, self.what = kind, req, txn_per_req, what
        self.result = {"unavailable": "did not finish"}

    def run(self):
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as O
            if not O.ref_available(self.kind):
                self.result = {"unavailable": "oracle/_ref not built"}
                return
            t0 = time.time()
            _, st = O.run_ref(self.kind, self.req, threads=1, repeat=1, want_out=False, timeout=240)
            self.result = {"value": st["req_per_s"] * self.txn_per_req, "unit": "txn/s", "cores": 1, "kind": "reference",
                           "req_per_s": st["req_per_s"], "sample": f"{self.what}: {st['requests']} requests in {st['seconds']:.2f} s "
                           f"through the oracle/_ref server under the replay shim (population + replay {time.time() - t0:.0f} s wall)"}
        except Exception as ex:
            self.result = {"unavailable": repr(ex)[:200]}


def run_txn(args, torch, rank, kind_name, rounds_timed=12, rounds_warm=8, clients=1 << 20):
    """Full transaction mixes driven by the reference's closed-loop client state machines
    (dint_b200/csrc/txn_workloads.cc) against THREE shard servers (primary key % 3 + 2 backups + log on all
    three, as {tatp,smallbank}/caladan/client_udp_shard.cc) -- here three engines resident on one GPU, each
    holding the reference's full population (tatp: 7,000,000 subscribers, mix 35/35/10/2/14/2/2; smallbank:
    24,000,000 accounts, 4 % hot accounts drawing 90 % of the transactions, mix 15/15/15/25/15/15)."""
    from dint_b200 import Engine, wire
    from dint_b200.txn_workloads import TxnWorkload, Cluster, partition_by_shard
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    kind = wire.TATP if kind_name == "tatp" else wire.SMALLBANK
    subscribers = 7_000_000 if kind == wire.TATP else 24_000_000
    msg, G = wire.MSG_SIZE[kind], 3
    t0 = time.time()

    def make():
        return [Engine(kind, device=dev.index, chunk=args.chunk, populate=True) for _ in range(G)]

    engs = make()
    t_pop = time.time() - t0
    wl = TxnWorkload(kind, n_clients=clients, n_shards=G, subscribers=subscribers, gid0=rank * clients)
    cl = Cluster([e.submit for e in engs], msg)
    rec, committed, nreq = [], [], []
    for r in range(rounds_warm + rounds_timed):
        before = wl.stats()["committed"]
        rq, dst = wl.next()
        rs = cl.submit(rq, dst)
        wl.feed(rs)
        order, counts, parts = partition_by_shard(rq, dst, G, msg)
        rparts = partition_by_shard(rs, dst, G, msg)[2]
        rec.append((parts, rparts))
        committed.append(wl.stats()["committed"] - before)
        nreq.append(int(dst.size))
    st = wl.stats()
    for e in engs:
        e.close()
    deferred = None
    if rank == 0:                                          # shard 0's request stream from the start of the recording
        sample = np.concatenate([np.ascontiguousarray(parts[0]).reshape(-1) for parts, _ in rec[: rounds_warm + 4]])
        deferred = DeferredCpuBaseline(kind, sample, st["committed"] / max(1, st["requests"]),
                                       f"shard 0's first {rounds_warm + 4} rounds of the recorded closed loop")
        deferred.start()
    # device-resident replay from freshly populated shards
    engs = make()
    d = [[torch.from_numpy(np.ascontiguousarray(p)).to(dev) for p in parts] for parts, _ in rec]
    outs = [[torch.empty_like(x) for x in row] for row in d]
    for r in range(rounds_warm):
        for s_ in range(G):
            if d[r][s_].numel():
                engs[s_].submit_tensor(d[r][s_], outs[r][s_])
    torch.cuda.synchronize(dev)
    for e in engs:
        e.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(rounds_warm, rounds_warm + rounds_timed):
        for s_ in range(G):
            if d[r][s_].numel():
                engs[s_].submit_tensor(d[r][s_], outs[r][s_])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    ok = all(bool((outs[r][s_].cpu().numpy() == rec[r][1][s_]).all()) for r in (rounds_warm - 1, rounds_warm + rounds_timed - 1) for s_ in range(G))
    launches = sum(e.stats()["kernel_launches"] for e in engs)
    conflicted = sum(e.stats()["conflicted"] for e in engs)
    for e in engs:
        e.close()
    tc = sum(committed[rounds_warm:])
    tr = sum(nreq[rounds_warm:])
    return {"workload": f"{kind_name} mix, {clients} closed-loop clients, 3 shard servers x {subscribers} "
                        f"{'subscribers' if kind == wire.TATP else 'accounts'} on one GPU, "
                        f"{rounds_timed} protocol rounds timed (device-resident replay of the recorded closed-loop trace)",
            "abort_rate": 1.0 - st["committed"] / max(1, st["txns"]),
            "txn_per_s": tc / (ms * 1e-3), "requests_per_s": tr / (ms * 1e-3), "requests_per_txn": st["requests"] / max(1, st["txns"]),
            "commit_rate_by_type": {k: round(v[1] / max(1, v[0]), 4) for k, v in st["by_type"].items()},
            "replies_bit_exact": ok, "gpu_launches": launches, "conflicted_fraction": conflicted / max(1, tr),
            "populate_s_per_3_shards": round(t_pop, 1), "_deferred_cpu": deferred}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dint_b200", choices=["dint_b200", "reference"])
    ap.add_argument("--chunk", type=int, default=1 << 20)
    ap.add_argument("--no-extra", action="store_true", help="skip the HOT and store GET side measurements")
    ap.add_argument("--extra-only", default=None, help=argparse.SUPPRESS)     # child-process mode for a side measurement
    args = ap.parse_args()
    if args.extra_only:
        import torch
        print(json.dumps(run_closed_loop_extra(args, torch, 0, args.extra_only)), flush=True)
        return
    args.warmup = max(args.warmup, 3) if args.impl == "dint_b200" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference":
        return main_reference(args, rank, world)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (dint_b200 has no CPU fallback)")
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist_mod.init_process_group("nccl")
        dist = dist_mod
    from dint_b200.workloads import REF, HOT

    if world > 1:
        from dint_b200 import shard
        res = shard.bench_fasst_sharded(args, torch, dist, rank, world, REF)
    else:
        res = run_fasst(args, torch, dist, rank, world, "REF", REF, args.steps, args.warmup)

    # reduce over ranks: time = max, work = sum
    ms, committed, reqs = res["ms"], res["committed"], res["requests"]
    e2e_s = res.get("e2e_s")
    if dist is not None:
        t = torch.tensor([ms, e2e_s or 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w = torch.tensor([committed, reqs, res["stats"]["kernel_launches"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        ms, e2e_s = float(t[0]), float(t[1]) or None
        committed, reqs, launches = int(w[0]), int(w[1]), int(w[2])
    else:
        launches = res["stats"]["kernel_launches"]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_how = peaks()
    kt = res["kernel_times"]
    name = "k_apply"
    nl, tot_ms = kt[name]
    avg_s = tot_ms / nl * 1e-3
    alg_per_launch = res["alg_bytes"] / nl
    achieved = alg_per_launch / avg_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(name + "<lock_fasst>")
    line = {
        "metric": "committed txns/sec (lock_fasst)", "value": committed / (ms * 1e-3), "unit": "txn/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "lock_fasst REF: 24,000,000 uniform lock ids, 5-10 ids/txn, p(write)=0.2, closed-loop "
                               "FaSST clients (read/acquire/validate/commit), 36,000,000-slot table; "
                               f"{CLIENTS} logical clients per GPU, {ROUNDS_PER_STEP} rounds = {STEP_REQS} requests per step per GPU",
                   "baseline_config": "BASELINE.json configs[1] (lock_fasst OCC validate/commit, 1 B200).  Its '4800 keys, "
                                      "Zipf-0.8, 24M-op' wording is, in the reference, 4800 trace FILES, read fraction 0.8 and 24 M "
                                      "uniform lock ids (BASELINE.md section 1 note, lock_fasst/caladan/trace_init.sh:9-27): the headline "
                                      "runs that reference shape; the literal reading (4800 ids, Zipf 0.8) is extra.lock_fasst_HOT",
                   "requests_per_step": STEP_REQS * world, "chunk": args.chunk,
                   "cache": "every step replays a different 37.7 MB trace segment (inputs larger than reuse distance; "
                            "lock/version tables 148.5 MB > 126 MB L2)",
                   "parallelism": (f"key-space sharded x{world}, exchange={res.get('exchange', 'slabs')} "
                                   "(dispatch/combine kernels over NVLink peer memory)") if world > 1 else "single GPU"},
        "requests_per_s": reqs / (ms * 1e-3),
        "replies_bit_exact_vs_closed_loop_recording": bool(res["parity_last_step"]),
        "abort_stats": {k: res["wl_stats"][k] for k in ("committed", "validation_aborts", "lock_rejects")},
        "conflicted_fraction": res["stats"]["conflicted"] / max(1, res["stats"]["requests"]),
        "clocks": res["clocks"],
        "gpu_launches": launches,
        "roofline": {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_how,
                     "avg_launch_us": avg_s * 1e6, "launches": nl, "algorithmic_bytes_per_launch": alg_per_launch,
                     "all_kernels_ms_profiled_pass": {k: round(v[1], 3) for k, v in res.get("all_kernel_times", kt).items()}},
    }
    if e2e_s:
        line["e2e"] = {"value": committed / e2e_s, "unit": "txn/s", "h2d_bytes_per_step": STEP_REQS * 9 * world,
                       "d2h_bytes_per_step": STEP_REQS * 9 * world, "requests_per_s": reqs / e2e_s,
                       "replies_bit_exact": bool(res.get("e2e_parity", False)),
                       "path": ("dint_submit(): pinned host wire structs -> H2D -> kernels -> D2H, per step" if world == 1 else
                                "pinned host wire structs -> H2D -> dint_shard_submit_many (dispatch, engine, combine) -> D2H, per step")}
    if "cpu_baseline" in res:
        line["cpu_baseline"] = res["cpu_baseline"]
    if not args.no_extra and world == 1:
        try:
            hot = run_fasst(args, torch, None, rank, world, "HOT", HOT, max(3, args.steps // 3), 3, do_e2e=False, do_cpu=False)
            line["extra"] = {"lock_fasst_HOT": {
                "workload": "BASELINE.json literal: 4800 lock ids, Zipf 0.8, same clients/protocol",
                "txn_per_s": hot["committed"] / (hot["ms"] * 1e-3), "requests_per_s": hot["requests"] / (hot["ms"] * 1e-3),
                "abort_stats": {k: hot["wl_stats"][k] for k in ("committed", "validation_aborts", "lock_rejects")},
                "conflicted_fraction": hot["stats"]["conflicted"] / max(1, hot["stats"]["requests"]),
                "replies_bit_exact": bool(hot["parity_last_step"])}}
            line["extra"]["store_get"] = run_store_get(args, torch, rank, max(3, args.steps // 2), 3)
            line["extra"]["tatp"] = run_txn(args, torch, rank, "tatp")
            line["extra"]["smallbank"] = run_txn(args, torch, rank, "smallbank")
        except Exception as ex:  # side measurements must never cost the headline line
            line.setdefault("extra", {})["error"] = repr(ex)
        for kn in ("lock_2pl", "log_server"):
            # newer side measurements run in a child process with a deadline: whatever happens there, the headline stands
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--extra-only", kn, "--chunk", str(args.chunk)],
                                   capture_output=True, timeout=240)
                rows = [l for l in r.stdout.decode().splitlines() if l.startswith("{")]
                line["extra"][kn] = json.loads(rows[-1]) if rows else {"error": f"exit {r.returncode}: {r.stderr.decode()[-300:]}"}
            except Exception as ex:
                line.setdefault("extra", {})[kn] = {"error": repr(ex)[:300]}
        try:
            torch.cuda.empty_cache()
            line["extra"]["udp_front_end"] = run_udp_front_end()
        except Exception as ex:
            line.setdefault("extra", {})["udp_front_end"] = {"error": repr(ex)[:300]}
    for v in line.get("extra", {}).values():               # background CPU baselines: collect (bounded wait)
        if isinstance(v, dict) and "_deferred_cpu" in v:
            d = v.pop("_deferred_cpu")
            if d is not None:
                d.join(timeout=120)
                v["cpu_baseline"] = d.result
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main_reference(args, rank, world):
    """--impl reference: the reference's own lock_fasst UDP server (oracle/_ref, built from
    /root/reference unmodified), all host threads, same trace shape / metric."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from dint_b200 import wire
    from dint_b200.workloads import Workload, REF
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # the same closed-loop trace shape, recorded against the CPU restatement (no GPU needed on this arm)
    clients, rounds = 1 << 16, 48          # long enough for transactions to complete (>= 11 rounds each)
    ora = O.Oracle(wire.FASST)
    wl = Workload(wire.FASST, n_clients=clients, seed=20230, **REF)
    reqs = []
    for _ in range(rounds):
        r = wl.next()
        wl.feed(ora.process(r))
        reqs.append(r.copy())
    sample = np.concatenate(reqs)
    st = wl.stats()
    txn_per_req = st["committed"] / st["requests"]
    n = sample.size // 9
    kind = "reference" if O.ref_available(wire.FASST) else "port"
    steps, times, reqs_done = args.steps, [], 0
    for s in range(args.warmup + steps):
        if kind == "reference":
            _, stt = O.run_ref(wire.FASST, sample, threads=cores, repeat=max(1, cores // 2), want_out=False, spread=True)
            dt, nn = stt["seconds"], stt["requests"]
        else:
            t0 = time.perf_counter(); O.Oracle(wire.FASST).process(sample); dt = time.perf_counter() - t0; nn = n
        if s >= args.warmup:
            times.append(dt); reqs_done += nn
    total = sum(times)
    val = reqs_done / total * txn_per_req
    line = {"impl": "reference", "metric": "committed txns/sec (lock_fasst)", "value": val, "unit": "txn/s",
            "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": total / steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "lock_fasst REF: 24,000,000 uniform lock ids, 5-10 ids/txn, p(write)=0.2, closed-loop "
                                   "FaSST clients, 36,000,000-slot table",
                       "note": "handler only: the reference server.cc under the LD_PRELOAD replay shim (no UDP syscalls), "
                               f"`server {cores}` re-pinned one thread per host core; a step = {n} requests x {max(1, cores // 2)} passes"},
            "requests_per_s": reqs_done / total,
            "cpu_baseline": {"value": val, "unit": "txn/s", "cores": cores if kind == "reference" else 1, "kind": kind,
                             "sample": f"{n}-request closed-loop trace x {max(1, cores // 2)} passes per step"},
            "e2e": {"value": val, "unit": "txn/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if kind == "reference":
        line["cpu_baseline"]["udp_as_shipped"] = udp_as_shipped(O, wire.FASST, sample, txn_per_req)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
