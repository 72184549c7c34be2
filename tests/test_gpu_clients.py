"""-m gpu: the lock_fasst closed-loop clients resident on the GPU (dint_b200/csrc/clients.cuh, dint_clients_*) must take,
round for round, the decisions of the host-side restatement of the reference's clients (workloads.cc, restating
lock_fasst/caladan/client.cc:183-280) driving the oracle server: same requests on the wire every round, same counters."""
import numpy as np
import pytest

import oracle_lib as O
from dint_b200 import Engine, GpuClients, wire
from dint_b200.workloads import Workload
from golden_util import first_diff

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fam", [dict(n_keys=50000, zipf_theta=0.0), dict(n_keys=4800, zipf_theta=0.8), dict(n_keys=7, zipf_theta=0.0)],
                         ids=["uniform", "zipf_hot", "tiny"])
def test_gpu_clients_reproduce_the_host_clients_round_for_round(fam):
    n, rounds = 3000, 120
    ora = O.Oracle(wire.FASST)
    wl = Workload(wire.FASST, n_clients=n, seed=77, **fam)
    with Engine(wire.FASST, chunk=2048) as eng:
        gc = GpuClients(eng, n, seed=77, **fam)
        for r in range(rounds):
            want_req = wl.next()
            got_req, _ = gc.peek()
            assert first_diff(got_req, want_req, 9) is None, f"round {r}: requests differ: {first_diff(got_req, want_req, 9)}"
            want_resp = ora.process(want_req)
            wl.feed(want_resp)
            gc.run(1)
            _, got_resp = gc.peek()
            assert first_diff(got_resp, want_resp, 9) is None, f"round {r}: replies differ"
        a, b = gc.stats(), wl.stats()
        assert a["rounds"] == rounds and a["requests"] == b["requests"]
        for k in ("committed", "validation_aborts", "lock_rejects"):
            assert a[k] == b[k], (k, a, b)
        assert a["committed"] > 0
        gc.close()


def test_gpu_clients_many_rounds_in_one_call():
    """run(k) = k rounds back to back on the stream; the counters equal k single rounds."""
    n = 20000
    with Engine(wire.FASST) as e1, Engine(wire.FASST) as e2:
        a, b = GpuClients(e1, n, seed=5), GpuClients(e2, n, seed=5)
        a.run(60)
        for _ in range(60):
            b.run(1)
        assert a.stats() == b.stats() and a.stats()["committed"] > 0
        ra, rb = a.peek(), b.peek()
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
        a.close(); b.close()
