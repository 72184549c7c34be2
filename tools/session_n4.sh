#!/bin/bash
# FOUR B200s: the cluster with one shard per device (tatp / smallbank placement needs >= 3), the N = 4 bench line incl. TATP / SmallBank.
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-n4}; mkdir -p $O
N=${2:-4}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 900 python -m pytest tests/test_gpu_cluster.py -m gpu -x -q 2>&1 | tail -4 ) | tee $O/pytest.txt
( DINT_BENCH_SECONDS=${3:-0.5} timeout 2400 $T --master-port 29640 bench.py --gpus $N --steps ${4:-8} --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"; tail -c 1500 $O/bench_n$N.err; head -c 3000 $O/bench_n$N.json ) 2>&1 | tee $O/bench_tail.txt
