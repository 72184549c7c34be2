#!/bin/bash
# TWO B200s: A/B of the step's stream layout, dispatch kernel alone.   usage: gpurun --gpus 2 -- bash tools/session_n2.sh <tag>
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-n2}; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_cluster.py -m gpu -x -q 2>&1 | tail -4 ) | tee $O/pytest.txt
cd tools
( timeout 200 python ab.py --route 2>&1 | grep "route" ) | tee ../$O/ab_route.txt
for os in 0 1; do
  ( echo "#### DINT_SHARD_ONE_STREAM=$os"; DINT_SHARD_ONE_STREAM=$os SANITY_MODES=p2p timeout 300 $T --master-port 2960$os p2p_sanity.py 2>&1 | grep "p2p:" ) | tee -a ../$O/p2p_ab.txt
done
( DINT_SHARD_TRACE=1 SANITY_MODES=p2p timeout 300 $T --master-port 29612 p2p_sanity.py 2>&1 | grep "dint_shard\|p2p:" ) | tee ../$O/p2p_trace.txt
