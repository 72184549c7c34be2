#!/bin/bash
# TWO B200s: the multi-process exchange (symmetric memory, K2 storing replies over NVLink), the one-process cluster with
# one shard per device, where the step spends its time.   usage: gpurun --gpus 2 -- bash tools/session_n2.sh <tag>
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-n2}; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_cluster.py -m gpu -x -q 2>&1 | tail -15 ) | tee $O/pytest.txt
cd tools
( timeout 300 $T --master-port 29601 step_probe.py 2>&1 | grep -v "Warning\|warn" ) | tee ../$O/step_probe.txt
( DINT_SHARD_TRACE=1 SANITY_MODES=p2p timeout 300 $T --master-port 29602 p2p_sanity.py 2>&1 | grep -v "Warning\|warn" ) | tee ../$O/p2p_trace.txt
( timeout 300 $T --master-port 29603 p2p_sanity.py 2>&1 | grep -v "Warning\|warn" ) | tee ../$O/p2p.txt
