#!/bin/bash
# TWO B200s: the step after the one-wave dispatch and the replay-only flush; A/B of K1 / K2 occupancy inside the step.
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-n2}; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_cluster.py -m gpu -x -q 2>&1 | tail -4 ) | tee $O/pytest.txt
cd tools
for sc in 0 12 8; do
  ( echo "#### DINT_STEP_CTAS=$sc"; DINT_STEP_CTAS=$sc SANITY_MODES=p2p timeout 300 $T --master-port 296$sc p2p_sanity.py 2>&1 | grep "p2p:" ) | tee -a ../$O/p2p_ab.txt
done
( DINT_SHARD_TRACE=1 SANITY_MODES=p2p timeout 300 $T --master-port 29612 p2p_sanity.py 2>&1 | grep "dint_shard\|p2p:" ) | tee ../$O/p2p_trace.txt
# where does one dispatch tile spend its time?  (one rank, local slabs)
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_route_dispatch' -s 20 -c 1 -f -o ../$O/prof_dispatch python ab.py --route > /dev/null 2>&1
ls -la ../$O/
