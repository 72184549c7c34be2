cd tools
export SANITY_MODES=p2p
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 p2p_sanity.py 2>&1 | grep "p2p:"
DINT_SHARD_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 p2p_sanity.py 2>&1 | grep -v Warn | grep "dint_shard\|p2p:" 
DINT_SHARD_STREAMS=3 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 p2p_sanity.py 2>&1 | grep "p2p:"
cd ..; timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q -k p2p 2>&1 | tail -2
