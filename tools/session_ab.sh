#!/bin/bash
# One B200: A/B of the flag-set retirement (K1 chasing words vs side-stream memset), parity under the variant.
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-ab}; mkdir -p $O
cd tools
for fm in 0 1 0 1; do
  ( echo "#### DINT_FLAG_MEMSET=$fm"; DINT_FLAG_MEMSET=$fm timeout 300 python ab.py --store --hot 2>&1 | grep -v Warning ) | tee -a ../$O/ab.txt
done
cd ..
( DINT_FLAG_MEMSET=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_cluster.py -m gpu -x -q -k "not udp" 2>&1 | tail -4 ) | tee $O/pytest_memset.txt
