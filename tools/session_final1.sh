#!/bin/bash
# One B200: the whole GPU test-suite, smoke, then the round-end capture (bench line + ncu evidence).
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-f1}; mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) | tee $O/pytest.txt
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) | tee $O/smoke.txt
bash tools/round_capture.sh r02 2>&1 | tail -40 | tee $O/capture.txt
