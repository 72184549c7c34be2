#!/bin/bash
# One B200: the whole GPU test-suite, then the per-kernel A/B probe.  usage: tools/session.sh <tag>
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-s}; mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) | tee $O/pytest.txt
cd tools
( timeout 300 python ab.py --store --hot 2>&1 | grep -v Warning ) | tee ../$O/ab.txt
