#!/bin/bash
# One B200: the whole GPU test-suite, the per-kernel A/B probe (engine, dispatch / combine), optionally the bench.
# usage: tools/session.sh <tag> [bench]
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-s}; mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) | tee $O/pytest.txt
cd tools
( timeout 400 python ab.py --store --hot --route 2>&1 | grep -v Warning ) | tee ../$O/ab.txt
( [ -x ./ubench2 ] && timeout 120 ./ubench2 ) 2>&1 | tee ../$O/ubench2.txt
cd ..
if [ "$2" = bench ]; then
  ( timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 1500 $O/bench.err; head -c 6000 $O/bench.json ) 2>&1 | tee $O/bench_tail.txt
fi
