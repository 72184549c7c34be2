#!/bin/bash
# Round-end evidence on one B200: the bench line, the ncu launch list of the bench command, --set full captures of the
# dominant kernels, single-pass steady-state DRAM traffic, SASS excerpt.  usage: tools/round_capture.sh r02
R=${1:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${R}_n1.json 2> gpurun_out/bench_${R}_n1.err
tail -c 400 gpurun_out/bench_${R}_n1.err
DINT_BENCH_SECONDS=0.005 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_apply|k_classify' -s 8 -c 2 -f -o gpurun_out/prof_fasst_${R} python tools/prof_run.py fasst > gpurun_out/prof_fasst.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_apply|k_classify' -s 4 -c 2 -f -o gpurun_out/prof_store_${R} python tools/prof_run.py store > gpurun_out/prof_store.log 2>&1
# steady-state DRAM traffic per launch: ONE metrics pass (no kernel replay), caches left alone
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum
ncu --metrics $M --cache-control none --clock-control none -k regex:'k_apply|k_classify|k_ordered' -s 9 -c 6 --csv --log-file gpurun_out/dram_fasst_${R}.csv python tools/prof_run.py fasst > /dev/null 2>&1
ncu --metrics $M --cache-control none --clock-control none -k regex:'k_apply|k_classify' -s 4 -c 4 --csv --log-file gpurun_out/dram_store_${R}.csv python tools/prof_run.py store > /dev/null 2>&1
ncu --metrics $M --cache-control none --clock-control none -k regex:'k_route' -s 12 -c 6 --csv --log-file gpurun_out/dram_route_${R}.csv python tools/ab.py --route > /dev/null 2>&1
cuobjdump -sass dint_b200/lib/libdint_b200.so | awk '/Function :/ {fn=$3} match($0, /(UBLKCP|SYNCS|MATCH|ATOMG|ATOMS|REDG|RED)[.A-Z0-9_]*/) {c[fn" "substr($0, RSTART, RLENGTH)]++} END {for (k in c) print c[k], k}' | c++filt | sort -k2 > gpurun_out/sass_${R}.txt
head -c 3000 gpurun_out/bench_${R}_n1.json
