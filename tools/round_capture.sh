#!/bin/bash
# Round-end evidence on one B200: bench line, ncu launch list of the bench command, --set full captures.
R=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${R}_n1.json 2> gpurun_out/bench_${R}_n1.err
tail -c 600 gpurun_out/bench_${R}_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_apply|k_classify|k_ordered' -s 6 -c 3 -f -o gpurun_out/prof_fasst_${R} python tools/prof_run.py fasst > gpurun_out/prof_fasst.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_apply|k_classify' -s 2 -c 2 -f -o gpurun_out/prof_store_${R} python tools/prof_run.py store > gpurun_out/prof_store.log 2>&1
cat gpurun_out/bench_${R}_n1.json
