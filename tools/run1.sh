export RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29511
cd tools
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 40 -c 260 --csv --log-file ../gpurun_out/route_launches.csv python route_prof.py > ../gpurun_out/route_prof.log 2>&1
tail -5 ../gpurun_out/route_prof.log
