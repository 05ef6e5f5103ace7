#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"forward_record_kernel|backward_cached_kernel" -s 6 -c 2 \
    -o gpurun_out/prof_ours3 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ours3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_ours3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch3.log 2>&1
tail -3 gpurun_out/ncu_ours3.log | cut -c1-300
