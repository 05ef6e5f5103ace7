#!/bin/bash
# Round 2, second GPU call (1 GPU): full GPU suite incl. the config-size parity tests, five consecutive bench
# processes (e2e reproducibility), CTA-size variants, ncu captures (full frame + the N = 8 shard size).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -25 > gpurun_out/r2_pytest_gpu.log
for i in 1 2 3 4 5; do
  extra="--no-cpu-baseline"; [ $i = 1 ] && extra=""
  timeout 400 python bench.py --steps 20 --warmup 5 $extra > gpurun_out/r2_bench_$i.json 2> gpurun_out/r2_bench_$i.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --emulate-shard 8 > gpurun_out/r2_bench_shard8.json 2> gpurun_out/r2_bench_shard8.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"forward_record_kernel|backward_cached_kernel" -s 6 -c 2 \
    -o gpurun_out/r2_prof_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"forward_record_kernel|backward_cached_kernel" -s 6 -c 2 \
    -o gpurun_out/r2_prof_shard8 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --emulate-shard 8 > gpurun_out/r2_ncu_shard8.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1
tail -12 gpurun_out/r2_pytest_gpu.log
for i in 1 2 3 4 5; do python - <<P
import json
try:
    b=json.load(open("gpurun_out/r2_bench_$i.json")); d=b["e2e"]["diag"]
    print($i, round(b["value"],1), round(b["e2e"]["value"],1), d["step_wall_ms"], d["torch_device_allocs"], b["phases_ms"])
except Exception as e: print($i, "ERR", e)
P
done
