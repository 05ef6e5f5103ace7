#!/bin/bash
# second GPU call: updated tests, bench both arms, ncu launch list + full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_ours.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"forward_kernel|backward_kernel" -s 2 -c 2 \
    -o gpurun_out/prof_ours -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ours.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"radfoam.*(forward|backward)" -s 2 -c 2 \
    -o gpurun_out/prof_ref -f python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/ncu_ref.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_ours.json; cat gpurun_out/bench_ref.json; ls -la gpurun_out
