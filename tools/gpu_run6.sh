#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err; cat gpurun_out/bench_ref.json | cut -c1-400
