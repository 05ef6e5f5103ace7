#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
python tools/fps_bench.py > gpurun_out/fps.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -18 gpurun_out/fps.log
