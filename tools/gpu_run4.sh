#!/bin/bash
mkdir -p gpurun_out
python tools/debug_bwd.py > gpurun_out/debug_random.log 2>&1
python tools/debug_bwd.py inside > gpurun_out/debug_inside.log 2>&1
RFB_BWD_MODE=cached timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config1_matches_reference_kernels and 3" > gpurun_out/racecheck.log 2>&1
RFB_BWD_MODE=cached timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config1_matches_reference_kernels and 3" > gpurun_out/memcheck.log 2>&1
cat gpurun_out/debug_random.log; cat gpurun_out/debug_inside.log; tail -15 gpurun_out/racecheck.log; tail -8 gpurun_out/memcheck.log
