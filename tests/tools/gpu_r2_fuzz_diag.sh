#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/tools/gpu_fuzz_bisect.py 3736 > gpurun_out/r2_fuzz_bisect.log 2>&1
tail -n 60 gpurun_out/r2_fuzz_bisect.log
