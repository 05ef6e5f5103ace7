#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/test_gpu_fuzz.py 1000 3000 > gpurun_out/r2_fuzz_campaign.log 2>&1
tail -n 15 gpurun_out/r2_fuzz_campaign.log
timeout 300 python -m pytest tests/test_gpu_fuzz.py -m gpu -q 2>&1 | tail -n 5 | tee gpurun_out/r2_fuzz_pytest.log
