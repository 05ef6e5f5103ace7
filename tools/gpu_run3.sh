#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
RFB_BWD_MODE=direct python tools/quick_bench.py --points 1000000 --out gpurun_out/quick_direct.json > gpurun_out/quick_direct.log 2>&1
RFB_BWD_MODE=cached python tools/quick_bench.py --points 1000000 --out gpurun_out/quick_cached.json > gpurun_out/quick_cached.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/quick_direct.json; cat gpurun_out/quick_cached.json
