#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python tools/debug_bwd.py > gpurun_out/debug_random.log 2>&1
RFB_BWD_MODE=direct python tools/quick_bench.py --points 1000000 --out gpurun_out/quick_direct.json > gpurun_out/quick_direct.log 2>&1
RFB_BWD_MODE=cached python tools/quick_bench.py --points 1000000 --out gpurun_out/quick_cached.json > gpurun_out/quick_cached.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"forward_kernel|backward_cached_kernel" -s 2 -c 2 \
    -o gpurun_out/prof_ours2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ours2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^(forward|backward)$" -s 2 -c 2 \
    -o gpurun_out/prof_ref -f python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/ncu_ref.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; grep "max|d|" gpurun_out/debug_random.log; grep -E "ours_fwd_ms\"|ours_bwd_ms\"|_err|speedup" gpurun_out/quick_direct.json gpurun_out/quick_cached.json
