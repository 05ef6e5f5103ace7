#!/bin/bash
# Round 2, first GPU call: e2e stall diagnosis (4 bench processes, the first one builds the foam), then everything
# round 1 built but never measured: opt-in GPU tests, backward / forward variants, farthest_neighbor variants.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err
RFB_BENCH_GC_ON=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_b3.json 2> gpurun_out/r2_b3.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_b4.json 2> gpurun_out/r2_b4.err
RFB_TEST_EXPERIMENTS=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu.log
VARIANTS=0,4,5,6,7,8 timeout 600 python tests/tools/variant_bench.py > gpurun_out/r2_variant_bench.log 2>&1
timeout 200 python tests/tools/farthest_bench.py > gpurun_out/r2_farthest.log 2>&1
tail -3 gpurun_out/r2_pytest_gpu.log; grep -E "replay_v[0-9]_ms|fwd_record" gpurun_out/variant_bench.json | head -24
tail -1 gpurun_out/r2_farthest.log | cut -c1-600
