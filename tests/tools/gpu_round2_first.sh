#!/bin/bash
# First GPU call of the next round (about 6-8 box minutes): everything that was built after round 1's GPU minutes ran
# out, in the order of what decides the next step.
#   1. the GPU suite incl. the opt-in tests of the experimental pooled-row backward (parity on hardware)
#   2. backward variants timed on the bench frame (shipped 0 vs pooled 4..8), with gradient error vs the direct kernel,
#      and the recording forward with / without the warp-voted face scan (fwd_record_ms vs fwd_record_voted_ms)
#   3. farthest_neighbor variants (3 shipped, 1/2/4 unmeasured) vs the reference's kernel
#   4. one ncu --set full capture of the pooled replay kernel, if it is within 10 % of the shipped one or faster
mkdir -p gpurun_out
RFB_TEST_EXPERIMENTS=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu.log
VARIANTS=0,4,5,6,7,8 timeout 600 python tests/tools/variant_bench.py > gpurun_out/r2_variant_bench.log 2>&1
timeout 200 python tests/tools/farthest_bench.py > gpurun_out/r2_farthest.log 2>&1
RFB_BWD_VARIANT=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"backward_pooled_kernel" \
    -s 4 -c 1 -o gpurun_out/r2_prof_pooled -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r2_ncu_pooled.log 2>&1
tail -3 gpurun_out/r2_pytest_gpu.log; grep -E "replay_v[0-9]_ms|err|fwd_record" gpurun_out/variant_bench.json | head -24
tail -1 gpurun_out/r2_farthest.log | cut -c1-600
