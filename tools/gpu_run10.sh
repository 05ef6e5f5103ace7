#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 1500 python tools/configs_bench.py 2 3 5 > gpurun_out/configs.log 2>&1
cat gpurun_out/smoke.log | tail -3; cat gpurun_out/configs.log | tail -8
