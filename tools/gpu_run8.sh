#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python tools/fps_bench.py > gpurun_out/fps.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; python -c "
import json
d=json.loads(open('gpurun_out/bench_ours.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernels_ms','clocks','gpu_launches')}, d['e2e'])
"; tail -20 gpurun_out/fps.log
