#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/pytest_gpu.log; python -c "
import json
d=json.loads(open('gpurun_out/bench_ours.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernels_ms','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'])
"
