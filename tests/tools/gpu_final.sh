#!/bin/bash
# what the driver runs at round end, in one call: smoke, GPU tests, both bench arms
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -1 gpurun_out/smoke.log; tail -3 gpurun_out/pytest_gpu.log; for f in bench_ref bench_ours; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', {k:d.get(k) for k in ('value','ms_per_step','kernels_ms','gpu_launches','clocks')}, d['e2e'], (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))
"; done
