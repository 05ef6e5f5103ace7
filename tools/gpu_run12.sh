#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours_b.json 2>> gpurun_out/bench_ours.err
tail -2 gpurun_out/smoke.log; for f in bench_ref bench_ours bench_ours_b; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', {k:d.get(k) for k in ('value','ms_per_step','kernels_ms','gpu_launches')}, d['e2e'])
"; done
