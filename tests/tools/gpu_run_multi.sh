#!/bin/bash
# multi-GPU call: N>1 parity test, then bench at 1 and N GPUs (both arms at 1)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_multi.txt
python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_multigpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  fi
done
cat gpurun_out/pytest_multigpu.log; for f in gpurun_out/bench_n*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','kernels_ms')}, d['e2e']['value'])
"; done; tail -3 gpurun_out/bench_n2.err
