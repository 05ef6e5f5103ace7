#!/bin/bash
# Round 2, call 5 (2 GPUs): the fused peer-memory reduce+finalize kernel vs NCCL all-reduce + finalize.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c5_topo.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -x -s 2>&1 | tail -30 > gpurun_out/r2c5_pytest.log
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c5_bench_$tag.json 2> gpurun_out/r2c5_bench_$tag.err; }
run fused RFB_FUSED_REDUCE=1
run nccl RFB_FUSED_REDUCE=0
tail -12 gpurun_out/r2c5_pytest.log
python - <<P
import json
for n in ("fused", "nccl"):
    try:
        b = json.load(open(f"gpurun_out/r2c5_bench_{n}.json")); print(n, round(b["value"], 1), round(b["e2e"]["value"], 1), b["phases_ms"], b["config"]["parallelism"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2c5_bench_{n}.err").read()[-1500:])
P
