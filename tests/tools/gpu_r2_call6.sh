#!/bin/bash
# Round 2, call 6 (8 GPUs): the N = 8 bench with the fused peer-memory exchange kernel, and with NCCL for comparison.
mkdir -p gpurun_out
run() { tag=$1; n=$2; shift 2; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 12 --warmup 4 --no-cpu-baseline 2> gpurun_out/r2c6_bench_$tag.err | grep '^{' > gpurun_out/r2c6_bench_$tag.json; }
run n8_fused 8 RFB_FUSED_REDUCE=1
run n8_nccl 8 RFB_FUSED_REDUCE=0
run n4_fused 4 RFB_FUSED_REDUCE=1
python - <<P
import json
for n in ("n8_fused", "n8_nccl", "n4_fused"):
    try:
        b = json.load(open(f"gpurun_out/r2c6_bench_{n}.json")); print(n, round(b["value"], 1), "ms", round(b["ms_per_step"], 3), "e2e", round(b["e2e"]["value"], 1), b["phases_ms"])
        print("   ", b["e2e"]["diag"]["step_wall_ms"], b["e2e"]["diag"]["phase_ms_median"])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2c6_bench_{n}.err").read()[-1500:])
P
