#!/bin/bash
# Round 2 (8 GPUs): N = 8 and N = 4 with the shipped defaults (exchange over NVSwitch multicast from 8 ranks on, peer
# pointers below).  Earlier runs of this script also compared RFB_MULTICAST=0/1 at N = 8 and ran BASELINE config 5
# (--points 4194304 --width 3840 --height 2160): profiles/r02_bench_n8_*.json.
mkdir -p gpurun_out
run() { tag=$1; n=$2; shift 2; env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --no-cpu-baseline --steps 12 --warmup 4 2> gpurun_out/r2_n8_bench_$tag.err | grep '^{' > gpurun_out/r2_n8_bench_$tag.json; }
run n8 8 RFB_NOP=1
run n4 4 RFB_NOP=1
python - <<P
import json
for n in ("n8", "n4"):
    try:
        b = json.load(open(f"gpurun_out/r2_n8_bench_{n}.json")); p = b["phases_ms"]
        print(n, round(b["value"], 1), "ms", round(b["ms_per_step"], 3), "e2e", round(b["e2e"]["value"], 1), b["e2e"].get("mode", "")[:10])
        print("   ", {k: v for k, v in p.items() if k != "per_rank"})
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2_n8_bench_{n}.err").read()[-1500:])
P
