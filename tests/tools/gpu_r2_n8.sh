#!/bin/bash
# Round 2, call 9 (8 GPUs): N = 8 with the exchange over peer pointers vs NVSwitch multicast, and BASELINE config 5
# (4,194,304 points, 3840x2160) on 8 GPUs -- the 872 MB gradient-accumulator exchange.
mkdir -p gpurun_out
run() { tag=$1; shift; extra=$1; shift; env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --no-cpu-baseline $extra 2> gpurun_out/r2_n8_bench_$tag.err | grep '^{' > gpurun_out/r2_n8_bench_$tag.json; }
run n8_peers "--steps 12 --warmup 4" RFB_MULTICAST=0
run n8_multicast "--steps 12 --warmup 4" RFB_MULTICAST=1
run n8_config5 "--points 4194304 --width 3840 --height 2160 --steps 6 --warmup 3" RFB_MULTICAST=0
python - <<P
import json
for n in ("n8_peers", "n8_multicast", "n8_config5"):
    try:
        b = json.load(open(f"gpurun_out/r2_n8_bench_{n}.json")); p = b["phases_ms"]
        print(n, round(b["value"], 1), "ms", round(b["ms_per_step"], 3), "e2e", round(b["e2e"]["value"], 1), b["e2e"].get("mode", "")[:10])
        print("   ", {k: v for k, v in p.items() if k != "per_rank"})
        print("    per rank", p.get("per_rank"))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2_n8_bench_{n}.err").read()[-1500:])
P
