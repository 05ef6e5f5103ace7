#!/bin/bash
# Round 2, call 7 (2 GPUs): what limits the fused exchange kernel (RFB_PEER_DEBUG: 1 = own-rank stores only, 2 = own-rank
# loads only), with the exchange split into barrier / kernel / barrier / copy-out.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2c7_pytest.log
run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 4 --no-cpu-baseline 2> gpurun_out/r2c7_bench_$tag.err | grep '^{' > gpurun_out/r2c7_bench_$tag.json; }
run normal RFB_PEER_DEBUG=0
run own_stores RFB_PEER_DEBUG=1 RFB_BENCH_E2E_GRAPH=0
run own_loads RFB_PEER_DEBUG=2 RFB_BENCH_E2E_GRAPH=0
cat gpurun_out/r2c7_pytest.log
python - <<P
import json
for n in ("normal", "own_stores", "own_loads"):
    try:
        b = json.load(open(f"gpurun_out/r2c7_bench_{n}.json")); p = b["phases_ms"]
        print(n, round(b["value"], 1), "e2e", round(b["e2e"]["value"], 1), b["e2e"].get("mode", "")[:10], p.get("grad_reduce_and_finalize"), p.get("grad_reduce_and_finalize_parts"))
        print("    per rank", p.get("per_rank"))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2c7_bench_{n}.err").read()[-1200:])
P
