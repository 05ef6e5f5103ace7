#!/bin/bash
# Round 2, call 8 (2 GPUs): the exchange kernel over NVSwitch multicast (multimem.ld_reduce / multimem.st) vs peer pointers.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -x -s 2>&1 | grep -E "passed|failed|fused path|Error|error" | tail -12 > gpurun_out/r2c8_pytest.log
run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 4 --no-cpu-baseline 2> gpurun_out/r2c8_bench_$tag.err | grep '^{' > gpurun_out/r2c8_bench_$tag.json; }
run multicast RFB_MULTICAST=1
run peers RFB_MULTICAST=0 RFB_BENCH_E2E_GRAPH=0
cat gpurun_out/r2c8_pytest.log
python - <<P
import json
for n in ("multicast", "peers"):
    try:
        b = json.load(open(f"gpurun_out/r2c8_bench_{n}.json")); p = b["phases_ms"]
        print(n, round(b["value"], 1), "e2e", round(b["e2e"]["value"], 1), b["e2e"].get("mode", "")[:10], p.get("grad_reduce_and_finalize"), p.get("grad_reduce_and_finalize_parts"), b["config"]["parallelism"][:90])
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/r2c8_bench_{n}.err").read()[-1200:])
P
