#!/bin/bash
# Round 2, call 4 (1 GPU): the twin-kernel forward and the longest-first replay schedule, measured.
mkdir -p gpurun_out
timeout 300 python tests/tools/kblock_bench.py default lpt > gpurun_out/r2c4_lpt.log 2>&1
RFB_REPLAY_ORDER=0 timeout 300 python tests/tools/kblock_bench.py default nolpt > gpurun_out/r2c4_nolpt.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --emulate-shard 8 > gpurun_out/r2c4_bench_shard8.json 2> gpurun_out/r2c4_bench_shard8.err
timeout 900 python -m pytest tests/test_gpu_reference_op.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2c4_pytest.log
grep -h '^{' gpurun_out/r2c4_lpt.log gpurun_out/r2c4_nolpt.log | cut -c1-420
python - <<P
import json
for n in ("r2c4_bench", "r2c4_bench_shard8"):
    b = json.load(open(f"gpurun_out/{n}.json")); print(n, round(b["value"], 1), round(b["e2e"]["value"], 1), b["phases_ms"])
P
tail -4 gpurun_out/r2c4_pytest.log
