#!/bin/bash
# Round 2, third GPU call (1 GPU): which build of the recording forward is fastest (exact-scan dispatch mode x
# occupancy hint x CTA size), and the tests added since call 2.
mkdir -p gpurun_out
for v in exact0 exact0_mb7 exact0_fat exact1 exact1_mb7 exact2 k64 k32; do
  timeout 300 python tests/tools/kblock_bench.py tests/tools/_variants/libradfoam_b200_$v.so $v > gpurun_out/r2c3_variant_$v.log 2>&1
done
timeout 900 python -m pytest tests/test_gpu_reference_op.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2c3_pytest.log
cat gpurun_out/r2c3_variant_*.log | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], {k: round(v, 3) for k, v in d.items() if k.endswith('_ms')})
"
tail -6 gpurun_out/r2c3_pytest.log
