#!/bin/bash
# Round 2, call 11 (1 GPU): the tests added since call 10 + latency-chain variants of the recording forward.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2c11_pytest.log
for v in default pre_row pre_nbr pre_both; do
  lib=tests/tools/_variants/libradfoam_b200_$v.so; [ $v = default ] && lib=default
  timeout 300 python tests/tools/kblock_bench.py $lib $v > gpurun_out/r2c11_variant_$v.log 2>&1
done
tail -4 gpurun_out/r2c11_pytest.log
cat gpurun_out/r2c11_variant_*.log | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], {k: round(v, 3) for k, v in d.items() if k.endswith('_ms')}, d['full_checksum'])
"
