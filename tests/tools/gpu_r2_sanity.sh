#!/bin/bash
# The round-end checks in one call: the whole GPU suite, smoke(), the default bench line, and a longer run of the
# tie-heavy differential campaign (tests/test_gpu_fuzz.py) against the reference's kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 6 > gpurun_out/r2_sanity_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanity_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/r2_sanity_bench.json 2> gpurun_out/r2_sanity_bench.err
tail -n 6 gpurun_out/r2_sanity_pytest.log; tail -n 1 gpurun_out/r2_sanity_smoke.log
python - <<P
import json
b = json.loads([l for l in open("gpurun_out/r2_sanity_bench.json") if l.startswith("{")][-1])
print(round(b["value"], 2), "e2e", round(b["e2e"]["value"], 2), b["e2e"].get("mode", "")[:10], b["steps"], b["warmup"], b["phases_ms"], b["cpu_baseline"]["value"])
P
timeout 300 python tests/test_gpu_fuzz.py 1000 4000 > gpurun_out/r2_fuzz_campaign.log 2>&1
tail -n 12 gpurun_out/r2_fuzz_campaign.log
