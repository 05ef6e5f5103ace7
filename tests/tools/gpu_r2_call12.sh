#!/bin/bash
# Round 2, call 12 (1 GPU): full GPU suite after the last changes + the rewritten streaming kernels in the launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2c12_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 90 --csv --log-file gpurun_out/r2c12_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c12_ncu_launch.log 2>&1
tail -4 gpurun_out/r2c12_pytest.log
python - <<P
import json, csv
b = json.loads([l for l in open("gpurun_out/r2c12_bench.json") if l.startswith("{")][-1])
print(round(b["value"], 2), "e2e", round(b["e2e"]["value"], 2), b["phases_ms"])
rows = list(csv.reader(open("gpurun_out/r2c12_launches.csv")))
hdr = [r for r in rows if "Kernel Name" in r][0]
seen = {}
for r in rows:
    if len(r) == len(hdr) and r != hdr:
        d = dict(zip(hdr, r)); seen.setdefault(d["Kernel Name"][:60], []).append(float(d["Metric Value"]) / 1e3)
for n, v in seen.items():
    if "rfb" in n: print(f"{n:62s} x{len(v):2d} median {sorted(v)[len(v)//2]:9.1f} us")
P
