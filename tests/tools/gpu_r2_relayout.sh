#!/bin/bash
# Per-launch times of the relayout kernels (ncu launch list; cold-cache, comparable between builds) + the default bench.
mkdir -p gpurun_out
RFB_BENCH_E2E_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"build_|finalize_|tile_steps|tape_order" -c 40 --csv \
  --log-file gpurun_out/r2_relayout_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_relayout_ncu.log 2>&1
python - <<P
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2_relayout_launches.csv") if l.startswith('"'))]
h = rows[0]; k = h.index("Kernel Name"); v = h.index("Metric Value"); u = h.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[k][:60]].append(float(r[v].replace(",", "")) / (1000.0 if r[u] in ("ns", "nsecond") else 1.0))
for n, t in d.items():
    print(n, len(t), "launches, median us", sorted(t)[len(t) // 2], "min", min(t))
P
timeout 400 python bench.py > gpurun_out/r2_relayout_bench.json 2> gpurun_out/r2_relayout_bench.err
python - <<P
import json
b = json.loads([l for l in open("gpurun_out/r2_relayout_bench.json") if l.startswith("{")][-1])
print(round(b["value"], 2), "e2e", round(b["e2e"]["value"], 2), b["phases_ms"])
P
