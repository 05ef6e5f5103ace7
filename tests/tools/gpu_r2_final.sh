#!/bin/bash
# Round 2, final 1-GPU call: the complete GPU suite, smoke, both bench arms, the final ncu captures of the two ray
# kernels (full set) + launch list, and the side measurements (random training batch incl. starting_points, FPS).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 2>&1 | tail -14 > gpurun_out/r2_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err
timeout 400 python tests/tools/random_batch_bench.py > gpurun_out/r2_final_random_batch.log 2>&1
timeout 400 python tests/tools/fps_bench.py > gpurun_out/r2_final_fps.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"forward_record_kernel<.*0>|backward_cached_kernel<.*1>" -s 4 -c 2 \
    -o gpurun_out/r2_prof_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 90 --csv --log-file gpurun_out/r2_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_ncu_launch.log 2>&1
tail -8 gpurun_out/r2_final_pytest_gpu.log; tail -2 gpurun_out/r2_final_smoke.log
python - <<P
import json, csv
for n in ("r2_final_bench", "r2_final_bench_ref"):
    try:
        b = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        print(n, round(b["value"], 2), "e2e", round(b["e2e"]["value"], 2), b["e2e"].get("mode", "")[:10], b.get("phases_ms"), b.get("gpu_launches"))
    except Exception as e:
        print(n, "ERR", e)
rows = list(csv.reader(open("gpurun_out/r2_launches_final.csv")))
hdr = [r for r in rows if "Kernel Name" in r][0]
seen = {}
for r in rows:
    if len(r) == len(hdr) and r != hdr:
        d = dict(zip(hdr, r)); seen.setdefault(d["Kernel Name"][:60], []).append(float(d["Metric Value"]) / 1e3)
for n, v in seen.items():
    if "rfb" in n: print(f"{n:62s} x{len(v):2d} median {sorted(v)[len(v)//2]:9.1f} us")
P
tail -3 gpurun_out/r2_final_fps.log | cut -c1-300
grep -E "starting_points_ms|ours_fwd_ms|ours_bwd_ms|ref_fwd_ms|ref_bwd_ms|speedup|start_cells" gpurun_out/r2_final_random_batch.log
tail -3 gpurun_out/r2_final_ncu_full.log | cut -c1-200
