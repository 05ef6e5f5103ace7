#!/bin/bash
# Round 2, call 10 (1 GPU): the complete GPU suite, both bench arms, the final ncu captures (full set on the two ray
# kernels + launch list) and the side measurements (random training batch incl. starting_points, trace_benchmark FPS).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 2>&1 | tail -20 > gpurun_out/r2_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err
timeout 400 python tests/tools/random_batch_bench.py > gpurun_out/r2_final_random_batch.log 2>&1
timeout 400 python tests/tools/fps_bench.py > gpurun_out/r2_final_fps.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"forward_record_kernel|backward_cached_kernel" -s 8 -c 2 \
    -o gpurun_out/r2_prof_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 90 --csv --log-file gpurun_out/r2_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_ncu_launch.log 2>&1
tail -14 gpurun_out/r2_final_pytest_gpu.log; tail -2 gpurun_out/r2_final_smoke.log
python - <<P
import json
for n in ("r2c10_bench", "r2c10_bench_ref"):
    try:
        b = json.loads([l for l in open(f"gpurun_out/{n}.json") if l.startswith("{")][-1])
        print(n, round(b["value"], 2), "e2e", round(b["e2e"]["value"], 2), b["e2e"].get("mode", "")[:10], b.get("phases_ms"), b.get("gpu_launches"))
    except Exception as e:
        print(n, "ERR", e)
P
tail -3 gpurun_out/r2_final_fps.log | cut -c1-300
grep -E "starting_points_ms|ours_fwd_ms|ours_bwd_ms|ref_fwd_ms|ref_bwd_ms|speedup|start_cells" gpurun_out/r2_final_random_batch.log
