"""Global reductions and warp collectives each backward variant executes, counted by RUNNING the kernels on the CPU
emulator (tests/emu) on a scene with the bench workload's rays/points ratio.  Counts, not times.

    python tests/tools/emu_reduction_counts.py [points=50000] [width=416] [height=240] > profiles/...json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import common  # noqa: E402
import emu  # noqa: E402

NAMES = {0: "shipped: 4-row cache, groups >= 6 lanes", 1: "8-row cache, groups >= 8", 2: "4 rows, groups >= 5",
         3: "2 rows, groups >= 6", 6: "pooled rows, 8-row cache", 4: "pooled rows, 16-row cache",
         5: "pooled rows, 32-row cache", 7: "pooled rows, 16-row cache, lone lanes direct",
         8: "pooled rows, 8-row cache, lone lanes direct"}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    width = int(sys.argv[2]) if len(sys.argv) > 2 else 416
    height = int(sys.argv[3]) if len(sys.argv) > 3 else 240
    case = common.scene_case(num_points=n, width=width, height=height, q=2)
    f = case.foam
    scene = (f.points, f.attributes, f.adjacency, f.offsets)
    res = {"points": int(f.num_points), "rays": width * height, "variants": {}}
    base = None
    for variant in (0, 1, 2, 3, 6, 4, 5, 8, 7):
        os.environ["RFB_BWD_VARIANT"] = str(variant)
        pipe = emu.EmuPipeline(3)
        for _ in range(2):
            fwd = pipe.trace_forward(*scene, case.rays, case.start, case.quantiles, scene_version=2, record_tape=True)
        emu.red_counters()
        bwd = pipe.trace_backward(*(None,) * 6, fwd["rgba"], case.grad_rgba, None, fwd["depth_indices"],
                                  case.grad_depth, scene_version=2, use_tape=True)
        c = emu.red_counters()
        base = base or (c, bwd)
        c["name"] = NAMES[variant]
        c["reduction_bytes_vs_shipped"] = round(c["bytes"] / base[0]["bytes"], 3)
        c["collectives_vs_shipped"] = round(c["warp_collectives"] / base[0]["warp_collectives"], 2)
        c["max_grad_diff_vs_shipped"] = max(common.grad_error(bwd[k], base[1][k]) for k in ("points_grad", "attr_grad"))
        res["variants"][str(variant)] = c
    res["mean_steps"] = float(fwd["num_intersections"].mean())
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
