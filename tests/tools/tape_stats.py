"""Gradient-routing statistics from a recorded walk tape, computed on the CPU through the kernel-logic emulator
(tests/emu): how many 16-byte reductions each accumulation strategy of the backward would issue on a scene whose
cell-to-pixel ratio matches the bench workload (rays / points = 2).  Counts, not times -- used to rank ideas before
spending GPU minutes on them.

    python tests/tools/tape_stats.py [points=50000] [width=416] [height=240]
"""
import json
import os
import sys
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import common  # noqa: E402
import emu  # noqa: E402

NONE = 0xFFFFFFFF
ROW_REDS = 13      # 52 floats = 13 x RED.128


def slot_of(cell, slots):
    return ((int(cell) * 2654435761) & 0xFFFFFFFF) >> (32 - int(np.log2(slots)))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    width = int(sys.argv[2]) if len(sys.argv) > 2 else 416
    height = int(sys.argv[3]) if len(sys.argv) > 3 else 240
    case = common.scene_case(num_points=n, width=width, height=height, q=2)
    f = case.foam
    pipe = emu.EmuPipeline(3)
    for _ in range(2):  # the second recording has a large enough pool
        out = pipe.trace_forward(f.points, f.attributes, f.adjacency, f.offsets, case.rays, case.start,
                                 case.quantiles, scene_version=1, record_tape=True)
    cells, t1, count = emu.tape_records(pipe, height, width)
    W, S, _ = cells.shape
    valid = cells != NONE
    t0 = np.maximum.accumulate(np.concatenate([np.zeros((W, 1, 32), np.float32), t1[:, :-1]], axis=1), axis=1)
    comp = valid & (t1 > t0)                                   # composited steps (pipeline.cu: if (t1 > t0))
    lane_steps = int(comp.sum())
    res = {"points": int(f.num_points), "rays": width * height, "mean_steps": float(count.sum() / (width * height)),
           "composited_lane_steps": lane_steps}

    # ---- shipped scheme: groups >= 6 lanes through a 4-slot direct-mapped row cache, others direct; position
    # gradients as one extra reduction per composited lane-step after the first
    group_hist = Counter()
    singles = evictions = staged = 0
    for w in range(W):
        tags = {}
        for k in range(S):
            m = comp[w, k]
            if not m.any():
                continue
            cs, cnt = np.unique(cells[w, k][m], return_counts=True)
            for c, g in zip(cs, cnt):
                group_hist[int(g)] += int(g)
                if g < 6:
                    singles += int(g)
                else:
                    staged += int(g)
                    s = slot_of(c, 4)
                    if tags.get(s, NONE) != c:
                        if s in tags:
                            evictions += 1
                        tags[s] = c
        evictions += len(tags)
    first = comp & (np.cumsum(comp, axis=1) == 1)
    flushes = lane_steps - int(first.sum())
    res["shipped"] = {"direct_lane_steps": singles, "staged_lane_steps": staged, "cache_evictions": evictions,
                      "flush_reds": flushes, "red128": ROW_REDS * (singles + evictions) + flushes}
    res["group_size_share"] = {str(g): round(v / lane_steps, 4) for g, v in sorted(group_hist.items())}

    # ---- every group (also single lanes) through a direct-mapped row cache of `slots` rows per warp, position
    # gradients inside the row: reductions = 13 per eviction
    for slots in (4, 8, 16, 32, 64):
        for min_group in (1, 2):
            ev = direct = rounds = 0
            for w in range(W):
                tags = {}
                for k in range(S):
                    m = comp[w, k]
                    if not m.any():
                        continue
                    cs, cnt = np.unique(cells[w, k][m], return_counts=True)
                    for c, g in zip(cs, cnt):
                        if g < min_group:
                            direct += int(g)
                            continue
                        rounds += 1
                        sl = slot_of(c, slots)
                        if tags.get(sl, NONE) != c:
                            if sl in tags:
                                ev += 1
                            tags[sl] = c
                ev += len(tags)
            total = ROW_REDS * (ev + direct)
            res[f"all_cached_{slots}_min{min_group}"] = {"evictions": ev, "direct_lane_steps": direct,
                                                         "group_rounds": rounds, "red128": total,
                                                         "vs_shipped": round(total / res["shipped"]["red128"], 3)}
    res["shipped"]["group_rounds"] = int(sum(v / int(g) for g, v in group_hist.items() if int(g) >= 6))
    res["groups_per_lane_step"] = {"all": sum(v / int(g) for g, v in group_hist.items()) / lane_steps}

    # ---- chunk-sorted accumulation: one row of reductions per distinct cell of a warp's chunk of S_c steps
    # (position gradients carried in the same row: elements 49..51)
    for sc in (4, 8, 16, 32, 64):
        distinct = 0
        for w in range(W):
            for k0 in range(0, S, sc):
                m = comp[w, k0:k0 + sc]
                if m.any():
                    distinct += len(np.unique(cells[w, k0:k0 + sc][m]))
        res[f"chunk_{sc}"] = {"distinct_rows": distinct, "red128": ROW_REDS * distinct,
                              "vs_shipped": round(ROW_REDS * distinct / res["shipped"]["red128"], 3)}
    whole = sum(len(np.unique(cells[w][comp[w]])) for w in range(W) if comp[w].any())
    res["whole_warp"] = {"distinct_rows": whole, "red128": ROW_REDS * whole,
                         "vs_shipped": round(ROW_REDS * whole / res["shipped"]["red128"], 3)}
    # CTA level (4 warps = a 16x8 pixel block)
    cta = 0
    for b in range(W // 4):
        sel = [cells[4 * b + i][comp[4 * b + i]] for i in range(4)]
        sel = np.concatenate(sel) if sel else np.zeros(0, np.uint32)
        cta += len(np.unique(sel))
    res["whole_cta"] = {"distinct_rows": cta, "red128": ROW_REDS * cta,
                        "vs_shipped": round(ROW_REDS * cta / res["shipped"]["red128"], 3)}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
