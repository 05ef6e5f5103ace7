"""Dissect one tie-heavy case (tests/fuzz_cases.py seed) on the GPU: gradients of our path (re-walk and tape replay),
of the reference's kernels (twice: their own run-to-run noise) and of the CPU oracle, pairwise, with the worst rows."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import fuzz_cases  # noqa: E402
import test_gpu_parity as parity  # noqa: E402

for seed in [int(a) for a in sys.argv[1:]]:
    scene_kind, ray_kind, f, rays, start, dq, kw = fuzz_cases.make_case(seed)
    case = common.Case(f, rays, start, dq, seed=seed)
    full = dict(weight_threshold=0.001, max_intersections=1024)
    full.update(kw)
    runs = {
        "ours_rewalk": parity.run_ours(torch, case, tape=False, **kw),
        "ours_tape": parity.run_ours(torch, case, tape=True, **kw),
        "ref_a": parity.run_ref_gpu(torch, case, **full),
        "ref_b": parity.run_ref_gpu(torch, case, **full),
        "oracle": parity.run_cpu_oracle(case, **full),
    }
    print("seed", seed, scene_kind, ray_kind, "points", f.points.shape[0], "rays", rays.shape[0], kw, "deg", f.sh_degree)
    names = list(runs)
    for k in ("points_grad", "attr_grad"):
        for i, a in enumerate(names):
            for b in names[i + 1:]:
                print("  %-11s %-11s vs %-11s %.3e" % (k, a, b, common.grad_error(runs[a][k], runs[b][k])))
    d = np.abs(runs["ours_rewalk"]["points_grad"].astype(np.float64) - runs["ref_a"]["points_grad"]).max(axis=1)
    for row in np.argsort(-d)[:4]:
        print("  row", int(row), "degree", int(f.offsets[row + 1] - f.offsets[row]),
              *["%s %s" % (n, runs[n]["points_grad"][row]) for n in names])
    hits = [(int(i), int(n)) for i, n in enumerate(runs["oracle"]["num_intersections"].reshape(-1))]
    print("  steps per ray: max", max(n for _, n in hits))
