"""Which rays of a tie-heavy case (tests/fuzz_cases.py seed) give different gradients from the reference's kernels?
Traces every ray of the case on its own (ours, direct and warp-aggregated backward, and the reference) and prints the
rays whose position gradients differ, with their inputs."""
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

seed = int(sys.argv[1])
scene_kind, ray_kind, f, rays, start, dq, kw = fuzz_cases.make_case(seed)
whole = common.Case(f, rays, start, dq, seed=seed)
full = dict(weight_threshold=0.001, max_intersections=1024)
full.update(kw)
for i in range(rays.shape[0]):
    case = common.Case(f, rays[i:i + 1], start[i:i + 1], None if dq is None else dq[i:i + 1], seed=seed)
    case.grad_rgba = whole.grad_rgba[i:i + 1]
    case.grad_depth = None if whole.grad_depth is None else whole.grad_depth[i:i + 1]
    ref = parity.run_ref_gpu(torch, case, **full)
    outs = {}
    for mode in ("cached", "direct"):
        os.environ["RFB_BWD_MODE"] = mode
        outs[mode] = parity.run_ours(torch, case, tape=False, **kw)
    os.environ.pop("RFB_BWD_MODE")
    errs = {m: common.grad_error(o["points_grad"], ref["points_grad"]) for m, o in outs.items()}
    if max(errs.values()) > 1e-5:
        got = outs["cached"]
        rows = np.nonzero(np.abs(got["points_grad"] - ref["points_grad"]).max(axis=1) > 1e-9)[0]
        print("ray", i, errs, "ray", rays[i].tolist(), "start", int(start[i]), "dq", None if dq is None else dq[i].tolist(),
              "depth_grad", None if case.grad_depth is None else case.grad_depth[0].tolist(),
              "grad_rgba", case.grad_rgba[0].tolist(),
              "steps", int(ref["num_intersections"].reshape(-1)[0]), int(got["num_intersections"].reshape(-1)[0]),
              "qidx", ref.get("depth_indices", np.zeros(0)).reshape(-1).tolist(),
              got.get("depth_indices", np.zeros(0)).reshape(-1).tolist(),
              "depth", ref.get("depth", np.zeros(0)).reshape(-1).tolist(), got.get("depth", np.zeros(0)).reshape(-1).tolist(),
              "rgba", ref["rgba"].reshape(-1).tolist(), got["rgba"].reshape(-1).tolist())
        for r in rows[:12]:
            print("   row", int(r), "ours", got["points_grad"][r].tolist(), "direct", outs["direct"]["points_grad"][r].tolist(),
                  "ref", ref["points_grad"][r].tolist())
print("done", seed)
