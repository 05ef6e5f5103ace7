import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
import test_gpu_parity as t
case = common.scene_case(num_points=60000, width=320, height=200, inside=True)
kw = dict(weight_threshold=0.0, max_intersections=1024)
cpu = t.run_cpu_oracle(case, **kw)
ref = t.run_ref_gpu(torch, case, **kw)
for tape, rep in ((False,1),(True,1),(True,3)):
    got = t.run_ours(torch, case, tape=tape, repeat=rep, **kw)
    for name, other in (("ref", ref), ("cpu", cpu)):
        print("tape", tape, "rep", rep, "vs", name,
              "nint mism", int((got["num_intersections"] != other["num_intersections"]).sum()),
              "didx mism", int((got["depth_indices"] != other["depth_indices"]).sum()),
              "rgba max", float(np.abs(got["rgba"] - other["rgba"]).max()),
              "pg", common.grad_error(got["points_grad"], other["points_grad"]),
              "ag", common.grad_error(got["attr_grad"], other["attr_grad"]))
bad = np.argwhere(got["depth_indices"] != cpu["depth_indices"])
print(bad[:5], got["depth_indices"][tuple(bad[0])], cpu["depth_indices"][tuple(bad[0])], got["depth"][tuple(bad[0])], cpu["depth"][tuple(bad[0])], case.quantiles[tuple(bad[0])] ) if len(bad) else None
print("mean n", cpu["num_intersections"].mean(), cpu["num_intersections"].max())
