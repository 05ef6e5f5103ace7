"""Kernel times on the bench frame: forward (plain / recording), backward (direct, re-walk with the warp cache,
tape replay), with the replay's gradient error against the direct kernel."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import radfoam_b200  # noqa: E402
from quick_bench import timeit  # noqa: E402

points = int(sys.argv[1]) if len(sys.argv) > 1 else 1_048_576
f = bench.load_or_build_foam(points, print)
frame = bench.make_frame(f, 1920, 1080)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
scene = [d(x) for x in (f.points, f.attributes, f.adjacency, f.offsets)]
fr = {k: d(v) for k, v in frame.items()}
pipe = radfoam_b200.create_pipeline(3)
pipe.record_tape = False
fwd = pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"])
res = {"fwd_ms": timeit(lambda: pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"]))}


def bwd():
    return pipe.trace_backward(*scene, fr["rays"], fr["start"], fwd["rgba"], fr["grad_rgba"], fr["dq"],
                               fwd["depth_indices"], fr["grad_depth"])


os.environ["RFB_BWD_MODE"] = "direct"
base = bwd()
res["direct_ms"] = timeit(bwd)
os.environ["RFB_BWD_MODE"] = "cached"
out = bwd()
res["rewalk_v0_ms"] = timeit(bwd)
print("rewalk", res["rewalk_v0_ms"], flush=True)
# walk tape: recording forward + replaying backward
scene[0].requires_grad_(True)
pipe.record_tape = True
fwd = pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"])
res["fwd_record_ms"] = timeit(lambda: pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"]))
res["tape"] = pipe.tape_status()
out = bwd()
res["replay_err_vs_direct"] = float((out["attr_grad"] - base["attr_grad"]).abs().max() / base["attr_grad"].abs().max())
res["replay_points_err_vs_direct"] = float((out["points_grad"] - base["points_grad"]).abs().max() / base["points_grad"].abs().max())
res["replay_ms"] = timeit(bwd)
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/variant_bench.json", "w"), indent=1)
