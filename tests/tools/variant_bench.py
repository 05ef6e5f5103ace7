"""Time backward variants (RFB_BWD_MODE / RFB_BWD_VARIANT env switches) on one foam."""
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
os.environ["RFB_BWD_VARIANT"] = "0"
out = bwd()
res["rewalk_v0_ms"] = timeit(bwd)
print("rewalk", res["rewalk_v0_ms"], flush=True)
# walk tape: recording forward + replaying backward
scene[0].requires_grad_(True)
pipe.record_tape = True
fwd = pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"])
res["fwd_record_ms"] = timeit(lambda: pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"]))
# experimental recording forwards (RFB_FWD_VARIANT=1 warp-voted face scan, 2 two-pass scan): time + bit-identity
os.environ["RFB_FWD_VARIANT"] = "0"
fwd = pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"])
for fv, name in ((1, "voted"), (2, "two_pass")):
    os.environ["RFB_FWD_VARIANT"] = str(fv)
    alt = pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"])
    res[f"fwd_record_{name}_ms"] = timeit(lambda: pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"]))
    res[f"fwd_record_{name}_identical"] = bool(all(
        torch.equal(alt[k].view(torch.int32) if alt[k].dtype != torch.float16 else alt[k],
                    fwd[k].view(torch.int32) if fwd[k].dtype != torch.float16 else fwd[k])
        for k in ("rgba", "depth", "depth_indices", "num_intersections")))
os.environ["RFB_FWD_VARIANT"] = "0"
fwd = pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"])
res["tape"] = pipe.tape_status()
# 0 shipped; 1-3 neighbouring cache configurations; 4/5/6 the experimental pooled-row kernel (16/32/8 rows);
# 7/8 the same with lone lanes reducing directly (16/8 rows)
for v in (int(x) for x in os.environ.get("VARIANTS", "0,1,2,3,4,5,6,7,8").split(",")):
    os.environ["RFB_BWD_VARIANT"] = str(v)
    out = bwd()
    res[f"replay_v{v}_err_vs_direct"] = float((out["attr_grad"] - base["attr_grad"]).abs().max() / base["attr_grad"].abs().max())
    res[f"replay_v{v}_points_err_vs_direct"] = float((out["points_grad"] - base["points_grad"]).abs().max() / base["points_grad"].abs().max())
    res[f"replay_v{v}_ms"] = timeit(bwd)
    print(v, res[f"replay_v{v}_ms"], flush=True)
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/variant_bench.json", "w"), indent=1)
