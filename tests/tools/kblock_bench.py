"""Time the recording forward + replaying backward of a build variant of the library (e.g. another RFB_KBLOCK)
on the bench frame and on one rank's shard of an 8-way split (interleaved 8-row bands = what each GPU traces at N = 8).
usage: python tests/tools/kblock_bench.py <path/to/libvariant.so> <tag>      (test infrastructure)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
lib_path, tag = sys.argv[1], sys.argv[2]
from radfoam_b200 import _lib  # noqa: E402

if lib_path != "default":
    _lib.library_path = lambda: os.path.abspath(lib_path)  # this process only
import bench  # noqa: E402
import radfoam_b200  # noqa: E402
from quick_bench import timeit  # noqa: E402
from radfoam_b200 import sharded  # noqa: E402

f = bench.load_or_build_foam(1_048_576, print)
frame = bench.make_frame(f, 1920, 1080)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
scene = [d(x) for x in (f.points, f.attributes, f.adjacency, f.offsets)]
scene[0].requires_grad_(True)
res = {"tag": tag, "lib": lib_path}
for name, world in (("full", 1), ("shard_1_of_8", 8)):
    fr = {k: d(sharded.shard_image(torch.from_numpy(v), 0, world).contiguous().numpy()) for k, v in frame.items()}
    pipe = radfoam_b200.create_pipeline(3)
    fwd = pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"])
    res[f"{name}_fwd_record_ms"] = timeit(
        lambda: pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"]), iters=9, warmup=3)
    fwd = pipe.trace_forward(*scene, fr["rays"], fr["start"], depth_quantiles=fr["dq"])
    bwd = lambda: pipe.trace_backward(*scene, fr["rays"], fr["start"], fwd["rgba"], fr["grad_rgba"], fr["dq"],  # noqa: E731
                                      fwd["depth_indices"], fr["grad_depth"])
    out = bwd()
    res[f"{name}_bwd_replay_ms"] = timeit(bwd, iters=9, warmup=3)
    res[f"{name}_checksum"] = [float(fwd["rgba"].double().sum()), float(out["attr_grad"].double().abs().sum())]
print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"kblock_{tag}.json"), "w") as fh:
    json.dump(res, fh, indent=1)
