"""Golden RGBA8 frames from the reference's OWN benchmark kernel (oracle/_ref, B200) for the CPU test
of the restatement's trace_benchmark.  Run on the GPU box: python tests/golden/make_golden_benchmark.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from oracle import ref_gpu  # noqa: E402
from radfoam_b200 import foam  # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
os.makedirs(out_dir, exist_ok=True)
f = common.scene_case(2000, 64, 48, 2).foam
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
rec = dict(points=f.points, adjacency=f.adjacency, offsets=f.offsets)
for dtype, tag in ((np.float16, "f16"), (np.float32, "f32")):
    attrs = f.attributes.astype(dtype)
    scene = [d(f.points), d(attrs), d(f.adjacency), d(f.offsets)]
    diff = ref_gpu.prefetch_adjacent_diff(scene[0], scene[2], scene[3])
    rec["attributes_" + tag] = attrs
    rec["adjacent_diff"] = diff.cpu().numpy()
    for model, fov in (("pinhole", 0.9), ("fisheye", 1.2)):
        pos = (2.5, 2.5, 2.5)
        cam = foam.camera_dict(pos, fov=fov, width=64, height=48, model=model)
        start = d(np.array([foam.nearest_point(f.points, pos)], dtype=np.uint32))
        img = torch.zeros((48, 64), dtype=torch.uint32, device="cuda")
        ref_gpu.trace_benchmark(*scene, diff, cam, start, img, weight_threshold=0.05)
        torch.cuda.synchronize()
        rec[f"image_{tag}_{model}"] = img.cpu().numpy()
        rec[f"start_{model}"] = start.cpu().numpy()
        for k in ("position", "forward", "right", "up"):
            rec[f"cam_{model}_{k}"] = cam[k]
        rec[f"cam_{model}_fov"] = np.float32(fov)
np.savez_compressed(os.path.join(out_dir, "benchmark2k.npz"), **rec)
print("written", {k: getattr(v, "shape", None) for k, v in rec.items() if k.startswith("image")})
