"""trace_benchmark FPS (the reference's own speed method, benchmark.py:95-139: fp16 attributes,
weight_threshold 0.05, in-kernel ray generation, RGBA8 output): ours vs the reference's kernel."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import radfoam_b200  # noqa: E402
from oracle import ref_gpu  # noqa: E402
from radfoam_b200 import foam  # noqa: E402
from quick_bench import timeit  # noqa: E402

points = int(sys.argv[1]) if len(sys.argv) > 1 else 1_048_576
f = bench.load_or_build_foam(points, print)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
res = {}
for dtype in (np.float16, np.float32):
    name = "f16" if dtype == np.float16 else "f32"
    scene = [d(f.points), d(f.attributes.astype(dtype)), d(f.adjacency), d(f.offsets)]
    pipe = radfoam_b200.create_pipeline(3, "float16" if dtype == np.float16 else "float32")
    diff = pipe.prefetch_adjacent_diff(scene[0], scene[2], scene[3])
    cams = []
    for k in range(4):
        ang = 0.4 * k
        pos = (2.5 * np.cos(ang) * 1.2, 2.5 * np.sin(ang) * 1.2, 2.0)
        cam = foam.camera_dict(pos, fov=0.9, width=1920, height=1080)
        start = d(np.array([foam.nearest_point(f.points, pos)], dtype=np.uint32))
        cams.append((cam, {k2: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k2, v in cam.items()}, start))
    out = torch.zeros((1080, 1920), dtype=torch.uint32, device="cuda")
    out_ref = torch.zeros_like(out)

    def ours():
        for cam, cam_t, start in cams:
            pipe.trace_benchmark(*scene, diff, cam_t, start, out, weight_threshold=0.05)

    def ref():
        for cam, cam_t, start in cams:
            ref_gpu.trace_benchmark(*scene, diff, cam, start, out_ref, weight_threshold=0.05)

    ours(); ref(); torch.cuda.synchronize()
    a = out.cpu().numpy().view(np.uint8).astype(np.int32)
    b = out_ref.cpu().numpy().view(np.uint8).astype(np.int32)
    res[name] = {"ours_fps": 4000.0 / timeit(ours), "ref_fps": 4000.0 / timeit(ref),
                 "max_channel_diff": int(np.abs(a - b).max()),
                 "pixels_differing": float((a.reshape(-1, 4) != b.reshape(-1, 4)).any(axis=1).mean())}
    res[name]["speedup"] = res[name]["ours_fps"] / res[name]["ref_fps"]
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/fps_bench.json", "w"), indent=1)
