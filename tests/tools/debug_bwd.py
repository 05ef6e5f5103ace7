"""Compare direct vs cached backward (and the reference) entry by entry on a failing case."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import radfoam_b200  # noqa: E402
from oracle import ref_gpu  # noqa: E402

case = common.random_ray_case(num_points=60000, num_rays=100000) if len(sys.argv) < 2 else \
    common.scene_case(num_points=60000, width=320, height=200, inside=True)
f = case.foam
d = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
scene = [d(x) for x in (f.points, f.attributes, f.adjacency, f.offsets)]
rays, start, dq, g, gd = d(case.rays), d(case.start), d(case.quantiles), d(case.grad_rgba), d(case.grad_depth)
res = {}
for mode in ("direct", "cached"):
    os.environ["RFB_BWD_MODE"] = mode
    pipe = radfoam_b200.create_pipeline(3)
    fwd = pipe.trace_forward(*scene, rays, start, depth_quantiles=dq)
    bwd = pipe.trace_backward(*scene, rays, start, fwd["rgba"], g, dq, fwd["depth_indices"], gd)
    torch.cuda.synchronize()
    res[mode] = {k: v.cpu().numpy() for k, v in bwd.items() if k != "ray_grad"}
rf = ref_gpu.trace_forward(*scene, rays, start, dq)
rb = ref_gpu.trace_backward(*scene, rays, start, rf["rgba"], g, dq, rf["depth_indices"], gd)
res["ref"] = {k: v.cpu().numpy() for k, v in rb.items() if k != "ray_grad"}
for k in ("points_grad", "attr_grad"):
    ref = res["ref"][k].astype(np.float64)
    for mode in ("direct", "cached"):
        got = res[mode][k].astype(np.float64)
        ok = np.isfinite(ref) & np.isfinite(got)
        diff = np.where(ok, np.abs(got - ref), 0)
        scale = np.abs(ref[ok]).max()
        idx = np.argsort(-diff.ravel())[:8]
        print(k, mode, "max|d|/max|ref| =", diff.max() / scale, "max|ref| =", scale,
              "entries with d > 1e-6*scale:", int((diff > 1e-6 * scale).sum()))
        for i in idx:
            r, c = divmod(int(i), ref.shape[1])
            print(f"   [{r},{c}] ref={ref[r, c]:.9g} got={got[r, c]:.9g} direct={res['direct'][k][r, c]:.9g} d={diff[r, c]:.3g}")
