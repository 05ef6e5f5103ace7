"""The reference's real training access pattern (train.py:61, batch_fetcher.cpp:65-70): 1,000,000
rays drawn at random over 64 cameras (spatially incoherent, per-ray start cells), fwd+bwd, ours vs
the reference's kernels -- SURVEY.md §8d 'coherence caveat' row."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import common  # noqa: E402
import radfoam_b200  # noqa: E402
from oracle import ref_gpu  # noqa: E402
from radfoam_b200 import foam  # noqa: E402
from quick_bench import timeit  # noqa: E402

f = bench.load_or_build_foam(1_048_576, print)
rng = np.random.default_rng(64)
cams = rng.normal(size=(64, 3))
cams = 2.5 * cams / np.linalg.norm(cams, axis=1, keepdims=True)
R = 1_000_000
which = rng.integers(0, 64, size=R)
target = rng.normal(0.0, 0.5, size=(R, 3))
d = target - cams[which]
d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([cams[which], d], axis=1).astype(np.float32)
dq = np.sort(rng.uniform(0, 1, size=(R, 2)).astype(np.float32), axis=-1)[..., ::-1].copy()
g = rng.normal(size=(R, 4)).astype(np.float32)
gd = (rng.normal(size=(R, 2)) * 1e-4).astype(np.float32)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
scene = [t(f.points).requires_grad_(True), t(f.attributes), t(f.adjacency), t(f.offsets)]
rays_d, dq_d, g_d, gd_d = t(rays), t(dq), t(g), t(gd)
start_d = radfoam_b200.starting_points(rays_d, scene[0])
want_start = np.array([foam.nearest_point(f.points, c.astype(np.float32)) for c in cams], dtype=np.uint32)[which]
pipe = radfoam_b200.create_pipeline(3)
res = {"rays": R, "cameras": 64, "start_cells_match_numpy": bool(np.array_equal(start_d.cpu().numpy(), want_start))}
res["starting_points_ms"] = timeit(lambda: radfoam_b200.starting_points(rays_d, scene[0]))


def ours_fwd():
    return pipe.trace_forward(*scene, rays_d, start_d, depth_quantiles=dq_d)


fwd = ours_fwd()
res["mean_cells"] = float(fwd["num_intersections"].to(torch.int64).float().mean())
res["ours_fwd_ms"] = timeit(ours_fwd)
fwd = ours_fwd()
bwd = lambda: pipe.trace_backward(*scene, rays_d, start_d, fwd["rgba"], g_d, dq_d, fwd["depth_indices"], gd_d)  # noqa: E731
ob = bwd()
res["ours_bwd_ms"] = timeit(bwd)
sc = [x.detach() for x in scene]
rf = ref_gpu.trace_forward(*sc, rays_d, start_d, dq_d)
res["ref_fwd_ms"] = timeit(lambda: ref_gpu.trace_forward(*sc, rays_d, start_d, dq_d), iters=3, warmup=1)
rb = ref_gpu.trace_backward(*sc, rays_d, start_d, rf["rgba"], g_d, dq_d, rf["depth_indices"], gd_d)
res["ref_bwd_ms"] = timeit(lambda: ref_gpu.trace_backward(*sc, rays_d, start_d, rf["rgba"], g_d, dq_d,
                                                          rf["depth_indices"], gd_d), iters=3, warmup=1)
res["nint_equal"] = bool(torch.equal(rf["num_intersections"], fwd["num_intersections"]))
res["rgba_maxdiff"] = float((rf["rgba"] - fwd["rgba"]).abs().max())
for k in ("points_grad", "attr_grad"):
    res[k + "_err"] = common.grad_error(ob[k].cpu().numpy(), rb[k].cpu().numpy())

# the same calls with the pipeline's coherent re-ordering switched off (rays traced in batch order)
pipe.reorder_rays = False
fwd_u = ours_fwd()
res["unordered_fwd_ms"] = timeit(ours_fwd)
fwd_u = ours_fwd()
bwd_u = lambda: pipe.trace_backward(*scene, rays_d, start_d, fwd_u["rgba"], g_d, dq_d, fwd_u["depth_indices"], gd_d)  # noqa: E731
obu = bwd_u()
res["unordered_bwd_ms"] = timeit(bwd_u)
res["reorder_rgba_equal"] = bool(torch.equal(fwd_u["rgba"], fwd["rgba"]))
res["reorder_attr_grad_err"] = common.grad_error(obu["attr_grad"].cpu().numpy(), ob["attr_grad"].cpu().numpy())
pipe.reorder_rays = True
res["mrays_fwdbwd_ours"] = R / (res["ours_fwd_ms"] + res["ours_bwd_ms"]) / 1e3
res["mrays_fwdbwd_ref"] = R / (res["ref_fwd_ms"] + res["ref_bwd_ms"]) / 1e3
res["speedup"] = res["mrays_fwdbwd_ours"] / res["mrays_fwdbwd_ref"]
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/random_batch.json", "w"), indent=1)
