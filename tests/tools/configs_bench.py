"""BASELINE.json configs 2, 3, 5 on one GPU: parity vs the reference's kernels + timings.
python tests/tools/configs_bench.py [2 3 5]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import radfoam_b200  # noqa: E402
from oracle import ref_gpu  # noqa: E402
from radfoam_b200 import foam  # noqa: E402
from quick_bench import timeit  # noqa: E402

CONFIGS = {
    2: dict(points=524_288, width=1920, height=1080, q=0, backward=False, pos=(0.3, 0.3, 0.3), target=(1.0, 0.2, -0.1)),
    3: dict(points=2_097_152, width=1920, height=1080, q=2, backward=True, pos=(2.5, 2.5, 2.5), target=(0, 0, 0)),
    5: dict(points=4_194_304, width=3840, height=2160, q=2, backward=True, pos=(2.5, 2.5, 2.5), target=(0, 0, 0)),
}


def run(cid):
    c = CONFIGS[cid]
    t0 = time.time()
    f = foam.scene_foam(c["points"])
    res = {"config": cid, "points": f.num_points, "edges": int(f.adjacency.size), "foam_build_s": time.time() - t0}
    W, H = c["width"], c["height"]
    rays = foam.pinhole_rays(W, H, c["pos"], target=c["target"], fov=0.9)
    start = np.full((H, W), foam.nearest_point(f.points, c["pos"]), dtype=np.uint32)
    rng = np.random.default_rng(cid)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    scene = [d(f.points).requires_grad_(c["backward"]), d(f.attributes), d(f.adjacency), d(f.offsets)]
    rays_d, start_d = d(rays), d(start)
    dq_d = g_d = gd_d = None
    if c["q"]:
        dq_d = d(np.sort(rng.uniform(0, 1, size=(H, W, c["q"])).astype(np.float32), axis=-1)[..., ::-1].copy())
        gd_d = d((rng.normal(size=(H, W, c["q"])) * 1e-4).astype(np.float32))
    g_d = d(rng.normal(size=(H, W, 4)).astype(np.float32))
    pipe = radfoam_b200.create_pipeline(3)
    R = W * H

    def ours_fwd():
        return pipe.trace_forward(*scene, rays_d, start_d, depth_quantiles=dq_d)

    fwd = ours_fwd()
    n = fwd["num_intersections"].to(torch.int64)
    res.update(rays=R, mean_cells=float(n.float().mean()), max_cells=int(n.max()))
    res["ours_fwd_ms"] = timeit(ours_fwd)
    fwd = ours_fwd()

    def ours_bwd():
        return pipe.trace_backward(*scene, rays_d, start_d, fwd["rgba"], g_d, dq_d, fwd.get("depth_indices"), gd_d)

    if c["backward"]:
        ob = ours_bwd()
        res["ours_bwd_ms"] = timeit(ours_bwd)
        try:
            res["tape"] = pipe.tape_status()
        except RuntimeError:
            pass
    sc = [t.detach() for t in scene]
    rf = ref_gpu.trace_forward(*sc, rays_d, start_d, dq_d)
    res["ref_fwd_ms"] = timeit(lambda: ref_gpu.trace_forward(*sc, rays_d, start_d, dq_d), iters=3, warmup=1)
    res["nint_equal"] = bool(torch.equal(rf["num_intersections"], fwd["num_intersections"]))
    res["rgba_maxdiff"] = float((rf["rgba"] - fwd["rgba"]).abs().max())
    if c["q"]:
        res["didx_equal"] = bool(torch.equal(rf["depth_indices"], fwd["depth_indices"]))
        res["depth_maxdiff"] = float((rf["depth"] - fwd["depth"]).abs().max())
    if c["backward"]:
        rb = ref_gpu.trace_backward(*sc, rays_d, start_d, rf["rgba"], g_d, dq_d, rf["depth_indices"], gd_d)
        res["ref_bwd_ms"] = timeit(lambda: ref_gpu.trace_backward(
            *sc, rays_d, start_d, rf["rgba"], g_d, dq_d, rf["depth_indices"], gd_d), iters=2, warmup=1)
        for k in ("points_grad", "attr_grad"):
            res[k + "_err"] = common.grad_error(ob[k].cpu().numpy(), rb[k].cpu().numpy())
        res["mrays_fwdbwd_ours"] = R / (res["ours_fwd_ms"] + res["ours_bwd_ms"]) / 1e3
        res["mrays_fwdbwd_ref"] = R / (res["ref_fwd_ms"] + res["ref_bwd_ms"]) / 1e3
    res["mrays_fwd_ours"] = R / res["ours_fwd_ms"] / 1e3
    res["mrays_fwd_ref"] = R / res["ref_fwd_ms"] / 1e3
    res["gpu_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
    del pipe
    ref_gpu.load().rfref_release_pool()
    torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    ids = [int(a) for a in sys.argv[1:]] or [2, 3, 5]
    out = []
    for cid in ids:
        r = run(cid)
        print(json.dumps(r), flush=True)
        out.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/configs_bench.json", "w"), indent=1)
