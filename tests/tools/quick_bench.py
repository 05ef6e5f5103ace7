"""Ad-hoc timing of ours vs the reference's own kernels on one GPU (not bench.py):
python tests/tools/quick_bench.py --points 1000000 --width 1920 --height 1080"""
import argparse
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


def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--out", default="gpurun_out/quick_bench.json")
    ap.add_argument("--no-ref", action="store_true")
    args = ap.parse_args()
    t0 = time.time()
    f = foam.scene_foam(args.points)
    print(f"foam: {f.num_points} points, E={f.adjacency.size}, {time.time() - t0:.1f}s", flush=True)
    pos = (2.5, 2.5, 2.5)
    rays = foam.pinhole_rays(args.width, args.height, pos, fov=0.9)
    start = np.full((args.height, args.width), foam.nearest_point(f.points, pos), dtype=np.uint32)
    rng = np.random.default_rng(0)
    dq = np.sort(rng.uniform(0.05, 0.95, size=(args.height, args.width, 2)).astype(np.float32), axis=-1)[..., ::-1].copy()
    g = rng.normal(size=(args.height, args.width, 4)).astype(np.float32)
    gd = (rng.normal(size=(args.height, args.width, 2)) * 1e-4).astype(np.float32)
    d = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    scene = [d(x) for x in (f.points, f.attributes, f.adjacency, f.offsets)]
    rays_d, start_d, dq_d, g_d, gd_d = d(rays), d(start), d(dq), d(g), d(gd)
    pipe = radfoam_b200.create_pipeline(3)
    res = {"points": f.num_points, "rays": args.width * args.height}

    fwd = pipe.trace_forward(*scene, rays_d, start_d, depth_quantiles=dq_d)
    nint = fwd["num_intersections"].cpu().numpy().astype(np.int64)
    res["mean_cells_per_ray"] = float(nint.mean())
    res["max_cells_per_ray"] = int(nint.max())
    res["alpha_mean"] = float(fwd["rgba"][..., 3].mean())
    R = res["rays"]

    def ours_fwd():
        return pipe.trace_forward(*scene, rays_d, start_d, depth_quantiles=dq_d)

    def ours_bwd():
        return pipe.trace_backward(*scene, rays_d, start_d, fwd["rgba"], g_d, dq_d, fwd["depth_indices"], gd_d)

    pipe.cache_scene = False
    res["ours_fwd_ms_nocache"] = timeit(ours_fwd)
    res["ours_bwd_ms_nocache"] = timeit(ours_bwd)
    pipe.cache_scene = True
    res["ours_fwd_ms"] = timeit(ours_fwd)
    res["ours_bwd_ms"] = timeit(ours_bwd)
    flat = [rays_d.reshape(-1, 6), start_d.reshape(-1), dq_d.reshape(-1, 2)]
    res["ours_fwd_ms_linear"] = timeit(lambda: pipe.trace_forward(*scene, flat[0], flat[1], depth_quantiles=flat[2]))
    print(json.dumps(res), flush=True)

    if not args.no_ref and ref_gpu.available():
        rf = ref_gpu.trace_forward(*scene, rays_d, start_d, dq_d)
        res["ref_fwd_ms"] = timeit(lambda: ref_gpu.trace_forward(*scene, rays_d, start_d, dq_d))
        res["ref_bwd_ms"] = timeit(lambda: ref_gpu.trace_backward(
            *scene, rays_d, start_d, rf["rgba"], g_d, dq_d, rf["depth_indices"], gd_d), iters=3, warmup=1)
        rb = ref_gpu.trace_backward(*scene, rays_d, start_d, rf["rgba"], g_d, dq_d, rf["depth_indices"], gd_d)
        ob = ours_bwd()
        torch.cuda.synchronize()
        res["nint_equal"] = bool(torch.equal(rf["num_intersections"], fwd["num_intersections"]))
        res["nint_mismatch"] = int((rf["num_intersections"].int() != fwd["num_intersections"].int()).sum())
        res["didx_equal"] = bool(torch.equal(rf["depth_indices"], fwd["depth_indices"]))
        res["rgba_maxdiff"] = float((rf["rgba"] - fwd["rgba"]).abs().max())
        res["rgba_bitequal"] = bool(torch.equal(rf["rgba"], fwd["rgba"]))
        res["depth_maxdiff"] = float((rf["depth"] - fwd["depth"]).abs().max())
        for k in ("points_grad", "attr_grad"):
            res[k + "_err"] = common.grad_error(ob[k].cpu().numpy(), rb[k].cpu().numpy())
        rb2 = ref_gpu.trace_backward(*scene, rays_d, start_d, rf["rgba"], g_d, dq_d, rf["depth_indices"], gd_d)
        for k in ("points_grad", "attr_grad"):
            res[k + "_ref_selfnoise"] = common.grad_error(rb2[k].cpu().numpy(), rb[k].cpu().numpy())
        res["speedup_fwd"] = res["ref_fwd_ms"] / res["ours_fwd_ms"]
        res["speedup_fwdbwd"] = (res["ref_fwd_ms"] + res["ref_bwd_ms"]) / (res["ours_fwd_ms"] + res["ours_bwd_ms"])
    res["ours_fwd_mrays"] = R / res["ours_fwd_ms"] / 1e3
    res["ours_fwdbwd_mrays"] = R / (res["ours_fwd_ms"] + res["ours_bwd_ms"]) / 1e3
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
