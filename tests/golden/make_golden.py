"""Generate golden vectors from the reference's OWN kernels (oracle/_ref) on a GPU box.

Run on a B200 via gpurun:  python tests/golden/make_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/.  Each file holds the seeded inputs
and the outputs of /root/reference/src/tracing/pipeline.cu (compiled unmodified against
oracle/eigen_shim) for one case; tests/test_oracle_golden.py checks the CPU restatement
against them on CPU-only boxes, tests/test_gpu_parity.py checks the CUDA path on the GPU.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
from oracle import ref_gpu  # noqa: E402


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run(case, attr_dtype=np.float32, **settings):
    f = case.foam
    attrs = f.attributes.astype(attr_dtype)
    scene = [dev(x) for x in (f.points, attrs, f.adjacency, f.offsets)]
    rays, start, dq = dev(case.rays), dev(case.start), dev(case.quantiles)
    fwd = ref_gpu.trace_forward(*scene, rays, start, dq, return_contribution=True, **settings)
    g = case.grad_rgba.astype(attr_dtype)
    bwd = ref_gpu.trace_backward(*scene, rays, start, fwd["rgba"], dev(g), dq, fwd.get("depth_indices"),
                                 dev(case.grad_depth), **settings)
    torch.cuda.synchronize()
    rec = dict(points=f.points, attributes=attrs, adjacency=f.adjacency, offsets=f.offsets,
               rays=case.rays, start=case.start, grad_rgba=g,
               sh_degree=np.int32(f.sh_degree),
               weight_threshold=np.float32(settings.get("weight_threshold", 0.001)),
               max_intersections=np.uint32(settings.get("max_intersections", 1024)))
    if case.quantiles is not None:
        rec.update(quantiles=case.quantiles, grad_depth=case.grad_depth)
    for k, v in list(fwd.items()) + list(bwd.items()):
        if k != "ray_grad":
            rec["out_" + k] = v.cpu().numpy()
    return rec


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    cases = {
        "config1_deg0_q2": (common.config1(0, 2), {}),
        "config1_deg1_q2": (common.config1(1, 2), {}),
        "config1_deg2_q0": (common.config1(2, 0), {}),
        "config1_deg3_q2": (common.config1(3, 2), {}),
        "config1_deg3_q3": (common.config1(3, 3, fixed_quantiles=False), {}),
        "config1_deg3_q2_maxint5": (common.config1(3, 2), dict(max_intersections=5)),
        "config1_deg3_q2_thr0p3": (common.config1(3, 2), dict(weight_threshold=0.3)),
        "scene2k_deg3_q2": (common.scene_case(2000, 64, 48, 2), {}),
        "scene2k_inside_deg3_q2": (common.scene_case(2000, 64, 48, 2, inside=True), {}),
        "random2k_deg3_q2": (common.random_ray_case(2000, 3000, 8, 2), {}),
    }
    for name, (case, settings) in cases.items():
        rec = run(case, **settings)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **rec)
        print(name, "mean n =", float(rec["out_num_intersections"].mean()))
    rec = run(common.config1(3, 2), np.float16)
    np.savez_compressed(os.path.join(outdir, "config1_deg3_q2_half.npz"), **rec)
    print("golden vectors written to", outdir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
