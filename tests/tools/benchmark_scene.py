"""FPS of a radfoam ``.pt`` checkpoint through this library -- the reference's benchmark.py:95-139 loop on
``radfoam_b200`` (fp16 attributes, weight_threshold 0.05, in-kernel ray generation, RGBA8 frames).

    python tests/tools/benchmark_scene.py path/to/model.pt [--sh-degree 3] [--width 1920 --height 1080]
    python tests/tools/benchmark_scene.py --synthetic 1048576          # no checkpoint: a synthetic foam saved to .pt first

Real checkpoints carry their test cameras in the dataset, which is out of scope here; the poses are an orbit around the
scene's centre at 2.5x its RMS radius, every 8th of 64 used as in benchmark.py:63-64."""
import argparse
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import radfoam_b200  # noqa: E402
from radfoam_b200 import foam, scene_io  # noqa: E402


def orbit_c2w(centre, radius, n=64, height_frac=0.4):
    c2w = torch.zeros((n, 4, 4))
    for i in range(n):
        ang = 2 * np.pi * i / n
        pos = centre + radius * np.array([np.cos(ang), np.sin(ang), height_frac])
        cam = foam.camera_dict(pos, target=centre)
        c2w[i, :3, 0] = torch.from_numpy(cam["right"])
        c2w[i, :3, 1] = -torch.from_numpy(cam["up"])
        c2w[i, :3, 2] = torch.from_numpy(cam["forward"])
        c2w[i, :3, 3] = torch.from_numpy(cam["position"])
        c2w[i, 3, 3] = 1.0
    return c2w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("checkpoint", nargs="?")
    ap.add_argument("--synthetic", type=int, default=0, help="points of a synthetic foam to use instead of a checkpoint")
    ap.add_argument("--sh-degree", type=int, default=3)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--fov", type=float, default=0.9)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    path = args.checkpoint
    if path is None:
        n = args.synthetic or 262144
        f = foam.scene_foam(n, sh_degree=args.sh_degree)
        path = os.path.join(tempfile.mkdtemp(), "model.pt")
        scene_io.FoamScene.from_foam(f, device="cpu").save_pt(path)
    scene = scene_io.FoamScene.load_pt(path, sh_degree=args.sh_degree, attr_dtype=torch.float16, device="cuda")
    pts = scene.primal_points.double()
    centre = pts.median(dim=0).values.cpu().numpy()
    radius = 2.5 * float((pts - pts.median(dim=0).values).norm(dim=-1).median())
    fy = args.height / (2.0 * np.tan(args.fov / 2.0))
    cameras, positions = scene_io.benchmark_cameras(orbit_c2w(centre, radius), fy, args.width, args.height)
    pipe = radfoam_b200.create_pipeline(args.sh_degree, "float16")
    res = scene_io.benchmark_fps(pipe, scene, cameras, positions, n_reps=args.reps)
    print(f"Total time: {res['total_ms']} ms")
    print(f"FPS: {res['fps']}")
    print(json.dumps({"checkpoint": path, "points": scene.num_points, "frames": res["frames"], "fps": res["fps"],
                      "width": args.width, "height": args.height}))


if __name__ == "__main__":
    main()
