"""Golden (indices, cell_radius) from the reference's OWN farthest_neighbor kernel (oracle/_ref, B200) for the
CPU test of the restatement.  Run on the GPU box: python tests/golden/make_golden_farthest.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from oracle import ref_gpu  # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
os.makedirs(out_dir, exist_ok=True)
rec = {}
for tag, f in (("scene20k", common.scene_case(20000, 8, 8, 0).foam), ("edge", common.farthest_edge_case())):
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    idx, radius = ref_gpu.farthest_neighbor(d(f.points), d(f.adjacency), d(f.offsets))
    torch.cuda.synchronize()
    rec[f"{tag}_points"], rec[f"{tag}_adjacency"], rec[f"{tag}_offsets"] = f.points, f.adjacency, f.offsets
    rec[f"{tag}_indices"], rec[f"{tag}_radius"] = idx.cpu().numpy(), radius.cpu().numpy()
np.savez_compressed(os.path.join(out_dir, "farthest_neighbor.npz"), **rec)
print("written", {k: v.shape for k, v in rec.items()})
