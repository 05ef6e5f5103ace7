"""farthest_neighbor at ~1M points: ours vs the reference's own kernel (oracle/_ref), L2 flushed
between timed calls, outputs compared bit-for-bit.  The point set is a 200k-point foam tiled 5x (ids shifted), so
building it costs seconds instead of a 1M-point Delaunay; row lengths and gather locality are the foam's."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import radfoam_b200  # noqa: E402
from oracle import ref_gpu  # noqa: E402
from radfoam_b200 import foam  # noqa: E402

base = foam.scene_foam(int(os.environ.get("FARTHEST_BASE", 200_000)), sh_degree=0)
tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n, e = base.num_points, base.adjacency.size
points = np.concatenate([base.points + np.float32(4.0 * k) for k in range(tiles)])
adjacency = np.concatenate([base.adjacency + np.uint32(n * k) for k in range(tiles)])
offsets = np.concatenate([base.offsets[:-1].astype(np.uint64) + e * k for k in range(tiles)] +
                         [np.array([e * tiles], dtype=np.uint64)]).astype(np.uint32)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
p, a, o = d(points), d(adjacency), d(offsets)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10):
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from oracle import oracle  # noqa: E402

edge = common.farthest_edge_case()
edge_ref = oracle.farthest_neighbor(edge.points, edge.adjacency, edge.offsets)
nbytes = 16 * adjacency.size + 24 * points.shape[0]
res = {"points": int(points.shape[0]), "edges": int(adjacency.size), "algorithmic_bytes": int(nbytes),
       "bytes_note": "adjacency 4E + gathered neighbour points 12E + own point 12N + offsets 4N + outputs 8N"}
ref_out = None
if ref_gpu.available():
    res["reference_ms"], ref_out = timed(lambda: ref_gpu.farthest_neighbor(p, a, o))
ms, (idx, radius) = timed(lambda: radfoam_b200.farthest_neighbor(p, a, o))
rec = {"ms": ms, "GBps": nbytes / ms / 1e6}
if ref_out is not None:
    rec["speedup_vs_reference"] = res["reference_ms"] / ms
    rec["identical_to_reference"] = bool(torch.equal(idx.view(torch.int32), ref_out[0].view(torch.int32)) and
                                         torch.equal(radius.view(torch.int32), ref_out[1].view(torch.int32)))
e_idx, e_radius = radfoam_b200.farthest_neighbor(d(edge.points), d(edge.adjacency), d(edge.offsets))
try:
    assert np.array_equal(e_idx.cpu().numpy(), edge_ref[0])
    common.assert_same_floats(e_radius.cpu().numpy(), edge_ref[1])
    rec["edge_case_matches_oracle"] = True
except AssertionError:
    rec["edge_case_matches_oracle"] = False
res["ours"] = rec
print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "farthest_bench.json"), "w") as fh:
    fh.write(json.dumps(res) + "\n")
