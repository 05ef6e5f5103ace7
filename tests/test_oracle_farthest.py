"""CPU tests of the farthest-neighbour restatement (oracle/radfoam_oracle.c:rfo_farthest_neighbor) -- SURVEY.md
§8f.4: against an independent numpy restatement, against golden outputs of the reference's own kernel
(tests/golden/farthest_neighbor.npz, made on a B200 by tests/golden/make_golden_farthest.py), and the host-side
validation of the product wrapper (no GPU needed for that)."""
import os

import numpy as np
import pytest

import common
from oracle import oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "farthest_neighbor.npz")


def test_edge_cases_match_numpy_restatement():
    f = common.farthest_edge_case()
    idx, radius = oracle.farthest_neighbor(f.points, f.adjacency, f.offsets)
    ref_idx, ref_radius = common.farthest_neighbor_numpy(f.points, f.adjacency, f.offsets)
    assert np.array_equal(idx, ref_idx)
    common.assert_same_floats(radius, ref_radius)
    # the documented corner cases
    assert idx[0] == common.NONE and radius[0] == 0.0            # every neighbour coincides
    assert idx[4] == 5                                            # exact tie: first in row order
    assert idx[13] == common.NONE and np.isnan(radius[13])        # empty row
    assert idx[11] == common.NONE and np.isnan(radius[11])        # every distance NaN
    assert idx[12] == 8 and np.isnan(radius[12])                  # inf wins, NaN poisons the sum
    assert idx[8] == 10 and np.isinf(radius[8])
    assert idx[10] == common.NONE and radius[10] == 0.0           # |d|^2 underflows to 0


def test_small_foam_matches_numpy_restatement():
    from radfoam_b200 import foam

    f = foam.small_foam(300, sh_degree=0, seed=5)
    idx, radius = oracle.farthest_neighbor(f.points, f.adjacency, f.offsets)
    ref_idx, ref_radius = common.farthest_neighbor_numpy(f.points, f.adjacency, f.offsets)
    assert np.array_equal(idx, ref_idx)
    common.assert_same_floats(radius, ref_radius)
    # sanity: the farthest neighbour is adjacent and no adjacent point is farther
    for i in (0, 17, 299):
        row = f.adjacency[f.offsets[i]:f.offsets[i + 1]]
        d = np.linalg.norm(f.points[row].astype(np.float64) - f.points[i], axis=1)
        assert idx[i] in row and d[list(row).index(idx[i])] >= d.max() * (1 - 1e-6)
        assert abs(radius[i] - 0.5 * d.mean()) <= 1e-5 * d.mean()


@pytest.mark.parametrize("tag", ["scene20k", "edge"])
def test_oracle_matches_reference_kernel_golden(tag):
    z = np.load(GOLDEN)
    idx, radius = oracle.farthest_neighbor(z[f"{tag}_points"], z[f"{tag}_adjacency"], z[f"{tag}_offsets"])
    assert np.array_equal(idx, z[f"{tag}_indices"])               # integers: bit-exact
    common.assert_same_floats(radius, z[f"{tag}_radius"])         # floats: bit-exact too (IEEE sqrt/div, pinned fma)


def test_wrapper_validation_without_gpu():
    import torch

    import radfoam_b200

    pts = torch.zeros((4, 3))
    adj = torch.zeros((2,), dtype=torch.uint32)
    off = torch.zeros((5,), dtype=torch.uint32)
    with pytest.raises(RuntimeError, match="points must be on CUDA device"):   # triangulation_bindings.cpp:194-196
        radfoam_b200.farthest_neighbor(pts, adj, off)
