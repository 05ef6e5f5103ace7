"""CPU tests of host-side logic: synthetic foam formats, ray sharding arithmetic, oracle
self-consistency properties (no GPU)."""
import numpy as np
import pytest
import torch

import common
from radfoam_b200 import sharded


def test_foam_arrays_follow_the_reference_contract():
    f = common.scene_case(2000, 64, 48).foam
    n = f.num_points
    assert f.points.dtype == np.float32 and f.points.shape == (n, 3)
    assert f.attributes.shape == (n, 49) and f.offsets.shape == (n + 1,)
    assert f.adjacency.dtype == np.uint32 and f.offsets.dtype == np.uint32
    off = f.offsets.astype(np.int64)
    assert off[0] == 0 and off[-1] == f.adjacency.size and (np.diff(off) >= 3).all()
    rows = np.repeat(np.arange(n), np.diff(off))
    adj = f.adjacency.astype(np.int64)
    assert (adj != rows).all()                      # no self loops
    key = rows * n + adj
    assert (np.diff(key) > 0).all()                 # rows ascending, no duplicates
    assert np.array_equal(np.sort(adj * n + rows), key)  # symmetric graph
    assert (f.attributes[:, -1] >= 0).all()         # density last, non-negative


@pytest.mark.parametrize("height,world,band", [(1080, 8, 8), (1080, 3, 8), (37, 4, 8), (5, 2, 8), (64, 1, 8)])
def test_band_sharding_is_a_partition_and_round_trips(height, world, band):
    img = torch.arange(height * 7 * 2, dtype=torch.float32).reshape(height, 7, 2)
    parts = [sharded.shard_image(img, r, world, band) for r in range(world)]
    assert sum(p.shape[0] for p in parts) == height
    assert torch.equal(sharded.unshard_image(parts, height, band), img)
    sizes = [p.shape[0] for p in parts]
    assert max(sizes) - min(sizes) <= band
    for r, p in enumerate(parts):  # every band except the image's last is whole: tiles stay aligned
        rows = sharded.band_rows(height, r, world, band)
        assert torch.equal(p, img[rows])


@pytest.mark.parametrize("n,world", [(10, 3), (7, 8), (1000, 4)])
def test_flat_sharding_is_a_partition(n, world):
    t = torch.arange(n)
    parts = [sharded.shard_flat(t, r, world) for r in range(world)]
    assert torch.equal(torch.cat(parts), t)
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_oracle_properties_on_config1():
    """Size-independent invariants the reference's algorithm has (checked on the CPU oracle
    here and on the CUDA path in the gpu tests)."""
    from oracle import oracle

    case = common.config1(3, 2)
    f = case.foam
    out = oracle.trace_forward(f.points, f.attributes, f.adjacency, f.offsets, case.rays, case.start,
                               case.quantiles, return_contribution=True)
    rgba, n = out["rgba"], out["num_intersections"]
    assert (rgba[..., 3] >= 0).all() and (rgba[..., 3] <= 1).all()
    assert (rgba[..., :3] >= 0).all()
    assert n.min() >= 1 and n.max() <= 1025
    valid = out["depth_indices"] != common.NONE
    assert (out["depth"][~valid] == -1).all() and (out["depth"][valid] >= 0).all()
    # quantile 0.7 is crossed no later than quantile 0.3
    both = valid.all(axis=-1)
    assert (out["depth"][both][:, 0] <= out["depth"][both][:, 1]).all()
    # sum of per-point contributions == sum of ray opacities (each weight is added once)
    np.testing.assert_allclose(out["contribution"].sum(), rgba[..., 3].sum(), rtol=1e-4)
    # determinism and thread-count independence of the forward
    again = oracle.trace_forward(f.points, f.attributes, f.adjacency, f.offsets, case.rays, case.start,
                                 case.quantiles, num_threads=0)
    assert np.array_equal(again["rgba"], rgba) and np.array_equal(again["num_intersections"], n)
    # max_intersections budget: n == budget + 1 exactly when the budget ran out
    capped = oracle.trace_forward(f.points, f.attributes, f.adjacency, f.offsets, case.rays, case.start,
                                  None, max_intersections=4)
    assert capped["num_intersections"].max() == 5
    assert np.array_equal(np.minimum(n, 5), capped["num_intersections"])


def test_oracle_backward_is_linear_and_ray_additive():
    from oracle import oracle

    case = common.config1(2, 2)
    f = case.foam
    args = (f.points, f.attributes, f.adjacency, f.offsets)
    fwd = oracle.trace_forward(*args, case.rays, case.start, case.quantiles)

    def bwd(rows, scale=1.0):
        sl = slice(*rows)
        return oracle.trace_backward(*args, case.rays[sl], case.start[sl], fwd["rgba"][sl],
                                     case.grad_rgba[sl] * scale, case.quantiles[sl],
                                     fwd["depth_indices"][sl], case.grad_depth[sl] * scale)

    full, top, bottom, twice = bwd((0, 32)), bwd((0, 16)), bwd((16, 32)), bwd((0, 32), 2.0)
    for k in ("points_grad", "attr_grad"):
        assert common.grad_error(top[k] + bottom[k], full[k]) < 1e-5   # rays shard additively
        assert common.grad_error(twice[k], 2.0 * full[k]) < 1e-6       # linear in dL/dout


def test_prefetch_matches_numpy_half_rounding():
    from oracle import oracle

    f = common.config1(3, 0).foam
    diff = oracle.prefetch_adjacent_diff(f.points, f.adjacency, f.offsets)
    rows = np.repeat(np.arange(f.num_points), np.diff(f.offsets.astype(np.int64)))
    want = (f.points[f.adjacency.astype(np.int64)] - f.points[rows]).astype(np.float16)
    assert np.array_equal(diff[:, :3].view(np.uint16), want.view(np.uint16))
    assert (diff[:, 3] == 0).all()
