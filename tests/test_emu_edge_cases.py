"""Edge cases of the tracing path, kernel logic (CPU emulator, tests/emu) against the oracle: budget exhaustion,
threshold extremes, ragged image / batch sizes, degenerate rays, zero / tiny / huge densities, quantiles at 0 and 1,
far-field points whose fp16 neighbour offsets overflow, wrong start cells.  Integers must agree exactly, floats to 1e-5
with identical NaN / inf patterns, gradients to 2e-5 of max with identical non-finite patterns."""
import os
import sys

import numpy as np
import pytest

import common
from oracle import oracle
from radfoam_b200 import foam

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))
import emu  # noqa: E402


def compare(f, rays, start, dq, weight_threshold=None, max_intersections=None, attributes=None, grads=True):
    attributes = f.attributes if attributes is None else attributes
    scene = (f.points, attributes, f.adjacency, f.offsets)
    kw = {k: v for k, v in (("weight_threshold", weight_threshold), ("max_intersections", max_intersections))
          if v is not None}
    pipe = emu.EmuPipeline(f.sh_degree)
    got = pipe.trace_forward(*scene, rays, start, dq, **kw)
    ref = oracle.trace_forward(*scene, rays, start, dq, **kw)
    for k in ("num_intersections", "depth_indices"):
        if k in got:
            assert np.array_equal(got[k].reshape(-1), np.asarray(ref[k]).reshape(-1)), k
    for k in ("rgba", "depth"):
        if k in got:
            a, b = got[k].astype(np.float32).reshape(-1), np.asarray(ref[k], dtype=np.float32).reshape(-1)
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b)), k
            m = np.isfinite(b)
            if m.any():
                assert np.abs(a[m] - b[m]).max() <= 1e-5 * max(1.0, float(np.abs(b[m]).max())), k
    if not grads:
        return got
    rng = np.random.default_rng(3)
    g = rng.normal(size=got["rgba"].shape).astype(np.float32)
    gd = None if dq is None else (rng.normal(size=dq.shape) * 1e-3).astype(np.float32)
    bwd = pipe.trace_backward(*scene, rays, start, got["rgba"], g, dq, got.get("depth_indices"), gd, **kw)
    rb = oracle.trace_backward(*scene, rays, start, np.asarray(ref["rgba"]), g, dq,
                               None if dq is None else np.asarray(ref["depth_indices"]), gd, **kw)
    for k in ("points_grad", "attr_grad"):
        a, b = bwd[k].astype(np.float64), np.asarray(rb[k], dtype=np.float64)
        assert np.array_equal(np.isfinite(a), np.isfinite(b)), k
        fin = np.isfinite(b)
        if fin.any():
            assert np.abs(a[fin] - b[fin]).max() <= 2e-5 * max(float(np.abs(b[fin]).max()), 1e-30), k
    return got


@pytest.fixture(scope="module")
def base():
    return common.config1(3, 2)


@pytest.mark.parametrize("max_intersections", [1, 2, 5])
def test_step_budget_exhaustion(base, max_intersections):
    got = compare(base.foam, base.rays, base.start, base.quantiles, max_intersections=max_intersections)
    assert int(got["num_intersections"].max()) == max_intersections + 1   # A.2: n = max_steps + 1 when the budget runs out


@pytest.mark.parametrize("weight_threshold", [0.0, 0.5, 0.999])
def test_weight_threshold_extremes(base, weight_threshold):
    compare(base.foam, base.rays, base.start, base.quantiles, weight_threshold=weight_threshold)


def test_ragged_image_and_batch_sizes(base):
    f = base.foam
    rays = foam.pinhole_rays(37, 13, (2.5, 2.5, 2.5), fov=0.9)          # not a multiple of the 16x8 CTA tile
    start = np.full((13, 37), foam.nearest_point(f.points, (2.5, 2.5, 2.5)), dtype=np.uint32)
    dq = np.sort(np.random.default_rng(1).uniform(0, 1, size=(13, 37, 3)).astype(np.float32), axis=-1)[..., ::-1].copy()
    compare(f, rays, start, dq)
    for n in (1, 31, 33, 129):
        compare(f, rays.reshape(-1, 6)[:n].copy(), start.reshape(-1)[:n].copy(), dq.reshape(-1, 3)[:n].copy())
    pipe = emu.EmuPipeline(3)
    out = pipe.trace_forward(f.points, f.attributes, f.adjacency, f.offsets, rays.reshape(-1, 6)[:0],
                             start.reshape(-1)[:0], None)
    assert out["rgba"].shape == (0, 4)                                   # empty batch: nothing launched


def test_density_extremes_and_quantile_corner_values(base):
    f = base.foam
    attrs = f.attributes.copy()
    attrs[::3, -1] = 0.0
    attrs[1::7, -1] = 1e-7          # below the 1e-6 colour gate (pipeline.cu:49-56)
    attrs[2::11, -1] = 1e4
    compare(f, base.rays, base.start, base.quantiles, attributes=attrs)
    dq = np.zeros(base.rays.shape[:-1] + (3,), np.float32)
    dq[..., 0], dq[..., 1] = 1.0, 0.5
    compare(f, base.rays, base.start, dq)                                # quantiles 1, 0.5, 0
    dq[...] = 0.5
    compare(f, base.rays, base.start, dq)                                # equal quantiles


def test_degenerate_rays_and_wrong_start_cells(base):
    f = base.foam
    rays = base.rays.copy()
    rays[0, 0, 3:] = 0.0                                                 # zero direction -> NaN after normalisation
    rays[0, 1, 3:] = [np.nan, 0, 1]
    rays[0, 2, 3:] = [np.inf, 0, 0]
    rays[0, 3, :3] = [1e30, 0, 0]
    rays[0, 4, 3:] = [1e-30, 0, 0]
    rays[0, 5, 3:] = [0, 0, 1e20]
    compare(f, rays, base.start, base.quantiles)
    start = np.random.default_rng(9).integers(0, f.num_points, size=base.start.shape).astype(np.uint32)
    compare(f, base.rays, start, base.quantiles)                         # start cell does not contain the origin


def test_far_points_overflowing_half_offsets(base):
    big = foam.small_foam(256, sh_degree=3, seed=0)
    pts = big.points.copy()
    pts[:8] *= 1e5                                                        # neighbour offsets beyond 65504 -> inf in fp16
    adj, off = foam.delaunay_adjacency(pts)
    f = foam.Foam(pts, big.attributes, adj, off, 3)
    start = np.full(base.start.shape, foam.nearest_point(pts, (2.5, 2.5, 2.5)), dtype=np.uint32)
    compare(f, base.rays, start, base.quantiles)


@pytest.mark.parametrize("sh_degree", [0, 1, 2])
def test_lower_sh_degrees_with_early_termination(sh_degree):
    case = common.config1(sh_degree, 2)
    compare(case.foam, case.rays, case.start, case.quantiles, weight_threshold=0.3)


def test_camera_inside_the_foam_without_quantiles(base):
    f = base.foam
    pos = (0.05, 0.02, -0.03)
    rays = foam.pinhole_rays(40, 24, pos, target=(1, 0.3, 0.2), fov=1.4)
    start = np.full((24, 40), foam.nearest_point(f.points, pos), dtype=np.uint32)
    compare(f, rays, start, None)


def test_tiny_direction_components_take_the_exact_twin(base):
    """Rays with a nonzero direction component below 2^-60 are outside the domain the ranked face scan is proven
    for (a front face's dp can be a positive subnormal, which MUFU.RCP flushes): the fast kernel raises the device
    flag and the exact twin re-does the launch -- plain forward, recording forward + replay, contribution
    (per-ray dispatch) and the benchmark kernel must all agree with the oracle, other rays of the batch included."""
    f = base.foam
    rays = base.rays.copy()
    rays[3, 5, 3:] = [1.0, 1e-30, 0.0]       # axis-aligned but for a denormal-sized component
    rays[9, 17, 3:] = [3e-25, -1.0, 2e-22]
    rays[20, 2, 3:] = [0.0, 0.0, 1.0]        # exactly axis-aligned: zeros are fine, stays on the fast path
    got = compare(f, rays, base.start, base.quantiles)
    scene = (f.points, f.attributes, f.adjacency, f.offsets)
    ref = oracle.trace_forward(*scene, rays, base.start, base.quantiles, return_contribution=True)
    pipe = emu.EmuPipeline(f.sh_degree)
    con = pipe.trace_forward(*scene, rays, base.start, base.quantiles, return_contribution=True)
    assert np.array_equal(con["num_intersections"].reshape(-1), np.asarray(ref["num_intersections"]).reshape(-1))
    assert common.grad_error(con["contribution"], np.asarray(ref["contribution"], dtype=np.float32)) <= 1e-5
    # recording forward (twice: the second one has a pool large enough) + replaying backward
    for _ in range(2):
        rec = pipe.trace_forward(*scene, rays, base.start, base.quantiles, scene_version=3, record_tape=True)
    assert not pipe.tape_status()["overflowed"]
    for k in ("num_intersections", "depth_indices"):
        assert np.array_equal(rec[k], got[k]), k
    assert np.array_equal(rec["rgba"].view(np.uint32), got["rgba"].view(np.uint32))
    g, gd = base.grad_rgba, base.grad_depth
    rb = oracle.trace_backward(*scene, rays, base.start, np.asarray(ref["rgba"]), g, base.quantiles,
                               np.asarray(ref["depth_indices"]), gd)
    bwd = pipe.trace_backward(*(None,) * 6, rec["rgba"], g, None, rec["depth_indices"], gd, scene_version=3,
                              use_tape=True)
    for k in ("points_grad", "attr_grad"):
        assert common.grad_error(bwd[k], np.asarray(rb[k])) <= 2e-5, k
