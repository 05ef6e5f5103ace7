"""Property tests (hypothesis) of the path's size-independent invariants on the CPU restatement:
random small foams, cameras, quantiles and settings.  The CUDA path is held to the same properties
at full size in tests/test_gpu_fullsize.py."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import common
from oracle import oracle
from radfoam_b200 import foam


def make_case(seed, n_points, n_rays, q, deg):
    rng = np.random.default_rng(seed)
    f = foam.small_foam(n_points, sh_degree=deg, seed=seed)
    cam = rng.normal(size=3)
    cam = 3.0 * cam / np.linalg.norm(cam)
    target = rng.uniform(-0.5, 0.5, size=(n_rays, 3))
    d = target - cam
    rays = np.concatenate([np.tile(cam, (n_rays, 1)), d], axis=1).astype(np.float32)  # un-normalised on purpose
    start = np.full(n_rays, foam.nearest_point(f.points, cam), dtype=np.uint32)
    dq = None
    if q:
        dq = np.sort(rng.uniform(0.02, 0.98, size=(n_rays, q)).astype(np.float32), axis=-1)[:, ::-1].copy()
    g = rng.normal(size=(n_rays, 4)).astype(np.float32)
    gd = (rng.normal(size=(n_rays, q)) * 1e-3).astype(np.float32) if q else None
    return f, rays, start, dq, g, gd


CASE = dict(seed=st.integers(0, 10_000), n_points=st.integers(24, 96), n_rays=st.integers(1, 40),
            q=st.sampled_from([0, 1, 2, 3]), deg=st.sampled_from([0, 1, 2, 3]))
SETTINGS = settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))


@SETTINGS
@given(**CASE, thr=st.sampled_from([0.0, 1e-3, 0.05, 0.5]), max_steps=st.sampled_from([1, 3, 1024]))
def test_forward_invariants(seed, n_points, n_rays, q, deg, thr, max_steps):
    f, rays, start, dq, _, _ = make_case(seed, n_points, n_rays, q, deg)
    scene = (f.points, f.attributes, f.adjacency, f.offsets)
    out = oracle.trace_forward(*scene, rays, start, dq, thr, max_steps, return_contribution=True)
    rgba, n = out["rgba"], out["num_intersections"].astype(np.int64)
    assert np.isfinite(rgba).all() and (rgba[:, :3] >= 0).all()
    assert (rgba[:, 3] >= -1e-7).all() and (rgba[:, 3] <= 1 + 1e-7).all()
    assert (n >= 1).all() and (n <= max_steps + 1).all()
    if q:
        idx = out["depth_indices"]
        valid = idx != common.NONE
        assert (idx[valid] < f.num_points).all() and (out["depth"][~valid] == -1).all()
        assert (np.diff(out["depth"], axis=1)[valid[:, 1:] & valid[:, :-1]] >= -1e-5).all()  # later quantile, deeper
        assert (np.diff(valid.astype(int), axis=1) <= 0).all()        # a missed quantile stays missed
    # scaling the (un-normalised) directions changes nothing: the kernel re-normalises them
    scaled = rays.copy()
    scaled[:, 3:] *= 2.0  # exact in binary floating point
    again = oracle.trace_forward(*scene, scaled, start, dq, thr, max_steps)
    assert np.array_equal(again["num_intersections"], out["num_intersections"])
    np.testing.assert_allclose(again["rgba"], rgba, rtol=1e-5, atol=1e-6)
    # rays are independent: any permutation of the batch permutes the outputs
    perm = np.random.default_rng(seed).permutation(n_rays)
    shuffled = oracle.trace_forward(*scene, rays[perm], start[perm], None if dq is None else dq[perm], thr, max_steps)
    assert np.array_equal(shuffled["rgba"], rgba[perm])
    # a lower budget truncates the walk and nothing else
    if max_steps > 1:
        short = oracle.trace_forward(*scene, rays, start, None, thr, 1)
        assert np.array_equal(np.minimum(n, 2), short["num_intersections"].astype(np.int64))
    # total contribution == total opacity
    np.testing.assert_allclose(out["contribution"].sum(dtype=np.float64), rgba[:, 3].sum(dtype=np.float64),
                               rtol=1e-4, atol=1e-5)


@SETTINGS
@given(**CASE)
def test_backward_is_linear_and_additive_over_rays(seed, n_points, n_rays, q, deg):
    f, rays, start, dq, g, gd = make_case(seed, n_points, n_rays, q, deg)
    scene = (f.points, f.attributes, f.adjacency, f.offsets)
    fwd = oracle.trace_forward(*scene, rays, start, dq)

    def bwd(sel, scale=1.0):
        return oracle.trace_backward(*scene, rays[sel], start[sel], fwd["rgba"][sel], g[sel] * scale,
                                     None if dq is None else dq[sel],
                                     fwd["depth_indices"][sel] if q else None, gd[sel] * scale if q else None)

    everything = np.arange(n_rays)
    full = bwd(everything)
    doubled = bwd(everything, 2.0)
    half = n_rays // 2
    a, b = bwd(everything[:half]), bwd(everything[half:])
    for k in ("points_grad", "attr_grad"):
        finite = np.isfinite(full[k])
        assert (np.isfinite(doubled[k]) == finite).all()
        assert common.grad_error(doubled[k], 2.0 * full[k]) < 1e-5
        if half:
            both = finite & np.isfinite(a[k]) & np.isfinite(b[k])
            assert common.grad_error(np.where(both, a[k] + b[k], 0), np.where(both, full[k], 0)) < 1e-4
    assert full["attr_grad"].shape == f.attributes.shape and full["points_grad"].shape == f.points.shape
