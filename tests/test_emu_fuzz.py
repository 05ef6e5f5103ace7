"""Tie-heavy differential cases for the tracing path: kernel logic (CPU emulator, tests/emu) against the oracle on
geometry built to produce equal face distances and degenerate denominators (tests/fuzz_cases.py): exact and jittered
cubic / BCC lattices, tight clusters, scenes scaled by 1e-3 / 1e3, crossed by axis-aligned rays, rays through cell
sites and edge midpoints, rays with zero / denormal-scale direction components, and cameras inside the foam; random
SH degree, quantile count, step budget and weight threshold.  The bar is tests/test_emu_edge_cases.compare (integers exact,
floats 1e-5, gradients 2e-5 of max, identical non-finite patterns).

`python tests/test_emu_fuzz.py FIRST_SEED COUNT` runs a longer campaign (8000 cases, seeds 1000-8999, ran clean at the
end of round 2: DESIGN.md section 2)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fuzz_cases  # noqa: E402
from oracle import oracle  # noqa: E402
import test_emu_edge_cases as edge  # noqa: E402
from test_emu_edge_cases import emu  # noqa: E402


def run_case(seed: int):
    scene_kind, ray_kind, f, rays, start, dq, kw = fuzz_cases.make_case(seed)
    edge.compare(f, rays, start, dq, **kw)
    return scene_kind, ray_kind


def run_case_as_image(seed: int):
    """The same case traced as an image (8x4 warp tiles over a ragged h x w grid), with the walk tape recorded and
    replayed, the per-point contribution and the ray-error scatter on: all of it against the oracle's flat trace."""
    _, _, f, rays, start, dq, kw = fuzz_cases.make_case(seed)
    m = rays.shape[0]
    w = next((c for c in (11, 10, 7) if m % c == 0), 1)
    shape = (m // w, w)
    full = dict(weight_threshold=0.001, max_intersections=1024)
    full.update(kw)
    scene = (f.points, f.attributes, f.adjacency, f.offsets)
    rng = np.random.default_rng(seed)
    g = rng.normal(size=(m, 4)).astype(np.float32)
    gd = None if dq is None else (rng.normal(size=dq.shape) * 1e-3).astype(np.float32)
    err = rng.uniform(0.0, 1.0, size=(m, 1)).astype(np.float32)

    ref = oracle.trace_forward(*scene, rays, start, dq, full["weight_threshold"], full["max_intersections"], True)
    rb = oracle.trace_backward(*scene, rays, start, np.asarray(ref["rgba"]), g, dq,
                               None if dq is None else np.asarray(ref["depth_indices"]), gd, err,
                               full["weight_threshold"], full["max_intersections"])
    pipe = emu.EmuPipeline(f.sh_degree)
    img = lambda a, k: None if a is None else a.reshape(shape + (k,))  # noqa: E731
    got = pipe.trace_forward(*scene, img(rays, 6), start.reshape(shape), img(dq, 0 if dq is None else dq.shape[1]),
                             return_contribution=True, scene_version=7, record_tape=True, **kw)
    bwd = pipe.trace_backward(*scene, img(rays, 6), start.reshape(shape), got["rgba"], img(g, 4),
                              img(dq, 0 if dq is None else dq.shape[1]), got.get("depth_indices"),
                              img(gd, 0 if gd is None else gd.shape[1]), img(err, 1), scene_version=7, use_tape=True,
                              **kw)
    assert np.array_equal(got["num_intersections"].reshape(-1), np.asarray(ref["num_intersections"]).reshape(-1))
    if dq is not None:
        assert np.array_equal(got["depth_indices"].reshape(-1), np.asarray(ref["depth_indices"]).reshape(-1))
    for k, a, b in (("rgba", got["rgba"], ref["rgba"]), ("contribution", got["contribution"], ref["contribution"]),
                    ("point_error", bwd["point_error"], rb["point_error"]),
                    ("points_grad", bwd["points_grad"], rb["points_grad"]), ("attr_grad", bwd["attr_grad"], rb["attr_grad"])):
        a, b = np.asarray(a, np.float64).reshape(-1), np.asarray(b, np.float64).reshape(-1)
        assert np.array_equal(np.isfinite(a), np.isfinite(b)), k
        fin = np.isfinite(b)
        if fin.any():
            assert np.abs(a[fin] - b[fin]).max() <= 2e-5 * max(float(np.abs(b[fin]).max()), 1e-30), k


@pytest.mark.parametrize("first_seed", [0, 100, 200])
def test_tie_heavy_cases_match_the_oracle(first_seed):
    seen = set()
    for seed in range(first_seed, first_seed + 100):
        seen.add(run_case(seed))
    assert len({s for s, _ in seen}) == 5 and len({r for _, r in seen}) == 5  # every generator was exercised


def test_tie_heavy_cases_as_images_with_tape_contribution_and_error_map():
    for seed in range(300, 400):
        run_case_as_image(seed)


if __name__ == "__main__":
    first, count = int(sys.argv[1]), int(sys.argv[2])
    bad = 0
    for seed in range(first, first + count):
        try:
            run_case(seed)
            run_case_as_image(seed)
        except AssertionError as e:
            bad += 1
            print("FAIL seed", seed, str(e)[:200].replace("\n", " "))
    print("cases", count, "from seed", first, "failures", bad)
