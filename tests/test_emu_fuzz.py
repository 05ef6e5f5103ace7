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

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fuzz_cases  # noqa: E402
import test_emu_edge_cases as edge  # noqa: E402


def run_case(seed: int):
    scene_kind, ray_kind, f, rays, start, dq, kw = fuzz_cases.make_case(seed)
    edge.compare(f, rays, start, dq, **kw)
    return scene_kind, ray_kind


@pytest.mark.parametrize("first_seed", [0, 100, 200])
def test_tie_heavy_cases_match_the_oracle(first_seed):
    seen = set()
    for seed in range(first_seed, first_seed + 100):
        seen.add(run_case(seed))
    assert len({s for s, _ in seen}) == 5 and len({r for _, r in seen}) == 5  # every generator was exercised


if __name__ == "__main__":
    first, count = int(sys.argv[1]), int(sys.argv[2])
    bad = 0
    for seed in range(first, first + count):
        try:
            run_case(seed)
        except AssertionError as e:
            bad += 1
            print("FAIL seed", seed, str(e)[:200].replace("\n", " "))
    print("cases", count, "from seed", first, "failures", bad)
