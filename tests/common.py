"""Shared seeded test cases (numpy) and comparison helpers."""
from __future__ import annotations

import functools

import numpy as np

from radfoam_b200 import foam

NONE = 0xFFFFFFFF


class Case:
    """One traced scene: foam + rays + start cells + quantiles + upstream gradients."""

    def __init__(self, f: foam.Foam, rays, start, quantiles, seed=1):
        self.foam = f
        self.rays = rays
        self.start = start
        self.quantiles = quantiles
        rng = np.random.default_rng(seed)
        batch = rays.shape[:-1]
        self.grad_rgba = rng.normal(size=batch + (4,)).astype(np.float32)
        q = 0 if quantiles is None else quantiles.shape[-1]
        self.grad_depth = (rng.normal(size=batch + (q,)) * 1e-4).astype(np.float32) if q else None


def _quantiles(batch, q, seed):
    if q == 0:
        return None
    rng = np.random.default_rng(seed)
    # training passes sorted-descending uniform quantiles (train.py:176-180)
    return np.sort(rng.uniform(0.05, 0.95, size=batch + (q,)).astype(np.float32), axis=-1)[..., ::-1].copy()


@functools.lru_cache(maxsize=None)
def config1(sh_degree: int = 3, q: int = 2, fixed_quantiles: bool = True) -> Case:
    """BASELINE config 1: 256-point foam, 32x32 pinhole camera at (0,0,-3), fov 0.7,
    quantiles (0.7, 0.3)."""
    f = foam.small_foam(256, sh_degree=sh_degree, seed=0)
    rays = foam.pinhole_rays(32, 32, (0, 0, -3), fov=0.7, up=(0, 1, 0))
    start = np.full((32, 32), foam.nearest_point(f.points, (0, 0, -3)), dtype=np.uint32)
    if q == 0:
        dq = None
    elif fixed_quantiles and q == 2:
        dq = np.tile(np.array([0.7, 0.3], dtype=np.float32), (32, 32, 1))
    else:
        dq = _quantiles((32, 32), q, 7)
    return Case(f, rays, start, dq)


@functools.lru_cache(maxsize=None)
def scene_case(num_points: int = 20000, width: int = 160, height: int = 96, q: int = 2,
               sh_degree: int = 3, inside: bool = False) -> Case:
    """Mid-size shell scene (SURVEY.md §8d recipe) seen from outside (or inside) the shell."""
    f = foam.scene_foam(num_points, sh_degree=sh_degree)
    pos = (0.3, 0.3, 0.3) if inside else (2.5, 2.5, 2.5)
    target = (1.0, 0.2, -0.1) if inside else (0.0, 0.0, 0.0)
    rays = foam.pinhole_rays(width, height, pos, target=target, fov=0.9)
    start = np.full((height, width), foam.nearest_point(f.points, pos), dtype=np.uint32)
    return Case(f, rays, start, _quantiles((height, width), q, 11))


@functools.lru_cache(maxsize=None)
def random_ray_case(num_points: int = 20000, num_rays: int = 20000, cameras: int = 16, q: int = 2) -> Case:
    """Unordered ray batch from several cameras, per-ray start cells: the access pattern of
    the reference's training batches (train.py:61, scene.py:224-234)."""
    f = foam.scene_foam(num_points)
    rng = np.random.default_rng(5)
    cams = rng.normal(size=(cameras, 3))
    cams = 2.5 * cams / np.linalg.norm(cams, axis=1, keepdims=True)
    starts = np.array([foam.nearest_point(f.points, c) for c in cams], dtype=np.uint32)
    which = rng.integers(0, cameras, size=num_rays)
    target = rng.normal(0.0, 0.4, size=(num_rays, 3))
    d = target - cams[which]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([cams[which], d], axis=1).astype(np.float32)
    return Case(f, rays, starts[which].copy(), _quantiles((num_rays,), q, 13))


def grad_error(got: np.ndarray, ref: np.ndarray) -> float:
    """max |got - ref| / max |ref| over entries finite in the reference (the reference zeroes
    non-finite gradient entries afterwards, render.py:98-99).  Gradients are float scatter-adds
    whose order differs run to run in the reference itself, hence a norm-scaled figure."""
    ref = ref.astype(np.float64)
    got = got.astype(np.float64)
    ok = np.isfinite(ref) & np.isfinite(got)
    scale = max(float(np.abs(ref[ok]).max()) if ok.any() else 0.0, 1e-30)
    return float(np.abs(got[ok] - ref[ok]).max() / scale) if ok.any() else 0.0


def nonfinite_mismatch(got: np.ndarray, ref: np.ndarray) -> int:
    return int((np.isfinite(got) != np.isfinite(ref)).sum())


# ---- comparison bars -----------------------------------------------------------------------
# Against the reference's own kernels (same libdevice expf/logf) the bar is the north star's:
# integers bit-exact, floats 1e-5, gradients 1e-5 of max|ref|.  Against the CPU restatement the
# integer traversal is still bit-exact, but composited floats go through libm's expf/logf
# instead of the GPU's MUFU-based ones (a few 1e-7 relative apart), and two formulas of the
# reference amplify that: alpha = 1 - exp(-x) loses x's relative accuracy for small x, and the
# quantile depth t0 + log(T/q)/sigma divides an O(1e-7) difference in T by the cell's density.
# The CPU-side bars below state that conditioning explicitly instead of hiding it in a loose
# global tolerance.
CPU_GRAD_TOL = 3e-4


def depth_tolerance(attributes: np.ndarray, depth: np.ndarray, depth_indices: np.ndarray) -> np.ndarray:
    """Per-entry bound for |depth - ref| when expf/logf implementations differ:
    1e-5 * max(1, |depth|) + 2e-6 / sigma(cell that crossed the quantile)."""
    idx = depth_indices.astype(np.int64)
    valid = depth_indices != NONE
    sigma = np.ones(depth.shape, dtype=np.float64)
    sigma[valid] = np.maximum(attributes[idx[valid], -1].astype(np.float64), 1e-30)
    return 1e-5 * np.maximum(1.0, np.abs(depth)) + np.where(valid, 2e-6 / sigma, 0.0)


def assert_forward_close_cpu(got: dict, ref: dict, attributes: np.ndarray) -> None:
    """CUDA path (or golden reference output) vs the CPU restatement."""
    assert np.array_equal(got["num_intersections"], ref["num_intersections"]), "num_intersections"
    np.testing.assert_allclose(got["rgba"].astype(np.float32), ref["rgba"].astype(np.float32),
                               rtol=1e-5, atol=1e-5)
    if "depth_indices" in ref:
        assert np.array_equal(got["depth_indices"], ref["depth_indices"]), "depth_indices"
        tol = depth_tolerance(attributes, ref["depth"], ref["depth_indices"])
        bad = np.abs(got["depth"].astype(np.float64) - ref["depth"]) > tol
        assert not bad.any(), f"{int(bad.sum())} depth entries outside the conditioned bound"
    if "contribution" in ref and "contribution" in got:
        # a float scatter-add like the gradients: norm-scaled
        assert grad_error(got["contribution"].astype(np.float32),
                          ref["contribution"].astype(np.float32)) <= 1e-5


def farthest_edge_case():
    """A hand-built CSR (not a triangulation) for farthest_neighbor's corner cases: an empty row, a row whose
    neighbours all coincide with the point, an exact tie (first wins), rows of 9 / 21 / 40 faces (more than one
    8-lane pass), overflowing / underflowing / NaN coordinates, and rows mixing magnitudes 1e-12..1e12 so that the
    fp64-add-then-round-to-fp32 accumulation is exercised."""
    from radfoam_b200 import foam

    rng = np.random.default_rng(77)
    n = 96
    pts = rng.uniform(-1.0, 1.0, size=(n, 3)).astype(np.float32)
    pts[1] = pts[2] = pts[3] = pts[0]                     # row 0: all neighbours coincide
    pts[4] = np.float32([0.25, 0.5, -0.75])
    pts[5] = pts[4] + np.float32([0.5, 0.0, 0.0])         # row 4: exact tie between 5 and 6
    pts[6] = pts[4] - np.float32([0.5, 0.0, 0.0])
    pts[7] = pts[4] + np.float32([0.25, 0.0, 0.0])
    pts[8] = np.float32([3e19, 0.0, 0.0])                 # |d|^2 overflows -> inf
    pts[9] = np.float32([1e-30, 1e-30, 0.0])              # |d|^2 underflows
    pts[10] = np.float32([0.0, 0.0, 0.0])
    pts[11] = np.float32([np.nan, 0.0, 0.0])
    scale = np.float32(10.0) ** rng.integers(-12, 13, size=(n - 40, 1)).astype(np.float32)
    pts[40:] *= scale
    rows = [[] for _ in range(n)]
    rows[0] = [1, 2, 3]
    rows[4] = [7, 5, 6]
    rows[8] = [10, 12, 13]
    rows[10] = [9, 9]
    rows[12] = [8, 11, 13]                                 # inf and NaN distances in one row
    rows[11] = [10, 12]                                    # every distance NaN
    rows[13] = []                                          # empty row -> UINT32_MAX, NaN
    for i, k in ((14, 8), (15, 9), (16, 21), (17, 40), (18, 1), (19, 16), (20, 17)):
        rows[i] = list(rng.choice(np.arange(20, 40), size=k, replace=k > 20))
    for i in range(40, n):
        rows[i] = list(rng.choice(np.arange(40, n), size=int(rng.integers(3, 30)), replace=False))
    off = np.zeros(n + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(r) for r in rows])
    adj = np.array([j for r in rows for j in r], dtype=np.uint32)
    return foam.Foam(pts, np.zeros((n, 4), dtype=np.float32), adj, off, 0)


def farthest_neighbor_numpy(points, adjacency, offsets):
    """Independent restatement of triangulation_ops.cu:9-44 in numpy scalars (slow; small cases only)."""
    n = points.shape[0]
    idx = np.full(n, NONE, dtype=np.uint32)
    radius = np.zeros(n, dtype=np.float32)
    with np.errstate(all="ignore"):
        for i in range(n):
            p = points[i]
            s, m = np.float32(0.0), np.float32(0.0)
            b, e = int(offsets[i]), int(offsets[i + 1])
            for f in range(b, e):
                d = points[adjacency[f]] - p
                # fma(dx,dx, fma(dy,dy, dz*dz)) with a single rounding per fma: exact in float64 then rounded
                inner = np.float32(np.float64(d[1]) * np.float64(d[1]) + np.float64(np.float32(d[2] * d[2])))
                sq = np.float32(np.float64(d[0]) * np.float64(d[0]) + np.float64(inner))
                dist = np.sqrt(sq, dtype=np.float32)
                s = np.float32(np.float64(s) + 0.5 * np.float64(dist))
                if dist > m:
                    m, idx[i] = dist, adjacency[f]
            radius[i] = s / np.float32(e - b) if e > b else np.float32(np.nan)
    return idx, radius


def assert_same_floats(got: np.ndarray, ref: np.ndarray) -> None:
    """Bit-for-bit equality of two float32 arrays, any NaN matching any NaN (payloads differ between libm/GPU)."""
    got, ref = np.asarray(got, dtype=np.float32), np.asarray(ref, dtype=np.float32)
    assert got.shape == ref.shape
    nan = np.isnan(ref)
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got[~nan].view(np.uint32), ref[~nan].view(np.uint32))
