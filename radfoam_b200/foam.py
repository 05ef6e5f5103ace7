"""Synthetic Voronoi foams for tests and benchmarks (numpy / scipy only, no CUDA).

There is no dataset or checkpoint on the box, so every scene the
tests and bench.py trace is generated here, deterministically from a seed, in the
exact array formats radfoam's tracer consumes:

  points        [N, 3]   float32
  attributes    [N, A]   float32/float16, A = 1 + 3*(deg+1)^2: SH coefficients
                         interleaved RGB (attr[3k+c]) then density LAST
                         (radfoam_model/scene.py:202-217, src/tracing/sh_utils.cuh:78-80)
  adjacency     [E]      uint32 CSR neighbour lists, every row ASCENDING
                         (src/delaunay/delaunay.cu:146-226)
  offsets       [N + 1]  uint32

The adjacency comes from scipy's Qhull Delaunay (unique for points in general
position, so it is the edge set radfoam's GPU Delaunay would produce); only the
row order has to be normalised.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np


def attr_dim(sh_degree: int) -> int:
    return 1 + 3 * (sh_degree + 1) ** 2


def softplus_beta10(x: np.ndarray) -> np.ndarray:
    """F.softplus(x, beta=10) (radfoam_model/scene.py:202-203), overflow-safe."""
    y = 10.0 * x
    return (np.logaddexp(0.0, y) / 10.0).astype(np.float32)


def delaunay_adjacency(points: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """CSR point adjacency of the Delaunay triangulation, rows ascending, uint32."""
    from scipy.spatial import Delaunay

    tri = Delaunay(points.astype(np.float64))
    indptr, indices = tri.vertex_neighbor_vertices
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    n = points.shape[0]
    if indptr.shape[0] != n + 1:
        raise RuntimeError("Qhull dropped points (duplicates / degenerate input)")
    # sort every row ascending with one global lexsort on (row, neighbour)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    order = np.lexsort((indices, rows))
    indices = indices[order]
    return indices.astype(np.uint32), indptr.astype(np.uint32)


def morton_codes(points: np.ndarray, bits: int = 21) -> np.ndarray:
    """63-bit Morton code of each point inside the bounding box of the set."""
    p = points.astype(np.float64)
    lo, hi = p.min(axis=0), p.max(axis=0)
    q = ((p - lo) / np.maximum(hi - lo, 1e-30) * ((1 << bits) - 1)).astype(np.uint64)

    def spread(v: np.ndarray) -> np.ndarray:  # insert two zero bits between the low 21 bits
        v = v & np.uint64(0x1FFFFF)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v

    return spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))


@dataclasses.dataclass
class Foam:
    points: np.ndarray       # [N,3] f32
    attributes: np.ndarray   # [N,A] f32
    adjacency: np.ndarray    # [E] u32
    offsets: np.ndarray      # [N+1] u32
    sh_degree: int

    @property
    def num_points(self) -> int:
        return int(self.points.shape[0])


def make_attributes(rng: np.random.Generator, n: int, sh_degree: int, density: np.ndarray,
                    dc_scale: float = 1.0, sh_sigma: float = 0.1) -> np.ndarray:
    a = attr_dim(sh_degree)
    attrs = np.empty((n, a), dtype=np.float32)
    attrs[:, 0:3] = rng.uniform(-dc_scale, dc_scale, size=(n, 3))
    if a > 4:
        attrs[:, 3:a - 1] = rng.normal(0.0, sh_sigma, size=(n, a - 4))
    attrs[:, a - 1] = density
    return attrs


def small_foam(num_points: int = 256, sh_degree: int = 3, seed: int = 0) -> Foam:
    """BASELINE config 1: points ~U[-1,1]^3, SH ~N(0,0.3^2),
    density = softplus_beta10(N(0,1))."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-1.0, 1.0, size=(num_points, 3)).astype(np.float32)
    adj, off = delaunay_adjacency(pts)
    dens = softplus_beta10(rng.normal(0.0, 1.0, size=num_points))
    a = attr_dim(sh_degree)
    attrs = np.empty((num_points, a), dtype=np.float32)
    attrs[:, :a - 1] = rng.normal(0.0, 0.3, size=(num_points, a - 1))
    attrs[:, a - 1] = dens
    return Foam(pts, attrs, adj, off, sh_degree)


def scene_foam(num_points: int, sh_degree: int = 3, seed: int | None = None, adjacency=None) -> Foam:
    """Configs 2-5 recipe: 70% of the points on a unit-sphere shell
    (radial noise sigma 0.01, dense), 25% ~U[-1.5,1.5]^3, 5% far field ~N(0,8^2)
    (both nearly empty), so rays from outside mostly terminate at the shell after
    O(10^2) cells and some run to the hull.
    ``adjacency`` = ``(adjacency, offsets)`` of an earlier build of the SAME foam (same arguments) skips the
    Delaunay triangulation -- the only slow part; everything else is regenerated from the seed."""
    seed = num_points if seed is None else seed
    rng = np.random.default_rng(seed)
    n_surf = int(0.70 * num_points)
    n_box = int(0.25 * num_points)
    n_far = num_points - n_surf - n_box
    u = rng.normal(size=(n_surf, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    surf = u * (1.0 + rng.normal(0.0, 0.01, size=(n_surf, 1)))
    box = rng.uniform(-1.5, 1.5, size=(n_box, 3))
    far = rng.normal(0.0, 8.0, size=(n_far, 3))
    pts = np.concatenate([surf, box, far], axis=0).astype(np.float32)
    pts = np.unique(pts, axis=0)
    # Morton (Z-curve) order: spatially coherent memory layout like the kd-order
    # the reference's sort_points() leaves points in (aabb_tree.cu:62-190), so the
    # cells a ray crosses sit near each other in HBM as they do in a real scene.
    pts = pts[np.argsort(morton_codes(pts), kind="stable")]
    # re-derive which points are "surface" from the radius so density matches position
    n = pts.shape[0]
    radius = np.linalg.norm(pts.astype(np.float64), axis=1)
    is_surf = np.abs(radius - 1.0) < 0.05
    dens = np.where(is_surf,
                    10.0 * softplus_beta10(rng.normal(1.0, 1.0, size=n)),
                    softplus_beta10(rng.normal(-1.0, 0.5, size=n))).astype(np.float32)
    attrs = make_attributes(rng, n, sh_degree, dens)
    if adjacency is not None:
        adj, off = (np.ascontiguousarray(a, dtype=np.uint32) for a in adjacency)
        if off.shape[0] != n + 1 or int(off[-1]) != adj.shape[0]:
            raise RuntimeError("the given adjacency does not belong to this foam")
    else:
        adj, off = delaunay_adjacency(pts)
    return Foam(pts, attrs, adj, off, sh_degree)


def pack_adjacency(adjacency: np.ndarray, offsets: np.ndarray) -> dict:
    """CSR adjacency as arrays that compress well (rows are ascending and points Morton-ordered, so neighbour ids
    sit near the row index): per-row first entry relative to the row, then in-row deltas; row lengths as bytes."""
    counts = np.diff(offsets.astype(np.int64))
    rows = np.repeat(np.arange(counts.shape[0], dtype=np.int64), counts)
    a = adjacency.astype(np.int64)
    delta = np.empty_like(a)
    delta[1:] = a[1:] - a[:-1]
    first = offsets[:-1].astype(np.int64)[counts > 0]
    delta[first] = a[first] - rows[first]
    if counts.max(initial=0) > 255 or np.abs(delta).max(initial=0) >= 2 ** 31:
        raise RuntimeError("adjacency not packable")
    return {"counts": counts.astype(np.uint8), "delta": delta.astype(np.int32)}


def unpack_adjacency(counts: np.ndarray, delta: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    counts = counts.astype(np.int64)
    offsets = np.zeros(counts.shape[0] + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    total = np.cumsum(delta.astype(np.int64))
    nonempty = counts > 0
    first = offsets[:-1][nonempty]
    rows = np.arange(counts.shape[0], dtype=np.int64)[nonempty]
    # within a row: value = row + (running sum - running sum before the row's first entry)
    before = total[first] - delta.astype(np.int64)[first]
    adjacency = total - np.repeat(before - rows, counts[nonempty])
    return adjacency.astype(np.uint32), offsets.astype(np.uint32)


def pinhole_rays(width: int, height: int, position, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0),
                 fov: float = 0.9) -> np.ndarray:
    """[H, W, 6] float32 rays (origin, unit direction) through pixel centres, built
    like data_loader/colmap.py:10-20, 93-100 (pixel + 0.5, normalised)."""
    pos = np.asarray(position, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - pos
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, dtype=np.float64))
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    f = 0.5 * height / math.tan(0.5 * fov)
    xs = (np.arange(width, dtype=np.float64) + 0.5 - 0.5 * width) / f
    ys = (np.arange(height, dtype=np.float64) + 0.5 - 0.5 * height) / f
    gx, gy = np.meshgrid(xs, ys)
    d = fwd[None, None, :] + gx[..., None] * right[None, None, :] - gy[..., None] * upv[None, None, :]
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rays = np.empty((height, width, 6), dtype=np.float32)
    rays[..., :3] = pos.astype(np.float32)
    rays[..., 3:] = d.astype(np.float32)
    return rays


def nearest_point(points: np.ndarray, query) -> int:
    """Entry cell of a camera: brute-force nearest point (stands in for radfoam.nn,
    src/aabb_tree/aabb_tree.cu:391-415, which is out of scope here)."""
    q = np.asarray(query, dtype=np.float64)[None, :]
    d2 = ((points.astype(np.float64) - q) ** 2).sum(axis=1)
    return int(np.argmin(d2))


def camera_dict(position, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0), fov: float = 0.9,
                width: int = 64, height: int = 48, model: str = "pinhole") -> dict:
    """Camera in the dict form Pipeline.trace_benchmark takes
    (torch_bindings/pipeline_bindings.cpp:526-547), numpy float32 vectors."""
    pos = np.asarray(position, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - pos
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, dtype=np.float64))
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    return {
        "position": pos.astype(np.float32), "forward": fwd.astype(np.float32),
        "right": right.astype(np.float32), "up": upv.astype(np.float32),
        "fov": float(fov), "width": int(width), "height": int(height), "model": model,
    }
