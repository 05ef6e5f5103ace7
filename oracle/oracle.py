"""ctypes wrapper of the CPU restatement oracle/radfoam_oracle.c (numpy in, numpy out).

TEST INFRASTRUCTURE ONLY -- see the header of radfoam_oracle.c.  Function names and
result keys follow the reference's Pipeline bindings
(torch_bindings/pipeline_bindings.cpp:107-585).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_float, c_int, c_uint32, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libradfoam_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "radfoam_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB)
        P = c_void_p
        lib.rfo_prefetch_adjacent_diff.argtypes = [P, c_uint32, P, P, P]
        lib.rfo_prefetch_adjacent_diff.restype = None
        lib.rfo_trace_forward.argtypes = [c_int, c_int, c_float, c_uint32, c_uint32, P, P, c_uint32, P,
                                          P, c_uint32, P, P, c_uint32, P, P, P, P, P, P, c_int]
        lib.rfo_trace_forward.restype = c_int
        lib.rfo_trace_backward.argtypes = [c_int, c_int, c_float, c_uint32, c_uint32, P, P, c_uint32,
                                           P, P, c_uint32, P, P, c_uint32, P, P, P, P, P, P, P, P, P,
                                           c_int]
        lib.rfo_trace_backward.restype = c_int
        lib.rfo_trace_benchmark.argtypes = [c_int, c_int, c_float, c_uint32, P, P, P, P, P, P, P, P,
                                            P, c_float, c_uint32, c_uint32, c_int, c_uint32, P, c_int]
        lib.rfo_trace_benchmark.restype = c_int
        lib.rfo_max_threads.restype = c_int
        lib.rfo_farthest_neighbor.argtypes = [P, c_uint32, P, P, P, P]
        lib.rfo_farthest_neighbor.restype = None
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def _attr_dtype(attributes):
    if attributes.dtype == np.float16:
        return np.float16, 1
    return np.float32, 0


def sh_degree_of(attr_dim: int) -> int:
    return {4: 0, 13: 1, 28: 2, 49: 3}[attr_dim]


def max_threads() -> int:
    return int(load().rfo_max_threads())


def prefetch_adjacent_diff(points, adjacency, offsets):
    points = _c(points, np.float32)
    adjacency = _c(adjacency, np.uint32)
    offsets = _c(offsets, np.uint32)
    out = np.zeros((adjacency.shape[0], 4), dtype=np.float16)
    load().rfo_prefetch_adjacent_diff(_p(points), points.shape[0], _p(adjacency), _p(offsets), _p(out))
    return out


def farthest_neighbor(points, adjacency, offsets):
    """(indices uint32[N], cell_radius float32[N]); triangulation_ops.cu:9-44."""
    points = _c(points, np.float32)
    adjacency = _c(adjacency, np.uint32)
    offsets = _c(offsets, np.uint32)
    n = points.shape[0]
    indices = np.zeros((n,), dtype=np.uint32)
    radius = np.zeros((n,), dtype=np.float32)
    load().rfo_farthest_neighbor(_p(points), n, _p(adjacency), _p(offsets), _p(indices), _p(radius))
    return indices, radius


def trace_forward(points, attributes, adjacency, offsets, rays, start_point, depth_quantiles=None,
                  weight_threshold=0.001, max_intersections=1024, return_contribution=False,
                  num_threads=1):
    adt, is_half = _attr_dtype(attributes)
    points = _c(points, np.float32)
    attributes = _c(attributes, adt)
    adjacency = _c(adjacency, np.uint32)
    offsets = _c(offsets, np.uint32)
    rays = _c(rays, np.float32)
    start = _c(start_point, np.uint32)
    batch = rays.shape[:-1]
    R = rays.size // 6
    n = points.shape[0]
    Q = 0
    dq = None
    if depth_quantiles is not None:
        dq = _c(depth_quantiles, np.float32)
        Q = dq.shape[-1]
    rgba = np.empty(batch + (4,), dtype=adt)
    nint = np.empty(batch + (1,), dtype=np.uint32)
    depth = np.zeros(batch + (Q,), dtype=np.float32) if dq is not None else None
    didx = np.zeros(batch + (Q,), dtype=np.uint32) if dq is not None else None
    contrib = np.zeros((n, 1), dtype=adt) if return_contribution else None
    rc = load().rfo_trace_forward(sh_degree_of(attributes.shape[-1]), is_half, weight_threshold,
                                  max_intersections, n, _p(points), _p(attributes), adjacency.shape[0],
                                  _p(adjacency), _p(offsets), R, _p(rays), _p(start), Q, _p(dq),
                                  _p(rgba), _p(depth), _p(didx), _p(nint), _p(contrib), num_threads)
    if rc:
        raise RuntimeError(f"oracle trace_forward failed ({rc})")
    out = {"rgba": rgba}
    if dq is not None:
        out["depth"] = depth
        out["depth_indices"] = didx
    if return_contribution:
        out["contribution"] = contrib
    out["num_intersections"] = nint
    return out


def trace_backward(points, attributes, adjacency, offsets, rays, start_point, rgb_out, grad_in,
                   depth_quantiles=None, depth_indices=None, depth_grad_in=None, ray_error=None,
                   weight_threshold=0.001, max_intersections=1024, num_threads=1):
    adt, is_half = _attr_dtype(attributes)
    points = _c(points, np.float32)
    attributes = _c(attributes, adt)
    adjacency = _c(adjacency, np.uint32)
    offsets = _c(offsets, np.uint32)
    rays = _c(rays, np.float32)
    start = _c(start_point, np.uint32)
    rgb_out = _c(rgb_out, adt)
    grad_in = _c(grad_in, adt)
    R = rays.size // 6
    n = points.shape[0]
    Q = 0
    dq = di = dg = None
    if depth_quantiles is not None:
        dq = _c(depth_quantiles, np.float32)
        di = _c(depth_indices, np.uint32)
        dg = _c(depth_grad_in, np.float32)
        Q = dq.shape[-1]
    err = _c(ray_error, adt)
    points_grad = np.zeros((n, 3), dtype=np.float32)
    attr_grad = np.zeros(attributes.shape, dtype=adt)
    point_error = np.zeros((n, 1), dtype=adt) if err is not None else None
    rc = load().rfo_trace_backward(sh_degree_of(attributes.shape[-1]), is_half, weight_threshold,
                                   max_intersections, n, _p(points), _p(attributes), adjacency.shape[0],
                                   _p(adjacency), _p(offsets), R, _p(rays), _p(start), Q, _p(dq), _p(di),
                                   _p(rgb_out), _p(grad_in), _p(dg), _p(err), _p(points_grad),
                                   _p(attr_grad), _p(point_error), num_threads)
    if rc:
        raise RuntimeError(f"oracle trace_backward failed ({rc})")
    out = {"points_grad": points_grad, "attr_grad": attr_grad}
    if err is not None:
        out["point_error"] = point_error
    return out


def trace_benchmark(points, attributes, adjacency, offsets, adjacent_diff, camera, start_point,
                    weight_threshold=0.001, max_intersections=1024, num_threads=1):
    adt, is_half = _attr_dtype(attributes)
    points = _c(points, np.float32)
    attributes = _c(attributes, adt)
    adjacency = _c(adjacency, np.uint32)
    offsets = _c(offsets, np.uint32)
    diff = _c(adjacent_diff, np.float16)
    w, h = int(camera["width"]), int(camera["height"])
    out = np.zeros((h, w), dtype=np.uint32)
    vec = {k: _c(np.asarray(camera[k]), np.float32) for k in ("position", "forward", "right", "up")}
    rc = load().rfo_trace_benchmark(sh_degree_of(attributes.shape[-1]), is_half, weight_threshold,
                                    max_intersections, _p(points), _p(attributes), _p(adjacency),
                                    _p(offsets), _p(diff), _p(vec["position"]), _p(vec["forward"]),
                                    _p(vec["right"]), _p(vec["up"]), float(camera["fov"]), w, h,
                                    0 if camera["model"] == "pinhole" else 1, int(start_point),
                                    _p(out), num_threads)
    if rc:
        raise RuntimeError(f"oracle trace_benchmark failed ({rc})")
    return out
