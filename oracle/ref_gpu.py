"""ctypes wrapper of oracle/_ref/libradfoam_ref.so: the reference's OWN CUDA kernels
(/root/reference/src/tracing/pipeline.cu compiled unmodified against oracle/eigen_shim,
see oracle/Makefile and oracle/ref_abi.cu), driven with torch CUDA tensors.

TEST INFRASTRUCTURE ONLY: the strongest available parity checker (same kernels the
upstream project ships, same libdevice expf/logf) and the baseline bench.py times
beside the product.  Host-side behaviour mirrors torch_bindings/pipeline_bindings.cpp
(output allocation incl. the zero-fills of :441-455, launch on the legacy default
stream).  /root/reference is not needed at run time, only the prebuilt .so.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_uint32, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libradfoam_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(LIB_PATH)
        P = c_void_p
        lib.rfref_last_error.restype = c_char_p
        lib.rfref_trace_forward.argtypes = [c_int, c_int, c_float, c_uint32, c_uint32, P, P, c_uint32,
                                            P, P, c_uint32, P, P, c_uint32, P, P, P, P, P, P]
        lib.rfref_trace_backward.argtypes = [c_int, c_int, c_float, c_uint32, c_uint32, P, P, c_uint32,
                                             P, P, c_uint32, P, P, c_uint32, P, P, P, P, P, P, P, P,
                                             P, P]
        lib.rfref_trace_benchmark.argtypes = [c_int, c_int, c_float, c_uint32, c_uint32, P, P, P, P, P,
                                              P, P, P, P, c_float, c_uint32, c_uint32, c_int, P, P]
        lib.rfref_prefetch_adjacent_diff.argtypes = [P, c_uint32, c_uint32, P, P, P]
        lib.rfref_release_pool.restype = None
        lib.rfref_farthest_neighbor.argtypes = [P, c_uint32, P, P, P, P]
        _lib = lib
    return _lib


def _check(rc):
    if rc:
        raise RuntimeError(load().rfref_last_error().decode())


def _ptr(t):
    return None if t is None else t.data_ptr()


def _deg(attr_dim):
    return {4: 0, 13: 1, 28: 2, 49: 3}[attr_dim]


def trace_forward(points, attributes, adjacency, offsets, rays, start_point, depth_quantiles=None,
                  weight_threshold=0.001, max_intersections=1024, return_contribution=False):
    lib = load()
    is_half = attributes.dtype == torch.float16
    dev = rays.device
    batch = list(rays.shape[:-1])
    R = rays.numel() // 6
    n = points.size(0)
    Q = depth_quantiles.size(-1) if depth_quantiles is not None else 0
    rgba = torch.empty(batch + [4], dtype=attributes.dtype, device=dev)
    nint = torch.empty(batch + [1], dtype=torch.uint32, device=dev)
    contrib = torch.zeros((n, 1), dtype=attributes.dtype, device=dev) if return_contribution else None
    depth = didx = None
    if depth_quantiles is not None:
        depth = torch.zeros(batch + [Q], dtype=torch.float32, device=dev)
        didx = torch.zeros(batch + [Q], dtype=torch.uint32, device=dev)
    _check(lib.rfref_trace_forward(_deg(attributes.size(-1)), int(is_half), weight_threshold,
                                   max_intersections, n, _ptr(points), _ptr(attributes),
                                   adjacency.numel(), _ptr(adjacency), _ptr(offsets), R, _ptr(rays),
                                   _ptr(start_point), Q, _ptr(depth_quantiles), _ptr(rgba), _ptr(depth),
                                   _ptr(didx), _ptr(nint), _ptr(contrib)))
    out = {"rgba": rgba}
    if depth_quantiles is not None:
        out["depth"] = depth
        out["depth_indices"] = didx
    if return_contribution:
        out["contribution"] = contrib
    out["num_intersections"] = nint
    return out


def trace_backward(points, attributes, adjacency, offsets, rays, start_point, rgb_out, grad_in,
                   depth_quantiles=None, depth_indices=None, depth_grad_in=None, ray_error=None,
                   weight_threshold=0.001, max_intersections=1024):
    lib = load()
    is_half = attributes.dtype == torch.float16
    dev = rays.device
    R = rays.numel() // 6
    n = points.size(0)
    Q = depth_quantiles.size(-1) if depth_quantiles is not None else 0
    attr_grad = torch.zeros_like(attributes)
    points_grad = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    ray_grad = torch.empty_like(rays)
    point_error = torch.zeros((n, 1), dtype=attributes.dtype, device=dev) if ray_error is not None else None
    _check(lib.rfref_trace_backward(_deg(attributes.size(-1)), int(is_half), weight_threshold,
                                    max_intersections, n, _ptr(points), _ptr(attributes),
                                    adjacency.numel(), _ptr(adjacency), _ptr(offsets), R, _ptr(rays),
                                    _ptr(start_point), Q, _ptr(depth_quantiles), _ptr(depth_indices),
                                    _ptr(rgb_out), _ptr(grad_in), _ptr(depth_grad_in), _ptr(ray_error),
                                    _ptr(ray_grad), _ptr(points_grad), _ptr(attr_grad),
                                    _ptr(point_error)))
    out = {"points_grad": points_grad, "attr_grad": attr_grad, "ray_grad": ray_grad}
    if ray_error is not None:
        out["point_error"] = point_error
    return out


def prefetch_adjacent_diff(points, adjacency, offsets):
    out = torch.zeros((adjacency.numel(), 4), dtype=torch.float16, device=points.device)
    _check(load().rfref_prefetch_adjacent_diff(_ptr(points), points.size(0), adjacency.numel(),
                                               _ptr(adjacency), _ptr(offsets), _ptr(out)))
    return out


def farthest_neighbor(points, adjacency, offsets):
    """The reference's own kernel (triangulation_ops.cu:9-44) -> (indices uint32[N], cell_radius f32[N])."""
    n = points.size(0)
    indices = torch.zeros((n,), dtype=torch.uint32, device=points.device)
    radius = torch.zeros((n,), dtype=torch.float32, device=points.device)
    _check(load().rfref_farthest_neighbor(_ptr(points), n, _ptr(adjacency), _ptr(offsets), _ptr(indices),
                                          _ptr(radius)))
    return indices, radius


def trace_benchmark(points, attributes, adjacency, offsets, adjacent_diff, camera, start_point,
                    output_rgba, weight_threshold=0.001, max_intersections=1024):
    lib = load()
    is_half = attributes.dtype == torch.float16
    vec = {k: (ctypes.c_float * 3)(*[float(v) for v in camera[k]])
           for k in ("position", "forward", "right", "up")}
    _check(lib.rfref_trace_benchmark(
        _deg(attributes.size(-1)), int(is_half), weight_threshold, max_intersections, points.size(0),
        _ptr(points), _ptr(attributes), _ptr(adjacency), _ptr(offsets), _ptr(adjacent_diff),
        ctypes.cast(vec["position"], c_void_p), ctypes.cast(vec["forward"], c_void_p),
        ctypes.cast(vec["right"], c_void_p), ctypes.cast(vec["up"], c_void_p), float(camera["fov"]),
        int(camera["width"]), int(camera["height"]), 0 if camera["model"] == "pinhole" else 1,
        _ptr(start_point), _ptr(output_rgba)))
