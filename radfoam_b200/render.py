"""The render autograd op: mirror of radfoam_model/render.py::TraceRays (lines 10-122).

Same forward/backward contract (inputs, outputs, the ErrorBox side channel), so
``RadFoamScene.forward`` (radfoam_model/scene.py:236-261) can call it unchanged.  The
one difference is internal: the non-finite gradient scrub the reference does with two
boolean-mask passes after the kernel (render.py:98-99) is folded into the backward
kernel's epilogue.
"""
from __future__ import annotations

import torch


class ErrorBox:
    def __init__(self):
        self.ray_error = None
        self.point_error = None


class TraceRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pipeline, _points, _attributes, _point_adjacency, _point_adjacency_offsets,
                rays, start_point, depth_quantiles, return_contribution):
        ctx.rays = rays
        ctx.start_point = start_point
        ctx.depth_quantiles = depth_quantiles
        ctx.pipeline = pipeline
        ctx.points = _points
        ctx.attributes = _attributes
        ctx.point_adjacency = _point_adjacency
        ctx.point_adjacency_offsets = _point_adjacency_offsets

        results = pipeline.trace_forward(
            _points, _attributes, _point_adjacency, _point_adjacency_offsets, rays, start_point,
            depth_quantiles=depth_quantiles, return_contribution=return_contribution)
        ctx.rgba = results["rgba"]
        ctx.depth_indices = results.get("depth_indices", None)
        errbox = ErrorBox()
        ctx.errbox = errbox
        return (results["rgba"], results.get("depth", None), results.get("contribution", None),
                results["num_intersections"], errbox)

    @staticmethod
    def backward(ctx, grad_rgba, grad_depth, grad_contribution, grad_num_intersections, errbox_grad):
        del grad_contribution, grad_num_intersections, errbox_grad
        results = ctx.pipeline.trace_backward(
            ctx.points, ctx.attributes, ctx.point_adjacency, ctx.point_adjacency_offsets, ctx.rays,
            ctx.start_point, ctx.rgba, grad_rgba, ctx.depth_quantiles, ctx.depth_indices, grad_depth,
            ctx.errbox.ray_error, scrub_nonfinite=True)
        points_grad = results["points_grad"]
        attr_grad = results["attr_grad"]
        ctx.errbox.point_error = results.get("point_error", None)
        del (ctx.rays, ctx.start_point, ctx.pipeline, ctx.rgba, ctx.points, ctx.attributes,
             ctx.point_adjacency, ctx.point_adjacency_offsets, ctx.depth_quantiles)
        return (None, points_grad, attr_grad, None, None, None, None, None, None)
