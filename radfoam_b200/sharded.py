"""Ray-sharded data parallelism over the GPUs of one box.

The reference is single-GPU; this is the one thing the build adds around the hot path.
One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  The scene buffers are
replicated on every rank; rays are dealt to ranks

  * image batches ``[H, W, 6]``: as interleaved bands of ``band`` (default 8) full-width
    pixel rows, band b going to rank ``b % world`` -- every rank samples the whole frame
    (load balance: sky and surface regions cost very different numbers of cells) and its
    shard is still a row-major image, so the kernels' 8x4 warp tiles stay coherent;
  * unordered batches ``[R, 6]``: as contiguous chunks.

Forward needs no collective (disjoint pixels).  Backward scatter-adds into the pipeline's
fp32 accumulator ``[N, grad_row]`` on each rank; ONE all-reduce (sum) of that accumulator
makes every rank hold the full per-point gradient, then the epilogue writes the reference
layout and zeroes non-finite entries -- after the reduction, as radfoam_model/render.py:98-99
does after its single-GPU kernel.

The partition / reassembly helpers are pure index arithmetic on torch tensors and are
exercised on CPU with gloo (tests/test_sharded_gloo.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def band_rows(height: int, rank: int, world: int, band: int = 8) -> torch.Tensor:
    """Row indices (ascending) of the image rows rank ``rank`` owns."""
    rows = torch.arange(height)
    return rows[(rows // band) % world == rank]


def shard_image(t: torch.Tensor, rank: int, world: int, band: int = 8) -> torch.Tensor:
    """``t`` is ``[H, W, ...]``; returns this rank's rows ``[h_r, W, ...]`` (a copy)."""
    if world == 1:
        return t
    idx = band_rows(t.shape[0], rank, world, band).to(t.device)
    return t.index_select(0, idx)


def shard_flat(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """``t`` is ``[R, ...]``; contiguous chunk ``rank`` of ``world`` (sizes differ by <= 1)."""
    if world == 1:
        return t
    n = t.shape[0]
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return t[lo:hi]


def unshard_image(parts: list[torch.Tensor], height: int, band: int = 8) -> torch.Tensor:
    """Inverse of :func:`shard_image` given every rank's part in rank order."""
    world = len(parts)
    if world == 1:
        return parts[0]
    out = torch.empty((height,) + tuple(parts[0].shape[1:]), dtype=parts[0].dtype, device=parts[0].device)
    for r, p in enumerate(parts):
        out.index_copy_(0, band_rows(height, r, world, band).to(p.device), p)
    return out


class ShardedTracer:
    """Wraps a Pipeline (or any object with its trace_* methods) for ray-sharded use.

    ``trace_forward`` / ``trace_backward`` take THIS RANK'S shard of the rays (see
    :func:`shard_image` / :func:`shard_flat`) and the full, replicated scene; backward
    returns the full (all-reduced) gradients on every rank.
    """

    def __init__(self, pipeline, group=None, band: int = 8):
        self.pipeline = pipeline
        self.group = group
        self.band = band
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.world = dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1

    # -- partition helpers bound to this rank
    def shard(self, t: torch.Tensor, image: bool) -> torch.Tensor:
        return (shard_image(t, self.rank, self.world, self.band) if image
                else shard_flat(t, self.rank, self.world))

    def gather_image(self, local: torch.Tensor, height: int) -> torch.Tensor:
        """All ranks receive the full ``[H, W, ...]`` image assembled from the shards."""
        if self.world == 1:
            return local
        sizes = [int(band_rows(height, r, self.world, self.band).numel()) for r in range(self.world)]
        rows = max(sizes)  # all_gather wants equal shapes: pad the short shards, trim after
        padded = local
        if local.shape[0] != rows:
            padded = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            padded[:local.shape[0]] = local
        parts = [torch.empty_like(padded) for _ in range(self.world)]
        dist.all_gather(parts, padded.contiguous(), group=self.group)
        return unshard_image([p[:s] for p, s in zip(parts, sizes)], height, self.band)

    # -- the hot path
    def trace_forward(self, *args, **kwargs):
        return self.pipeline.trace_forward(*args, **kwargs)

    def trace_backward(self, points, attributes, point_adjacency, point_adjacency_offsets, rays,
                       start_point, rgb_out, grad_in, depth_quantiles=None, depth_indices=None,
                       depth_grad_in=None, ray_error=None, weight_threshold=None,
                       max_intersections=None, scrub_nonfinite=True):
        if self.world == 1:
            return self.pipeline.trace_backward(
                points, attributes, point_adjacency, point_adjacency_offsets, rays, start_point,
                rgb_out, grad_in, depth_quantiles, depth_indices, depth_grad_in, ray_error,
                weight_threshold, max_intersections, scrub_nonfinite=scrub_nonfinite)
        acc, point_error = self.pipeline.trace_backward_accumulate(
            points, attributes, point_adjacency, point_adjacency_offsets, rays, start_point, rgb_out,
            grad_in, depth_quantiles, depth_indices, depth_grad_in, ray_error, weight_threshold,
            max_intersections)
        # the single collective of the path: per-point gradient sum over ranks
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
        if point_error is not None:
            dist.all_reduce(point_error, op=dist.ReduceOp.SUM, group=self.group)
        points_grad, attr_grad = self.pipeline.trace_backward_finalize(
            points.shape[0], rays.device, scrub_nonfinite=scrub_nonfinite)
        out = {"points_grad": points_grad, "attr_grad": attr_grad, "ray_grad": torch.empty_like(rays)}
        if point_error is not None:
            out["point_error"] = point_error
        return out


class ShardedTraceRays(torch.autograd.Function):
    """The render op (radfoam_b200/render.py) over a ShardedTracer: forward on the local ray shard,
    backward yields the all-reduced scene gradients on every rank.  Returns
    ``(rgba, depth, contribution, num_intersections)`` of the local shard."""

    @classmethod
    def apply(cls, tracer, *args):
        pipe = tracer.pipeline
        noted = hasattr(pipe, "autograd_recording")
        if noted:
            before, pipe.autograd_recording = pipe.autograd_recording, torch.is_grad_enabled()
        try:
            return super().apply(tracer, *args)
        finally:
            if noted:
                pipe.autograd_recording = before

    @staticmethod
    def forward(ctx, tracer, points, attributes, point_adjacency, point_adjacency_offsets, rays, start_point,
                depth_quantiles, return_contribution):
        out = tracer.trace_forward(points, attributes, point_adjacency, point_adjacency_offsets, rays,
                                   start_point, depth_quantiles=depth_quantiles,
                                   return_contribution=return_contribution)
        saved = [points, attributes, point_adjacency, point_adjacency_offsets, rays, start_point, out["rgba"]]
        ctx.with_depth = depth_quantiles is not None
        if ctx.with_depth:
            saved += [depth_quantiles, out["depth_indices"]]
        ctx.save_for_backward(*saved)  # no output tensor on ctx itself: that would be a reference cycle
        ctx.tracer = tracer
        return out["rgba"], out.get("depth"), out.get("contribution"), out["num_intersections"]

    @staticmethod
    def backward(ctx, grad_rgba, grad_depth, _grad_contribution, _grad_num_intersections):
        saved = ctx.saved_tensors
        pts, attrs, adj, off, rays, start, rgba = saved[:7]
        dq, didx = (saved[7], saved[8]) if ctx.with_depth else (None, None)
        res = ctx.tracer.trace_backward(pts, attrs, adj, off, rays, start, rgba, grad_rgba, dq, didx, grad_depth,
                                        scrub_nonfinite=True)
        ctx.tracer = None
        return (None, res["points_grad"], res["attr_grad"]) + (None,) * 6
