"""world_size-2 gloo tests (CPU) of the N>1 path's host logic: shard -> trace -> gather and
accumulate -> all-reduce -> finalize, with a pure-torch stand-in for the CUDA pipeline (the
real kernels are covered by the -m gpu tests; the collective wiring is what runs here)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common  # noqa: F401
from radfoam_b200 import sharded


class FakePipeline:
    """Additive toy 'tracer': rgba depends only on the ray; gradients are per-ray scatter-adds
    into a [N, row] accumulator, so sharded-sum == unsharded."""

    row = 8

    def __init__(self, n):
        self.n = n
        self.acc = None

    def trace_forward(self, points, attributes, adj, off, rays, start, depth_quantiles=None,
                      return_contribution=False, **kw):
        rgba = torch.stack([rays[..., 0], rays[..., 1], rays[..., 3], rays[..., :3].sum(-1)], dim=-1)
        return {"rgba": rgba, "num_intersections": torch.ones(rays.shape[:-1] + (1,), dtype=torch.int32)}

    def _accumulate(self, rays, start, grad_in):
        acc = torch.zeros(self.n, self.row, dtype=torch.float64)
        idx = start.reshape(-1).long()
        contrib = torch.cat([grad_in.reshape(-1, 4), rays.reshape(-1, 6)[:, :4]], dim=1).double()
        acc.index_add_(0, idx, contrib)
        return acc

    def trace_backward_accumulate(self, points, attributes, adj, off, rays, start, rgb_out, grad_in, *a, **kw):
        self.acc = self._accumulate(rays, start, grad_in)
        return self.acc, None

    def trace_backward_finalize(self, num_points, device, scrub_nonfinite=False):
        return self.acc[:, 5:8].clone(), self.acc[:, :5].clone()

    def trace_backward(self, points, attributes, adj, off, rays, start, rgb_out, grad_in, *a, **kw):
        self.acc = self._accumulate(rays, start, grad_in)
        pg, ag = self.trace_backward_finalize(self.n, None)
        return {"points_grad": pg, "attr_grad": ag}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, height, width, n, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        rays = torch.rand(height, width, 6, generator=g)
        start = torch.randint(0, n, (height, width), generator=g)
        grad = torch.rand(height, width, 4, generator=g)
        tracer = sharded.ShardedTracer(FakePipeline(n))
        assert tracer.world == world and tracer.rank == rank
        my_rays, my_start, my_grad = (tracer.shard(t, image=True) for t in (rays, start, grad))
        fwd = tracer.trace_forward(None, None, None, None, my_rays, my_start)
        full = tracer.gather_image(fwd["rgba"], height)
        points = torch.zeros(n, 3)
        bwd = tracer.trace_backward(points, None, None, None, my_rays, my_start, fwd["rgba"], my_grad)
        # flat-batch path
        flat = tracer.shard(rays.reshape(-1, 6), image=False)
        results[rank] = (full, bwd["points_grad"], bwd["attr_grad"], flat.shape[0])
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_gather_and_grad_allreduce():
    world, height, width, n = 2, 37, 5, 11
    port = _free_port()
    with mp.Manager() as manager:
        results = manager.dict()
        mp.spawn(_worker, args=(world, port, height, width, n, results), nprocs=world, join=True)
        results = dict(results)
    g = torch.Generator().manual_seed(0)
    rays = torch.rand(height, width, 6, generator=g)
    start = torch.randint(0, n, (height, width), generator=g)
    grad = torch.rand(height, width, 4, generator=g)
    single = FakePipeline(n)
    want_rgba = single.trace_forward(None, None, None, None, rays, start)["rgba"]
    want = single.trace_backward(None, None, None, None, rays, start, want_rgba, grad)
    total_flat = 0
    for rank in range(world):
        full, pg, ag, nflat = results[rank]
        assert torch.equal(full, want_rgba)                     # forward: bit-identical reassembly
        assert torch.allclose(pg, want["points_grad"], rtol=1e-12, atol=1e-12)
        assert torch.allclose(ag, want["attr_grad"], rtol=1e-12, atol=1e-12)
        total_flat += nflat
    assert total_flat == height * width
