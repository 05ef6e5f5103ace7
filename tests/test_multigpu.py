"""N>1 parity on real GPUs: ray-sharded forward/backward over NCCL == single-GPU result
(SURVEY.md §8e parity statement).  Skipped on boxes with fewer than 2 GPUs; the host logic
of the same path is covered on CPU by test_sharded_gloo.py."""
import os
import socket

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, mode, results):
    fused = mode != "nccl_all_reduce"
    os.environ["RFB_MULTICAST"] = "1" if mode == "fused_nvswitch_multicast" else "0"
    import torch
    import torch.distributed as dist

    import radfoam_b200
    from radfoam_b200 import sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        case = common.scene_case(num_points=60000, width=320, height=200)
        f = case.foam
        dev = torch.device("cuda", rank)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        scene = [d(x) for x in (f.points, f.attributes, f.adjacency, f.offsets)]
        tracer = sharded.ShardedTracer(radfoam_b200.create_pipeline(3), fused_reduce=fused)
        full = {k: d(v) for k, v in dict(rays=case.rays, start=case.start, dq=case.quantiles,
                                         g=case.grad_rgba, gd=case.grad_depth).items()}
        mine = {k: tracer.shard(v, image=True).contiguous() for k, v in full.items()}
        fwd = tracer.trace_forward(*scene, mine["rays"], mine["start"], depth_quantiles=mine["dq"])
        H = case.rays.shape[0]
        rgba = tracer.gather_image(fwd["rgba"], H)
        nint = tracer.gather_image(fwd["num_intersections"].to(torch.int32), H)
        for _ in range(2):  # twice: the second step re-uses (and re-zeroes) the peer-mapped accumulators
            bwd = tracer.trace_backward(*scene, mine["rays"], mine["start"], fwd["rgba"], mine["g"], mine["dq"],
                                        fwd["depth_indices"], mine["gd"], scrub_nonfinite=False)
        torch.cuda.synchronize()
        # every rank must hold the same bits
        mine_bits = torch.cat([bwd["points_grad"].reshape(-1), bwd["attr_grad"].reshape(-1)]).view(torch.int32)
        all_bits = [torch.empty_like(mine_bits) for _ in range(world)]
        dist.all_gather(all_bits, mine_bits)
        if rank == 0:
            results["same_on_all_ranks"] = bool(all(torch.equal(all_bits[0], b) for b in all_bits))
            results["fused_active"] = bool(tracer._peer)
            results["multicast_active"] = bool(tracer._peer and tracer._peer["multicast"])
            results["fused_error"] = tracer.fused_reduce_error
            single = radfoam_b200.create_pipeline(3)
            sf = single.trace_forward(*scene, full["rays"], full["start"], depth_quantiles=full["dq"])
            sb = single.trace_backward(*scene, full["rays"], full["start"], sf["rgba"], full["g"], full["dq"],
                                       sf["depth_indices"], full["gd"])
            torch.cuda.synchronize()
            results["rgba_equal"] = bool(torch.equal(rgba, sf["rgba"]))
            results["nint_equal"] = bool(torch.equal(nint, sf["num_intersections"].to(torch.int32)))
            for k in ("points_grad", "attr_grad"):
                results[k] = common.grad_error(bwd[k].cpu().numpy(), sb[k].cpu().numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["fused_nvswitch_multicast", "fused_peer_pointers", "nccl_all_reduce"])
def test_two_gpu_sharded_matches_single_gpu(mode):
    fused = mode != "nccl_all_reduce"
    import torch
    import torch.multiprocessing as mp

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    with mp.Manager() as manager:
        results = manager.dict()
        mp.spawn(_worker, args=(world, _free_port(), mode, results), nprocs=world, join=True)
        results = dict(results)
    print("fused path active:", results["fused_active"], "multicast:", results["multicast_active"],
          results["fused_error"])
    if mode == "fused_peer_pointers":
        assert not results["multicast_active"]
    assert results["fused_active"] == fused or (fused and results["fused_error"])  # falls back only with a reason
    assert results["same_on_all_ranks"]
    assert results["rgba_equal"] and results["nint_equal"]     # same per-ray code: bit-identical
    assert results["points_grad"] < 1e-5 and results["attr_grad"] < 1e-5  # fp32 summation order only
