"""Kernel LOGIC on the CPU: the product's CUDA sources compiled for host threads by tests/emu (a small emulation of
warps, shared memory, atomics and the synchronous part of the runtime API) and compared with the oracle.

What this covers without a GPU: the cell walk with the ranked face scan, the compositing / quantile code, both backward
kernels' gradient routing (MATCH.ANY groups, shared-memory staging, the warp's row cache, direct reductions), the walk
tape (record, replay, overflow -> re-walk, pool growth), the scene-mirror cache keys, the re-layout kernels, and the
small CSR passes.  What it does NOT cover: the compiled SASS (nvcc's contraction, libdevice, MUFU.RCP) -- GPU parity is
tests/test_gpu_*.py.  The emulated library is test infrastructure; the product never loads it and has no fallback."""
import os
import sys

import numpy as np
import pytest

import common
from radfoam_b200 import foam
from oracle import oracle

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))
import emu  # noqa: E402

TOL = dict(rtol=1e-5, atol=1e-5)   # floats: the emulation and the oracle share libm, so they agree far inside this
GRAD_TOL = 2e-5                    # of max|ref|; float scatter-adds in thread order


def scene(case):
    f = case.foam
    return f.points, f.attributes, f.adjacency, f.offsets


def check_forward(got, ref):
    for k in ("num_intersections", "depth_indices"):
        if k in got:
            assert np.array_equal(got[k].reshape(-1), np.asarray(ref[k]).reshape(-1)), k   # integers: bit-exact
    for k in ("rgba", "depth"):
        if k in got:
            np.testing.assert_allclose(got[k].astype(np.float32).reshape(-1),
                                       np.asarray(ref[k], dtype=np.float32).reshape(-1), **TOL)


def check_backward(got, ref):
    for k in ("points_grad", "attr_grad"):
        assert common.grad_error(got[k].astype(np.float32), np.asarray(ref[k], dtype=np.float32)) <= GRAD_TOL, k


@pytest.fixture(scope="module")
def small_scene():
    return common.scene_case(num_points=1500, width=48, height=32, q=2)


@pytest.mark.parametrize("sh_degree", [0, 1, 2, 3])
def test_forward_backward_config1_all_degrees(sh_degree):
    case = common.config1(sh_degree, 2)
    pipe = emu.EmuPipeline(sh_degree)
    fwd = pipe.trace_forward(*scene(case), case.rays, case.start, case.quantiles, return_contribution=True)
    ref = oracle.trace_forward(*scene(case), case.rays, case.start, case.quantiles, return_contribution=True)
    check_forward(fwd, ref)
    np.testing.assert_allclose(fwd["contribution"].reshape(-1), np.asarray(ref["contribution"]).reshape(-1),
                               rtol=1e-4, atol=1e-4)
    bwd = pipe.trace_backward(*scene(case), case.rays, case.start, fwd["rgba"], case.grad_rgba, case.quantiles,
                              fwd["depth_indices"], case.grad_depth)
    rb = oracle.trace_backward(*scene(case), case.rays, case.start, np.asarray(ref["rgba"]), case.grad_rgba,
                               case.quantiles, np.asarray(ref["depth_indices"]), case.grad_depth)
    check_backward(bwd, rb)


def test_flat_ray_batch_and_no_quantiles():
    case = common.config1(3, 0)
    pipe = emu.EmuPipeline(3)
    rays, start = case.rays.reshape(-1, 6), case.start.reshape(-1)
    fwd = pipe.trace_forward(*scene(case), rays, start)
    ref = oracle.trace_forward(*scene(case), rays, start)
    check_forward(fwd, ref)
    tiled = pipe.trace_forward(*scene(case), case.rays, case.start)   # 8x4 warp tiles: same result per ray
    assert np.array_equal(tiled["rgba"].reshape(-1, 4), fwd["rgba"])
    assert np.array_equal(tiled["num_intersections"].reshape(-1), fwd["num_intersections"].reshape(-1))


@pytest.fixture(scope="module")
def long_walk_scene():
    return common.scene_case(num_points=8000, width=32, height=16, q=2)   # walks of up to ~50 cells


def test_walk_tape_record_replay_overflow_and_growth(long_walk_scene):
    """First recording overflows the initial pool (one 32-step chunk per warp): the backward must fall back to
    re-walking (device-side flag) and give the same gradients; the next recording has a grown pool and is replayed."""
    case = long_walk_scene
    pipe = emu.EmuPipeline(3)
    plain = pipe.trace_forward(*scene(case), case.rays, case.start, case.quantiles)
    ref = oracle.trace_forward(*scene(case), case.rays, case.start, case.quantiles)
    check_forward(plain, ref)
    assert int(plain["num_intersections"].max()) > 40          # so one chunk per warp cannot hold the walks
    rb = oracle.trace_backward(*scene(case), case.rays, case.start, np.asarray(ref["rgba"]), case.grad_rgba,
                               case.quantiles, np.asarray(ref["depth_indices"]), case.grad_depth)
    args = (None,) * 6
    rec = pipe.trace_forward(*scene(case), case.rays, case.start, case.quantiles, scene_version=3, record_tape=True)
    for k in plain:
        assert np.array_equal(rec[k], plain[k]), k               # recording changes nothing
    first = pipe.tape_status()
    assert first["overflowed"] and first["used_chunks"] > first["capacity_chunks"]
    bwd = pipe.trace_backward(*args, rec["rgba"], case.grad_rgba, None, rec["depth_indices"], case.grad_depth,
                              scene_version=3, use_tape=True)
    check_backward(bwd, rb)                                      # overflowed tape -> the re-walk kernel did the work
    rec = pipe.trace_forward(*scene(case), case.rays, case.start, case.quantiles, scene_version=3, record_tape=True)
    second = pipe.tape_status()
    assert not second["overflowed"] and second["capacity_chunks"] >= first["used_chunks"]
    assert second["used_chunks"] == first["used_chunks"]
    replay = pipe.trace_backward(*args, rec["rgba"], case.grad_rgba, None, rec["depth_indices"], case.grad_depth,
                                 scene_version=3, use_tape=True)
    check_backward(replay, rb)
    # a stale tape (other scene version) must not be replayed: still correct through the re-walk path
    stale = pipe.trace_backward(*scene(case), case.rays, case.start, rec["rgba"], case.grad_rgba, case.quantiles,
                                rec["depth_indices"], case.grad_depth, scene_version=4, use_tape=False)
    check_backward(stale, rb)


def test_direct_backward_mode_and_point_error(small_scene, monkeypatch):
    case = small_scene
    rng = np.random.default_rng(5)
    ray_error = rng.uniform(0, 1, size=case.rays.shape[:-1] + (1,)).astype(np.float32)
    ref = oracle.trace_forward(*scene(case), case.rays, case.start, case.quantiles)
    rb = oracle.trace_backward(*scene(case), case.rays, case.start, np.asarray(ref["rgba"]), case.grad_rgba,
                               case.quantiles, np.asarray(ref["depth_indices"]), case.grad_depth, ray_error=ray_error)
    for mode in ("cached", "direct"):
        monkeypatch.setenv("RFB_BWD_MODE", mode)
        pipe = emu.EmuPipeline(3)
        fwd = pipe.trace_forward(*scene(case), case.rays, case.start, case.quantiles)
        bwd = pipe.trace_backward(*scene(case), case.rays, case.start, fwd["rgba"], case.grad_rgba, case.quantiles,
                                  fwd["depth_indices"], case.grad_depth, ray_error=ray_error)
        check_backward(bwd, rb)
        np.testing.assert_allclose(bwd["point_error"].reshape(-1), np.asarray(rb["point_error"]).reshape(-1),
                                   rtol=1e-4, atol=1e-5)


def test_half_precision_attributes():
    case = common.config1(3, 2)
    f = case.foam
    attrs = f.attributes.astype(np.float16)
    pipe = emu.EmuPipeline(3, np.float16)
    fwd = pipe.trace_forward(f.points, attrs, f.adjacency, f.offsets, case.rays, case.start, case.quantiles)
    ref = oracle.trace_forward(f.points, attrs, f.adjacency, f.offsets, case.rays, case.start, case.quantiles)
    assert fwd["rgba"].dtype == np.float16
    for k in ("num_intersections", "depth_indices"):
        assert np.array_equal(fwd[k].reshape(-1), np.asarray(ref[k]).reshape(-1))
    np.testing.assert_allclose(fwd["rgba"].astype(np.float32).reshape(-1),
                               np.asarray(ref["rgba"], dtype=np.float32).reshape(-1), rtol=2e-3, atol=1e-3)


def test_small_csr_passes_bit_exact(small_scene):
    f = small_scene.foam
    assert np.array_equal(emu.prefetch_adjacent_diff(f.points, f.adjacency, f.offsets).view(np.uint16),
                          oracle.prefetch_adjacent_diff(f.points, f.adjacency, f.offsets).view(np.uint16))
    for g in (f, common.farthest_edge_case()):
        idx, radius = emu.farthest_neighbor(g.points, g.adjacency, g.offsets)
        ref_idx, ref_radius = oracle.farthest_neighbor(g.points, g.adjacency, g.offsets)
        assert np.array_equal(idx, ref_idx)
        common.assert_same_floats(radius, ref_radius)
    rng = np.random.default_rng(1)
    queries = rng.normal(0, 2, size=(9, 3)).astype(np.float32)
    d2 = ((f.points[None].astype(np.float64) - queries[:, None].astype(np.float64)) ** 2).sum(-1)
    assert np.array_equal(emu.nearest_point(f.points, queries), d2.argmin(axis=1).astype(np.uint32))


@pytest.mark.parametrize("num_queries", [1, 2, 31, 257, 600])
def test_nearest_point_tiled_layouts(num_queries):
    """Every CTA layout of the tiled brute force (query lanes x sub-slices; more than 256 queries = several
    batches), ragged last tile, an exact tie (lowest index wins) and a NaN query (-> UINT32_MAX)."""
    rng = np.random.default_rng(num_queries)
    pts = rng.normal(0, 1, size=(2500, 3)).astype(np.float32)
    pts[1700] = pts[300]                       # duplicate point: the tie goes to index 300
    q = rng.normal(0, 1.5, size=(num_queries, 3)).astype(np.float32)
    q[0] = pts[1700]
    if num_queries > 2:
        q[2, 1] = np.nan
    got = emu.nearest_point(pts, q)
    d = pts[None, :, :] - q[:, None, :]
    # float32 distances in the kernel's own order: fma(dx,dx, fma(dy,dy, dz*dz)), each fma rounded once
    d64 = d.astype(np.float64)
    inner = (d64[..., 1] * d64[..., 1] + (d[..., 2] * d[..., 2]).astype(np.float64)).astype(np.float32)
    dist = (d64[..., 0] * d64[..., 0] + inner.astype(np.float64)).astype(np.float32)
    want = np.where(np.isnan(dist).all(axis=1), 0xFFFFFFFF, np.nanargmin(np.where(np.isnan(dist), np.inf, dist), axis=1))
    assert np.array_equal(got, want.astype(np.uint32))
    assert got[0] == 300


def test_start_points_dedupes_origins_on_the_device():
    """rfb_start_points == nearest point of every ray's origin, for one camera (a frame), a few cameras
    (a training batch), per-ray distinct origins, and odd origins (+0 / -0, NaN)."""
    rng = np.random.default_rng(12)
    f = common.scene_case(num_points=8000, width=32, height=16, q=2).foam
    cams = rng.normal(0, 2, size=(7, 3)).astype(np.float32)
    cams[5] = [0.0, 1.0, -0.0]
    cams[6] = [-0.0, 1.0, 0.0]                 # same point as cams[5], different bits: two queries, one answer
    ref = emu.nearest_point(f.points, cams)
    which = rng.integers(0, 7, size=(5, 211))
    rays = np.concatenate([cams[which], rng.normal(size=(5, 211, 3)).astype(np.float32)], axis=-1)
    assert np.array_equal(emu.start_points(f.points, rays), ref[which])
    frame = np.broadcast_to(np.concatenate([cams[1], [0, 0, 1]]).astype(np.float32), (9, 13, 6)).copy()
    assert np.all(emu.start_points(f.points, frame) == ref[1])
    scattered = rng.normal(0, 2, size=(300, 6)).astype(np.float32)      # every origin distinct
    assert np.array_equal(emu.start_points(f.points, scattered), emu.nearest_point(f.points, scattered[:, :3]))
    bad = scattered[:4].copy()
    bad[2, 0] = np.nan
    got = emu.start_points(f.points, bad)
    assert got[2] == 0xFFFFFFFF and np.array_equal(got[[0, 1, 3]], emu.nearest_point(f.points, bad[[0, 1, 3], :3]))


@pytest.mark.parametrize("tag", ["f16", "f32"])
@pytest.mark.parametrize("model", ["pinhole", "fisheye"])
def test_trace_benchmark_against_reference_kernel_frames(tag, model):
    """The benchmark kernel (in-kernel ray generation, caller-supplied half4 offsets, RGBA8 packing) emulated on the
    CPU vs frames the REFERENCE's own kernel rendered on a B200 (tests/golden/benchmark2k.npz): at most one level on a
    few pixels (expf of libm vs libdevice before a truncating pack)."""
    z = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "benchmark2k.npz")))
    cam = {k: z[f"cam_{model}_{k}"] for k in ("position", "forward", "right", "up")}
    cam.update(fov=float(z[f"cam_{model}_fov"]), width=64, height=48, model=model)
    diff = emu.prefetch_adjacent_diff(z["points"], z["adjacency"], z["offsets"])
    assert np.array_equal(diff.view(np.uint16), z["adjacent_diff"].view(np.uint16))
    attrs = z["attributes_" + tag]
    pipe = emu.EmuPipeline(3, attrs.dtype)
    img = pipe.trace_benchmark(z["points"], attrs, z["adjacency"], z["offsets"], diff, cam,
                               int(z[f"start_{model}"][0]), weight_threshold=0.05)
    a = img.view(np.uint8).reshape(48, 64, 4).astype(np.int32)
    b = z[f"image_{tag}_{model}"].view(np.uint8).reshape(48, 64, 4).astype(np.int32)
    assert np.abs(a - b).max() <= 1
    assert (a != b).any(axis=-1).mean() < 5e-3


import glob  # noqa: E402

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("benchmark", "farthest")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_emulated_kernels_against_reference_kernel_golden_vectors(path):
    """The product kernels' logic (CPU emulator) against outputs of the REFERENCE's own kernels on a B200
    (tests/golden/*.npz): traversal integers bit-exact; floats / gradients to the CPU tolerances of the oracle's own
    golden test (libm vs libdevice conditioning, tests/common.py)."""
    z = dict(np.load(path))
    half = z["attributes"].dtype == np.float16
    q = z.get("quantiles")
    kw = dict(weight_threshold=float(z["weight_threshold"]), max_intersections=int(z["max_intersections"]))
    scene_ = (z["points"], z["attributes"], z["adjacency"], z["offsets"])
    deg = oracle.sh_degree_of(z["attributes"].shape[-1])
    pipe = emu.EmuPipeline(deg, z["attributes"].dtype)
    fwd = pipe.trace_forward(*scene_, z["rays"], z["start"], q, **kw)
    assert np.array_equal(fwd["num_intersections"].reshape(-1), z["out_num_intersections"].reshape(-1))
    if q is not None:
        assert np.array_equal(fwd["depth_indices"].reshape(-1), z["out_depth_indices"].reshape(-1))
    if half:
        np.testing.assert_allclose(fwd["rgba"].astype(np.float32), z["out_rgba"].astype(np.float32).reshape(fwd["rgba"].shape),
                                   rtol=2e-3, atol=1e-3)
        return
    ref = {k[4:]: v for k, v in z.items() if k.startswith("out_")}
    got = {k: v.reshape(np.asarray(ref[k]).shape) for k, v in fwd.items() if k in ref}
    common.assert_forward_close_cpu(got, ref, z["attributes"])
    bwd = pipe.trace_backward(*scene_, z["rays"], z["start"], z["out_rgba"], z["grad_rgba"], q,
                              z.get("out_depth_indices"), z.get("grad_depth"), **kw)
    for k in ("points_grad", "attr_grad"):
        assert common.nonfinite_mismatch(bwd[k], ref[k]) == 0
        assert common.grad_error(bwd[k], ref[k]) <= common.CPU_GRAD_TOL, k


@pytest.mark.parametrize("seed", ["1", "2", "3"])
def test_backward_under_shuffled_lane_schedules(seed, long_walk_scene, monkeypatch, tmp_path):
    """A poor man's racecheck for warp-level synchronisation: the emulator resumes the lanes of a CTA in a different
    pseudo-random order on every scheduling pass (RFB_EMU_SHUFFLE), so a shared-memory read that is not separated from
    another lane's write by a __syncwarp gives wrong gradients.  Runs in a subprocess: the order is fixed at load."""
    import subprocess
    code = f"""
import os, sys
sys.path.insert(0, {os.path.dirname(os.path.dirname(__file__))!r}); sys.path.insert(0, {os.path.dirname(__file__)!r})
sys.path.insert(0, {os.path.join(os.path.dirname(__file__), "emu")!r})
import numpy as np, common, emu
from oracle import oracle
case = common.scene_case(num_points=8000, width=32, height=16, q=2)
f = case.foam; sc = (f.points, f.attributes, f.adjacency, f.offsets)
ref = oracle.trace_forward(*sc, case.rays, case.start, case.quantiles)
rb = oracle.trace_backward(*sc, case.rays, case.start, np.asarray(ref["rgba"]), case.grad_rgba, case.quantiles,
                           np.asarray(ref["depth_indices"]), case.grad_depth)
pipe = emu.EmuPipeline(3)
for _ in range(2):
    rec = pipe.trace_forward(*sc, case.rays, case.start, case.quantiles, scene_version=5, record_tape=True)
assert np.array_equal(rec["num_intersections"].reshape(-1), np.asarray(ref["num_intersections"]).reshape(-1))
for use_tape in (True, False):
    bwd = pipe.trace_backward(*sc, case.rays, case.start, rec["rgba"], case.grad_rgba, case.quantiles,
                              rec["depth_indices"], case.grad_depth, scene_version=5) if not use_tape else \
          pipe.trace_backward(*(None,) * 6, rec["rgba"], case.grad_rgba, None, rec["depth_indices"], case.grad_depth,
                              scene_version=5, use_tape=True)
    for k in ("points_grad", "attr_grad"):
        assert common.grad_error(bwd[k], np.asarray(rb[k])) <= 2e-5, (k, use_tape)
print("ok")
"""
    env = dict(os.environ, RFB_EMU_SHUFFLE=seed)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_replay_schedule_is_a_longest_first_permutation(long_walk_scene):
    """The replay's tile order (small launches): tile_steps = the longest recorded walk of each 16x8 tile, order = a
    permutation of the tiles with non-increasing tile_steps."""
    case = long_walk_scene
    pipe = emu.EmuPipeline(3)
    for _ in range(2):
        rec = pipe.trace_forward(*scene(case), case.rays, case.start, case.quantiles, scene_version=2, record_tape=True)
    sched = emu.tape_schedule(pipe)
    assert sched is not None
    steps, order = sched
    h, w = case.rays.shape[:2]
    bx, by = (w + 15) // 16, (h + 7) // 8
    assert steps.shape == (bx * by,) and sorted(order.tolist()) == list(range(bx * by))
    assert np.all(np.diff(steps[order].astype(np.int64)) <= 0)
    cells, t1, count = emu.tape_records(pipe, h, w)          # [warps][steps][32], 4 warps per tile
    want = count.reshape(bx * by, 4 * 32).max(axis=1)
    assert np.array_equal(steps, want.astype(np.uint32))
    assert int(rec["num_intersections"].max()) >= int(steps.max())


@pytest.mark.parametrize("deg,dtype", [(3, np.float32), (1, np.float32), (0, np.float32), (3, np.float16)])
def test_parameter_form_scene(deg, dtype):
    """SURVEY.md §8f.2 on the CPU: with a bound parameter-form scene the re-layout kernel evaluates
    attributes = cat(att_dc, att_sh, scale * softplus(density, beta=10)).to(dtype) itself and the finalize kernel
    returns the parameters' gradients (chain rule through cat and softplus) -- against the attribute-form path fed
    with the same values."""
    f = foam.small_foam(200, sh_degree=deg, seed=4)
    rng = np.random.default_rng(8)
    n, adim = f.num_points, f.attributes.shape[1]
    att_dc, att_sh = f.attributes[:, :3].copy(), f.attributes[:, 3:adim - 1].copy()
    raw = rng.normal(0.0, 0.4, size=(n, 1)).astype(np.float32)
    raw[:3, 0] = [2.5, -3.0, 0.0]          # beyond softplus' linear threshold (x * beta > 20), deep in the tail, at 0
    scale = np.float32(1.7)
    xb = raw * np.float32(10.0)
    sigma = scale * np.where(xb > 20.0, raw, np.log1p(np.exp(xb.astype(np.float64))).astype(np.float32) / np.float32(10.0))
    attrs = np.concatenate([att_dc, att_sh, sigma.astype(np.float32)], axis=1).astype(dtype)
    rays = foam.pinhole_rays(24, 16, (0, 0, -3), fov=0.7, up=(0, 1, 0))
    start = np.full((16, 24), foam.nearest_point(f.points, (0, 0, -3)), dtype=np.uint32)
    dq = np.tile(np.array([0.7, 0.3], dtype=np.float32), (16, 24, 1))
    g = rng.normal(size=(16, 24, 4)).astype(dtype)
    gd = (rng.normal(size=(16, 24, 2)) * 1e-3).astype(np.float32)
    plain = emu.EmuPipeline(deg, dtype)
    want = plain.trace_forward(f.points, attrs, f.adjacency, f.offsets, rays, start, dq)
    wb = plain.trace_backward(f.points, attrs, f.adjacency, f.offsets, rays, start, want["rgba"], g, dq,
                              want["depth_indices"], gd, scrub_nonfinite=True)
    pipe = emu.EmuPipeline(deg, dtype)
    pipe.bind_scene_params(att_dc, att_sh, raw, scale)
    got = pipe.trace_forward(f.points, attrs, f.adjacency, f.offsets, rays, start, dq)   # `attrs` is ignored while bound
    assert np.array_equal(got["num_intersections"], want["num_intersections"])
    assert np.array_equal(got["depth_indices"], want["depth_indices"])
    np.testing.assert_allclose(got["rgba"].astype(np.float32), want["rgba"].astype(np.float32), rtol=2e-6, atol=2e-6)
    pipe.trace_backward_accumulate(f.points, attrs, f.adjacency, f.offsets, rays, start, got["rgba"], g, dq,
                                   got["depth_indices"], gd)
    gb = pipe.trace_backward_finalize_params(n, scrub_nonfinite=True)
    tol = 2e-3 if dtype == np.float16 else 2e-5
    ag = wb["attr_grad"].astype(np.float32)
    assert common.grad_error(gb["points_grad"], wb["points_grad"]) <= 2e-5
    assert common.grad_error(gb["att_dc_grad"], ag[:, :3]) <= tol
    if adim > 4:
        assert common.grad_error(gb["att_sh_grad"], ag[:, 3:adim - 1]) <= tol
    z = np.exp(xb.astype(np.float64))
    chain = np.where(xb > 20.0, 1.0, z / (z + 1.0)) * float(scale)
    assert common.grad_error(gb["density_grad"], (ag[:, -1:].astype(np.float64) * chain).astype(np.float32)) <= tol
    pipe.bind_scene_params(None, None, None)


@pytest.mark.parametrize("multicast", [False, True], ids=["peer_pointers", "nvswitch_multicast"])
@pytest.mark.parametrize("world,num_points,deg,dtype", [(2, 64, 3, np.float32), (3, 45, 3, np.float32),
                                                        (8, 150, 3, np.float32), (2, 37, 1, np.float16),
                                                        (4, 16, 0, np.float32), (5, 3, 2, np.float16)])
def test_fused_peer_reduce_finalize(world, num_points, deg, dtype, multicast):
    """The multi-GPU exchange kernel (rfb_reduce_finalize_peers) on the CPU: every rank sums its row blocks over all
    ranks' accumulators in rank order, finalizes, and writes them into every rank's outputs; after all ranks ran,
    every rank holds exactly what all-reduce + finalize gives (ragged last block, fp16 outputs, scrub)."""
    rng = np.random.default_rng(world * 1000 + num_points)
    pipe = emu.EmuPipeline(deg, dtype)
    sr = ((3 * (deg + 1) ** 2 + 3) // 4) * 4
    gr, adim = sr + 4, 1 + 3 * (deg + 1) ** 2
    accs = [rng.normal(size=(num_points, gr)).astype(np.float32) for _ in range(world)]
    accs[0][1, 2] = np.inf          # becomes non-finite in the sum -> scrubbed
    accs[-1][num_points - 1, sr + 2] = np.nan
    total = accs[0].copy()
    for a in accs[1:]:              # rank order, float32: bit-exact expectation
        total = total + a
    want_attr = np.concatenate([total[:, :adim - 1], total[:, sr:sr + 1]], axis=1).astype(dtype)
    want_pts = total[:, sr + 1:sr + 4].copy()
    want_attr[~np.isfinite(want_attr.astype(np.float32))] = 0
    want_pts[~np.isfinite(want_pts)] = 0
    got_attr = [np.full((num_points, adim), 7.0, dtype) for _ in range(world)]
    got_pts = [np.full((num_points, 3), 7.0, np.float32) for _ in range(world)]
    written = np.zeros(num_points, dtype=np.int64)
    for rank in range(world):
        attr, pts = pipe.reduce_finalize_peers(rank, accs, scrub_nonfinite=True, multicast=multicast)
        mine = (np.arange(num_points) // 16) % world == rank      # blocks of 16 rows, round-robin
        for w in range(world):
            assert np.all(attr[w][~mine] == 7.0) and np.all(pts[w][~mine] == 7.0)   # nothing outside its share
            got_attr[w][mine] = attr[w][mine]
            got_pts[w][mine] = pts[w][mine]
        written += mine
    assert np.all(written == 1)
    for w in range(world):
        assert np.array_equal(got_attr[w].view(np.uint16 if dtype == np.float16 else np.uint32),
                              want_attr.view(np.uint16 if dtype == np.float16 else np.uint32))
        assert np.array_equal(got_pts[w].view(np.uint32), want_pts.view(np.uint32))


@pytest.mark.parametrize("world", [2, 3])
def test_ray_sharded_split_backward_matches_single_rank(world, long_walk_scene):
    """The multi-GPU data path on the CPU: the frame is dealt to `world` ranks as interleaved 8-row bands
    (radfoam_b200.sharded), every rank runs forward + rfb_trace_backward_accumulate on its shard, the accumulators are
    summed (what the one NCCL all-reduce does) and rfb_trace_backward_finalize writes the gradients -- which must equal
    the single-rank backward (up to float summation order) and the oracle."""
    import torch

    from radfoam_b200 import sharded

    case = long_walk_scene
    height = case.rays.shape[0]
    ref = oracle.trace_forward(*scene(case), case.rays, case.start, case.quantiles)
    rb = oracle.trace_backward(*scene(case), case.rays, case.start, np.asarray(ref["rgba"]), case.grad_rgba,
                               case.quantiles, np.asarray(ref["depth_indices"]), case.grad_depth)
    single = emu.EmuPipeline(3)
    fwd1 = single.trace_forward(*scene(case), case.rays, case.start, case.quantiles)
    total = None
    pipes = []
    rgba_parts = []
    for rank in range(world):
        rows = sharded.band_rows(height, rank, world).numpy()
        part = lambda a: None if a is None else np.ascontiguousarray(a[rows])  # noqa: E731
        pipe = emu.EmuPipeline(3)
        fwd = pipe.trace_forward(*scene(case), part(case.rays), part(case.start), part(case.quantiles))
        rgba_parts.append(torch.from_numpy(fwd["rgba"]))
        acc = pipe.trace_backward_accumulate(*scene(case), part(case.rays), part(case.start), fwd["rgba"],
                                             part(case.grad_rgba), part(case.quantiles), fwd["depth_indices"],
                                             part(case.grad_depth))
        total = acc.copy() if total is None else total + acc
        pipes.append((pipe, acc))
    # forward needs no collective: the shards reassemble to the single-rank image bit for bit
    assert np.array_equal(sharded.unshard_image(rgba_parts, height).numpy(), fwd1["rgba"])
    for pipe, acc in pipes:      # every rank holds the reduced accumulator after the all-reduce
        acc[...] = total
        got = pipe.trace_backward_finalize(case.foam.num_points, scrub_nonfinite=True)
        check_backward(got, rb)
