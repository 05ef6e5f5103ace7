"""Host-side tests (CPU tensors) of the scene container and the reference's ``.pt`` checkpoint format
(radfoam_model/scene.py:202-217, 614-656; benchmark.py:57-84) -- SURVEY.md §8f.3."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from radfoam_b200 import foam, scene_io


def make_scene(attr_dtype=torch.float32, sh_degree=3, n=200):
    f = foam.small_foam(n, sh_degree=sh_degree, seed=3)
    return f, scene_io.FoamScene.from_foam(f, attr_dtype=attr_dtype, device="cpu")


def test_pt_round_trip_has_the_reference_layout(tmp_path):
    f, scene = make_scene()
    path = tmp_path / "model.pt"
    scene.save_pt(path)
    raw = torch.load(path)
    assert set(raw) == set(scene_io.PT_KEYS)
    n, e = f.num_points, f.adjacency.size
    assert raw["xyz"].shape == (n, 3) and raw["xyz"].dtype == torch.float32
    assert raw["density"].shape == (n, 1) and raw["color_dc"].shape == (n, 3) and raw["color_sh"].shape == (n, 45)
    assert raw["adjacency"].dtype == torch.int64 and raw["adjacency"].shape == (e,)          # scene.py:627-628
    assert raw["adjacency_offsets"].dtype == torch.int64 and raw["adjacency_offsets"].shape == (n + 1,)
    back = scene_io.FoamScene.load_pt(path, sh_degree=3, device="cpu")
    assert back.point_adjacency.dtype == torch.uint32 and back.point_adjacency_offsets.dtype == torch.uint32
    for a, b in zip(scene.get_trace_data(), back.get_trace_data()):
        assert a.dtype == b.dtype and torch.equal(a.to(torch.int64) if a.dtype == torch.uint32 else a,
                                                  b.to(torch.int64) if b.dtype == torch.uint32 else b)


def test_trace_data_matches_scene_py_formula():
    f, scene = make_scene()
    points, attributes, adjacency, offsets = scene.get_trace_data()
    want = torch.cat([scene.att_dc, scene.att_sh, 1.0 * F.softplus(scene.density, beta=10)], dim=-1)
    assert torch.equal(attributes, want) and attributes.shape == (f.num_points, 49)
    # from_foam inverts the activation: the traced attributes are the synthetic foam's again
    np.testing.assert_allclose(attributes.numpy(), f.attributes, rtol=2e-6, atol=1e-7)
    assert np.array_equal(adjacency.to(torch.int64).numpy(), f.adjacency.astype(np.int64))
    assert np.array_equal(points.numpy(), f.points)
    scaled = scene_io.FoamScene(scene.primal_points, scene.density, scene.att_dc, scene.att_sh, scene.point_adjacency,
                                scene.point_adjacency_offsets, activation_scale=2.5)
    assert torch.equal(scaled.get_primal_density(), 2.5 * F.softplus(scene.density, beta=10))


def test_half_precision_attributes_like_benchmark_py():
    _, scene = make_scene(attr_dtype=torch.float16)          # benchmark.py:36-38
    _, attributes, _, _ = scene.get_trace_data()
    assert attributes.dtype == torch.float16 and scene.att_dc.dtype == torch.float16
    assert scene.density.dtype == torch.float32               # the raw density stays fp32 (scene.py:632)


def test_sh_degree_mismatch_is_the_reference_assertion(tmp_path):
    _, scene = make_scene(sh_degree=2, n=120)
    path = tmp_path / "deg2.pt"
    scene.save_pt(path)
    with pytest.raises(AssertionError, match="Expected 45 SH coeffs per-point, got 24"):
        scene_io.FoamScene.load_pt(path, sh_degree=3, device="cpu")
    assert scene_io.FoamScene.load_pt(path, sh_degree=2, device="cpu").get_trace_data()[1].shape[1] == 28
    torch.save({"xyz": torch.zeros(1, 3)}, tmp_path / "bad.pt")
    with pytest.raises(KeyError, match="missing"):
        scene_io.FoamScene.load_pt(tmp_path / "bad.pt", device="cpu")
    with pytest.raises(RuntimeError, match="do not fit uint32"):
        scene_io._to_uint32(torch.tensor([-1], dtype=torch.int64))


def test_benchmark_cameras_follow_benchmark_py():
    rng = np.random.default_rng(0)
    c2w = torch.from_numpy(rng.normal(size=(20, 4, 4)).astype(np.float32))
    cams, positions = scene_io.benchmark_cameras(c2w, fy=600.0, width=64, height=48)
    assert len(cams) == 3 and positions.shape == (3, 3)       # poses 0, 8, 16
    cam = cams[1]
    assert torch.equal(cam["position"], c2w[8, :3, 3]) and torch.equal(cam["right"], c2w[8, :3, 0])
    assert torch.equal(cam["up"], -c2w[8, :3, 1]) and torch.equal(cam["forward"], c2w[8, :3, 2])
    assert cam["fov"] == pytest.approx(2 * math.atan(48 / 1200.0)) and cam["model"] == "pinhole"
    assert all(v.is_contiguous() for v in (cam["position"], cam["right"], cam["up"], cam["forward"]))


def test_benchmark_fps_loop_on_the_cpu_emulator(monkeypatch, tmp_path):
    """The FPS loop of scene_io.benchmark_fps (benchmark.py:86-139) end to end without a GPU: a scene saved to .pt and
    reloaded with fp16 attributes, the kernels emulated on the CPU (tests/emu) behind a stand-in pipeline object, CUDA
    events / synchronize stubbed; frames must equal the oracle's trace_benchmark up to one level on a few pixels."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))
    import emu
    from oracle import oracle
    from radfoam_b200 import pipeline as product_pipeline

    f = foam.scene_foam(2000, sh_degree=3)
    path = tmp_path / "model.pt"
    scene_io.FoamScene.from_foam(f, device="cpu").save_pt(path)
    scene = scene_io.FoamScene.load_pt(path, sh_degree=3, attr_dtype=torch.float16, device="cpu")

    def as_u32(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).view(torch.uint32)

    class EmulatedPipeline:
        def __init__(self):
            self.inner = emu.EmuPipeline(3, np.float16)

        def prefetch_adjacent_diff(self, points, adjacency, offsets):
            return torch.from_numpy(emu.prefetch_adjacent_diff(points.numpy(), adjacency.view(torch.int32).numpy().view(np.uint32),
                                                               offsets.view(torch.int32).numpy().view(np.uint32)))

        def trace_benchmark(self, points, attributes, adjacency, offsets, adjacent_diff, camera, start_point, output,
                            weight_threshold=None):
            cam = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in camera.items()}
            img = self.inner.trace_benchmark(points.numpy(), attributes.numpy(), adjacency.view(torch.int32).numpy().view(np.uint32),
                                             offsets.view(torch.int32).numpy().view(np.uint32), adjacent_diff.numpy(), cam,
                                             int(start_point.view(torch.int32).item()), weight_threshold=weight_threshold)
            output.view(torch.int32).copy_(torch.from_numpy(img.view(np.int32)))

    class FakeEvent:
        def __init__(self, enable_timing=False):
            pass

        def record(self):
            pass

        def elapsed_time(self, other):
            return 1.0

    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(product_pipeline, "nearest_point",
                        lambda points, queries: as_u32(emu.nearest_point(points.numpy(), queries.numpy())))

    width, height, fov = 48, 32, 0.9
    c2w = torch.zeros((9, 4, 4))
    for i in range(9):
        cam = foam.camera_dict((3.0 * np.cos(0.5 * i), 3.0 * np.sin(0.5 * i), 1.2), fov=fov, width=width, height=height)
        c2w[i, :3, 0], c2w[i, :3, 1] = torch.from_numpy(cam["right"]), -torch.from_numpy(cam["up"])
        c2w[i, :3, 2], c2w[i, :3, 3] = torch.from_numpy(cam["forward"]), torch.from_numpy(cam["position"])
    cameras, positions = scene_io.benchmark_cameras(c2w, height / (2 * math.tan(fov / 2)), width, height)
    res = scene_io.benchmark_fps(EmulatedPipeline(), scene, cameras, positions, n_reps=1)
    assert res["frames"] == 2 and res["output"].shape == (2, height, width) and res["fps"] > 0
    points, attributes, adjacency, offsets = (t.numpy() if t.dtype != torch.uint32 else t.view(torch.int32).numpy().view(np.uint32)
                                              for t in scene.get_trace_data())
    diff = oracle.prefetch_adjacent_diff(points, adjacency, offsets)
    for k, camera in enumerate(cameras):
        cam = {key: (v.numpy() if isinstance(v, torch.Tensor) else v) for key, v in camera.items()}
        start = foam.nearest_point(points, cam["position"])
        want = oracle.trace_benchmark(points, attributes, adjacency, offsets, diff, cam, start, weight_threshold=0.05)
        a = res["output"][k].view(torch.int32).numpy().view(np.uint8).reshape(height, width, 4).astype(np.int32)
        b = want.view(np.uint8).reshape(height, width, 4).astype(np.int32)
        assert (b[..., :3].sum(axis=-1) > 0).mean() > 0.05
        assert np.abs(a - b).max() <= 1 and (a != b).any(axis=-1).mean() < 1e-2
