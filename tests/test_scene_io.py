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
