"""Drop-in boundary: the REFERENCE'S OWN autograd op -- radfoam_model/render.py::TraceRays, unmodified --
driving radfoam_b200.Pipeline, compared with the same op driving the reference's own kernels (oracle/_ref).

The file is loaded from /root/reference when that exists (this container) or from its byte-compiled copy
oracle/_ref/radfoam_model_render.pyc (built by `make -C oracle ref_py`; travels to the GPU box like the
reference's .so).  Nothing of it is re-typed here."""
import importlib.machinery
import importlib.util
import os

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/radfoam_model/render.py"
REF_PYC = os.path.join(ROOT, "oracle", "_ref", "radfoam_model_render.pyc")


def load_reference_op():
    if os.path.exists(REF_SRC):
        loader = importlib.machinery.SourceFileLoader("radfoam_reference_render", REF_SRC)
    elif os.path.exists(REF_PYC):
        loader = importlib.machinery.SourcelessFileLoader("radfoam_reference_render", REF_PYC)
    else:
        pytest.skip("neither /root/reference nor oracle/_ref/radfoam_model_render.pyc is present")
    spec = importlib.util.spec_from_loader(loader.name, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


class ReferencePipeline:
    """The reference's kernels behind the dict API its pybind module gives render.py."""

    def __init__(self):
        from oracle import ref_gpu

        if not ref_gpu.available():
            pytest.skip("oracle/_ref not built")
        self.ref = ref_gpu

    def trace_forward(self, *args, **kwargs):
        return self.ref.trace_forward(*args, **kwargs)

    def trace_backward(self, *args, **kwargs):
        return self.ref.trace_backward(*args, **kwargs)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def drive(torch, op, pipeline, case, with_error, contribution=False):
    """One training-shaped step through `op` (an autograd Function with the reference's contract)."""
    f = case.foam
    dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    points = dev(f.points).requires_grad_(True)
    base = dev(f.attributes).requires_grad_(True)
    attributes = base * 1.0  # a non-leaf, like get_trace_data()'s cat (scene.py:202-217)
    out = op.apply(pipeline, points, attributes, dev(f.adjacency), dev(f.offsets), dev(case.rays), dev(case.start),
                   dev(case.quantiles), contribution)
    rgba, depth, contrib, nint, box = out
    if with_error:
        rng = np.random.default_rng(3)
        box.ray_error = dev(rng.uniform(0.0, 1.0, size=case.rays.shape[:-1] + (1,)).astype(np.float32))
    loss = (rgba * dev(case.grad_rgba)).sum()
    if depth is not None:
        loss = loss + (depth * dev(case.grad_depth)).sum()
    loss.backward()
    res = {"rgba": rgba.detach(), "num_intersections": nint, "points_grad": points.grad, "attr_grad": base.grad}
    if depth is not None:
        res["depth"] = depth.detach()
    if contrib is not None:
        res["contribution"] = contrib.detach()
    if with_error:
        res["point_error"] = box.point_error
    return {k: v.cpu().numpy() for k, v in res.items()}


def compare(got, ref):
    assert np.array_equal(got["num_intersections"], ref["num_intersections"])
    np.testing.assert_allclose(got["rgba"], ref["rgba"], rtol=1e-5, atol=1e-5)
    if "depth" in ref:
        np.testing.assert_allclose(got["depth"], ref["depth"], rtol=1e-5, atol=1e-5)
    for k in ("points_grad", "attr_grad", "point_error", "contribution"):
        if k in ref:
            assert common.nonfinite_mismatch(got[k], ref[k]) == 0, k
            assert common.grad_error(got[k], ref[k]) <= 1e-5, k


@pytest.mark.parametrize("with_error", [False, True])
def test_reference_render_op_runs_unmodified_on_our_pipeline(torch_cuda, with_error):
    import radfoam_b200

    mod = load_reference_op()
    case = common.scene_case(20000, 160, 96, q=2)
    ours = drive(torch_cuda, mod.TraceRays, radfoam_b200.create_pipeline(3, "float32"), case, with_error)
    ref = drive(torch_cuda, mod.TraceRays, ReferencePipeline(), case, with_error)
    compare(ours, ref)


def test_reference_render_op_records_the_walk_tape(torch_cuda):
    """Inside the reference's Function.forward grad mode is off; the pipeline must still see that a backward is
    coming (attributes is a non-leaf that requires grad) and replay the tape -- and must NOT record under no_grad."""
    import radfoam_b200

    torch = torch_cuda
    mod = load_reference_op()
    case = common.scene_case(20000, 160, 96, q=2)
    pipe = radfoam_b200.create_pipeline(3, "float32")
    drive(torch, mod.TraceRays, pipe, case, False)
    assert pipe.tape_status()["used_chunks"] > 0
    pipe2 = radfoam_b200.create_pipeline(3, "float32")
    f = case.foam
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    points = torch.nn.Parameter(dev(f.points))
    with torch.no_grad():  # eval render of the reference: attributes are built under no_grad
        attributes = dev(f.attributes) * 1.0
        mod.TraceRays.apply(pipe2, points, attributes, dev(f.adjacency), dev(f.offsets), dev(case.rays),
                            dev(case.start), None, False)
    with pytest.raises(RuntimeError, match="no recorded tape"):
        pipe2.tape_status()


@pytest.mark.parametrize("with_error", [False, True])
def test_native_op_matches_reference_op(torch_cuda, with_error):
    """radfoam_b200.TraceRays (written independently: save_for_backward, in-kernel scrub) == the reference's op."""
    import radfoam_b200

    mod = load_reference_op()
    case = common.scene_case(20000, 160, 96, q=2)
    ours = drive(torch_cuda, radfoam_b200.TraceRays, radfoam_b200.create_pipeline(3, "float32"), case, with_error,
                 contribution=True)
    ref = drive(torch_cuda, mod.TraceRays, ReferencePipeline(), case, with_error, contribution=True)
    compare(ours, ref)


def test_native_op_leaves_no_reference_cycle(torch_cuda):
    """The step's graph must die by reference counting (no tensor -> grad_fn -> ctx -> tensor cycle)."""
    import gc
    import weakref

    import radfoam_b200

    torch = torch_cuda
    case = common.config1()
    f = case.foam
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    pipe = radfoam_b200.create_pipeline(3, "float32")
    points, attrs = dev(f.points).requires_grad_(True), dev(f.attributes).requires_grad_(True)
    gc.disable()
    try:
        rgba, depth, _, _, _ = radfoam_b200.TraceRays.apply(pipe, points, attrs, dev(f.adjacency), dev(f.offsets),
                                                            dev(case.rays), dev(case.start), dev(case.quantiles), False)
        ref = weakref.ref(rgba)
        (rgba.sum() + depth.sum()).backward()
        del rgba, depth
        assert ref() is None, "rgba survived: a reference cycle keeps the step's tensors alive until gc runs"
    finally:
        gc.enable()
