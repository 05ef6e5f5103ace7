"""CPU tests of the host-side mirror of the reference's pipeline bindings (radfoam_b200/pipeline.py):
validation errors (pipeline_bindings.cpp:8-71, 139-175), the scene-mirror cache key and the walk-tape
key.  The C library loads without a GPU; nothing here launches a kernel."""
import pytest
import torch

import common  # noqa: F401
import radfoam_b200
from radfoam_b200 import _lib


def scene(n=8, e=24, a=49, dtype=torch.float32):
    return (torch.zeros(n, 3), torch.zeros(n, a, dtype=dtype), torch.zeros(e, dtype=torch.uint32),
            torch.zeros(n + 1, dtype=torch.uint32))


def test_create_pipeline_accepts_the_reference_dtype_spellings():
    for spelling in ("float32", torch.float32, "float"):
        assert radfoam_b200.create_pipeline(3, spelling).attribute_type() == torch.float32
    for spelling in ("float16", torch.float16, "half"):
        assert radfoam_b200.create_pipeline(0, spelling).attribute_type() == torch.float16
    assert [radfoam_b200.create_pipeline(d).attribute_dim() for d in range(4)] == [4, 13, 28, 49]
    for bad in (-1, 4, 2.0):
        with pytest.raises(RuntimeError, match="Unsupported SH degree"):
            radfoam_b200.create_pipeline(bad)
    with pytest.raises(RuntimeError, match="Unsupported attribute type"):
        radfoam_b200.create_pipeline(3, "float64")
    with pytest.raises(RuntimeError, match="dtype must be a string or torch.dtype"):
        radfoam_b200.create_pipeline(3, 32)


def test_validation_order_and_messages_match_the_bindings():
    pipe = radfoam_b200.create_pipeline(3)
    pts, attrs, adj, off = scene()
    rays, start = torch.zeros(4, 6), torch.zeros(4, dtype=torch.uint32)
    cases = [
        ((pts[:, :2], attrs, adj, off), "points had dimension 2 along axis -1, expected 3"),
        ((pts.double(), attrs, adj, off), "points had dtype float64, expected float32"),
        ((pts, attrs, adj, off), "points must be on CUDA device"),
    ]
    for sc, msg in cases:
        with pytest.raises(RuntimeError, match=msg):
            pipe.trace_forward(*sc, rays, start)
    # the remaining scene checks, reached with the device check patched out of the way
    class OnCuda(torch.Tensor):
        @property
        def device(self):
            return torch.device("cuda", 0)

    def cuda(t):
        return t.as_subclass(OnCuda)

    p, a, j, o = map(cuda, (pts, attrs, adj, off))
    for sc, msg in [
        ((p, cuda(torch.zeros(8, 28)), j, o), "attributes had dimension 28 along axis -1, expected 49"),
        ((p, cuda(torch.zeros(7, 49)), j, o), "attributes must have the same number of rows as points"),
        ((p, cuda(attrs.half()), j, o), "attributes had dtype float16, expected float32"),
        ((p, a, j, cuda(off.to(torch.int32))), "point_adjacency_offsets must have uint32 dtype"),
        ((p, a, j, cuda(off[:-1])), r"point_adjacency_offsets must have num_points \+ 1 elements"),
        ((p, a, cuda(adj.to(torch.int64)), o), "point_adjacency must have uint32 dtype"),
    ]:
        with pytest.raises(RuntimeError, match=msg):
            pipe.trace_forward(*sc, cuda(rays), cuda(start))
    for kw, msg in [
        (dict(rays=cuda(torch.zeros(4, 5))), "rays must have 6 as the last dimension"),
        (dict(rays=cuda(torch.zeros(4, 6, dtype=torch.float64))), "rays must have float32 dtype"),
        (dict(start=cuda(torch.zeros(3, dtype=torch.uint32))), "start_point must have the same batch size as rays"),
        (dict(start=cuda(torch.zeros(4, dtype=torch.int64))), "start_point must have uint32 dtype"),
    ]:
        args = dict(rays=cuda(rays), start=cuda(start))
        args.update(kw)
        with pytest.raises(RuntimeError, match=msg):
            pipe.trace_forward(p, a, j, o, args["rays"], args["start"])


def test_scene_cache_key_follows_identity_and_version_counters():
    pipe = radfoam_b200.create_pipeline(3)
    sc = scene()
    rays = torch.zeros(2, 3, 6)
    v1 = pipe._opts(sc, rays).scene_version
    assert v1 != 0 and pipe._opts(sc, rays).scene_version == v1            # same tensors: reuse
    assert pipe._opts(sc, rays).image_width == 3 and pipe._opts(sc, rays.reshape(-1, 6)).image_width == 0
    sc[1].add_(1.0)                                                         # optimizer.step()-style update
    v2 = pipe._opts(sc, rays).scene_version
    assert v2 != v1
    new_attrs = sc[1].clone()                                               # a new tensor, same values
    v3 = pipe._opts((sc[0], new_attrs, sc[2], sc[3]), rays).scene_version
    assert v3 != v2
    pipe.invalidate_cache()
    assert pipe._opts((sc[0], new_attrs, sc[2], sc[3]), rays).scene_version != v3
    pipe.cache_scene = False
    assert pipe._opts(sc, rays).scene_version == 0                          # 0 = always rebuild


def test_tape_is_offered_only_for_the_tensors_of_the_recording_forward():
    import weakref

    pipe = radfoam_b200.create_pipeline(3)
    rays, start = torch.zeros(4, 6), torch.zeros(4, dtype=torch.uint32)
    assert pipe._tape_flag(rays, start, 5) == 0                             # nothing recorded yet
    pipe._tape_refs = [(weakref.ref(rays), rays._version), (weakref.ref(start), start._version), 5]
    assert pipe._tape_flag(rays, start, 5) == _lib.FLAG_USE_TAPE
    assert pipe._tape_flag(rays, start, 6) == 0                             # scene changed since
    assert pipe._tape_flag(rays.clone(), start, 5) == 0                     # other ray tensor
    rays.add_(1.0)
    assert pipe._tape_flag(rays, start, 5) == 0                             # rays modified in place
