"""Driver of the CPU kernel-logic emulator (tests/emu/cuda_emu.h) -- TEST INFRASTRUCTURE ONLY.

Builds the product's C-ABI translation unit (radfoam_b200/csrc/radfoam_b200.cu with its kernels) for host threads
into tests/emu/libradfoam_b200_emu.so and calls it through the same ctypes table as the product
(radfoam_b200/_lib.py: SIGNATURES), with numpy arrays standing in for device memory.  The product never loads this
library (radfoam_b200/_lib.py only knows libradfoam_b200.so) and has no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from radfoam_b200 import _lib as product_abi  # the ctypes TABLE only; product_abi.load() is never called here

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
# RFB_EMU_ASAN=1: an AddressSanitizer build (run python under LD_PRELOAD=libasan.so, ASAN_OPTIONS=detect_leaks=0):
# out-of-bounds accesses of the kernels to "shared" or "global" memory abort with a report
ASAN = bool(os.environ.get("RFB_EMU_ASAN"))
LIB = os.path.join(HERE, "libradfoam_b200_emu_asan.so" if ASAN else "libradfoam_b200_emu.so")
_lib = None


def _sources():
    csrc = os.path.join(ROOT, "radfoam_b200", "csrc")
    return ([os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, "include", "radfoam_b200.h")]
            + [os.path.join(HERE, f) for f in ("cuda_emu.h", "emu_lib.cpp")])


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in _sources()):
        return LIB
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    tmp = f"{LIB}.{os.getpid()}.tmp"  # atomic replace: concurrent test processes may build at the same time
    extra = ["-fsanitize=address", "-fno-omit-frame-pointer", "-g"] if ASAN else ["-g0"]
    subprocess.check_call([cxx, "-std=c++17", "-O1", *extra, "-DRFB_EMU", "-ffp-contract=off", "-mfma", "-mf16c",
                           "-fPIC", "-shared", "-pthread", "-w", f"-I{cuda}/include",
                           os.path.join(HERE, "emu_lib.cpp"), "-o", tmp])
    os.replace(tmp, LIB)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        for name, (restype, argtypes) in product_abi.SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = lib
    return _lib


def _check(rc):
    if rc:
        raise RuntimeError(load().rfb_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype=None):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class EmuPipeline:
    """numpy-in / numpy-out calls into the emulated library, argument for argument like radfoam_b200.Pipeline."""

    def __init__(self, sh_degree: int, attr_dtype=np.float32):
        self.lib = load()
        self.half = np.dtype(attr_dtype) == np.float16
        self.dtype = np.float16 if self.half else np.float32
        self.handle = ctypes.c_void_p()
        _check(self.lib.rfb_create_pipeline(int(sh_degree), 1 if self.half else 0, ctypes.byref(self.handle)))
        self.attr_dim = int(self.lib.rfb_attribute_dim(self.handle))
        self._keep = []

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.rfb_destroy_pipeline(self.handle)
            self.handle = None

    def _scene(self, points, attributes, adjacency, offsets):
        return (_c(points, np.float32), _c(attributes, self.dtype), _c(adjacency, np.uint32), _c(offsets, np.uint32))

    @staticmethod
    def _settings(weight_threshold, max_intersections):
        return product_abi.TraceSettings(0.001 if weight_threshold is None else weight_threshold,
                                         1024 if max_intersections is None else max_intersections)

    def trace_forward(self, points, attributes, adjacency, offsets, rays, start, depth_quantiles=None,
                      weight_threshold=None, max_intersections=None, return_contribution=False, scene_version=0,
                      record_tape=False):
        pts, att, adj, off = self._scene(points, attributes, adjacency, offsets)
        rays_c, start_c, dq = _c(rays, np.float32), _c(start, np.uint32), _c(depth_quantiles, np.float32)
        batch, n, r = rays_c.shape[:-1], pts.shape[0], rays_c.size // 6
        q = 0 if dq is None else dq.shape[-1]
        out = {"rgba": np.empty(batch + (4,), self.dtype), "num_intersections": np.empty(batch + (1,), np.uint32)}
        if dq is not None:
            out["depth"] = np.empty(batch + (q,), np.float32)
            out["depth_indices"] = np.empty(batch + (q,), np.uint32)
        if return_contribution:
            out["contribution"] = np.zeros((n, 1), self.dtype)
        width = rays_c.shape[-2] if rays_c.ndim >= 3 else 0
        opts = product_abi.LaunchOpts(scene_version, width, product_abi.FLAG_RECORD_TAPE if record_tape else 0)
        settings = self._settings(weight_threshold, max_intersections)
        self._keep = [pts, att, adj, off, rays_c, start_c, dq]  # the tape key compares these pointers
        _check(self.lib.rfb_trace_forward(
            self.handle, ctypes.byref(settings), n, _p(pts), _p(att), adj.size, _p(adj), _p(off), r, _p(rays_c),
            _p(start_c), q, _p(dq), _p(out["rgba"]), _p(out.get("depth")), _p(out.get("depth_indices")),
            _p(out["num_intersections"]), _p(out.get("contribution")), ctypes.byref(opts), None))
        return out

    def trace_backward(self, points, attributes, adjacency, offsets, rays, start, rgba, grad_rgba,
                       depth_quantiles=None, depth_indices=None, grad_depth=None, ray_error=None,
                       weight_threshold=None, max_intersections=None, scene_version=0, use_tape=False,
                       scrub_nonfinite=False):
        if use_tape:  # same buffers as the recording forward, as the product's Python layer guarantees
            pts, att, adj, off, rays_c, start_c, dq = self._keep
        else:
            pts, att, adj, off = self._scene(points, attributes, adjacency, offsets)
            rays_c, start_c, dq = _c(rays, np.float32), _c(start, np.uint32), _c(depth_quantiles, np.float32)
        n, r = pts.shape[0], rays_c.size // 6
        q = 0 if dq is None else dq.shape[-1]
        rgba_c, g_c = _c(rgba, self.dtype), _c(grad_rgba, self.dtype)
        di, gd, err = _c(depth_indices, np.uint32), _c(grad_depth, np.float32), _c(ray_error, self.dtype)
        out = {"points_grad": np.empty((n, 3), np.float32), "attr_grad": np.empty((n, self.attr_dim), self.dtype)}
        if err is not None:
            out["point_error"] = np.zeros((n, 1), self.dtype)
        width = rays_c.shape[-2] if rays_c.ndim >= 3 else 0
        flags = (product_abi.FLAG_USE_TAPE if use_tape else 0) | (product_abi.FLAG_SCRUB_NONFINITE if scrub_nonfinite else 0)
        opts = product_abi.LaunchOpts(scene_version, width, flags)
        settings = self._settings(weight_threshold, max_intersections)
        ray_grad = np.zeros_like(rays_c)
        _check(self.lib.rfb_trace_backward(
            self.handle, ctypes.byref(settings), n, _p(pts), _p(att), adj.size, _p(adj), _p(off), r, _p(rays_c),
            _p(start_c), q, _p(dq), _p(di), _p(rgba_c), _p(g_c), _p(gd), _p(err), _p(ray_grad),
            _p(out["points_grad"]), _p(out["attr_grad"]), _p(out.get("point_error")), ctypes.byref(opts), None))
        return out

    def trace_backward_accumulate(self, points, attributes, adjacency, offsets, rays, start, rgba, grad_rgba,
                                  depth_quantiles=None, depth_indices=None, grad_depth=None,
                                  weight_threshold=None, max_intersections=None):
        """The first half of the split backward (ray-sharded multi-GPU use): returns a VIEW of the pipeline's fp32
        accumulator [N, grad_row] that the caller may sum across ranks before trace_backward_finalize."""
        pts, att, adj, off = self._scene(points, attributes, adjacency, offsets)
        rays_c, start_c, dq = _c(rays, np.float32), _c(start, np.uint32), _c(depth_quantiles, np.float32)
        n, r = pts.shape[0], rays_c.size // 6
        q = 0 if dq is None else dq.shape[-1]
        rgba_c, g_c = _c(rgba, self.dtype), _c(grad_rgba, self.dtype)
        di, gd = _c(depth_indices, np.uint32), _c(grad_depth, np.float32)
        width = rays_c.shape[-2] if rays_c.ndim >= 3 else 0
        opts = product_abi.LaunchOpts(0, width, 0)
        settings = self._settings(weight_threshold, max_intersections)
        _check(self.lib.rfb_trace_backward_accumulate(
            self.handle, ctypes.byref(settings), n, _p(pts), _p(att), adj.size, _p(adj), _p(off), r, _p(rays_c),
            _p(start_c), q, _p(dq), _p(di), _p(rgba_c), _p(g_c), _p(gd), None, None, ctypes.byref(opts), None))
        ptr, count = ctypes.c_void_p(), ctypes.c_uint64()
        _check(self.lib.rfb_grad_accumulator(self.handle, ctypes.byref(ptr), ctypes.byref(count)))
        row = int(self.lib.rfb_grad_row_floats(self.handle))
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), (count.value // row, row))

    def trace_backward_finalize(self, num_points, scrub_nonfinite=False):
        out = {"points_grad": np.empty((num_points, 3), np.float32),
               "attr_grad": np.empty((num_points, self.attr_dim), self.dtype)}
        _check(self.lib.rfb_trace_backward_finalize(
            self.handle, num_points, _p(out["points_grad"]), _p(out["attr_grad"]),
            product_abi.FLAG_SCRUB_NONFINITE if scrub_nonfinite else 0, None))
        return out

    def bind_scene_params(self, att_dc, att_sh, density, activation_scale=1.0):
        """rfb_bind_scene_params (None unbinds); the arrays are kept alive here, like the product's Python layer does."""
        if att_dc is None:
            self._params = None
            _check(self.lib.rfb_bind_scene_params(self.handle, None))
            return
        self._params = [_c(a, np.float32) for a in (att_dc, att_sh, density)]
        sp = product_abi.SceneParams(self._params[0].ctypes.data, self._params[1].ctypes.data if self._params[1].size else None,
                                     self._params[2].ctypes.data, float(activation_scale))
        _check(self.lib.rfb_bind_scene_params(self.handle, ctypes.byref(sp)))

    def trace_backward_finalize_params(self, num_points, scrub_nonfinite=False):
        out = {"points_grad": np.empty((num_points, 3), np.float32), "att_dc_grad": np.empty((num_points, 3), np.float32),
               "att_sh_grad": np.empty((num_points, self.attr_dim - 4), np.float32),
               "density_grad": np.empty((num_points, 1), np.float32)}
        _check(self.lib.rfb_trace_backward_finalize_params(
            self.handle, num_points, _p(out["points_grad"]), _p(out["att_dc_grad"]),
            _p(out["att_sh_grad"]) if out["att_sh_grad"].size else None, _p(out["density_grad"]),
            product_abi.FLAG_SCRUB_NONFINITE if scrub_nonfinite else 0, None))
        return out

    def reduce_finalize_peers(self, rank, accumulators, scrub_nonfinite=False, multicast=False):
        """rfb_reduce_finalize_peers with `accumulators` (one [N, grad_row] float32 array per rank) standing in for
        the peer-mapped buffers; returns THIS rank's writes into every rank's outputs as (attr_grads, points_grads),
        lists of arrays pre-filled with a sentinel so that the rows a rank does not own are recognisable.
        `multicast`: go through the emulated NVSwitch multicast group instead of the peer pointers."""
        world, n = len(accumulators), accumulators[0].shape[0]
        gr = accumulators[0].shape[1]
        itemsize = np.dtype(self.dtype).itemsize
        acc_b, attr_b, pts_b = n * gr * 4, (n * self.attr_dim * itemsize + 15) // 16 * 16, (n * 3 * 4 + 15) // 16 * 16
        # one allocation per rank laid out like radfoam_b200/sharded.py's symmetric buffer: [acc | attr | pts]
        bufs = [np.zeros(acc_b + attr_b + pts_b, np.uint8) for _ in range(world)]
        accs, attr, pts = [], [], []
        for w in range(world):
            a = bufs[w][:acc_b].view(np.float32).reshape(n, gr)
            a[...] = accumulators[w]
            accs.append(a)
            t = bufs[w][acc_b:acc_b + n * self.attr_dim * itemsize].view(self.dtype).reshape(n, self.attr_dim)
            t[...] = 7.0
            attr.append(t)
            q = bufs[w][acc_b + attr_b:acc_b + attr_b + n * 12].view(np.float32).reshape(n, 3)
            q[...] = 7.0
            pts.append(q)
        table = lambda arrs: (ctypes.c_void_p * world)(*[a.ctypes.data for a in arrs])  # noqa: E731
        mc = None
        if multicast:
            self.lib.rfe_set_multicast.restype = ctypes.c_void_p
            self.lib.rfe_set_multicast.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint32, ctypes.c_uint64]
            base = self.lib.rfe_set_multicast(table(bufs), world, acc_b + attr_b + pts_b)
            mc = ctypes.byref(product_abi.Multicast(base, base + acc_b, base + acc_b + attr_b))
        _check(self.lib.rfb_reduce_finalize_peers(
            self.handle, world, rank, n, table(accs), table(attr), table(pts), mc,
            product_abi.FLAG_SCRUB_NONFINITE if scrub_nonfinite else 0, None))
        self._keep_peers = bufs
        return attr, pts

    def trace_benchmark(self, points, attributes, adjacency, offsets, adjacent_diff, camera, start_point,
                        weight_threshold=None, max_intersections=None, scene_version=0):
        pts, att, adj, off = self._scene(points, attributes, adjacency, offsets)
        diff = _c(adjacent_diff, np.float16)
        cam = product_abi.Camera()
        for key in ("position", "forward", "right", "up"):
            getattr(cam, key)[:] = [float(v) for v in np.asarray(camera[key]).reshape(-1)[:3]]
        cam.fov, cam.width, cam.height = float(camera["fov"]), int(camera["width"]), int(camera["height"])
        cam.model = 0 if camera["model"] == "pinhole" else 1
        start = np.array([start_point], dtype=np.uint32)
        out = np.zeros((cam.height, cam.width), dtype=np.uint32)
        opts = product_abi.LaunchOpts(scene_version, 0, 0)
        settings = self._settings(weight_threshold, max_intersections)
        _check(self.lib.rfb_trace_benchmark(self.handle, ctypes.byref(settings), pts.shape[0], _p(pts), _p(att),
                                            _p(adj), _p(off), _p(diff), ctypes.byref(cam), _p(start), _p(out),
                                            ctypes.byref(opts), None))
        return out

    def tape_status(self):
        cap, used, over = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        _check(self.lib.rfb_tape_status(self.handle, ctypes.byref(cap), ctypes.byref(used), ctypes.byref(over)))
        return {"capacity_chunks": cap.value, "used_chunks": used.value, "overflowed": bool(over.value)}


def prefetch_adjacent_diff(points, adjacency, offsets):
    pts, adj, off = _c(points, np.float32), _c(adjacency, np.uint32), _c(offsets, np.uint32)
    out = np.zeros((adj.size, 4), np.float16)
    _check(load().rfb_prefetch_adjacent_diff(_p(pts), pts.shape[0], adj.size, _p(adj), _p(off), _p(out), None))
    return out


def nearest_point(points, queries):
    pts, q = _c(points, np.float32), _c(queries, np.float32).reshape(-1, 3)
    out = np.zeros((q.shape[0],), np.uint32)
    _check(load().rfb_nearest_point(_p(pts), pts.shape[0], _p(q), q.shape[0], _p(out), None))
    return out


def start_points(points, rays):
    pts, r = _c(points, np.float32), _c(rays, np.float32)
    out = np.zeros(r.shape[:-1], np.uint32)
    _check(load().rfb_start_points(_p(pts), pts.shape[0], _p(r), r.size // 6, _p(out), None))
    return out


def farthest_neighbor(points, adjacency, offsets):
    pts, adj, off = _c(points, np.float32), _c(adjacency, np.uint32), _c(offsets, np.uint32)
    idx, radius = np.zeros((pts.shape[0],), np.uint32), np.zeros((pts.shape[0],), np.float32)
    _check(load().rfb_farthest_neighbor(_p(pts), pts.shape[0], _p(adj), _p(off), _p(idx), _p(radius), None))
    return idx, radius


def tape_records(pipe: EmuPipeline, height: int, width: int):
    """Decode the walk tape of the last recording forward of an image-shaped batch (debug / analysis only).
    Returns (cells[W][S][32] uint32, t1[W][S][32] float32, count[W][32] int) in the backward's own layout:
    W warps (8x4-pixel tiles, 4 per 16x8 CTA), S = longest walk, lane-major; cells == 0xFFFFFFFF past a lane's end."""
    lib = load()
    lib.rfe_tape_buffers.restype = ctypes.c_int
    pool, table, per_ray = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    stride, cap = ctypes.c_uint32(), ctypes.c_uint32()
    if lib.rfe_tape_buffers(pipe.handle, ctypes.byref(pool), ctypes.byref(table), ctypes.byref(per_ray),
                            ctypes.byref(stride), ctypes.byref(cap)):
        raise RuntimeError("no recorded tape")
    blocks_x, blocks_y = (width + 15) // 16, (height + 7) // 8
    nwarps = blocks_x * blocks_y * 4
    pool_a = np.ctypeslib.as_array(ctypes.cast(pool, ctypes.POINTER(ctypes.c_uint32)), (cap.value, 32, 32, 2))
    table_a = np.ctypeslib.as_array(ctypes.cast(table, ctypes.POINTER(ctypes.c_uint32)), (nwarps, stride.value))
    per_ray_a = np.ctypeslib.as_array(ctypes.cast(per_ray, ctypes.POINTER(ctypes.c_uint32)), (height * width, 2))
    count = np.zeros((nwarps, 32), dtype=np.int64)
    for w in range(nwarps):
        block, warp = divmod(w, 4)
        bx, by = block % blocks_x, block // blocks_x
        for lane in range(32):
            x, y = bx * 16 + (warp & 1) * 8 + (lane & 7), by * 8 + (warp >> 1) * 4 + (lane >> 3)
            if x < width and y < height:
                count[w, lane] = per_ray_a[y * width + x, 0]
    steps = int(count.max())
    cells = np.full((nwarps, steps, 32), 0xFFFFFFFF, dtype=np.uint32)
    t1 = np.zeros((nwarps, steps, 32), dtype=np.float32)
    for w in range(nwarps):
        n = int(count[w].max())
        for c in range((n + 31) // 32):
            chunk = pool_a[table_a[w, c]]                      # [32 steps][32 lanes][2]
            hi = min(32, n - 32 * c)
            cells[w, 32 * c:32 * c + hi] = chunk[:hi, :, 0]
            t1[w, 32 * c:32 * c + hi] = chunk[:hi, :, 1].view(np.float32)
        valid = np.arange(steps)[:, None] < count[w][None, :]
        cells[w][~valid] = 0xFFFFFFFF
        t1[w][~valid] = 0.0
    return cells, t1, count


def tape_schedule(pipe: EmuPipeline):
    """(tile_steps, order) of the last recording forward's replay schedule, or None when it was not scheduled."""
    lib = load()
    steps, order, blocks = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint32()
    if lib.rfe_tape_schedule(pipe.handle, ctypes.byref(steps), ctypes.byref(order), ctypes.byref(blocks)):
        return None
    as_arr = lambda p: np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint32)), (blocks.value,)).copy()  # noqa: E731
    return as_arr(steps), as_arr(order)


def red_counters(reset: bool = True) -> dict:
    """16-byte / 8-byte global reductions and warp-collective operations (shuffles, votes, matches, syncs) the
    emulated kernels executed since the last reset."""
    v4, v2, coll = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    load().rfe_counters(ctypes.byref(v4), ctypes.byref(v2), ctypes.byref(coll), 1 if reset else 0)
    return {"red_v4": v4.value, "red_v2": v2.value, "bytes": 16 * v4.value + 8 * v2.value,
            "warp_collectives": coll.value}
