"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and
exports every symbol include/radfoam_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

import common  # noqa: F401  (sets up sys.path via conftest)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "radfoam_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rfb_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_reference_entry_points():
    syms = header_symbols()
    for needed in ("rfb_create_pipeline", "rfb_trace_forward", "rfb_trace_backward",
                   "rfb_trace_benchmark", "rfb_prefetch_adjacent_diff", "rfb_attribute_dim",
                   "rfb_attribute_type", "rfb_last_error"):
        assert needed in syms


def test_library_exports_every_declared_symbol():
    from radfoam_b200 import _lib

    lib = ctypes.CDLL(_lib.library_path())
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_pipeline_lifecycle_without_gpu():
    from radfoam_b200 import _lib

    lib = _lib.load()
    assert lib.rfb_abi_version() == 1
    for deg, dim in ((0, 4), (1, 13), (2, 28), (3, 49)):
        for dtype in (0, 1):
            h = ctypes.c_void_p()
            assert lib.rfb_create_pipeline(deg, dtype, ctypes.byref(h)) == 0
            assert lib.rfb_attribute_dim(h) == dim  # pipeline.cu:767-769
            assert lib.rfb_attribute_type(h) == dtype
            assert lib.rfb_grad_row_floats(h) == ((3 * (deg + 1) ** 2 + 3) // 4) * 4 + 4
            lib.rfb_destroy_pipeline(h)
    h = ctypes.c_void_p()
    assert lib.rfb_create_pipeline(4, 0, ctypes.byref(h)) != 0
    assert b"Unsupported SH degree" in lib.rfb_last_error()
    assert lib.rfb_create_pipeline(3, 7, ctypes.byref(h)) != 0
    assert b"Unsupported attribute type" in lib.rfb_last_error()


def test_missing_library_is_a_hard_error(monkeypatch):
    from radfoam_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_LIB_NAME", "libdoes_not_exist.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    """oracle/ is a checker: nothing under radfoam_b200/ may import, link or call it."""
    pkg = os.path.join(ROOT, "radfoam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "radfoam_oracle" not in text and "libradfoam_ref" not in text, f
                # nor the CPU kernel-logic emulator of tests/emu (the RFB_EMU guards in csrc are compile-time only)
                assert "libradfoam_b200_emu" not in text and "cuda_emu" not in text, f


def test_only_tests_smoke_and_bench_touch_the_oracle():
    """Outside tests/ and oracle/ itself, only bench.py (cpu_baseline / --impl reference legs) and
    __graft_entry__.py (build + smoke's check) may mention the oracle."""
    allowed = {"bench.py", "__graft_entry__.py"}
    for dirpath, dirs, files in os.walk(ROOT):
        rel = os.path.relpath(dirpath, ROOT)
        dirs[:] = [d for d in dirs if not d.startswith(".") and d not in ("gpurun_out", "__pycache__")
                   and os.path.join(rel, d) not in ("./tests", "./oracle", "./baseline")]
        for f in files:
            if f.endswith(".py") and os.path.join(rel, f).lstrip("./") not in allowed:
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(rel, f)


def test_pipeline_methods_take_the_reference_binding_arguments():
    """Argument names / order / defaults of the pybind11 Pipeline the reference's Python code calls
    (torch_bindings/pipeline_bindings.cpp:626-672), so radfoam_model/render.py:33-42, 80-93 and
    benchmark.py:99-108 work unchanged against this class."""
    import inspect

    from radfoam_b200.pipeline import Pipeline, create_pipeline

    def names(fn):
        return [p.name for p in inspect.signature(fn).parameters.values()][1:]

    assert names(Pipeline.trace_forward) == [
        "points", "attributes", "point_adjacency", "point_adjacency_offsets", "rays", "start_point",
        "depth_quantiles", "weight_threshold", "max_intersections", "return_contribution"]
    assert names(Pipeline.trace_backward)[:14] == [
        "points", "attributes", "point_adjacency", "point_adjacency_offsets", "rays", "start_point",
        "rgb_out", "grad_in", "depth_quantiles", "depth_indices", "depth_grad_in", "ray_error",
        "weight_threshold", "max_intersections"]
    assert names(Pipeline.trace_benchmark) == [
        "points", "attributes", "point_adjacency", "point_adjacency_offsets", "adjacent_diff", "camera",
        "start_point", "output_rgba", "weight_threshold", "max_intersections"]
    sig = inspect.signature(Pipeline.trace_forward).parameters
    assert sig["depth_quantiles"].default is None and sig["return_contribution"].default is False
    assert inspect.signature(create_pipeline).parameters["attr_dtype"].default == "float32"


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++ in the signatures) and a C program that
    binds the path's entry points must link against the library and run without a GPU (argument checks only)."""
    import subprocess

    from radfoam_b200 import _lib

    src = tmp_path / "bind.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "radfoam_b200.h"
int main(void) {
    rfb_pipeline *p = NULL;
    rfb_trace_settings s = {0.001f, 1024};
    rfb_launch_opts o = {0, 0, 0};
    rfb_scene_params sp = {NULL, NULL, NULL, 1.0f};
    rfb_multicast mc = {NULL, NULL, NULL};
    (void)sp; (void)mc;
    if (rfb_abi_version() != RFB_ABI_VERSION) return 1;
    if (rfb_create_pipeline(3, RFB_FLOAT32, &p) != 0 || rfb_attribute_dim(p) != 49) return 2;
    /* NULL scene arrays are rejected before anything touches the device */
    if (rfb_trace_forward(p, &s, 1, NULL, NULL, 0, NULL, NULL, 1, NULL, NULL, 0, NULL, NULL, NULL, NULL, NULL, NULL, &o, NULL) == 0) return 3;
    if (!strstr(rfb_last_error(), "NULL")) return 4;
    if (rfb_create_pipeline(9, RFB_FLOAT32, &p) == 0) return 5;
    puts(rfb_last_error());
    return 0;
}
''')
    exe = tmp_path / "bind"
    lib_dir = os.path.dirname(_lib.library_path())
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{os.path.join(ROOT, 'include')}", str(src),
                           "-o", str(exe), f"-L{lib_dir}", "-lradfoam_b200", f"-Wl,-rpath,{lib_dir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "Unsupported SH degree" in out.stdout
