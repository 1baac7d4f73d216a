"""The C-ABI library loads, exports exactly what include/ptb200.h declares, and refuses to compute
without a CUDA device (no CPU fallback).  CPU only: no compute call is made with a GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ptb_[a-z0-9_]+|path_tracer_kernel[a-z_]*)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "include/ptb200.h declares %s but libptb200.so does not export it" % n
    out = subprocess.run(["nm", "-D", "--defined-only", api._LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (ptb_[a-z0-9_]+|path_tracer_kernel[a-z_]*)\n", out))
    assert exported == set(names), (exported ^ set(names))


def test_struct_layouts_match_reference_sizes():
    assert ctypes.sizeof(ptb.Camera) == 64 and ptb.MATERIAL_DTYPE.itemsize == 84
    assert ptb.SPHERE_DTYPE.itemsize == 100 and ptb.CONFIG_DTYPE.itemsize == 96


def test_product_does_not_link_or_import_the_oracle():
    deps = subprocess.run(["ldd", api._LIB_PATH], capture_output=True, text=True).stdout
    assert "ptoracle" not in deps and "ptref" not in deps
    for d, _, files in os.walk(os.path.join(ROOT, "pathtracerwithcuda_b200")):
        for fn in files:
            if fn.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                src = open(os.path.join(d, fn), errors="replace").read()
                assert "from oracle" not in src and "import oracle" not in src and "pt_oracle" not in src, fn


def test_no_cpu_fallback(workload_root):
    root, w = workload_root("mix", width=96, height=72)
    r = ptb.Renderer(w["config"], device=-1)          # host-only handle: loaders only
    r.load_scene(w["scene"], root)
    for call in (lambda: r.render(1), lambda: r.trace_batch(np.zeros((1, 6), np.float32)), lambda: r.image_f32(),
                 lambda: r.generate_rays(1), lambda: r.render_strided(1, 2, 1), lambda: r.finalize(1)):
        with pytest.raises(ptb.PtbError):
            call()
    assert r.image_device_ptr() is None
    # the multi-GPU entry points refuse a host-only handle as loudly
    for call in (lambda: r.dist_init(0, 1, b"\0" * 128), lambda: r.dist_render(1), lambda: r.dist_reduce(0), lambda: r.dist_broadcast_scene(0),
                 lambda: r.merged_image_f32(), lambda: r.scene_blob_roundtrip()):
        with pytest.raises(ptb.PtbError):
            call()
    if ptb.device_count() == 0:
        with pytest.raises(ptb.PtbError, match="no CPU fallback"):
            ptb.Renderer(w["config"], device=0)
        with pytest.raises(ptb.PtbError, match="no CPU fallback"):
            ptb.MultiRenderer(w["config"], 2)


def test_path_tracer_class_mirror(workload_root):
    root, w = workload_root("mix", width=96, height=72)
    pt = ptb.PathTracer(device=-1)
    files = pt.init(None, w["config"], "res\\scene", asset_root=root)
    assert any(f.endswith("ptb_mix.json") for f in files)
    assert pt.render() is None                        # no scene initiated -> nullptr (path_tracer.cpp:95-98)
    assert not pt.init_scene_device_data(len(files))  # out-of-range index -> false (scene_parser.cpp:39-42)
    idx = [i for i, f in enumerate(files) if f.endswith("ptb_mix.json")][0]
    assert pt.init_scene_device_data(idx)
    pt.clear()
    pt.release_scene_device_data()
    assert pt.render() is None
