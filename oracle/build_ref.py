#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — builds the UNMODIFIED reference (BlauHimmel/PathTracerWithCuda) headless
for sm_100a into oracle/_ref/libptref.so, compiling its sources WHERE THEY LIE under
/root/reference (nothing is copied into this repo).

The reference is a Visual-Studio project: its `#include`s use '\\' as the
path separator and it needs <Windows.h>, <io.h> and FreeImage.  We deal with that without touching
the sources:
  * for every `#include "A\\B.h"` / `<thrust\\x.h>` spelled with a backslash we generate a one-line
    forwarding header whose FILE NAME literally contains the backslash (legal on Linux) in a
    scratch include directory, pointing at the real file;
  * oracle/ref_shim/ supplies Windows.h, io.h (glob-based _findfirst), a FreeImage stand-in and the
    C-ABI harness ref_driver.cu (which textually includes Kernel/path_tracer_kernel.cu).
Flags follow the reference project (gpu_path_tracer.vcxproj: -O2/Full, default -fmad, NO fast-math)
retargeted to compute_100a.

Outputs: oracle/_ref/libptref.so only (git-ignored, travels to the GPU box with gpurun).
Usage: python oracle/build_ref.py [--force]
"""
import os
import re
import subprocess
import sys
import hashlib
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PTB_REFERENCE_ROOT", "/root/reference")
GPT = os.path.join(REF, "gpu_path_tracer")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libptref.so")
SHIM = os.path.join(HERE, "ref_shim")

SOURCES = [
    (os.path.join(SHIM, "ref_driver.cu"), "ref_driver"),
    (os.path.join(GPT, "Kernel/parallel_function.cu"), "parallel_function"),
    (os.path.join(GPT, "Kernel/bvh_morton_code_kernel.cu"), "bvh_morton_code_kernel"),
    (os.path.join(GPT, "Bvh/bvh.cpp"), "bvh"),
    (os.path.join(GPT, "Bvh/bvh_build_config.cpp"), "bvh_build_config"),
    (os.path.join(GPT, "Core/material.cpp"), "material"),
    (os.path.join(GPT, "Core/image.cpp"), "image"),
    (os.path.join(GPT, "Core/camera.cpp"), "camera"),
    (os.path.join(GPT, "Core/config_parser.cpp"), "config_parser"),
    (os.path.join(GPT, "Core/triangle_mesh.cpp"), "triangle_mesh"),
    (os.path.join(GPT, "Core/cube_map_loader.cpp"), "cube_map_loader"),
    (os.path.join(GPT, "Core/scene_parser.cpp"), "scene_parser"),
    (os.path.join(GPT, "Others/image_loader.cpp"), "image_loader"),
    (os.path.join(GPT, "lib/tiny_obj_loader/tiny_obj_loader.cc"), "tiny_obj_loader"),
    (os.path.join(GPT, "lib/lodepng/lodepng.cpp"), "lodepng"),
    (os.path.join(SHIM, "freeimage_shim.cpp"), "freeimage_shim"),
]

INC_RE = re.compile(r'^\s*#\s*include\s*([<"])([^>"]*\\[^>"]*)[>"]', re.M)


def make_forwarders(fwd_dir):
    """One forwarding header per backslash-spelled include found in the reference's own code."""
    names = set()
    for sub in ("Kernel", "Core", "Bvh", "Math", "Others", "Main"):
        d = os.path.join(GPT, sub)
        for fn in os.listdir(d):
            p = os.path.join(d, fn)
            try:
                txt = open(p, "r", encoding="latin-1").read()
            except OSError:
                continue
            for m in INC_RE.finditer(txt):
                names.add((m.group(1), m.group(2)))
    for fn in os.listdir(SHIM):
        txt = open(os.path.join(SHIM, fn), "r", encoding="latin-1").read()
        for m in INC_RE.finditer(txt):
            names.add((m.group(1), m.group(2)))
    for kind, name in sorted(names):
        real = name.replace("\\", "/")
        target = os.path.join(fwd_dir, name)  # file name contains literal backslashes
        if kind == '"':
            body = '#include "%s"\n' % os.path.join(GPT, real)
        else:
            body = "#include <%s>\n" % real
        with open(target, "w") as f:
            f.write(body)
    return len(names)


def main():
    force = "--force" in sys.argv
    if not os.path.isdir(GPT):
        print("[build_ref] %s not present; keeping prebuilt %s" % (GPT, OUT))
        return 0 if os.path.exists(OUT) else 1
    os.makedirs(OUT_DIR, exist_ok=True)
    h = hashlib.sha1()
    for src, _ in SOURCES:
        h.update(open(src, "rb").read())
    for fn in sorted(os.listdir(SHIM)):
        h.update(open(os.path.join(SHIM, fn), "rb").read())
    h.update(open(__file__, "rb").read())
    stamp = os.path.join(OUT_DIR, "libptref.stamp")
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        print("[build_ref] up to date:", OUT)
        return 0

    work = os.path.join(tempfile.gettempdir(), "ptb_ref_build")
    fwd = os.path.join(work, "fwd")
    os.makedirs(fwd, exist_ok=True)
    n = make_forwarders(fwd)
    print("[build_ref] %d forwarding headers in %s" % (n, fwd))

    base = [
        "nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
        "-x", "cu", "-Xcompiler", "-fPIC,-fopenmp,-w", "-w", "-diag-suppress", "20012",
        "-I", fwd, "-I", GPT, "-I", os.path.join(GPT, "lib/json"), "-I", SHIM,
        # thrust/CUB first: the reference's `#define E 2.718...` (Math/basic_math.hpp:20) breaks them otherwise
        "-include", "thrust/device_vector.h", "-include", "thrust/sort.h", "-include", "thrust/remove.h",
        "-include", "thrust/random.h", "-include", "thrust/execution_policy.h",
        "-DGLM_ENABLE_EXPERIMENTAL", "-DFREEIMAGE_LIB",
    ]
    # two variants: the real one (runs on the GPU box) and a host-only one whose managed
    # allocations live on the heap, so the reference's scene pipeline can run in this container
    variants = [("libptref.so", []), ("libptref_host.so", ["-include", os.path.join(SHIM, "host_cuda_stub.h")])]

    jobs = []
    for lib, extra in variants:
        vdir = os.path.join(work, lib.replace(".so", ""))
        os.makedirs(vdir, exist_ok=True)
        for src, name in SOURCES:
            jobs.append((lib, base + extra + ["-c", src, "-o", os.path.join(vdir, name + ".o")], os.path.join(vdir, name + ".o"), name))

    def compile_one(job):
        lib, cmd, obj, name = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return lib, obj, name, r

    objs = {lib: [] for lib, _ in variants}
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for lib, obj, name, r in ex.map(compile_one, jobs):
            if r.returncode != 0:
                sys.stderr.write("[build_ref] FAILED %s (%s)\n%s\n%s\n" % (name, lib, r.stdout[-4000:], r.stderr[-8000:]))
                return 1
            objs[lib].append(obj)
    for lib, _ in variants:
        link = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC,-fopenmp",
                "-o", os.path.join(OUT_DIR, lib)] + objs[lib] + ["-lgomp"]
        r = subprocess.run(link, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write("[build_ref] LINK FAILED %s\n%s\n%s\n" % (lib, r.stdout, r.stderr))
            return 1
        print("[build_ref] built", os.path.join(OUT_DIR, lib))
    stage_assets()
    open(stamp, "w").write(h.hexdigest())
    return 0


# Reference data files (scene/config JSON, OBJ, textures) that the parity tests and the
# reference bench arm read on the GPU box, where /root/reference does not exist. Staged into the
# git-ignored oracle/_ref/res (never committed); big 2048^2 JPG cube maps are left out.
STAGE = ["scene", "configuration", "obj", "texture/lobby", "texture/lake", "texture/tex_cube", "texture/vanille"]


def stage_assets():
    import shutil
    dst_root = os.path.join(OUT_DIR, "res")
    for rel in STAGE:
        src = os.path.join(GPT, "res", rel)
        dst = os.path.join(dst_root, rel)
        if not os.path.isdir(src):
            continue
        os.makedirs(dst, exist_ok=True)
        for fn in os.listdir(src):
            s, d = os.path.join(src, fn), os.path.join(dst, fn)
            if os.path.isfile(s) and not (os.path.exists(d) and os.path.getsize(d) == os.path.getsize(s)):
                shutil.copyfile(s, d)
    print("[build_ref] staged reference res/ ->", dst_root)


if __name__ == "__main__":
    sys.exit(main())
