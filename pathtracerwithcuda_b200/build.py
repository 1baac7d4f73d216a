"""Builds pathtracerwithcuda_b200/libptb200.so IN-TREE for sm_100a (nvcc cross-compiles without a GPU).

    python -m pathtracerwithcuda_b200.build [--force]

Host sources are compiled with -ffp-contract=off: the scene front-end must reproduce the
reference's un-fused IEEE binary32 transform arithmetic bit for bit (csrc/scene_io.cpp).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libptb200.so")
OBJ = os.path.join(HERE, "build")

HOST_SOURCES = ["scene_io.cpp", "bvh_host.cpp", "image_out.cpp", "jpeg_decode.cpp"]
CUDA_SOURCES = ["render.cu", "bvh_build.cu"]
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _hash_sources():
    h = hashlib.sha1()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for fn in sorted(os.listdir(root)):
            with open(os.path.join(root, fn), "rb") as f:
                h.update(fn.encode() + f.read())
    with open(__file__, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False, extra_nvcc=None, out=None):
    """extra_nvcc / out: build an experimental variant (e.g. ["-DPTB_PERSISTENT_MIN_BLOCKS=12"]) next to the default library."""
    global OUT, OBJ
    if out:
        OUT = out
        OBJ = os.path.join(HERE, "build", os.path.basename(out) + ".d")
        force = True
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    digest = _hash_sources()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == digest:
        return OUT
    jobs = []
    for src in HOST_SOURCES:
        obj = os.path.join(OBJ, src + ".o")
        jobs.append((obj, ["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fopenmp", "-Wall",
                           "-c", os.path.join(CSRC, src), "-o", obj]))
    for src in CUDA_SOURCES:
        obj = os.path.join(OBJ, src + ".o")
        jobs.append((obj, ["nvcc", "-std=c++17", "-O3", "-lineinfo"] + NVCC_ARCH +
                     (extra_nvcc or []) + ["-Xcompiler", "-fPIC,-fopenmp", "-Xptxas", "-v" if verbose else "-warn-spills",
                      "-c", os.path.join(CSRC, src), "-o", obj]))

    def run(job):
        obj, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return obj, cmd, r

    objs = []
    with ThreadPoolExecutor(max_workers=4) as ex:
        for obj, cmd, r in ex.map(run, jobs):
            if r.returncode != 0:
                raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
            if verbose:
                sys.stderr.write(r.stderr)
            objs.append(obj)
    link = ["nvcc", "-shared"] + NVCC_ARCH + ["-Xcompiler", "-fPIC,-fopenmp", "-o", OUT] + objs + ["-lgomp"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return OUT


ROOFS_SRC = os.path.join(HERE, "..", "tools", "roofs", "roofs.cu")
ROOFS_OUT = os.path.join(HERE, "..", "tools", "roofs", "libptbroofs.so")


def build_roofs(force=False):
    """tools/roofs/libptbroofs.so: the FP32 / L2-gather / L1-gather roof microbenchmarks bench.py runs before its timed region
    (measurement tooling; the product library neither links nor loads it)."""
    src, out = os.path.normpath(ROOFS_SRC), os.path.normpath(ROOFS_OUT)
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    cmd = ["nvcc", "-std=c++17", "-O3", "-lineinfo"] + NVCC_ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", out, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return out


def build_fastobj(force=False):
    """tools/fastobj/libfastobj.so: C stdio writer for the procedural OBJ fixtures (byte-identical to the numpy writer, much faster)."""
    src = os.path.normpath(os.path.join(HERE, "..", "tools", "fastobj", "fastobj.c"))
    out = os.path.normpath(os.path.join(HERE, "..", "tools", "fastobj", "libfastobj.so"))
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    r = subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", out, src], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: fastobj\n%s\n%s" % (r.stdout, r.stderr))
    return out


if __name__ == "__main__":
    extra = []
    for a in sys.argv[1:]:
        if a.startswith("-D"):
            extra.append(a)
        elif a.startswith("--nvcc="):      # e.g. --nvcc=-Xptxas,-fmad=false  ->  -Xptxas -fmad=false
            extra += a.split("=", 1)[1].split(",")
    out = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    build_roofs(force="--force" in sys.argv)
    build_fastobj(force="--force" in sys.argv)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, extra_nvcc=extra or None, out=out[0] if out else None))
