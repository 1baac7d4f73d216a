#!/usr/bin/env python3
"""The scene loader under shrinking address-space limits (CPU only): whatever runs out — an allocation in a worker thread, the start
of a thread — has to come back to the caller as an exception (the C ABI turns it into a load error), never std::terminate.
    python tools/fuzz/oom_sweep.py [work_dir]"""
import os, resource, subprocess, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from pathtracerwithcuda_b200 import procedural as pr

work = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="ptb_oom_")
os.makedirs(work, exist_ok=True)
src = os.path.join(REPO, "pathtracerwithcuda_b200", "csrc")
exe = os.path.join(work, "oom")
subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-pthread", "-I" + src, "-I" + os.path.join(REPO, "include"),
                os.path.join(HERE, "fuzz_oom.cpp"), os.path.join(src, "scene_io.cpp"), os.path.join(src, "jpeg_decode.cpp"), "-o", exe], check=True)
root = os.path.join(work, "root")
os.makedirs(root, exist_ok=True)
w = pr.make_workload(root, "c4", width=64, height=36)          # 1 M triangles, 75 MB of OBJ text: 8-16 slices, threaded fill
bad = 0
for mb in (4096, 1200, 900, 800, 750, 700, 650, 600, 560, 520, 480, 440, 400, 300, 200):
    def limit():
        resource.setrlimit(resource.RLIMIT_AS, (mb << 20, mb << 20))
    p = subprocess.run([exe, w["scene"], root], capture_output=True, text=True, preexec_fn=limit)
    last = (p.stdout.strip().splitlines() or [""])[-1][:100]
    abnormal = p.returncode != 0 or "terminate called" in p.stderr
    bad += abnormal
    print("%5d MB: rc %d  %s%s" % (mb, p.returncode, last, "   <-- ABNORMAL EXIT" if abnormal else ""))
print("oom sweep:", "no findings" if not bad else "%d abnormal exits" % bad)
sys.exit(1 if bad else 0)
