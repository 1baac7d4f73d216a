#!/usr/bin/env python3
"""ncu-friendly invocation of the GPU BVH builder only: small image (little device memory to save/restore between
replays), full-size mesh.   python tools/profile_build.py [workload]"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
root = tempfile.mkdtemp(prefix="ptb_prof_")
w = pr.make_workload(root, name, width=160, height=90)
r = ptb.Renderer(w["config"], device=0)
r.set_option("passes_in_flight", 1)
r.set_option("streams_in_flight", 1)
t0 = time.time()
r.load_scene(w["scene"], root)
print("load %.3f s" % (time.time() - t0), r.bvh_info())
