#!/usr/bin/env python3
"""Short, ncu-friendly invocation of the hot path: load workload, warm up one step, run one step.
    python tools/profile_step.py [workload] [passes] [layout]
Kernel launch order per step: k_generate, then (k_extend, k_shade) x MaxDepth, k_accumulate, k_tonemap."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
layout = sys.argv[3] if len(sys.argv) > 3 else "2"
persistent = sys.argv[4] if len(sys.argv) > 4 else "1"
root = tempfile.mkdtemp(prefix="ptb_prof_")
w = pr.make_workload(root, name)
r = ptb.Renderer(w["config"], device=0)
in_flight = int(os.environ.get("PTB_IN_FLIGHT", passes))
r.set_option("passes_in_flight", in_flight)
r.set_option("bvh_layout", layout)
r.set_option("extend_persistent", persistent)
for kv in sys.argv[5:]:
    k, v = kv.split("=")
    r.set_option(k, v)
r.load_scene(w["scene"], root)
if w["aperture"] >= 0:
    r.set_camera(ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
r.render(passes)            # warm-up step
r.set_option("profile_stages", 1)
t0 = time.time()
r.render(passes)            # profiled step
st = r.stats()
print("step: %.3f ms gpu (%.3f ms in k_extend), %d segments, %.1f Msamples/s, %.1f Mrays/s extend" % (
    st["gpu_ms_total"], st["gpu_ms_extend"], st["ray_segments"], w["width"] * w["height"] * passes / st["gpu_ms_total"] / 1e3,
    st["ray_segments"] / max(st["gpu_ms_extend"], 1e-9) / 1e3))
seg, ms = r.depth_profile()
print("depth  segments   extend_ms  Mrays/s")
for d in range(len(seg)):
    print("%5d %9d %10.3f %8.1f" % (d, seg[d], ms[d], seg[d] / max(ms[d], 1e-9) / 1e3))
r.set_option("count_traversal", 1)
r.render(passes)
st = r.stats()
print("nodes/segment %.2f  tris/segment %.2f" % ((st["nodes_visited"] + st["wide_nodes_visited"]) / st["ray_segments"], st["tris_tested"] / st["ray_segments"]))
mx, hist = r.traversal_histogram()
print("max node visits of one ray:", mx); print("log2 histogram of node visits per ray:", hist.tolist())
