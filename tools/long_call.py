#!/usr/bin/env python3
"""Does a longer call pay?  One synchronous ptb_render(n) drains the 4-stream pipeline at its end; this times n passes as one call and
as n/k calls of k passes (GPU box).    python tools/long_call.py c4 8 4 32 256"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
name, pif, streams, short, long_ = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
root = tempfile.mkdtemp(prefix="ptb_long_")
w = pr.make_workload(root, name)
r = ptb.Renderer(w["config"], device=0)
r.set_option("passes_in_flight", pif); r.set_option("streams_in_flight", streams)
r.load_scene(w["scene"], root)
if w["aperture"] >= 0:
    r.set_camera(ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
px = w["width"] * w["height"]
r.render(short)
for label, per_call in (("short", short), ("long", long_), ("short", short), ("long", long_)):
    t0 = time.perf_counter()
    for _ in range(long_ // per_call):
        r.render(per_call)
    dt = time.perf_counter() - t0
    print("%s %s (%d x %d in flight): %d passes as calls of %d: %.1f ms, %.0f Msamples/s" % (name, label, pif, streams, long_, per_call, dt * 1e3, px * long_ / dt / 1e6), flush=True)
