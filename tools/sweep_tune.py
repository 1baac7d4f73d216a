#!/usr/bin/env python3
"""Sweeps the voting thresholds of the persistent extend kernels (refill / leaf / node repetitions) on the shipped schedule."""
import itertools, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
root = tempfile.mkdtemp(prefix="ptb_tune_")
w = pr.make_workload(root, name)
r = ptb.Renderer(w["config"], device=0)
r.set_option("passes_in_flight", 16); r.set_option("streams_in_flight", 4)
r.load_scene(w["scene"], root)
if w["aperture"] >= 0:
    r.set_camera(ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
r.render(64)
res = []
grid = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 12, 16, 20]
leafs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4, 8, 12, 16]
repss = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [2, 3, 4]
for refill, leaf, reps in itertools.product(grid, leafs, repss):
    r.set_option("tune_refill", refill); r.set_option("tune_leaf", leaf); r.set_option("tune_reps", reps)
    best = min((r.render(64), r.stats()["gpu_ms_total"])[1] for _ in range(3))
    res.append((best, refill, leaf, reps))
    print("refill %2d leaf %2d reps %d: %.2f ms / 64 passes  %.0f Msamples/s" % (refill, leaf, reps, best, w["width"] * w["height"] * 64 / best / 1e3), flush=True)
res.sort()
print("best:", res[:5])
