#!/usr/bin/env python3
"""Sweeps the persistent extend kernel's tuning knobs on one workload (GPU box)."""
import itertools, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
passes = 8
root = tempfile.mkdtemp(prefix="ptb_sweep_")
w = pr.make_workload(root, name)
r = ptb.Renderer(w["config"], device=0)
r.set_option("passes_in_flight", passes)
r.load_scene(w["scene"], root)
r.set_option("profile_stages", 1)
r.render(passes)
def run(**kw):
    for k, v in kw.items():
        r.set_option(k, v)
    best = None
    for _ in range(3):
        r.render(passes)
        st = r.stats()
        if best is None or st["gpu_ms_extend"] < best[0]:
            best = (st["gpu_ms_extend"], st["gpu_ms_total"])
    print("%-60s extend %.3f ms  step %.3f ms" % (kw, best[0], best[1]), flush=True)
run(extend_persistent=0)
run(extend_persistent=1, tune_refill=8, tune_leaf=10, tune_reps=1)
for refill, leaf, reps in itertools.product([4, 8, 12, 16], [6, 10, 14], [1, 2, 3]):
    run(tune_refill=refill, tune_leaf=leaf, tune_reps=reps)
for grid in [148 * 4, 148 * 6, 148 * 8, 148 * 12, 148 * 16]:
    run(tune_refill=8, tune_leaf=10, tune_reps=2, persistent_grid=grid)
