#!/usr/bin/env python3
"""Persistent extend grid size (blocks) on the shipped schedule."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
for name in sys.argv[1:] or ["c2"]:
    root = tempfile.mkdtemp(); w = pr.make_workload(root, name)
    r = ptb.Renderer(w["config"], device=0)
    r.set_option("passes_in_flight", 16); r.set_option("streams_in_flight", 4)
    r.load_scene(w["scene"], root)
    if w["aperture"] >= 0: r.set_camera(ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
    r.render(64)
    for per_sm in (8, 7, 6, 5, 4):
        r.set_option("persistent_grid", 148 * per_sm)
        best = min((r.render(64), r.stats()["gpu_ms_total"])[1] for _ in range(3))
        print(name, "blocks/SM", per_sm, "%.2f ms  %.0f Msamples/s" % (best, w["width"] * w["height"] * 64 / best / 1e3), flush=True)
