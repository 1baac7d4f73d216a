#!/usr/bin/env python3
"""Sweeps (passes_in_flight, streams_in_flight) on one workload: ms per pass of a 64- or 128-pass render call."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
total = int(sys.argv[2]) if len(sys.argv) > 2 else 128
root = tempfile.mkdtemp(prefix="ptb_sweep_")
w = pr.make_workload(root, name)
for pif, streams in [(8, 4), (8, 8), (16, 2), (16, 4), (32, 1), (32, 2), (32, 4), (64, 1), (64, 2)]:
    r = ptb.Renderer(w["config"], device=0)
    r.set_option("passes_in_flight", pif)
    r.set_option("streams_in_flight", streams)
    r.load_scene(w["scene"], root)
    if w["aperture"] >= 0:
        r.set_camera(ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
    r.render(total)
    best = min((r.render(total), r.stats()["gpu_ms_total"])[1] for _ in range(3))
    print("%s in flight %2d x streams %d: %.3f ms/pass  %.0f Msamples/s  (state %.1f GB)" % (name, pif, streams, best / total,
          w["width"] * w["height"] * total / best / 1e3, w["width"] * w["height"] * pif * streams * 88 / 1e9), flush=True)
    r.close()
