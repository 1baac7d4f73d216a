#!/usr/bin/env python3
"""Counts non-finite rays in the live batches of a workload (they walk the whole tree in a NaN-blind traversal)."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
root = tempfile.mkdtemp()
w = pr.make_workload(root, name)
r = ptb.Renderer(w["config"], device=0)
r.load_scene(w["scene"], root)
for d in range(w["depth"]):
    pix, rays = r.capture_rays(1, d)
    bad = ~np.isfinite(rays).all(axis=1)
    zero = (np.abs(rays[:, 3:]).sum(axis=1) == 0)
    print("depth", d, "rays", len(pix), "non-finite", int(bad.sum()), "zero-dir", int(zero.sum()))
    if bad.any():
        print("  e.g.", rays[bad][:3])
