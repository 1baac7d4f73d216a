#!/usr/bin/env python3
"""Fraction of live rays at depths 1..3 that are bit-identical to the reference kernels' (golden fixtures)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
for name, kw in (("mix", dict(width=96, height=72)), ("c1", dict(width=64, height=64))):
    root = tempfile.mkdtemp()
    w = pr.make_workload(root, name, **kw)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_gpu_%s.npz" % name))
    r = ptb.Renderer(w["config"], device=0)
    r.load_scene(w["scene"], root)
    r.set_camera(g["camera"].view(np.float32))
    out = []
    for d in range(1, 4):
        rays = g["depth%d_rays" % d].view(np.float32)
        pix, mine = r.capture_rays(1, d)
        order = np.argsort(g["depth%d_pixels" % d])
        out.append("d%d %.4f" % (d, (mine.view(np.uint32) == rays[order].view(np.uint32)).all(axis=1).mean()))
    r.render(4)
    ref = g["image_sum"].view(np.float32)
    img = r.image_f32()
    rel = np.abs(img.astype(np.float64) - ref) / np.maximum(np.abs(ref), 1e-3)
    print(os.environ.get("PTB200_LIB", "default").split("/")[-1], name, " ".join(out), "image bit-equal %.4f outliers>1e-3 %.2e" % ((img == ref).mean(), (rel > 1e-3).mean()), flush=True)
