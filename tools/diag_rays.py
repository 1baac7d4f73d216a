#!/usr/bin/env python3
"""GPU diagnostic: where do the wavefront's depth-d rays differ from the reference's (golden)?"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr

name = sys.argv[1] if len(sys.argv) > 1 else "mix"
kw = {"mix": dict(width=96, height=72), "c1": dict(width=64, height=64)}[name]
g = np.load(os.path.join(ROOT, "tests", "golden", "ref_gpu_%s.npz" % name))
root = tempfile.mkdtemp()
w = pr.make_workload(root, name, **kw)
r = ptb.Renderer(w["config"], device=0)
r.load_scene(w["scene"], root)
r.set_camera(g["camera"].view(np.float32))
mats = r.scene_materials(); tri, tmat = r.scene_triangles(); sph = r.scene_spheres()
prim0 = dict(zip(g["depth0_pixels"].tolist(), g["depth0_prim"].tolist()))
for d in (1, 2):
    pix, mine = r.capture_rays(1, d)
    order = np.argsort(g["depth%d_pixels" % d])
    ref = g["depth%d_rays" % d].view(np.float32)[order]
    neq = (mine.view(np.uint32) != ref.view(np.uint32))
    bad = np.nonzero(neq.any(axis=1))[0]
    print("depth", d, "rays", len(pix), "differ", len(bad), "origin-differs", int(neq[:, :3].any(axis=1).sum()), "dir-differs", int(neq[:, 3:].any(axis=1).sum()))
    if d == 1:
        kinds = {}
        for i in bad:
            p = prim0[int(pix[i])]
            if p < -1:
                m = sph["mat"][-(p + 2)]; key = "sphere%d" % (-(p + 2))
            else:
                m = mats[tmat[p]]; key = "tri mat%d" % tmat[p]
            key += " T%d k%.2f" % (m["is_transparent"], m["extinction_coefficient"])
            kinds[key] = kinds.get(key, 0) + 1
        tot = {}
        for i in range(len(pix)):
            p = prim0[int(pix[i])]
            key = ("sphere%d" % (-(p + 2))) if p < -1 else ("tri mat%d" % tmat[p])
            tot[key] = tot.get(key, 0) + 1
        print(" by first-hit material (differ):", kinds)
        print(" totals:", tot)
        for i in bad[:12]:
            ulp = np.abs(mine[i].view(np.int32).astype(np.int64) - ref[i].view(np.int32).astype(np.int64))
            print("  pix", pix[i], "prim", prim0[int(pix[i])], "ulp", ulp.tolist(), "mine", mine[i], "ref", ref[i])
