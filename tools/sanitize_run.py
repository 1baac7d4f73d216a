#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck): GPU build + collapse, render (binary + wide kernels), edits, NEE."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
import make_golden as mg
root = tempfile.mkdtemp(prefix="ptb_san_")
for name, kw in (("mix", dict(width=96, height=72)), ("c2", dict(width=64, height=36, tri_scale=0.2)), ("c2", dict(width=100, height=70, tri_scale=0.1))):
    w = pr.make_workload(root, name, **kw)
    for opts in (dict(), dict(bvh_layout=8), dict(estimator="nee"), dict(sort_by_material=1, octant_order=1, extend_persistent=0),
                 dict(fused_from_depth=0), dict(fused_from_depth=1, hybrid_from_depth=99), dict(inline_scatter=0), dict(sampler="pcg", sss="per_channel"),
                 dict(texture_filter="hardware"), dict(extend_variant=1), dict(extend_variant=3), dict(extend_variant=4),
                 # where a search starts (csrc/kernels_entry.cuh): entry cuts of every length / tile shape, leaf starts on tiny and on every tree
                 # level, the sky fast path off, the fused walk with leaf starts
                 dict(entry_k=1), dict(entry_k=31, entry_tile="2x2"), dict(entry_tile="64x1", upwalk_min_nodes=1, hybrid_from_depth=99), dict(sky_fast=0, upwalk=0),
                 dict(fused_upwalk=1, upwalk_min_nodes=1), dict(fused_upwalk=1, fused_from_depth=0, upwalk_min_nodes=1), dict(entry_cuts=0, upwalk_min_nodes=1, bvh_max_leaf=1)):
        r = ptb.Renderer(w["config"], device=0)
        for k, v in opts.items():
            r.set_option(k, v)
        r.load_scene(w["scene"], root)
        r.render(3)
        rays = r.generate_rays(1)[::5]
        p, t = r.trace_batch(rays)
        if name == "mix" and not opts:
            for op, a in mg.EDIT_SCRIPT:
                mg.apply_edit(r, op, a)
            r.render(2)
            print("bvh after edits", r.bvh_info()["valid"])
        print(name, opts, float(r.image_f32().mean()), int((p >= 0).sum()), flush=True)
        r.close()
print("done")
