#!/usr/bin/env python3
"""Full-size check of a large workload (default c5: 3840x2160, ~5M triangles in two meshes) on one GPU:
scene load breakdown, GPU BVH build facts, traversal == exhaustive scan on sampled rays, determinism, rate.
    python tools/check_scale.py [workload] [passes] > profiles/..json"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 16
root = tempfile.mkdtemp(prefix="ptb_scale_")
t0 = time.perf_counter(); w = pr.make_workload(root, name); gen_s = time.perf_counter() - t0
r = ptb.Renderer(w["config"], device=0)
r.set_option("passes_in_flight", 8)
t0 = time.perf_counter(); r.load_scene(w["scene"], root); load_s = time.perf_counter() - t0
info = r.bvh_info()
out = {"workload": name, "resolution": [w["width"], w["height"]], "triangles": r.scene_counts()["triangles"], "meshes": r.scene_counts()["meshes"],
       "generate_obj_s": gen_s, "load_scene_s": load_s, "bvh": info}
rays = np.concatenate([r.generate_rays(1)[::20011], r.capture_rays(1, 2)[1][::4001]], 0)
prim, t = r.trace_batch(rays)
bp, bt = r.trace_batch(rays, bruteforce=True)
out["rays_checked_against_exhaustive_scan"] = int(rays.shape[0])
out["id_mismatches"] = int((prim != bp).sum()); out["t_bit_mismatches"] = int((t.view(np.uint32) != bt.view(np.uint32)).sum())
out["hit_fraction"] = float((prim != -1).mean())
r.render(passes)
a = r.image_f32().copy(); st = r.stats()
r.clear(); r.render(passes)
out["deterministic"] = bool(np.array_equal(a.view(np.uint32), r.image_f32().view(np.uint32)))
r.render(passes); r.render(passes)
st = r.stats()
out["passes"] = passes; out["gpu_ms"] = st["gpu_ms_total"]; out["ray_segments"] = st["ray_segments"]
out["Msamples_s"] = w["width"] * w["height"] * passes / st["gpu_ms_total"] / 1e3
out["Mrays_s"] = st["ray_segments"] / st["gpu_ms_total"] / 1e3
out["finite"] = bool(np.isfinite(a).all()); out["mean_radiance"] = float(a.mean() / passes)
print(json.dumps(out))
