#!/usr/bin/env python3
"""Sweeps builder parameters / queue order on one workload (GPU box): serial extend ms and overlapped step ms.
    python tools/sweep_tree.py c2 "tile_order=0" "tile_order=1" "bvh_max_leaf=6,bvh_intersect_cost=1.0" ...
Each argument is one configuration (comma-separated key=value options applied before load_scene)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr

name = sys.argv[1]
configs = sys.argv[2:] or [""]
root = tempfile.mkdtemp(prefix="ptb_sweep_")
w = pr.make_workload(root, name)
for cfg in configs:
    r = ptb.Renderer(w["config"], device=0)
    r.set_option("passes_in_flight", int(os.environ.get("PIF", 8)))
    r.set_option("streams_in_flight", int(os.environ.get("STREAMS", 4)))
    for kv in [c for c in cfg.split(",") if c]:
        k, v = kv.split("=")
        r.set_option(k, v)
    r.load_scene(w["scene"], root)
    if w["aperture"] >= 0:
        r.set_camera(ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
    try:
        info = r.bvh_info()
    except ptb.PtbError:
        info = {"sah_cost": float("nan"), "leaves": 0, "build_ms": float("nan")}
    r.render(32)
    r.set_option("active_streams", 1); r.set_option("profile_stages", 1)
    best = None
    for _ in range(3):
        r.render(32)
        st = r.stats()
        if best is None or st["gpu_ms_extend"] < best[0]:
            best = (st["gpu_ms_extend"], st["gpu_ms_total"], st["ray_segments"])
    seg, ms = r.depth_profile()
    r.set_option("active_streams", 0); r.set_option("profile_stages", 0)
    over = min(_t for _t in [(r.render(32), r.stats()["gpu_ms_total"])[1] for _ in range(4)])
    r.set_option("count_traversal", 1)
    r.render(8)
    st = r.stats()
    print("%-46s serial: extend %.2f step %.2f ms | overlapped step %.2f ms (%.0f Msamples/s) | d0 %.2f d1 %.2f d2 %.2f | nodes %.1f tris %.2f | sah %.2f leaves %d build %.1f ms"
          % (cfg or "(default)", best[0], best[1], over, w["width"] * w["height"] * 32 / over / 1e3, ms[0], ms[1], ms[2],
             (st["nodes_visited"] + st["wide_nodes_visited"]) / st["ray_segments"], st["tris_tested"] / st["ray_segments"], info["sah_cost"], info["leaves"], info["build_ms"]), flush=True)
    r.close()
