import os, sys, tempfile
sys.path.insert(0, "/root/repo")
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
for name in ("c2", "c3"):
    root = tempfile.mkdtemp(); w = pr.make_workload(root, name)
    r = ptb.Renderer(w["config"], device=0)
    r.set_option("passes_in_flight", 16); r.set_option("streams_in_flight", 4)
    r.load_scene(w["scene"], root)
    if w["aperture"] >= 0: r.set_camera(ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
    r.render(64)
    for u in (1, 0, 1, 0):
        r.set_option("unroll_reps", u)
        best = min((r.render(64), r.stats()["gpu_ms_total"])[1] for _ in range(3))
        print(name, "unroll", u, "%.2f ms  %.0f Msamples/s" % (best, w["width"] * w["height"] * 64 / best / 1e3), flush=True)
