import os, sys, tempfile
sys.path.insert(0, "/root/repo")
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
root = tempfile.mkdtemp(prefix="ptb_sweep_")
w = pr.make_workload(root, "c2")
for pif, streams in [(16, 4), (32, 2), (32, 4), (8, 8), (16, 6), (64, 1), (64,2)]:
    r = ptb.Renderer(w["config"], device=0)
    r.set_option("passes_in_flight", pif); r.set_option("streams_in_flight", streams)
    r.load_scene(w["scene"], root)
    for _ in range(3): r.render(64)
    best = min((r.render(64), r.stats()["gpu_ms_total"])[1] for _ in range(6))
    print("64-pass steps: in flight %2d x streams %d: %.2f ms/step %.0f Msamples/s" % (pif, streams, best, w["width"]*w["height"]*64/best/1e3), flush=True)
    r.close()
