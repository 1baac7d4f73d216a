import os, sys, tempfile, time
sys.path.insert(0, "/root/repo")
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
print("cpus", os.cpu_count(), len(os.sched_getaffinity(0)))
root = tempfile.mkdtemp(prefix="ptb_lt_")
w = pr.make_workload(root, "c5")
for th in (16, 16, 24, 32, 48, 64, 16):
    r = ptb.Renderer(w["config"], device=-1)
    r.set_option("loader_threads", th)
    t0 = time.perf_counter(); r.load_scene(w["scene"], root); dt = time.perf_counter() - t0
    print("loader_threads %d: host parse %.3f s" % (th, dt), flush=True)
    r.close()
