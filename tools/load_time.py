import sys, time, tempfile
sys.path.insert(0, "/root/repo")
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
for name in sys.argv[1:]:
    root = tempfile.mkdtemp()
    w = pr.make_workload(root, name, width=160, height=90)
    for hyb in (0, 1):
        r = ptb.Renderer(w["config"], device=0)
        r.set_option("passes_in_flight", 1); r.set_option("streams_in_flight", 1); r.set_option("bvh_hybrid", hyb)
        t0 = time.perf_counter(); r.load_scene(w["scene"], root); dt = time.perf_counter() - t0
        i = r.bvh_info()
        print(name, "hybrid", hyb, "load %.3f s, upload+build %.1f ms, gpu build %.1f ms" % (dt, i["upload_ms"], i["build_ms"]), flush=True)
        r.close()
