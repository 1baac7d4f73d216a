#!/usr/bin/env python3
"""GPU BVH build time (CUDA events inside the library) for a few workloads; best of 3 loads."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
for name in sys.argv[1:] or ["c2", "c4"]:
    root = tempfile.mkdtemp(prefix="ptb_tb_")
    w = pr.make_workload(root, name, width=160, height=90)
    best, info = None, None
    for _ in range(3):
        r = ptb.Renderer(w["config"], device=0)
        r.set_option("passes_in_flight", 1); r.set_option("streams_in_flight", 1)
        t0 = time.perf_counter(); r.load_scene(w["scene"], root); dt = time.perf_counter() - t0
        i = r.bvh_info()
        if best is None or i["build_ms"] < best:
            best, info = i["build_ms"], i
        r.close()
    print("%s %s: build %.2f ms (levels %d, sub-trees %d, depth %d, sah %.3f, valid %s), load %.2f s" % (
        os.environ.get("PTB200_LIB", "default").split("/")[-1], name, best, info["levels"], info["small_subtrees"], info["depth"], info["sah_cost"], info["valid"], dt), flush=True)
