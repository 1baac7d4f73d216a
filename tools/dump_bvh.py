"""Debug: print the binary BVH of a small workload built by both builders."""
import sys, tempfile
import numpy as np
sys.path.insert(0, ".")
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr

def dump(nodes, order, ref=0, depth=0, out=None):
    if ref < 0:
        lr = ~ref; first, cnt = lr >> 3, (lr & 7) + 1
        out.append("%sleaf %s" % ("  " * depth, sorted(order[first:first + cnt].tolist())))
        return
    d = nodes[ref]; refs = d[12:14].view(np.int32)
    out.append("%snode c0[%.3f %.3f|%.3f %.3f|%.3f %.3f] c1[%.3f %.3f|%.3f %.3f|%.3f %.3f]" % (("  " * depth,) + tuple(d[[0,1,2,3,8,9,4,5,6,7,10,11]])))
    dump(nodes, order, int(refs[0]), depth + 1, out)
    if not d[4] > d[5]:
        dump(nodes, order, int(refs[1]), depth + 1, out)

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
root = tempfile.mkdtemp()
w = pr.make_workload(root, name, width=64, height=64)
outs = {}
for b in ("gpu_sah", "host_sah"):
    r = ptb.Renderer(w["config"], device=0)
    r.set_option("bvh_builder", b)
    r.load_scene(w["scene"], root)
    nodes, order = r.bvh_download()
    o = []
    dump(nodes, order, 0, 0, o)
    outs[b] = o
    print(b, r.bvh_info())
for a, b in zip(outs["gpu_sah"], outs["host_sah"]):
    print(("  " if a == b else "!!") + a)
    if a != b:
        print("  " + b)
