import sys, tempfile
sys.path.insert(0, "/root/repo")
import numpy as np
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
root = tempfile.mkdtemp(); w = pr.make_workload(root, "c5", width=480, height=270, tri_scale=0.02)
r = ptb.Renderer(w["config"], device=0); r.load_scene(w["scene"], root)
means = []
for first, n in ((1, 256), (257, 256), (513, 512), (1, 1024)):
    r.clear(); r.render_strided(first, 1, n); means.append((first, n, float(r.image_f32().mean() / n)))
# sharded emulation: 8 "ranks" on one GPU, summed on the host
acc = np.zeros((270, 480, 3), np.float64)
for rank in range(8):
    r.clear(); r.render_strided(rank + 1, 8, 128); acc += r.image_f32()
means.append(("8x128 strided sum", 1024, float(acc.mean() / 1024)))
print(means)
