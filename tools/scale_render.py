#!/usr/bin/env python3
"""Strong-scaling record of BASELINE.json configs[4]: a fixed number of passes of workload c5 (3840x2160, ~5 M triangles in
two meshes) sharded over the ranks, one NCCL sum-reduce of the accumulation buffers at the end.
    torchrun --nproc-per-node N tools/scale_render.py [workload] [total_passes]        (N = 1: python tools/scale_render.py ...)
Prints one JSON line on rank 0: device time (max over ranks) of render + reduce + finalize, samples/s, image checksum."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
from pathtracerwithcuda_b200.distributed import CudaBackend, ShardedRenderer

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
total = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dist = None
sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)          # NCCL banner -> stderr
torch.cuda.set_device(local_rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
root = tempfile.mkdtemp(prefix="ptb_scale_%d_" % rank)
w = pr.make_workload(root, name)
r = ptb.Renderer(w["config"], device=local_rank)
r.set_option("passes_in_flight", 8)
t0 = time.perf_counter(); r.load_scene(w["scene"], root); load_s = time.perf_counter() - t0
sr = ShardedRenderer(CudaBackend(r), rank, world, dist)
sr.begin(); sr.render_local(8); sr.reduce()                    # warm-up incl. the first collective
def barrier():
    torch.cuda.synchronize()
    if dist is not None: dist.barrier()
    torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(r.stream(), device=torch.device("cuda", local_rank))
sr.begin(); barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
sr.render_total(total)
sr.reduce(total)
e1.record(stream); barrier()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
if dist is not None: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    img = r.image_f32()
    out = {"workload": name, "resolution": [w["width"], w["height"]], "triangles": w["triangles"], "total_passes": total, "n_gpus": world,
           "ms": float(ms.item()), "Msamples_s": w["width"] * w["height"] * total / float(ms.item()) / 1e3, "scene_load_s": load_s,
           "mean_radiance": float(img.mean() / total), "bvh": r.bvh_info()}
    sys.stdout.flush(); os.dup2(saved, 1)
    print(json.dumps(out)); sys.stdout.flush()
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
