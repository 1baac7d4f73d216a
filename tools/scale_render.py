#!/usr/bin/env python3
"""Strong-scaling record of BASELINE.json configs[4]: a FIXED number of passes of workload c5 (3840x2160, ~5 M triangles in two
meshes) sharded over the ranks through the C ABI (ptb_dist_*: the library's own NCCL communicator, scene broadcast and reduce;
SURVEY.md §8e), timed from ptb_load_scene to the finished image on rank 0.

    torchrun --nproc-per-node N tools/scale_render.py [workload] [total_passes]        (N = 1: python tools/scale_render.py ...)

Rank 0 alone generates / reads / parses the scene files; the other ranks receive the PARSED scene by ncclBroadcast and build their own
BVH.  Prints one JSON line on rank 0: wall seconds of load (parse, broadcast, upload + BVH build), device time (max over ranks) of
render + reduce + finalize, time to image, samples/s with and without the load, image checksum."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr
from pathtracerwithcuda_b200.distributed import DistRenderer, torch_exchange


def run(name, total, rank, local_rank, world, dist, passes_in_flight=8, warm=True, root_dir=None):
    """returns the record (rank 0) or None"""
    torch.cuda.set_device(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    # the scene files exist on rank 0 only; the generated config (a few hundred bytes) is handed to the others
    w = None
    if rank == 0:
        root_dir = root_dir or tempfile.mkdtemp(prefix="ptb_scale_")
        w = pr.make_workload(root_dir, name)
    if dist is not None:
        box = [None if w is None else {k: w[k] for k in ("width", "height", "depth", "triangles", "aperture", "focal", "name")} | {"config_text": open(w["config"]).read()}]
        dist.broadcast_object_list(box, src=0)
        meta = box[0]
    else:
        meta = {k: w[k] for k in ("width", "height", "depth", "triangles", "aperture", "focal", "name")} | {"config_text": open(w["config"]).read()}
    if rank == 0:
        cfg_path = w["config"]
    else:
        cfg_path = os.path.join(tempfile.mkdtemp(prefix="ptb_scale_cfg_%d_" % rank), "config.json")
        with open(cfg_path, "w") as f:
            f.write(meta["config_text"])
    r = ptb.Renderer(cfg_path, device=local_rank)
    r.set_option("passes_in_flight", passes_in_flight)
    dr = DistRenderer(r, rank, world, torch_exchange(dist) if dist is not None else None)
    barrier()
    # ---- time to image starts here: load (rank 0 parses, everyone else receives), build, render, reduce
    t0 = time.perf_counter()
    dr.load_scene(w["scene"] if rank == 0 else "", root_dir if rank == 0 else "")
    if meta["aperture"] >= 0 or meta["focal"] >= 0:
        r.set_camera(ptb.default_camera(meta["width"], meta["height"], meta["aperture"], meta["focal"]))
    barrier()
    load_s = time.perf_counter() - t0
    timings = [dict(dr.timing, rank=rank, bvh_build_ms=r.bvh_info()["build_ms"], upload_ms=r.bvh_info()["upload_ms"])]
    if dist is not None:
        box = [None] * world
        dist.all_gather_object(box, timings[0])
        timings = box
    if warm:
        dr.begin(); dr.render(8 * world); dr.reduce()          # warm-up incl. the first collective (outside the time-to-image below)
        barrier()
    stream = torch.cuda.ExternalStream(r.stream(), device=torch.device("cuda", local_rank))
    dr.begin(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t1 = time.perf_counter()
    e0.record(stream)
    dr.render(total)
    dr.reduce()
    e1.record(stream); barrier()
    render_wall_s = time.perf_counter() - t1
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out = None
    if rank == 0:
        img, passes = dr.image_f32()
        px = meta["width"] * meta["height"]
        out = {"workload": name, "resolution": [meta["width"], meta["height"]], "triangles": meta["triangles"], "total_passes": total, "passes_merged": passes,
               "n_gpus": world, "ms": float(ms.item()), "Msamples_s": px * total / float(ms.item()) / 1e3,
               "scene_load_s": load_s, "render_wall_s": render_wall_s, "time_to_image_s": load_s + render_wall_s,
               "Msamples_s_incl_load": px * total / (load_s + render_wall_s) / 1e6,
               "load_breakdown": timings[:2] + ([timings[-1]] if world > 2 else []),
               "mean_radiance": float(img.mean() / max(passes, 1)), "bvh": r.bvh_info(), "nccl_version": ptb.nccl_version() if world > 1 else None}
    dr.close()
    r.close()
    return out


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "c5"
    total = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dist = None
    sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)          # NCCL banner -> stderr
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out = run(name, total, rank, local_rank, world, dist)
    if rank == 0:
        sys.stdout.flush(); os.dup2(saved, 1)
        print(json.dumps(out)); sys.stdout.flush()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
