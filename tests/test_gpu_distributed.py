"""Single-GPU check of the multi-GPU plumbing: two renderers on cuda:0 play ranks 0 and 1 (one
process — separate rank processes on one GPU must not wait on one another), their device
accumulation buffers are summed with torch, and the result must equal the single-renderer image."""
import numpy as np
import pytest

import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200.distributed import CudaBackend, ShardedRenderer

pytestmark = pytest.mark.gpu


def test_sharded_sum_equals_single(workload_root):
    import torch
    root, w = workload_root("mix", width=96, height=72)
    total = 6
    single = ptb.Renderer(w["config"], device=0)
    single.load_scene(w["scene"], root)
    single.render(total)
    ref = single.image_f32()
    ranks = []
    for rank in range(2):
        r = ptb.Renderer(w["config"], device=0)
        r.load_scene(w["scene"], root)
        sr = ShardedRenderer(CudaBackend(r), rank, 2, None)
        sr.begin()
        sr.render_total(total)
        ranks.append(sr)
    t0 = ranks[0].backend.accumulation_tensor()
    t1 = ranks[1].backend.accumulation_tensor()
    assert t0.is_cuda and t0.numel() == 96 * 72 * 3
    torch.cuda.synchronize()
    t0 += t1                                    # what the NCCL sum-reduce does on rank 0
    torch.cuda.synchronize()
    ranks[0].backend.finalize(total)
    got = ranks[0].backend.image_f32()
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)
    assert np.abs(ranks[0].backend.image_u8().astype(int) - single.image_u8().astype(int)).max() <= 1


def test_two_process_nccl_reduce_equals_single(tmp_path):
    """Real two-GPU run (skipped on single-GPU boxes): torchrun x 2 ranks, passes sharded by index, ONE NCCL sum-reduce,
    image finished on rank 0 — the mean radiance must equal the one-process render's up to float re-association.
    (Guards the host-side wait after the collective: NCCL runs on torch's stream, the renderer on its own.)"""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tool = os.path.join(repo, "tools", "scale_render.py")
    env = dict(os.environ, NCCL_DEBUG="WARN")
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29577", tool, "mix", "64"], capture_output=True, text=True, env=env, timeout=600)
    assert two.returncode == 0, two.stderr[-2000:]
    one = subprocess.run([sys.executable, tool, "mix", "64"], capture_output=True, text=True, env=dict(env, CUDA_VISIBLE_DEVICES="0"), timeout=600)
    assert one.returncode == 0, one.stderr[-2000:]
    a = json.loads([l for l in two.stdout.splitlines() if l.startswith("{")][-1])
    b = json.loads([l for l in one.stdout.splitlines() if l.startswith("{")][-1])
    assert a["n_gpus"] == 2 and b["n_gpus"] == 1
    assert abs(a["mean_radiance"] - b["mean_radiance"]) <= 1e-5 * b["mean_radiance"], (a["mean_radiance"], b["mean_radiance"])
