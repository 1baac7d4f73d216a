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


# the real two-process NCCL run (library-owned communicator, scene broadcast, reduce) lives in tests/test_gpu_multi.py
