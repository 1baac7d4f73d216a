"""N>1 host logic on CPU: two gloo ranks shard passes, render them with the ORACLE as the stand-in
compute backend (there is no GPU here and the product has no CPU path), sum-reduce, and must
reproduce the single-process image.  Also unit-tests the shard arithmetic."""
import os
import socket
import sys

import numpy as np
import pytest

from pathtracerwithcuda_b200.distributed import ShardedRenderer, shard_passes


def test_shard_passes_partition():
    for total in (0, 1, 2, 7, 8, 9, 256, 4096):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                first, stride, count = shard_passes(total, r, world)
                seen += [first + k * stride for k in range(count)]
            assert sorted(seen) == list(range(1, total + 1)), (total, world)


def test_c_abi_shards_partition_the_passes():
    """csrc/multi.inc: the shard arithmetic behind ptb_multi_render / ptb_dist_render (host-only hook), incl. progressive calls."""
    import ctypes
    import pathtracerwithcuda_b200 as ptb
    L = ptb.load_library()
    L.ptb_test_shard.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_int)] * 2
    for world in (1, 2, 3, 8):
        done, seen = 0, []
        for total in (0, 1, 5, 8, 64, 7):
            for r in range(world):
                first, count = ctypes.c_int(), ctypes.c_int()
                assert L.ptb_test_shard(done, total, r, world, ctypes.byref(first), ctypes.byref(count)) == 0
                seen += [first.value + k * world for k in range(count.value)]
                assert (first.value, world, count.value) == tuple(x + (done if i == 0 else 0) for i, x in enumerate(shard_passes(total, r, world)))
            done += total
        assert sorted(seen) == list(range(1, done + 1)), world      # every pass exactly once, seeds as on one GPU
    assert L.ptb_test_shard(0, 4, 2, 2, ctypes.byref(ctypes.c_int()), ctypes.byref(ctypes.c_int())) == 1


class OracleBackend:
    def __init__(self, scene):
        import torch
        self.scene = scene
        self.t = torch.zeros(scene.height * scene.width * 3, dtype=torch.float32)
        self.final = None

    def clear(self):
        self.t.zero_()

    def render_strided(self, first, stride, count):
        from oracle import oracle as orc
        import ctypes
        img = self.t.numpy()
        for k in range(count):
            rad, _ = self.scene.render_pass(first + k * stride)
            # pass index 2 => "+=" branch of the accumulate step (the buffer starts zeroed)
            orc.lib().ptbo_accumulate(img.ctypes.data_as(ctypes.c_void_p), rad.ctypes.data_as(ctypes.c_void_p), img.size, 2, self.scene.max_depth)

    def accumulation_tensor(self):
        return self.t

    def synchronize(self):
        pass

    def finalize(self, total):
        self.final = self.scene.tonemap(self.t.numpy().reshape(self.scene.height, self.scene.width, 3), total)


def _worker(rank, world, port, root, cfg, scene_json, total, out_path):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import pathtracerwithcuda_b200 as ptb
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    host = ptb.Renderer(cfg, device=-1)
    host.load_scene(scene_json, root)
    backend = OracleBackend(orc.OracleScene.from_renderer(host))
    sr = ShardedRenderer(backend, rank, world, dist)
    sr.begin()
    sr.render_total(total)
    sr.reduce(total)
    if rank == 0:
        np.savez(out_path, acc=backend.t.numpy(), u8=backend.final)
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(workload_root, tmp_path):
    import torch.multiprocessing as mp
    import pathtracerwithcuda_b200 as ptb
    from oracle import oracle as orc
    root, w = workload_root("mix", width=48, height=36)
    total = 5  # odd: ranks render 3 and 2 passes
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, port, root, w["config"], w["scene"], total, out), nprocs=2, join=True)
    got = np.load(out)
    host = ptb.Renderer(w["config"], device=-1)
    host.load_scene(w["scene"], root)
    single, _ = orc.OracleScene.from_renderer(host).render(total)
    acc = got["acc"].reshape(single.shape)
    # identical multiset of per-pass images; only float summation order differs
    assert np.allclose(acc, single, rtol=1e-5, atol=1e-6)
    assert np.abs(got["u8"].astype(int) - orc.OracleScene.from_renderer(host).tonemap(single, total).astype(int)).max() <= 1
