"""Multi-GPU behind the C ABI (include/ptb200.h: ptb_multi_*, ptb_dist_*; csrc/multi.inc).

On a one-GPU box: the one-device ptb_multi path must equal ptb_render bit for bit, the scene must survive the broadcast format
byte for byte, and the library's own NCCL path is exercised with a world of one rank.  With two or more GPUs: the N-device image equals
the one-device image within 1e-5 relative (same multiset of per-pass images, float re-association only — SURVEY.md §8e), rendering
in two calls equals one call, and a torchrun job whose rank 1 never sees the scene files receives the parsed scene by ncclBroadcast."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import pathtracerwithcuda_b200 as ptb

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    return ptb.device_count()


def test_multi_one_device_is_ptb_render(workload_root):
    root, w = workload_root("mix", width=96, height=72)
    single = ptb.Renderer(w["config"], device=0)
    single.load_scene(w["scene"], root)
    single.set_camera(ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
    single.render(6)
    m = ptb.MultiRenderer(w["config"], 1)
    m.load_scene(w["scene"], root)
    m.set_camera(ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
    m.render(4)
    m.render(2)
    assert m.pass_counter() == 6
    assert np.array_equal(m.image_f32().view(np.uint32), single.image_f32().view(np.uint32))
    assert np.array_equal(m.image_u8(), single.image_u8())
    m.clear()
    assert m.pass_counter() == 0
    m.render(6)
    assert np.array_equal(m.image_f32().view(np.uint32), single.image_f32().view(np.uint32))
    # the per-device handle is reachable (options, stats, introspection)
    assert m.renderer(0).scene_counts()["triangles"] == single.scene_counts()["triangles"]
    m.close()


def test_scene_survives_the_broadcast_format(workload_root):
    for name, kw in (("mix", dict(width=96, height=72)), ("c1", dict(width=64, height=64)), ("c3", dict(width=160, height=90, tri_scale=0.05))):
        root, w = workload_root(name, **kw)
        r = ptb.Renderer(w["config"], device=0)
        r.load_scene(w["scene"], root)
        r.scene_blob_roundtrip()
        r.close()


def test_dist_world_of_one_runs_the_nccl_path(workload_root):
    """ptb_dist_* with a one-rank communicator: init, render share (= everything), reduce onto the merged image, progressive second reduce."""
    if ptb.nccl_version() == 0:
        pytest.skip("NCCL cannot be loaded on this box")
    root, w = workload_root("mix", width=96, height=72)
    single = ptb.Renderer(w["config"], device=0)
    single.load_scene(w["scene"], root)
    single.render(5)
    r = ptb.Renderer(w["config"], device=0)
    r.load_scene(w["scene"], root)
    r.dist_init(0, 1, ptb.dist_unique_id())
    r.dist_broadcast_scene(0)           # root of a one-rank world: a no-op that must not disturb the scene
    r.dist_clear()
    r.dist_render(3)
    r.dist_reduce(0)
    img, passes = r.merged_image_f32()
    assert passes == 3
    r.dist_render(2)
    r.dist_reduce(0)                    # a second reduce must not double-count (the reduce is never in place)
    img, passes = r.merged_image_f32()
    assert passes == 5
    assert np.array_equal(img.view(np.uint32), single.image_f32().view(np.uint32))
    assert np.array_equal(r.merged_image_u8(), single.image_u8())
    # the collective load: parse on the root (host only), broadcast, build from the DEVICE copy of the triangles
    r.dist_load_scene(w["scene"], root, 0)
    tri_a, mat_a = r.scene_triangles()
    tri_b, mat_b = single.scene_triangles()
    assert np.array_equal(tri_a.view(np.uint32), tri_b.view(np.uint32)) and np.array_equal(mat_a, mat_b)
    r.dist_clear()
    r.dist_render(5)
    r.dist_reduce(0)
    img, passes = r.merged_image_f32()
    assert passes == 5 and np.array_equal(img.view(np.uint32), single.image_f32().view(np.uint32))
    with pytest.raises(ptb.PtbError):
        r.dist_load_scene("/nonexistent/scene.json", root, 0)      # a root that cannot parse fails every rank, nobody hangs
    r.dist_shutdown()
    with pytest.raises(ptb.PtbError):
        r.dist_render(1)                # no communicator any more: fails loudly


@pytest.mark.skipif(n_gpus() < 2, reason="needs two GPUs")
def test_multi_n_devices_equal_one_device(workload_root):
    root, w = workload_root("c2", width=320, height=180, tri_scale=0.1)
    single = ptb.Renderer(w["config"], device=0)
    single.load_scene(w["scene"], root)
    single.render(13)
    ref = single.image_f32()
    n = min(n_gpus(), 8)
    m = ptb.MultiRenderer(w["config"], n)
    m.set_option("passes_in_flight", 2)
    m.load_scene(w["scene"], root)
    m.render(13)                        # 13 passes over n devices: uneven shares
    got = m.image_f32()
    assert m.pass_counter() == 13
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)
    assert np.abs(m.image_u8().astype(int) - single.image_u8().astype(int)).max() <= 1
    # every pass rendered exactly once: the shares add up
    assert sum(m.renderer(i).pass_counter() for i in range(n)) == 13
    m.clear()
    m.render(6)
    m.render(7)                         # progressive: a second merge sees all 13 passes once
    assert np.allclose(m.image_f32(), ref, rtol=1e-5, atol=1e-6)
    m.close()


@pytest.mark.skipif(n_gpus() < 2, reason="needs two GPUs")
def test_two_process_dist_with_scene_broadcast(tmp_path):
    """torchrun x 2 through tools/scale_render.py: only rank 0 generates and parses the scene, rank 1 receives it over NCCL; the merged
    image's mean radiance equals the one-process render's."""
    tool = os.path.join(REPO, "tools", "scale_render.py")
    env = dict(os.environ, NCCL_DEBUG="WARN")
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29578", tool, "mix", "64"], capture_output=True, text=True, env=env, timeout=600)
    assert two.returncode == 0, two.stderr[-2000:]
    one = subprocess.run([sys.executable, tool, "mix", "64"], capture_output=True, text=True, env=dict(env, CUDA_VISIBLE_DEVICES="0"), timeout=600)
    assert one.returncode == 0, one.stderr[-2000:]
    a = json.loads([l for l in two.stdout.splitlines() if l.startswith("{")][-1])
    b = json.loads([l for l in one.stdout.splitlines() if l.startswith("{")][-1])
    assert a["n_gpus"] == 2 and b["n_gpus"] == 1 and a["passes_merged"] == 64
    assert abs(a["mean_radiance"] - b["mean_radiance"]) <= 1e-5 * b["mean_radiance"], (a["mean_radiance"], b["mean_radiance"])
