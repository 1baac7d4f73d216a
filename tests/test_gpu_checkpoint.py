"""Checkpoint / resume of the accumulation (Core/image.h:10-23 state: pixels + pass_counter) and the handle-level
writers: a render interrupted after k passes and resumed in a NEW renderer must end bit-identical to an
uninterrupted one (pass index = seed, per-pass clamp, pass-ordered summation)."""
import numpy as np
import pytest

import pathtracerwithcuda_b200 as ptb
from test_image_out import decode_png

pytestmark = pytest.mark.gpu


def test_resume_is_bit_identical(workload_root, tmp_path):
    root, w = workload_root("mix", width=96, height=72)
    cam = ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"])
    full = ptb.Renderer(w["config"], device=0)
    full.load_scene(w["scene"], root)
    full.set_camera(cam)
    full.render(11)
    want, want8 = full.image_f32().copy(), full.image_u8().copy()

    a = ptb.Renderer(w["config"], device=0)
    a.load_scene(w["scene"], root)
    a.set_camera(cam)
    a.render(4)
    ck = str(tmp_path / "render.ptbck")
    a.save_checkpoint(ck)
    a.close()

    b = ptb.Renderer(w["config"], device=0)
    b.load_scene(w["scene"], root)
    b.load_checkpoint(ck)                       # restores accumulation, pass counter and camera
    assert b.pass_counter() == 4
    assert np.array_equal(b.camera().as_array().view(np.uint32)[:9], cam.as_array().view(np.uint32)[:9])
    b.render(7)
    assert b.pass_counter() == 11
    assert np.array_equal(b.image_f32().view(np.uint32), want.view(np.uint32))
    assert np.array_equal(b.image_u8(), want8)

    # writers on the handle
    png, pfm = str(tmp_path / "out.png"), str(tmp_path / "out.pfm")
    b.save_png(png)
    b.save_pfm(pfm)
    assert np.array_equal(decode_png(png)[..., :3], want8)
    data = open(pfm, "rb").read().split(b"\n", 3)
    mean = np.frombuffer(data[3], "<f4").reshape(72, 96, 3)[::-1]
    assert np.array_equal(mean, want * np.float32(1.0 / 11.0))


def test_checkpoint_rejects_mismatch_and_corruption(workload_root, tmp_path):
    root, w = workload_root("mix", width=96, height=72)
    r = ptb.Renderer(w["config"], device=0)
    r.load_scene(w["scene"], root)
    r.render(2)
    ck = str(tmp_path / "c.ptbck")
    r.save_checkpoint(ck)
    root2, w2 = workload_root("c1", width=64, height=64)
    other = ptb.Renderer(w2["config"], device=0)
    other.load_scene(w2["scene"], root2)
    with pytest.raises(ptb.PtbError):
        other.load_checkpoint(ck)               # resolution / depth differ
    good = open(ck, "rb").read()
    # a flipped pixel byte, a flipped header byte (pass counter at offset 16, camera from offset 24: the checksum covers the
    # header too), a truncated file and trailing bytes are all rejected
    for bad in (good[:len(good) // 2] + bytes([good[len(good) // 2] ^ 0x40]) + good[len(good) // 2 + 1:],
                good[:16] + bytes([good[16] ^ 1]) + good[17:], good[:30] + bytes([good[30] ^ 0x10]) + good[31:],
                good[:-8], good + b"\0\0\0\0"):
        open(ck, "wb").write(bad)
        with pytest.raises(ptb.PtbError):
            r.load_checkpoint(ck)
    open(ck, "wb").write(good)
    r.load_checkpoint(ck)                       # the untouched file still loads
    with pytest.raises(ptb.PtbError):
        r.load_checkpoint(str(tmp_path / "missing.ptbck"))
    assert r.pass_counter() == 2                # a failed load leaves the render untouched
