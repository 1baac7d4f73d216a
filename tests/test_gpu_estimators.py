"""Estimator / filtering options beside the parity mode — none is the reference's estimator, each must
have the reference estimator's EXPECTATION (or, for the hardware texture filter, its image up to the texture unit's 9-bit weights):

  sampler=pcg            pcg streams instead of the reference's hash-product + minstd streams: same mean image;
  sss=per_channel        free flight drawn from a uniformly picked channel of sigma_s' with single-sample MIS weights (the reference
                         samples from sigma_s'.x only, TODO at Kernel/path_tracer_kernel.cu:456-464): equal to the reference mode when
                         the three channels agree, and equal to the channel-wise composite of three reference-mode renders when not;
  texture_filter=hardware  bilinear lookups by the texture unit on cudaArray copies (software filter = Core/texture.h:15-79 stays default).
"""
import json
import os

import numpy as np
import pytest

import pathtracerwithcuda_b200 as ptb

pytestmark = pytest.mark.gpu


def render(w, root, scene=None, spp=64, **options):
    r = ptb.Renderer(w["config"], device=0)
    for k, v in options.items():
        r.set_option(k, v)
    r.load_scene(scene or w["scene"], root)
    r.set_camera(ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
    r.render(spp)
    img = r.image_f32() / spp
    u8 = r.image_u8()
    r.close()
    return img, u8


def box_blur(img, k=4):
    h, w, _ = img.shape
    return img[: h // k * k, : w // k * k].reshape(h // k, k, w // k, k, 3).mean(axis=(1, 3))


def scene_variant(w, root, name, sigma_s):
    """copy of the workload's scene with the subsurface material's ReducedScatteringCoef replaced"""
    doc = json.load(open(w["scene"]))
    hit = 0
    for m in doc.get("Material", []):
        if m["Name"] == "mix_sss":
            m["ReducedScatteringCoef"] = "%g %g %g" % tuple(sigma_s)
            hit += 1
    assert hit == 1
    path = os.path.join(os.path.dirname(w["scene"]), name + ".json")
    with open(path, "w") as f:
        json.dump(doc, f)
    return path


def test_pcg_sampler_has_the_reference_expectation(workload_root):
    root, w = workload_root("mix", width=96, height=72)
    a, _ = render(w, root, spp=768)
    b, _ = render(w, root, spp=768, sampler="pcg")
    assert not np.array_equal(a, b)                       # different samples ...
    assert abs(a.mean() - b.mean()) <= 0.01 * a.mean()    # ... same mean image
    d = np.abs(box_blur(a) - box_blur(b)) / np.maximum(box_blur(a), 0.05)
    assert np.quantile(d, 0.95) <= 0.12, float(np.quantile(d, 0.95))
    # the option also composes with the other estimator options
    c, _ = render(w, root, spp=768, sampler="pcg", russian_roulette=1)
    assert abs(a.mean() - c.mean()) <= 0.02 * a.mean()


def test_per_channel_sss_equals_reference_when_channels_agree(workload_root):
    root, w = workload_root("mix", width=96, height=72)           # mix_sss: sigma_s' = (4, 4, 4)
    a, ua = render(w, root, spp=8)
    b, ub = render(w, root, spp=8, sss="per_channel")
    rel = np.abs(a - b) / np.maximum(a, 1e-3)
    assert (rel > 1e-3).mean() <= 1e-3, float((rel > 1e-3).mean())   # weights are 1 up to rounding; the paths are the same paths
    assert np.abs(ua.astype(int) - ub.astype(int)).max() <= 1


def test_per_channel_sss_matches_the_channelwise_composite(workload_root):
    root, w = workload_root("mix", width=48, height=36)
    sig = (3.0, 9.0, 1.5)
    spp = 2048
    # ground truth: channel c of a reference-mode render whose (scalar) sigma_s'.x is sig[c] — transport is linear per channel
    truth = np.zeros((36, 48, 3), np.float32)
    for c in range(3):
        img, _ = render(w, root, scene=scene_variant(w, root, "ptb_mix_sss_c%d" % c, (sig[c],) * 3), spp=spp)
        truth[..., c] = img[..., c]
    coloured = scene_variant(w, root, "ptb_mix_sss_rgb", sig)
    got, _ = render(w, root, scene=coloured, spp=spp, sss="per_channel")
    ref_mode, _ = render(w, root, scene=coloured, spp=spp)            # the reference estimator ignores sigma_s'.y / .z
    for c in range(3):
        assert abs(got[..., c].mean() - truth[..., c].mean()) <= 0.02 * truth[..., c].mean(), (c, got[..., c].mean(), truth[..., c].mean())
    d = np.abs(box_blur(got) - box_blur(truth)) / np.maximum(box_blur(truth), 0.05)
    assert np.quantile(d, 0.9) <= 0.15, float(np.quantile(d, 0.9))
    # and it is a different image from the reference estimator's in the channels that one gets wrong
    d_ref = np.abs(box_blur(ref_mode) - box_blur(truth)) / np.maximum(box_blur(truth), 0.05)
    assert d_ref[..., 1].mean() > 2.0 * d[..., 1].mean()


def test_hardware_texture_filter_matches_the_software_filter(workload_root):
    root, w = workload_root("c3", width=320, height=180, tri_scale=0.05)      # 8 textures, sky box, thin lens
    a, ua = render(w, root, spp=4)
    b, ub = render(w, root, spp=4, texture_filter="hardware")
    assert not np.array_equal(a, b)                                       # the texture unit really filtered (9-bit weights)
    rel = np.abs(a - b) / np.maximum(a, 1e-2)
    assert rel.mean() <= 2e-3 and np.quantile(rel, 0.999) <= 0.05, (float(rel.mean()), float(np.quantile(rel, 0.999)))
    assert (np.abs(ua.astype(int) - ub.astype(int)) > 2).mean() <= 1e-3
