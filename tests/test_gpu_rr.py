"""Optional Russian roulette (NOT in the reference, off by default): from bounce 3 on a path survives
with probability clamp(max throughput, 0.05, 1) and is re-weighted.  Unbiased: it must converge to the reference estimator's
image; it must trace fewer deep segments; switched off it must leave the default estimator bit-identical."""
import os
import sys

import numpy as np
import pytest

import pathtracerwithcuda_b200 as ptb

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


def test_roulette_converges_to_the_reference_estimator():
    import nee_check
    # closed box, emissive triangles only, no per-pass clamp (re-weighted samples are brighter and rarer: the reference's
    # clamp to 2*MaxDepth would bite the two estimators differently, as it does for NEE)
    rep = nee_check.run("c1", 64, 64, 4096, 4096, sky=False, depth=8, clamp=1e30, alt_options={"russian_roulette": 1})
    assert rep["rel_mean_diff"] <= 0.03 and rep["block_rel_rmse"] <= 0.12, rep
    assert rep["segments_per_pass"][1] < rep["segments_per_pass"][0], rep
    # sky-lit scene with textures, media and glass, default clamp, depth 16: fewer segments, same mean
    rep = nee_check.run("mix", 96, 72, 2048, 2048, sky=True, depth=16, alt_options={"russian_roulette": 1})
    assert rep["rel_mean_diff"] <= 0.02, rep
    assert rep["segments_per_pass"][1] < 0.9 * rep["segments_per_pass"][0], rep
    # together with next-event estimation
    rep = nee_check.run("c1", 64, 64, 4096, 2048, sky=False, depth=8, clamp=1e30, alt_options={"russian_roulette": 1, "estimator": "nee"})
    assert rep["rel_mean_diff"] <= 0.05, rep


def test_roulette_off_is_the_default_estimator(workload_root):
    root, w = workload_root("mix", width=96, height=72)
    cam = ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"])
    r = ptb.Renderer(w["config"], device=0)
    r.load_scene(w["scene"], root)
    r.set_camera(cam)
    r.render(4)
    base = r.image_f32().copy()
    seg0 = r.depth_profile()[0].copy()
    r.set_option("russian_roulette", 1)           # restarts the accumulation
    assert r.pass_counter() == 0
    r.render(4)
    rr = r.image_f32().copy()
    seg1 = r.depth_profile()[0].copy()
    assert not np.array_equal(base, rr)
    assert np.array_equal(seg0[:4], seg1[:4]) and seg1[4:].sum() < seg0[4:].sum()      # bounces 0-3 are traced as before
    r.set_option("russian_roulette", 0)
    r.render(4)
    assert np.array_equal(r.image_f32().view(np.uint32), base.view(np.uint32))
