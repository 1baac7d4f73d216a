"""Optional estimator "nee" (NOT the reference's estimator, off by default): next-event
estimation of the emissive triangles with a shadow-ray stage.  It must converge to the SAME image as the
reference estimator (same expected value: Fresnel-gated emission, bounce limit, energy cut), it must leave the
default estimator untouched, and the path itself (ray segments per depth) must be the reference's."""
import os
import sys

import numpy as np
import pytest

import pathtracerwithcuda_b200 as ptb

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


def test_nee_converges_to_the_reference_estimator():
    """Compared WITHOUT the per-pass clamp (option pass_clamp): the reference clamps every pass to 2*MaxDepth
    (path_tracer_kernel.cu:644-651), which cuts its rare bright samples (a 13-unit emitter seen through one white bounce)
    much harder than NEE's many small ones, so under the clamp the two estimators converge to different images by design
    (measured: 85 % apart at MaxDepth 2, 5 % at MaxDepth 5, 0.4 % unclamped)."""
    import nee_check
    # closed-box lighting only (no sky): every photon comes from the emissive box -> the strictest check
    rep = nee_check.run("c1", 64, 64, 4096, 1024, sky=False, depth=2, clamp=1e30)       # direct light only
    assert rep["rel_mean_diff"] <= 0.015 and rep["block_rel_rmse"] <= 0.04, rep
    rep = nee_check.run("c1", 64, 64, 4096, 1024, sky=False, depth=5, clamp=1e30)       # + indirect, spheres, glass
    assert rep["rel_mean_diff"] <= 0.05 and rep["block_rel_rmse"] <= 0.12, rep
    # textures, media, DOF, sky — with the reference clamp (sky-dominated, the clamp hardly bites; the scene's NaN samples,
    # which the clamp maps to its upper bound exactly as the reference's clamp does, would swamp an unclamped mean)
    rep = nee_check.run("mix", 96, 72, 2048, 512, sky=True)
    assert rep["rel_mean_diff"] <= 0.03, rep


def test_nee_keeps_the_paths_and_the_default_estimator(workload_root):
    root, w = workload_root("mix", width=96, height=72)
    cam = ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"])
    imgs, segs = {}, {}
    for mode in ("reference", "nee", "reference"):
        r = ptb.Renderer(w["config"], device=0)
        r.set_option("estimator", mode)
        r.load_scene(w["scene"], root)
        r.set_camera(cam)
        r.render(4)
        imgs.setdefault(mode, []).append(r.image_f32().copy())
        segs[mode] = r.depth_profile()[0].copy()
        r.close()
    assert np.array_equal(imgs["reference"][0].view(np.uint32), imgs["reference"][1].view(np.uint32))
    assert np.array_equal(segs["reference"], segs["nee"])          # same paths, only the light accounting differs
    assert not np.array_equal(imgs["reference"][0], imgs["nee"][0])
    with pytest.raises(ptb.PtbError):
        ptb.Renderer(w["config"], device=0).set_option("estimator", "bidirectional")
