#!/usr/bin/env python3
"""Convergence check of the optional "nee" estimator against the reference estimator (same expected value)."""
import os, sys, tempfile, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr

def block_means(a, b=8):
    h, w, _ = a.shape
    return a[:h // b * b, :w // b * b].reshape(h // b, b, w // b, b, 3).mean(axis=(1, 3))

def run(name, width, height, spp_ref, spp_nee, sky=True, depth=None, clamp=None, alt_options=None, **workload_kw):
    root = tempfile.mkdtemp(prefix="ptb_nee_")
    w = pr.make_workload(root, name, width=width, height=height, depth=depth, **workload_kw)
    if not sky:
        pr.write_config(w["config"], Width=width, Height=height, MaxDepth=w["depth"], Skybox=False, Sky=False)
    out = {}
    for mode, spp in (("reference", spp_ref), ("nee", spp_nee)):
        r = ptb.Renderer(w["config"], device=0)
        if alt_options is None:
            r.set_option("estimator", mode)
        elif mode == "nee":                      # the second render uses the given options instead of estimator=nee
            for k, v in alt_options.items():
                r.set_option(k, v)
        if clamp is not None:
            r.set_option("pass_clamp", clamp)
        r.load_scene(w["scene"], root)
        if w["aperture"] >= 0:
            r.set_camera(ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
        r.render(spp)
        st = r.stats()
        out[mode] = (r.image_f32() / spp, st["gpu_ms_total"], int(r.depth_profile()[0].sum()))
        r.close()
    a, b = out["reference"][0].astype(np.float64), out["nee"][0].astype(np.float64)
    ba, bb = block_means(a), block_means(b)
    rep = {"scene": name, "sky": sky, "depth": w["depth"], "clamp": clamp, "spp": [spp_ref, spp_nee], "mean_ref": a.mean(), "mean_nee": b.mean(), "rel_mean_diff": abs(a.mean() - b.mean()) / a.mean(),
           "block_rel_rmse": float(np.sqrt(np.mean((ba - bb) ** 2)) / np.sqrt(np.mean(ba ** 2))), "block_max_rel": float(np.max(np.abs(ba - bb) / np.maximum(ba, 0.05))),
           "ms": [out["reference"][1], out["nee"][1]], "segments_per_pass": [out["reference"][2] / spp_ref, out["nee"][2] / spp_nee]}
    # noise: variance of each estimator from two half renders would need more runs; report per-pixel rmse as a proxy
    rep["pixel_rel_rmse"] = float(np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(a ** 2)))
    return rep

if __name__ == "__main__":
    for kw in (dict(depth=2), dict(depth=5), dict(depth=2, clamp=1e30), dict(depth=5, clamp=1e30)):
        print(json.dumps(run("c1", 64, 64, 4096, 1024, False, **kw)), flush=True)
    print(json.dumps(run("mix", 96, 72, 2048, 512, True)), flush=True)
