#!/usr/bin/env python3
"""GPU-box parity + timing report: ptb200 (CUDA product path) vs the headless reference
(oracle/_ref/libptref.so).  Writes one JSON document; used interactively through gpurun and by
tests/test_gpu_parity.py (which asserts on the same numbers).

    python tools/parity_report.py --scene cornell_box_simple --width 128 --height 128 --depth 5 --spp 4
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refharness as rh  # noqa: E402
import pathtracerwithcuda_b200 as ptb  # noqa: E402


def classify_prim_mismatch(ref_prim, ref_t, my_prim, my_t, brute_prim, brute_t):
    """split mismatches into exact-t ties, reference misses
    (brute force over all triangles agrees with us), and others."""
    diff = ref_prim != my_prim
    idx = np.nonzero(diff)[0]
    out = {"rays": int(ref_prim.size), "mismatch": int(idx.size), "exact_t_ties": 0, "reference_missed_hit": 0,
           "near_tie_1e-5": 0, "other": 0, "other_examples": []}
    for i in idx:
        if ref_t[i] == my_t[i]:
            out["exact_t_ties"] += 1
        elif brute_prim[i] == my_prim[i] and (my_t[i] < ref_t[i]):
            # we found a strictly nearer hit that an exhaustive scan confirms: the reference's
            # non-conservative slab test culled it (bounding_box.h:83-102)
            out["reference_missed_hit"] += 1
        elif np.isfinite(ref_t[i]) and np.isfinite(my_t[i]) and abs(ref_t[i] - my_t[i]) <= 1e-5 * max(abs(ref_t[i]), 1e-30):
            out["near_tie_1e-5"] += 1
        else:
            out["other"] += 1
            if len(out["other_examples"]) < 8:
                out["other_examples"].append([int(i), int(ref_prim[i]), float(ref_t[i]), int(my_prim[i]), float(my_t[i]), int(brute_prim[i]), float(brute_t[i])])
    return out


def image_stats(a, b, floor=1e-3):
    """relative error per pixel-channel where the reference value is above `floor`."""
    a = a.astype(np.float64).ravel()
    b = b.astype(np.float64).ravel()
    denom = np.maximum(np.abs(a), floor)
    rel = np.abs(a - b) / denom
    return {"max_rel": float(rel.max()), "p999_rel": float(np.quantile(rel, 0.999)), "p99_rel": float(np.quantile(rel, 0.99)),
            "mean_rel": float(rel.mean()), "outlier_frac_1e-3": float((rel > 1e-3).mean()), "bit_equal_frac": float((a == b).mean()),
            "rmse": float(np.sqrt(np.mean((a - b) ** 2))), "ref_mean": float(a.mean()), "new_mean": float(b.mean())}


def run(scene, width, height, depth, spp, root=None, config_extra=None, camera=None, time_passes=0, ids_depths=(0, 1, 2), keep=False,
        workload=None, tri_scale=1.0, options=None, edits=False):
    """scene: name of a reference scene JSON (staged copy of res/scene), or workload: a procedural one."""
    root = root or tempfile.mkdtemp(prefix="ptb_scratch_")
    cfg_rel = "res/configuration/parity.json"
    extra = dict(config_extra or {})
    if workload:
        from pathtracerwithcuda_b200 import procedural as pr
        w = pr.make_workload(root, workload, width=width, height=height, depth=depth, tri_scale=tri_scale)
        scene, width, height, depth = w["scene_name"], w["width"], w["height"], w["depth"]
        if camera is None and (w["aperture"] >= 0 or w["focal"] >= 0):
            camera = ptb.default_camera(width, height, w["aperture"], w["focal"]).as_array()
    else:
        rh.make_scratch_root(root)
    rh.write_config(os.path.join(root, cfg_rel), Width=width, Height=height, MaxDepth=depth, **extra)
    rh.link_backslash_names(root)
    report = {"scene": scene, "width": width, "height": height, "depth": depth, "spp": spp}

    ref = rh.RefLib(host_only=False)
    t0 = time.time()
    ref.open(root, config_rel=cfg_rel.replace("/", "\\"), scene=scene)
    report["ref_load_s"] = time.time() - t0
    mine = ptb.Renderer(os.path.join(root, cfg_rel), device=0)
    for k, v in (options or {}).items():
        mine.set_option(k, v)
    t0 = time.time()
    mine.load_scene(os.path.join(root, "res/scene", scene + ".json"), root)
    report["new_load_s"] = time.time() - t0
    if camera is not None:
        ref.set_camera(camera)
        mine.set_camera(camera)
    if edits:
        # the same live-edit script on both sides (reference setters vs ptb_set_* / ptb_apply_*), then compare as usual
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import make_golden as mg
        # "safe": without the rotate steps — apply_rotate makes the reference REBUILD a mesh tree inside a live
        # process, which trips its stateful Morton builder (incomplete tree -> missed hits)
        script = [e for e in mg.EDIT_SCRIPT if edits != "safe" or e[0] != "rotate"]
        for op, a in script:
            mg.apply_edit(ref, op, a)
            mg.apply_edit(mine, op, a)
        report["edits"] = [e[0] for e in script]

    # scene arrays
    rt, rm = ref.triangles()
    mt, mm = mine.scene_triangles()
    report["triangles"] = int(rt.shape[0])
    report["triangles_bit_equal"] = bool(np.array_equal(rt.view(np.uint32), mt.view(np.uint32)) and np.array_equal(rm, mm))
    report["ref_bvh_nodes"] = int(ref.lib.ref_num_bvh_nodes())

    # camera rays + closest-hit ids at several depths, on the reference's own live ray batches
    ids = {}
    for d in ids_depths:
        pix, rays = ref.capture_rays(1, d)
        if rays.shape[0] == 0:
            continue
        if d == 0:
            my_rays = mine.generate_rays(1)
            report["camera_rays_bit_equal"] = bool(np.array_equal(rays.view(np.uint32), my_rays[pix].view(np.uint32)))
            if not report["camera_rays_bit_equal"]:
                report["camera_rays_max_abs_diff"] = float(np.abs(rays - my_rays[pix]).max())
        rp, rtt = ref.trace_batch(rays)
        mp, mtt = mine.trace_batch(rays)
        bp, btt = mine.trace_batch(rays, bruteforce=True) if rt.shape[0] <= 200000 else (mp, mtt)
        c = classify_prim_mismatch(rp, rtt, mp, mtt, bp, btt)
        same = rp == mp
        c["t_bit_equal_on_same_prim"] = float((rtt[same].view(np.uint32) == mtt[same].view(np.uint32)).mean()) if same.any() else 1.0
        c["bvh_vs_bruteforce_mismatch"] = int((bp != mp).sum())
        ids["depth%d" % d] = c
    report["prim_ids"] = ids

    # fixed-spp per-pixel radiance
    ref.clear()
    mine.clear()
    ref.render(spp)
    mine.render(spp)
    ri = ref.image_f32()
    mi = mine.image_f32()
    report["image_sum"] = image_stats(ri, mi)
    mean_ref, mean_new = ri.astype(np.float64) / spp, mi.astype(np.float64) / spp
    report["mean_image_rmse"] = float(np.sqrt(np.mean((mean_ref - mean_new) ** 2)))
    report["mean_image_rel_rmse"] = float(report["mean_image_rmse"] / max(np.sqrt(np.mean(mean_ref ** 2)), 1e-30))
    report["last_pass"] = image_stats(ref.last_pass_f32(), mine.last_pass_f32())
    r8, m8 = ref.image_u8().astype(np.int32), mine.image_u8().astype(np.int32)
    report["image_u8_max_abs_diff"] = int(np.abs(r8 - m8).max())
    report["image_u8_equal_frac"] = float((r8 == m8).mean())
    report["new_stats"] = mine.stats()

    if time_passes > 0:
        ref.clear(); mine.clear()
        ref.render(1); mine.render(1)  # warm-up
        ref.prefetch()
        t_ref = ref.render(time_passes)
        t0 = time.time()
        mine.render(time_passes)
        t_new = time.time() - t0
        st = mine.stats()
        seg, trace_ms = ref.pass_instrumented(ref.lib.ref_pass_counter() + 1)
        report["timing"] = {"passes": time_passes, "ref_s": t_ref, "new_s": t_new, "speedup": t_ref / t_new if t_new > 0 else None,
                            "ref_Msamples_s": width * height * time_passes / t_ref / 1e6, "new_Msamples_s": width * height * time_passes / t_new / 1e6,
                            "new_gpu_ms": st["gpu_ms_total"], "new_segments": st["ray_segments"], "ref_segments_per_pass": seg,
                            "ref_trace_ms_per_pass": trace_ms}
    ref.close()
    mine.close()
    if not keep:
        import shutil
        shutil.rmtree(root, ignore_errors=True)
    return report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="cornell_box_simple")
    ap.add_argument("--workload", default="")
    ap.add_argument("--tri-scale", type=float, default=1.0)
    ap.add_argument("--passes-in-flight", type=int, default=0)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--depth", type=int, default=None)
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--time-passes", type=int, default=0)
    ap.add_argument("--out", default="")
    ap.add_argument("--edits", default="", choices=["", "all", "safe"], help="apply tests/golden/make_golden.py EDIT_SCRIPT to both sides first")
    args = ap.parse_args()
    opts = {"passes_in_flight": args.passes_in_flight} if args.passes_in_flight else None
    if args.workload:
        rep = run(None, args.width, args.height, args.depth, args.spp, time_passes=args.time_passes, workload=args.workload,
                  tri_scale=args.tri_scale, options=opts, edits=args.edits)
    else:
        rep = run(args.scene, args.width or 128, args.height or 128, args.depth or 5, args.spp, time_passes=args.time_passes, options=opts)
    txt = json.dumps(rep, indent=1)
    print(txt)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(txt)


if __name__ == "__main__":
    main()
