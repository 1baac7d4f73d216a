#!/usr/bin/env python3
"""Parity sweep over every scene file of the reference's own res/scene whose assets exist in its checkout:
each in a fresh process (the reference's builder keeps state between loads), 320x180, 2 spp, ptb200 vs the live reference kernels.
    python tools/parity_ref_scenes.py > profiles/..._ref_scenes.json"""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scene_dir = os.path.join(ROOT, "oracle", "_ref", "res", "scene")
rows = []
for f in sorted(os.listdir(scene_dir)):
    name = f[:-5]
    out = tempfile.mktemp(suffix=".json")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "parity_report.py"), "--scene", name, "--width", "320", "--height", "180", "--depth", "8",
                        "--spp", "2", "--out", out], capture_output=True, text=True, timeout=600)
    if p.returncode != 0 or not os.path.exists(out):
        rows.append({"scene": name, "loaded": False, "why": (p.stderr or p.stdout)[-200:].strip().splitlines()[-1:] })
        continue
    d = json.load(open(out))
    ids = d["prim_ids"]
    rows.append({"scene": name, "loaded": True, "triangles": d["triangles"], "triangles_bit_equal": d["triangles_bit_equal"], "camera_rays_bit_equal": d.get("camera_rays_bit_equal"),
                 "rays": sum(c["rays"] for c in ids.values()), "id_mismatch": sum(c["mismatch"] for c in ids.values()),
                 "exact_t_ties": sum(c["exact_t_ties"] for c in ids.values()), "reference_missed_hit": sum(c["reference_missed_hit"] for c in ids.values()),
                 "other": sum(c["other"] + c["near_tie_1e-5"] for c in ids.values()), "t_bit_equal": min(c["t_bit_equal_on_same_prim"] for c in ids.values()),
                 "bvh_vs_bruteforce_mismatch": sum(c["bvh_vs_bruteforce_mismatch"] for c in ids.values()),
                 "sum_outliers_1e-3": d["image_sum"]["outlier_frac_1e-3"], "last_pass_bit_equal": d["last_pass"]["bit_equal_frac"], "u8_max_diff": d["image_u8_max_abs_diff"]})
    print(rows[-1], file=sys.stderr, flush=True)
print(json.dumps(rows, indent=1))
