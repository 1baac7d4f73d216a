#!/usr/bin/env python3
"""Why do the reference arm's steps vary several-fold?  (VERDICT r01 "weak" 3)

    python tools/ref_variance.py [workload] [passes]      # GPU box with oracle/_ref/libptref.so

1. per-pass wall time of N synchronous passes through the UNMODIFIED path_tracer_kernel() after prefetching its managed memory;
2. the same loop re-issued launch by launch (oracle/ref_shim/ref_driver.cu: ref_pass_instrumented) with CUDA events around
   trace_ray_kernel and a host clock around thread_shrink (thrust::remove_if + its temporary cudaMalloc / cudaFree), so a slow
   pass can be attributed to kernel time, to the compaction call, or to neither.
Prints one JSON object (committed as profiles/r02_ref_variance_*.json).
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refharness as rh  # noqa: E402
from pathtracerwithcuda_b200 import procedural as pr  # noqa: E402


def stats(x):
    x = np.asarray(x, np.float64)
    return {"n": int(x.size), "min": float(x.min()), "median": float(np.median(x)), "mean": float(x.mean()), "p90": float(np.quantile(x, 0.9)),
            "p99": float(np.quantile(x, 0.99)), "max": float(x.max()), "above_2x_median": int((x > 2 * np.median(x)).sum()),
            "share_of_time_above_2x_median": float(x[x > 2 * np.median(x)].sum() / x.sum())}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 640
    root = tempfile.mkdtemp(prefix="ptb_refvar_")
    w = pr.make_workload(root, name)
    rh.link_backslash_names(root)
    saved = os.dup(1)
    os.dup2(2, 1)      # the reference prints its progress on stdout
    try:
        ref = rh.RefLib(host_only=False)
        ref.open(root, config_rel=w["config_rel"], scene=w["scene_name"])
        ref.render(4)
        ref.prefetch()
        ref.render(64)
        ref.prefetch()
        per_pass = ref.render_per_pass(n) * 1e3
        rows = np.array([ref.pass_breakdown(ref.lib.ref_pass_counter() + 1) for _ in range(min(n, 320))])
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
    wall, trace, shrink = rows[:, 0], rows[:, 1], rows[:, 2]
    other = wall - trace - shrink
    slow = wall > 2 * np.median(wall)
    px = w["width"] * w["height"]
    out = {"workload": name, "pixels": px,
           "unmodified_pass_ms": stats(per_pass),
           "Msamples_s_at_median_pass": px / (np.median(per_pass) / 1e3) / 1e6, "Msamples_s_at_mean_pass": px / (per_pass.mean() / 1e3) / 1e6,
           "Msamples_s_at_best_pass": px / (per_pass.min() / 1e3) / 1e6,
           "slowest_passes_ms": sorted(per_pass.tolist())[-8:],
           "reissued_loop": {"pass_wall_ms": stats(wall), "trace_ray_kernel_ms": stats(trace), "thread_shrink_wall_ms": stats(shrink), "other_ms": stats(other),
                             "slow_passes": int(slow.sum()),
                             "slow_pass_excess_ms_by_part": None if not slow.any() else {
                                 "trace_ray_kernel": float((trace[slow] - np.median(trace)).sum()),
                                 "thread_shrink": float((shrink[slow] - np.median(shrink)).sum()),
                                 "other": float((other[slow] - np.median(other)).sum())}}}
    print(json.dumps(out))
    ref.close()


if __name__ == "__main__":
    main()
