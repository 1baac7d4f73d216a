#!/usr/bin/env python3
"""Drop-in timing (GPU box): the UNMODIFIED reference host code (oracle/_ref/libptref.so: scene_parser, triangle_mesh, BVH build,
image, config_parser) builds its managed AoS scene once; then the SAME 18 arguments of Core/path_tracer.cpp:48-67 are handed,
one synchronous call per pass, first to the reference's own `path_tracer_kernel` and then to libptb200.so's symbol of that name.
    python tools/compat_bench.py [workload=c2] [passes=64] [steps=4] [reference_steps=steps]
PTB_COMPAT_LOOKAHEAD=1 switches the look-ahead of the drop-in symbol off (csrc/compat.inc).
Prints one JSON line: Msamples/s of both and their ratio.  With look-ahead the steps of the symbol alternate (batches of 8 passes on three
contexts against steps of `passes` calls): its rate is the MEAN over the steps."""
import ctypes, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refharness as rh
from pathtracerwithcuda_b200 import procedural as pr, api

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
root = tempfile.mkdtemp(prefix="ptb_compat_bench_")
w = pr.make_workload(root, name)
rh.link_backslash_names(root)
sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)      # the reference prints its progress lines on stdout
ref = rh.RefLib(host_only=False)
ref.open(root, config_rel=w["config_rel"], scene=w["scene_name"])
ref.set_camera(ref.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
px = w["width"] * w["height"]

# (a) the reference's kernel, managed memory prefetched (its most favourable condition, as in bench.py --impl reference)
ref_steps = int(sys.argv[4]) if len(sys.argv) > 4 else steps      # 0: skip the reference's timing (its image is still rendered for the comparison)
ref.clear(); ref.render(2); ref.prefetch()
if ref_steps:
    ref.render(passes); ref.prefetch()
ref_ms = [ref.render(passes) * 1e3 for _ in range(ref_steps)] or [float("nan")]
ref.clear(); ref.render(passes)
a_sum = ref.image_f32().copy()

# (b) our symbol behind the same arguments
args = (ctypes.c_void_p * 18)()
ref.lib.ref_kernel_args(args)
L = api.load_library()
vp, ci = ctypes.c_void_p, ctypes.c_int
L.path_tracer_kernel.argtypes = [ci, vp, vp, ci, vp, ci, vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp]
L.path_tracer_kernel.restype = None
iv = lambda k: int(args[k] or 0)


def call(p):
    L.path_tracer_kernel(iv(0), args[1], args[2], iv(3), args[4], iv(5), args[6], args[7], p, args[9], args[10], args[11], args[12],
                         args[13], args[14], args[15], args[16], args[17])


ref.clear()
t0 = time.perf_counter(); call(1); first_ms = (time.perf_counter() - t0) * 1e3        # ingest + BVH build + first pass
p = 1
for _ in range(passes - 1):                    # warm-up, then straight into the timed steps: the look-ahead is in steady state at
    p += 1; call(p)                            # both ends of the timed region (a host pause before it would hand it finished batches)
ours_ms = []
for _ in range(steps):
    t0 = time.perf_counter()
    for _ in range(passes):
        p += 1; call(p)
    ours_ms.append((time.perf_counter() - t0) * 1e3)
ref.clear()
for q in range(1, passes + 1):
    call(q)
b_sum = ref.image_f32().copy()
rel = np.abs(a_sum.astype(np.float64) - b_sum) / np.maximum(np.abs(a_sum), 1e-3)
# (c) for scale: the native C ABI on the same workload in the same process (8 passes x 3 streams in flight, like the look-ahead)
import pathtracerwithcuda_b200 as ptb
nr = ptb.Renderer(w["config"], device=0)
nr.set_option("passes_in_flight", 8); nr.set_option("streams_in_flight", 3)
nr.load_scene(w["scene"], root)
if w["aperture"] >= 0 or w["focal"] >= 0:
    nr.set_camera(ptb.default_camera(w["width"], w["height"], w["aperture"], w["focal"]))
nr.render(passes)
native_ms = []
for _ in range(steps):
    t0 = time.perf_counter(); nr.render(passes); native_ms.append((time.perf_counter() - t0) * 1e3)
native_segments = nr.stats()["ray_segments"]
nr.close()
sys.stdout.flush(); os.dup2(saved, 1)
rate = lambda ms: px * passes / (ms / 1e3) / 1e6
print(json.dumps({"workload": name, "lookahead": os.environ.get("PTB_COMPAT_LOOKAHEAD", "default (8)"), "resolution": [w["width"], w["height"]], "passes_per_step": passes, "steps": steps,
                  "reference_kernel_Msamples_s": {"median": rate(float(np.median(ref_ms))), "best": rate(min(ref_ms)), "step_ms": ref_ms},
                  "ptb200_symbol_Msamples_s": {"mean": rate(float(np.mean(ours_ms))), "median": rate(float(np.median(ours_ms))), "best": rate(min(ours_ms)), "step_ms": ours_ms,
                                               "first_call_ms_ingest_build_pass": first_ms},
                  "ptb200_native_abi_Msamples_s": {"median": rate(float(np.median(native_ms))), "step_ms": native_ms, "ray_segments_per_step": native_segments},
                  "ratio_median": rate(float(np.mean(ours_ms))) / rate(float(np.median(ref_ms))), "ratio_best_vs_best": rate(min(ours_ms)) / rate(min(ref_ms)),
                  "image_after_%d_passes" % passes: {"outliers_1e-3": float((rel > 1e-3).mean()), "p999_rel": float(np.quantile(rel, 0.999))}}))
