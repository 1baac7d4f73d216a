#!/usr/bin/env python3
"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref).

The reference ships no tests or golden vectors, so the pins are manufactured from
the reference's own code:

  --host   (build container, no GPU; needs oracle/_ref/libptref_host.so)
           kat_host.json      the reference's __host__ __device__ header functions and its thrust RNG
                              evaluated on the host on seeded inputs (bit patterns stored as uint32)
           scene_mix.npz      world-space triangles / materials / spheres the reference's own scene
                              pipeline (scene_parser + triangle_mesh + tinyobj + glm) produces for the
                              procedural 'mix' and 'c1' scenes, plus the default camera
  --gpu    (B200 box via gpurun; needs oracle/_ref/libptref.so)
           ref_gpu_<scene>.npz camera rays, closest-hit ids/distances at depths 0..2 on the reference's
                              own live ray batches, per-pass un-clamped radiance for passes 1..4, the
                              float accumulation image and the 8-bit image; device hash() vectors

    python tests/golden/make_golden.py --host
    gpurun -- 'python tests/golden/make_golden.py --gpu --out gpurun_out/golden'   # then copy *.npz here
"""
import argparse
import ctypes
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refharness as rh  # noqa: E402
from pathtracerwithcuda_b200 import procedural as pr  # noqa: E402

GOLDEN_SCENES = {"mix": dict(width=96, height=72), "c1": dict(width=64, height=64)}


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32).tolist()


def make_scene_root(root, name):
    w = pr.make_workload(root, name, **GOLDEN_SCENES[name])
    rh.link_backslash_names(root)
    return w


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def host_kats(ref):
    L = ref.lib
    rng = np.random.RandomState(1234)
    out = {}
    # RNG: the Appendix-B seeds + random ones, both distributions the kernels construct
    seeds = [1765928508, 517078614, 2686462232, 1495423408, 2654095160, 930103280, 866313100, 1571428960, 0, 2147483647, 4294967295] + \
        rng.randint(0, 2 ** 32, size=16, dtype=np.uint64).tolist()
    rows = []
    for s in seeds:
        a = np.zeros(6, np.float32); b = np.zeros(6, np.float32)
        L.ref_kat_rng(ctypes.c_uint(int(s)), -0.5, 0.5, 6, _p(a))
        L.ref_kat_rng(ctypes.c_uint(int(s)), 0.0, 1.0, 6, _p(b))
        rows.append({"seed": int(s), "m05_05": u32(a), "u01": u32(b)})
    out["rng"] = rows

    tri_rows = []
    for k in range(64):
        v = rng.uniform(-2, 2, 9).astype(np.float32)
        o = rng.uniform(-4, 4, 3).astype(np.float32)
        target = (v[0:3] + v[3:6] + v[6:9]) / 3 + rng.normal(0, 0.6 if k % 2 else 0.05, 3).astype(np.float32)
        d = (target - o); d = (d / np.linalg.norm(d)).astype(np.float32)
        ray = np.concatenate([o, d]).astype(np.float32)
        res = np.zeros(3, np.float32)
        hit = L.ref_kat_triangle(_p(v), _p(ray), _p(res))
        tri_rows.append({"v": u32(v), "ray": u32(ray), "hit": int(hit), "out": u32(res)})
    out["triangle"] = tri_rows

    sph_rows = []
    for k in range(48):
        c = rng.uniform(-2, 2, 3).astype(np.float32)
        r = np.float32(rng.uniform(0.2, 1.5))
        o = (c + rng.normal(0, 0.3 if k % 4 == 0 else 4.0, 3)).astype(np.float32)
        d = (c + rng.normal(0, 0.7, 3) - o); d = (d / np.linalg.norm(d)).astype(np.float32)
        cr = np.array([c[0], c[1], c[2], r], np.float32)
        ray = np.concatenate([o, d]).astype(np.float32)
        res = np.zeros(7, np.float32)
        hit = L.ref_kat_sphere(_p(cr), _p(ray), _p(res))
        sph_rows.append({"cr": u32(cr), "ray": u32(ray), "hit": int(hit), "out": u32(res)})
    out["sphere"] = sph_rows

    box_rows = []
    for k in range(48):
        lo = rng.uniform(-3, 0, 3).astype(np.float32); hi = (lo + rng.uniform(0.1, 4, 3)).astype(np.float32)
        o = rng.uniform(-6, 6, 3).astype(np.float32)
        d = rng.normal(0, 1, 3).astype(np.float32)
        if k % 8 == 0:
            d[k % 3] = 0.0
        d = (d / np.linalg.norm(d)).astype(np.float32)
        box = np.concatenate([lo, hi]).astype(np.float32); ray = np.concatenate([o, d]).astype(np.float32)
        t = np.array([np.inf], np.float32)
        hit = L.ref_kat_box(_p(box), _p(ray), _p(t))
        box_rows.append({"box": u32(box), "ray": u32(ray), "hit": int(hit), "t": u32(t)})
    out["box"] = box_rows

    fr_rows = []
    for k in range(48):
        n = rng.normal(0, 1, 3).astype(np.float32); n = (n / np.linalg.norm(n)).astype(np.float32)
        d = rng.normal(0, 1, 3).astype(np.float32); d = (d / np.linalg.norm(d)).astype(np.float32)
        if np.dot(n, d) > 0:
            n = -n
        n_in, n_out = (1.000293, 1.5319) if k % 3 == 0 else ((1.5319, 1.000293) if k % 3 == 1 else (1.000293, 1.33))
        # refraction direction per path_tracer_kernel.cu:54-83 (host evaluation in float32)
        i = -d; ndi = np.float32(np.dot(n, i)); ratio = np.float32(n_in / n_out)
        b = np.float32(1.0) - ratio * ratio * (np.float32(1.0) - ndi * ndi)
        refr = np.zeros(3, np.float32) if b < 0 else (n * (ratio * ndi - np.sqrt(b)) - ratio * i).astype(np.float32)
        refl = (d - 2 * np.dot(n, d) * n).astype(np.float32)
        fd = L.ref_kat_fresnel_dielectric(_p(n), _p(d), n_in, n_out, _p(refl), _p(refr))
        kk = [(2.5845, 2.7670), (0.04, 2.6484), (1.0220, 0.782)][k % 3]
        fc = L.ref_kat_fresnel_conductor(_p(n), _p(d), kk[0], kk[1])
        fr_rows.append({"n": u32(n), "d": u32(d), "n_in": float(np.float32(n_in)), "n_out": float(np.float32(n_out)), "refr": u32(refr),
                        "F_dielectric": u32([fd]), "nk": [float(np.float32(kk[0])), float(np.float32(kk[1]))], "F_conductor": u32([fc])})
    out["fresnel"] = fr_rows

    cube_rows = []
    dirs = [(1, 0.2, -0.3), (-0.2, -0.9, 0.1), (0.3, 0.3, -0.8), (0.5, 0.5, 0.5), (-0.5, -0.5, -0.5), (0, 1, 0), (0, 0, -1)] + rng.normal(0, 1, (40, 3)).tolist()
    for x, y, z in dirs:
        uv = np.zeros(2, np.float32)
        idx = L.ref_kat_cube_uv(ctypes.c_float(x), ctypes.c_float(y), ctypes.c_float(z), _p(uv))
        cube_rows.append({"d": u32([x, y, z]), "index": int(idx), "uv": u32(uv)})
    out["cube_uv"] = cube_rows

    tex = rng.randint(0, 256, (5, 7, 4)).astype(np.uint8)
    tex_rows = []
    for k in range(40):
        u, v = rng.uniform(-1.5, 2.5, 2).astype(np.float32)
        for bil in (0, 1):
            c = np.zeros(3, np.float32)
            L.ref_kat_texture(7, 5, _p(tex), ctypes.c_float(u), ctypes.c_float(v), bil, _p(c))
            tex_rows.append({"uv": u32([u, v]), "bilinear": bil, "rgb": u32(c)})
    out["texture"] = {"width": 7, "height": 5, "rgba": tex.reshape(-1).tolist(), "samples": tex_rows}

    faces = rng.randint(0, 256, (6, 4, 4, 4)).astype(np.uint8)
    ptrs = (ctypes.c_void_p * 6)(*[faces[f].ctypes.data for f in range(6)])
    bg_rows = []
    for k in range(48):
        d = rng.normal(0, 1, 3).astype(np.float32); d = (d / np.linalg.norm(d)).astype(np.float32)
        for mode in ((1, 0, 1), (1, 0, 0), (0, 1, 0), (0, 0, 0)):
            c = np.zeros(3, np.float32)
            L.ref_kat_background(4, ptrs, _p(d), mode[0], mode[1], mode[2], _p(c))
            bg_rows.append({"d": u32(d), "sky_box": mode[0], "sky": mode[1], "bilinear": mode[2], "rgb": u32(c)})
    out["background"] = {"length": 4, "faces": faces.reshape(-1).tolist(), "samples": bg_rows}

    mats = {}
    for name in ref.builtin_material_names():
        mats[name] = ref.builtin_material(name).tolist()
    out["builtin_materials"] = mats
    cams = []
    for w, h, ap, fo in [(640, 640, -1, -1), (1920, 1080, -1, -1), (96, 72, 0.08, 13.0), (3840, 2160, 0.05, 40.0)]:
        cams.append({"args": [w, h, ap, fo], "cam": u32(ref.default_camera(w, h, ap, fo))})
    out["default_camera"] = cams
    return out


# one edit of every kind the reference UI can issue, in an order that makes them interact
# (rotate after a non-uniform scale, a second transform after the rotation was applied)
EDIT_SCRIPT = [
    ("transform", (1, [-1.0, 0.45, 0.3], [1.2, 1.9, 1.4])),
    ("rotate", (1, [40.0, -25.0, 10.0])),
    ("transform", (2, [2.1, -0.2, 0.6], [0.9, 0.9, 1.7])),
    ("rotate", (2, [0.0, 90.0, -33.0])),
    ("transform", (1, [-1.6, 0.1, -0.4], [1.0, 1.0, 1.0])),
    ("sphere", (1, [2.5, 0.3, 1.2], 0.65, dict(diffuse_color=[0.2, 0.9, 0.4], roughness=0.35, is_transparent=0, extinction_coefficient=2.5))),
    ("material", (1, 2, dict(diffuse_color=[0.1, 0.2, 0.9], specular_color=[0.9, 0.8, 0.7], roughness=0.6, refraction_index=1.7))),
    ("transform", (0, [0.2, 4.2, 0.1], [1.0, 0.000000001, 2.0])),     # scale below the UI clamp
]


def apply_edit(target, op, args, spheres=None, mesh_materials=None, material_counts=None):
    """Runs one EDIT_SCRIPT step on `target`: a RefLib (reference) or a pathtracerwithcuda_b200.Renderer."""
    from pathtracerwithcuda_b200.api import MATERIAL_DTYPE, SPHERE_DTYPE
    is_ref = hasattr(target, "set_mesh_materials")

    def structured(a, dt):
        return np.ascontiguousarray(a).view(np.uint8).reshape(-1).view(dt).copy()
    if op == "transform":
        target.set_mesh_transform(args[0], args[1], args[2])
    elif op == "rotate":
        target.apply_mesh_rotate(args[0], args[1])
    elif op == "sphere":
        index, center, radius, changes = args
        sp = structured(target.spheres() if is_ref else target.scene_spheres(), SPHERE_DTYPE)
        one = sp[index:index + 1].copy()
        one["center"][0] = center
        one["radius"][0] = radius
        for k, v in changes.items():
            one["mat"][k][0] = v
        target.set_sphere(index, one)
    elif op == "material":
        mesh, which, changes = args
        mats = structured(target.mesh_materials() if is_ref else target.scene_materials(), MATERIAL_DTYPE)
        counts = target.mesh_material_counts()
        first = int(sum(counts[:mesh]))
        mine = mats[first:first + counts[mesh]].copy()
        for k, v in changes.items():
            mine[k][which] = v
        if is_ref:
            target.set_mesh_materials(mesh, mine)
        else:
            target.set_mesh_material(mesh, mine)
    else:
        raise ValueError(op)


def do_host(out_dir):
    ref = rh.RefLib(host_only=True)
    kats = host_kats(ref)
    with open(os.path.join(out_dir, "kat_host.json"), "w") as f:
        json.dump(kats, f)
    for name in GOLDEN_SCENES:
        root = tempfile.mkdtemp(prefix="ptb_golden_")
        w = make_scene_root(root, name)
        # host-only reference: its CPU Morton builder (NaiveCPU mis-handles planar groups, App. G.6)
        pr.write_config(w["config"], Width=w["width"], Height=w["height"], MaxDepth=w["depth"], BvhBuildMethod="MortonCodeCPU")
        ref.open(root, config_rel=w["config_rel"], scene=w["scene_name"])
        tri, mat = ref.triangles()
        np.savez_compressed(os.path.join(out_dir, "scene_%s.npz" % name), triangles=tri.view(np.uint32), triangle_material=mat,
                            materials=ref.mesh_materials(), spheres=ref.spheres(), camera=ref.camera().view(np.uint32))
        if name == "mix":
            # live edits through the reference's own setters (scene_parser.cpp:645-673): the world-space result
            # after each step of EDIT_SCRIPT is a fixture for the product's ptb_set_* / ptb_apply_* entry points
            steps = {}
            for k, (op, args) in enumerate(EDIT_SCRIPT):
                apply_edit(ref, op, args)
                tri_e, _ = ref.triangles()
                steps["step%d_triangles" % k] = tri_e.view(np.uint32)
                steps["step%d_materials" % k] = ref.mesh_materials()
                steps["step%d_spheres" % k] = ref.spheres()
            np.savez_compressed(os.path.join(out_dir, "scene_mix_edits.npz"), **steps)
            ref.close()
            ref.open(root, config_rel=w["config_rel"], scene=w["scene_name"])
            # the awkward-OBJ scene reuses this root (it needs the 64^2 cube map the mix workload wrote)
            sys.path.insert(0, HERE)
            import objedge
            objedge.write(root)
            rh.link_backslash_names(root)
            ref.close()
            ref.open(root, config_rel=w["config_rel"], scene="objedge")
            tri, mat = ref.triangles()
            np.savez_compressed(os.path.join(out_dir, "scene_objedge.npz"), triangles=tri.view(np.uint32), triangle_material=mat,
                                materials=ref.mesh_materials(), spheres=ref.spheres(), camera=ref.camera().view(np.uint32))
        ref.close()
        shutil.rmtree(root, ignore_errors=True)
    print("host fixtures written to", out_dir)


def gpu_scene(out_dir, name):
    """One scene per PROCESS: the reference's Morton builder keeps state between scene loads and
    produced an incomplete tree for the second scene of a process."""
    from oracle import oracle as orc
    import pathtracerwithcuda_b200 as ptb
    ref = rh.RefLib(host_only=False)
    root = tempfile.mkdtemp(prefix="ptb_golden_")
    w = make_scene_root(root, name)
    ref.open(root, config_rel=w["config_rel"], scene=w["scene_name"])
    cam = ref.default_camera(w["width"], w["height"], w["aperture"], w["focal"])
    ref.set_camera(cam)
    # exhaustive-scan cross-check of the reference tree ("tree reachability validated")
    host = ptb.Renderer(w["config"], device=-1)
    host.load_scene(w["scene"], root)
    host.set_camera(cam)
    brute = orc.OracleScene.from_renderer(host)
    data = {"camera": cam.view(np.uint32)}
    missed = 0
    for d in (0, 1, 2, 3):
        pix, rays = ref.capture_rays(1, d)
        prim, t = ref.trace_batch(rays)
        bp, bt, _ = brute.trace(rays, brute=True)
        missed += int(((prim != bp) & (bt < t)).sum())
        data["depth%d_pixels" % d] = pix
        data["depth%d_rays" % d] = rays.view(np.uint32)
        data["depth%d_prim" % d] = prim
        data["depth%d_t" % d] = t.view(np.uint32)
    data["reference_missed_hits"] = np.array([missed], np.int64)
    pix, rays2 = ref.capture_rays(2, 0)
    data["pass2_rays"] = rays2.view(np.uint32)
    ref.clear()
    per_pass = []
    for k in range(4):
        ref.render(1)
        per_pass.append(ref.last_pass_f32().view(np.uint32))
    data["pass_radiance"] = np.stack(per_pass)
    data["image_sum"] = ref.image_f32().view(np.uint32)
    data["image_u8"] = ref.image_u8()
    seg, _ = ref.pass_instrumented(5)
    data["segments_pass5"] = np.array([seg], np.int64)
    np.savez_compressed(os.path.join(out_dir, "ref_gpu_%s.npz" % name), **data)
    ref.close()
    shutil.rmtree(root, ignore_errors=True)
    print("scene", name, "reference_missed_hits", missed)
    return missed


def do_gpu(out_dir):
    import subprocess
    ref = rh.RefLib(host_only=False)
    ints = np.concatenate([np.arange(0, 64), np.array([2 ** 31 - 1, -1, -2 ** 31, 123456789, 262143, 2073599, 8294399]),
                           np.random.RandomState(7).randint(-2 ** 31, 2 ** 31 - 1, 64)]).astype(np.int32)
    hashed = np.zeros_like(ints)
    if ref.lib.ref_device_hash(_p(ints), ints.size, _p(hashed)) != 0:
        raise RuntimeError("ref_device_hash failed")
    np.savez_compressed(os.path.join(out_dir, "ref_gpu_hash.npz"), inputs=ints, outputs=hashed)
    for name in GOLDEN_SCENES:
        subprocess.run([sys.executable, os.path.abspath(__file__), "--gpu-scene", name, "--out", out_dir], check=True)
    print("gpu fixtures written to", out_dir)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--host", action="store_true")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--gpu-scene", default="")
    ap.add_argument("--out", default=HERE)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    if a.gpu_scene:
        gpu_scene(a.out, a.gpu_scene)
    if a.host:
        do_host(a.out)
    if a.gpu:
        do_gpu(a.out)


if __name__ == "__main__":
    main()
