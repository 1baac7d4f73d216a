"""TEST INFRASTRUCTURE — Python side of the reference harness (see oracle/build_ref.py).

* `make_scratch_root` lays out a scratch asset root the UNMODIFIED reference can run in on Linux:
  the staged copy of its `res/` tree (oracle/_ref/res, git-ignored), a synthesised `yokohama` cube
  map (missing from the checkout), `.rgba8` side-cars for images the FreeImage
  stand-in cannot decode, and — because the reference opens paths spelled with '\\' — a symlink
  whose NAME contains the backslashes for every file (legal on Linux), so its own fopen/ifstream
  calls succeed without any source change.
* `RefLib` is a ctypes binding of oracle/_ref/libptref.so (GPU box) / libptref_host.so (host-only).

Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs import this.
"""
import ctypes
import json
import os
import shutil
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
STAGED_RES = os.path.join(REF_DIR, "res")

DEFAULT_CONFIG = {
    "Width": "1440", "Height": "900", "FullScreen": "false", "BlockSize": "64", "MaxBlockSize": "608",
    "MaxDepth": "20", "BiasLength": "0.0002", "EnergyThreshold": "0.000001", "SSSThreshold": "0.000001",
    "Skybox": "true", "BilinearSample": "true", "Sky": "false", "GammaCorrection": "true", "AntiAlias": "true",
    "FOV": "45.0", "BvhLeafNodeTriangleNum": "1", "BvhBucketMaxDivideInternalNum": "12", "BvhBuildBlockSize": "32",
    "BvhBuildMethod": "MortonCodeCUDA", "AirRefractionIndex": "1.000293", "AirAbsorptionCoef": "0.0 0.0 0.0",
    "AirReducedScatteringCoef": "0.0 0.0 0.0", "CUDAAcceleration": "true",
}


def write_config(path, **overrides):
    """Config JSON in the reference's format: every value is a string (config_parser.cpp:49-117)."""
    cfg = dict(DEFAULT_CONFIG)
    for k, v in overrides.items():
        if isinstance(v, bool):
            v = "true" if v else "false"
        cfg[k] = str(v)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(cfg, f, indent=1)
    return path


def write_bmp24(path, rgb):
    """rgb: (H, W, 3) uint8, row 0 = top. Writes a bottom-up uncompressed 24-bit BMP."""
    h, w, _ = rgb.shape
    pitch = (w * 3 + 3) & ~3
    rows = np.zeros((h, pitch), dtype=np.uint8)
    rows[:, : w * 3] = rgb[::-1, :, ::-1].reshape(h, w * 3)
    header = b"BM" + struct.pack("<IHHI", 54 + pitch * h, 0, 0, 54)
    info = struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, pitch * h, 2835, 2835, 0, 0)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(header + info + rows.tobytes())


def synth_cube_face(face, n, seed=0):
    """Deterministic procedural cube-map face: per-face tint + gradient + checker (fixed formula)."""
    y, x = np.mgrid[0:n, 0:n].astype(np.float32)
    u, v = x / (n - 1), y / (n - 1)
    tint = np.array([[0.9, 0.5, 0.4], [0.4, 0.9, 0.5], [0.5, 0.6, 1.0], [0.35, 0.3, 0.25], [0.9, 0.85, 0.5], [0.6, 0.4, 0.9]],
                    dtype=np.float32)[face]
    checker = (((x // max(1, n // 16)).astype(np.int32) + (y // max(1, n // 16)).astype(np.int32) + seed) & 1).astype(np.float32)
    base = 0.35 + 0.45 * (1.0 - v) + 0.2 * checker * u
    img = np.clip(base[..., None] * tint[None, None, :], 0.0, 1.0)
    return (img * 255.0 + 0.5).astype(np.uint8)


def write_synth_cubemap(dir_path, n=512, seed=0):
    for i, name in enumerate(["xpos", "xneg", "ypos", "yneg", "zpos", "zneg"]):
        write_bmp24(os.path.join(dir_path, name + ".bmp"), synth_cube_face(i, n, seed))


def write_sidecar(img_path):
    """Pre-decode an image FreeImage would decode (JPG/TGA/PNG) into '<file>.rgba8' (top-down RGBA8)."""
    from PIL import Image
    im = Image.open(img_path).convert("RGB")
    a = np.asarray(im, dtype=np.uint8)
    h, w, _ = a.shape
    rgba = np.concatenate([a, np.full((h, w, 1), 255, np.uint8)], axis=2)
    with open(img_path + ".rgba8", "wb") as f:
        f.write(struct.pack("<II", w, h) + rgba.tobytes())


def link_backslash_names(root):
    """For every file under root/res create root/'res\\a\\b.ext' -> res/a/b.ext."""
    for d, _, files in os.walk(os.path.join(root, "res")):
        for fn in files:
            rel = os.path.relpath(os.path.join(d, fn), root)
            alias = os.path.join(root, rel.replace("/", "\\"))
            if "/" in rel and not os.path.lexists(alias):
                os.symlink(rel, alias)


def make_scratch_root(dst, with_reference_res=True, yokohama=True):
    os.makedirs(os.path.join(dst, "res"), exist_ok=True)
    if with_reference_res and os.path.isdir(STAGED_RES):
        shutil.copytree(STAGED_RES, os.path.join(dst, "res"), dirs_exist_ok=True)
    if yokohama and not os.path.exists(os.path.join(dst, "res/texture/yokohama/xpos.bmp")):
        write_synth_cubemap(os.path.join(dst, "res/texture/yokohama"), 512, seed=0)
    for d, _, files in os.walk(os.path.join(dst, "res", "texture")):
        for fn in files:
            if fn.lower().endswith((".jpg", ".jpeg", ".tga", ".png")) and not os.path.exists(os.path.join(d, fn + ".rgba8")):
                write_sidecar(os.path.join(d, fn))
    link_backslash_names(dst)
    return dst


class RefLib:
    """ctypes view of the headless reference. All calls run with CWD = scratch root."""

    def __init__(self, host_only=False):
        name = "libptref_host.so" if host_only else "libptref.so"
        path = os.path.join(REF_DIR, name)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run python oracle/build_ref.py where /root/reference exists)")
        self.lib = L = ctypes.CDLL(path)
        self.host_only = host_only
        L.ref_open.argtypes = [ctypes.c_char_p] * 3
        L.ref_render.restype = ctypes.c_double
        L.ref_render.argtypes = [ctypes.c_int]
        L.ref_last_trace_ms.restype = ctypes.c_double
        for fn in ("ref_last_shrink_ms", "ref_last_pass_ms"):
            if hasattr(L, fn):
                getattr(L, fn).restype = ctypes.c_double
        if hasattr(L, "ref_render_per_pass"):
            L.ref_render_per_pass.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.ref_num_bvh_nodes.restype = ctypes.c_longlong
        L.ref_pass_instrumented.restype = ctypes.c_longlong
        L.ref_pass_instrumented.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.ref_kat_fresnel_dielectric.restype = ctypes.c_float
        L.ref_kat_fresnel_dielectric.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
        L.ref_kat_fresnel_conductor.restype = ctypes.c_float
        L.ref_kat_fresnel_conductor.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_float]
        L.ref_kat_cube_uv.argtypes = [ctypes.c_float] * 3 + [ctypes.c_void_p]
        L.ref_kat_texture.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
        L.ref_kat_rng.argtypes = [ctypes.c_uint, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
        L.ref_default_camera.argtypes = [ctypes.c_float] * 4 + [ctypes.c_void_p]
        self.root = None

    # -- scene lifecycle -------------------------------------------------------------------
    def open(self, root, config_rel="res\\configuration\\config.json", scene_dir="res\\scene", scene="cornell_box_simple"):
        self.root = root
        cwd = os.getcwd()
        os.chdir(root)
        try:
            rc = self.lib.ref_open(config_rel.encode(), scene_dir.encode(), scene.encode())
        finally:
            os.chdir(cwd)
        if rc != 0:
            raise RuntimeError("ref_open failed rc=%d" % rc)
        self.w, self.h = self.lib.ref_width(), self.lib.ref_height()
        return self

    def close(self):
        self.lib.ref_close()

    def camera(self):
        cam = np.zeros(16, np.float32)
        self.lib.ref_get_camera(cam.ctypes.data_as(ctypes.c_void_p))
        return cam

    def set_camera(self, cam16):
        cam = np.ascontiguousarray(cam16, np.float32)
        self.lib.ref_set_camera(cam.ctypes.data_as(ctypes.c_void_p))

    def render(self, n):
        return self.lib.ref_render(int(n))

    def prefetch(self, advise=True):
        """Moves all managed buffers to the GPU (and marks the config read-mostly): steady-state timing."""
        return self.lib.ref_prefetch(1 if advise else 0)

    def clear(self):
        self.lib.ref_clear()

    def image_f32(self):
        out = np.zeros((self.h, self.w, 3), np.float32)
        self.lib.ref_image_f32(out.ctypes.data_as(ctypes.c_void_p))
        return out

    def image_u8(self):
        out = np.zeros((self.h, self.w, 3), np.uint8)
        self.lib.ref_image_u8(out.ctypes.data_as(ctypes.c_void_p))
        return out

    def last_pass_f32(self):
        out = np.zeros((self.h, self.w, 3), np.float32)
        self.lib.ref_last_pass_f32(out.ctypes.data_as(ctypes.c_void_p))
        return out

    def triangles(self):
        n = self.lib.ref_num_triangles()
        tri = np.zeros((n, 24), np.float32)
        mat = np.zeros(n, np.int32)
        if n:
            self.lib.ref_get_triangles(tri.ctypes.data_as(ctypes.c_void_p), mat.ctypes.data_as(ctypes.c_void_p))
        return tri, mat

    def mesh_materials(self):
        n = self.lib.ref_num_mesh_materials()
        out = np.zeros((n, 21), np.uint32)
        if n:
            self.lib.ref_get_mesh_materials(out.ctypes.data_as(ctypes.c_void_p))
        return out

    def spheres(self):
        n = self.lib.ref_num_spheres()
        out = np.zeros((n, 25), np.uint32)
        if n:
            self.lib.ref_get_spheres(out.ctypes.data_as(ctypes.c_void_p))
        return out

    # -- live edits, as path_tracer::render_ui issues them (Core/path_tracer.cpp:109-369) -------
    def set_sphere(self, index, sphere100):
        a = np.ascontiguousarray(sphere100)
        assert a.nbytes == 100
        if self.lib.ref_set_sphere(int(index), a.ctypes.data_as(ctypes.c_void_p)) != 0:
            raise RuntimeError("ref_set_sphere failed")

    def set_mesh_materials(self, mesh, mats84):
        a = np.ascontiguousarray(mats84)
        assert a.nbytes % 84 == 0
        if self.lib.ref_set_mesh_materials(int(mesh), a.ctypes.data_as(ctypes.c_void_p), a.nbytes // 84) != 0:
            raise RuntimeError("ref_set_mesh_materials failed")

    def set_mesh_transform(self, mesh, position, scale):
        p, s = np.ascontiguousarray(position, np.float32), np.ascontiguousarray(scale, np.float32)
        if self.lib.ref_set_mesh_transform(int(mesh), p.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p)) != 0:
            raise RuntimeError("ref_set_mesh_transform failed")

    def apply_mesh_rotate(self, mesh, rotate_degrees):
        a = np.ascontiguousarray(rotate_degrees, np.float32)
        if self.lib.ref_apply_mesh_rotate(int(mesh), a.ctypes.data_as(ctypes.c_void_p)) != 0:
            raise RuntimeError("ref_apply_mesh_rotate failed")

    def mesh_material_counts(self):
        out = np.zeros(self.lib.ref_num_meshes(), np.int32)
        self.lib.ref_mesh_material_counts(out.ctypes.data_as(ctypes.c_void_p))
        return out.tolist()

    def trace_batch(self, rays6):
        rays = np.ascontiguousarray(rays6, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        rc = self.lib.ref_trace_batch(rays.ctypes.data_as(ctypes.c_void_p), n, prim.ctypes.data_as(ctypes.c_void_p), t.ctypes.data_as(ctypes.c_void_p))
        if rc != 0:
            raise RuntimeError("ref_trace_batch failed")
        return prim, t

    def capture_rays(self, pass_counter, depth, max_out=None):
        max_out = max_out or self.w * self.h
        pix = np.zeros(max_out, np.int32)
        rays = np.zeros((max_out, 6), np.float32)
        n = self.lib.ref_pass_instrumented(int(pass_counter), int(depth), pix.ctypes.data_as(ctypes.c_void_p), rays.ctypes.data_as(ctypes.c_void_p), max_out)
        return pix[:n].copy(), rays[:n].copy()

    def render_per_pass(self, n_passes):
        """Wall seconds of each of n synchronous passes through the unmodified path_tracer_kernel()."""
        out = np.zeros(n_passes, np.float64)
        if self.lib.ref_render_per_pass(int(n_passes), out.ctypes.data_as(ctypes.c_void_p)):
            raise RuntimeError("ref_render_per_pass failed")
        return out

    def pass_breakdown(self, pass_counter):
        """One pass of the re-issued loop: (wall ms of the pass, ms in trace_ray_kernel by CUDA events, wall ms in thread_shrink)."""
        self.lib.ref_pass_instrumented(int(pass_counter), -1, None, None, 0)
        return float(self.lib.ref_last_pass_ms()), float(self.lib.ref_last_trace_ms()), float(self.lib.ref_last_shrink_ms())

    def pass_instrumented(self, pass_counter):
        """Runs one full pass; returns (ray segments, ms spent in trace_ray_kernel)."""
        seg = self.lib.ref_pass_instrumented(int(pass_counter), -1, None, None, 0)
        return int(seg), float(self.lib.ref_last_trace_ms())

    # -- KAT helpers (host-evaluated reference header functions) ----------------------------
    def builtin_material(self, name):
        out = np.zeros(21, np.uint32)
        rc = self.lib.ref_builtin_material(name.encode(), out.ctypes.data_as(ctypes.c_void_p))
        return None if rc else out

    def builtin_material_names(self):
        buf = ctypes.create_string_buffer(4096)
        self.lib.ref_builtin_material_names(buf, 4096)
        return [s for s in buf.value.decode().split("\n") if s]

    def default_camera(self, w, h, aperture=-1.0, focal=-1.0):
        cam = np.zeros(16, np.float32)
        self.lib.ref_default_camera(float(w), float(h), float(aperture), float(focal), cam.ctypes.data_as(ctypes.c_void_p))
        return cam
