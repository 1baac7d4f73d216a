"""TEST INFRASTRUCTURE — ctypes binding of the C oracle (oracle/pt_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libptoracle.so")


def build(force=False):
    src = os.path.join(HERE, "pt_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB


class _Texture(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("pixels", ctypes.c_void_p)]


class _Scene(ctypes.Structure):
    _fields_ = [("n_triangles", ctypes.c_int32), ("triangles", ctypes.c_void_p), ("triangle_material", ctypes.c_void_p),
                ("n_materials", ctypes.c_int32), ("materials", ctypes.c_void_p),
                ("n_spheres", ctypes.c_int32), ("spheres", ctypes.c_void_p),
                ("n_textures", ctypes.c_int32), ("textures", ctypes.c_void_p),
                ("cube_length", ctypes.c_int32), ("cube_faces", ctypes.c_void_p * 6),
                ("n_nodes", ctypes.c_int32), ("nodes", ctypes.c_void_p), ("order", ctypes.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        L.ptbo_render_pass.restype = ctypes.c_longlong
        L.ptbo_render_pass.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.ptbo_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.ptbo_generate_rays.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.ptbo_accumulate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.ptbo_tonemap.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.ptbo_rng.argtypes = [ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
        L.ptbo_hash.restype = ctypes.c_int32
        L.ptbo_hash.argtypes = [ctypes.c_int32]
        L.ptbo_triangle.argtypes = [ctypes.c_void_p] * 3
        L.ptbo_sphere.argtypes = [ctypes.c_void_p] * 3
        L.ptbo_box.argtypes = [ctypes.c_void_p] * 3
        L.ptbo_fresnel_dielectric.restype = ctypes.c_float
        L.ptbo_fresnel_dielectric.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_void_p]
        L.ptbo_fresnel_conductor.restype = ctypes.c_float
        L.ptbo_fresnel_conductor.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_float]
        L.ptbo_cube_uv.argtypes = [ctypes.c_float] * 3 + [ctypes.c_void_p]
        L.ptbo_texture.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
        L.ptbo_background.argtypes = [ctypes.c_void_p] * 4
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleScene:
    """Flat scene arrays (same layouts the product exports through ptb_scene_*) + oracle BVH."""

    def __init__(self, triangles24, triangle_material, materials84, spheres100, textures, cube_faces, config96, camera16):
        L = lib()
        self.tri = np.ascontiguousarray(triangles24, np.float32).reshape(-1, 24)
        self.tri_mat = np.ascontiguousarray(triangle_material, np.int32)
        self.mats = np.ascontiguousarray(materials84).view(np.uint8).reshape(-1, 84) if len(materials84) else np.zeros((0, 84), np.uint8)
        self.spheres = np.ascontiguousarray(spheres100).view(np.uint8).reshape(-1, 100) if len(spheres100) else np.zeros((0, 100), np.uint8)
        self.textures = [np.ascontiguousarray(t, np.uint8) for t in textures]
        self.cube = np.ascontiguousarray(cube_faces, np.uint8)
        self.config = np.ascontiguousarray(config96).view(np.uint8).reshape(96).copy()
        self.camera = np.ascontiguousarray(camera16, np.float32).reshape(16).copy()
        self._tex_structs = (_Texture * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            self._tex_structs[i].width = t.shape[1]
            self._tex_structs[i].height = t.shape[0]
            self._tex_structs[i].pixels = t.ctypes.data
        s = _Scene()
        s.n_triangles = self.tri.shape[0]
        s.triangles = self.tri.ctypes.data
        s.triangle_material = self.tri_mat.ctypes.data
        s.n_materials = self.mats.shape[0]
        s.materials = self.mats.ctypes.data
        s.n_spheres = self.spheres.shape[0]
        s.spheres = self.spheres.ctypes.data
        s.n_textures = len(self.textures)
        s.textures = ctypes.cast(self._tex_structs, ctypes.c_void_p)
        s.cube_length = self.cube.shape[1] if self.cube.size else 0
        for f in range(6):
            s.cube_faces[f] = self.cube[f].ctypes.data if self.cube.size else None
        self.s = s
        L.ptbo_build(ctypes.byref(self.s))
        self.width = int(self.config[0:4].view(np.int32)[0])
        self.height = int(self.config[4:8].view(np.int32)[0])
        self.max_depth = int(self.config[20:24].view(np.int32)[0])

    @staticmethod
    def from_renderer(r):
        """Builds the oracle's input from a (host-only or GPU) product Renderer's scene export."""
        tri, mat = r.scene_triangles()
        counts = r.scene_counts()
        textures = [r.scene_texture(i) for i in range(counts["textures"])]
        return OracleScene(tri, mat, r.scene_materials(), r.scene_spheres(), textures, r.scene_cubemap(), np.array([r.config()]), r.camera().as_array())

    def __del__(self):
        try:
            lib().ptbo_free(ctypes.byref(self.s))
        except Exception:
            pass

    def set_camera(self, cam16):
        self.camera = np.ascontiguousarray(cam16, np.float32).reshape(16).copy()

    def generate_rays(self, pass_index):
        out = np.zeros((self.width * self.height, 6), np.float32)
        lib().ptbo_generate_rays(_p(self.camera), _p(self.config), int(pass_index), _p(out))
        return out

    def trace(self, rays6, brute=False):
        rays = np.ascontiguousarray(rays6, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        prim, t, bary = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros((n, 2), np.float32)
        lib().ptbo_trace(ctypes.byref(self.s), _p(rays), n, 1 if brute else 0, _p(prim), _p(t), _p(bary))
        return prim, t, bary

    def render_pass(self, pass_index, pixel_begin=0, pixel_end=None):
        """Un-clamped radiance of one pass (H, W, 3) and the number of ray segments traced."""
        out = np.zeros((self.height, self.width, 3), np.float32)
        pixel_end = self.width * self.height if pixel_end is None else pixel_end
        seg = lib().ptbo_render_pass(ctypes.byref(self.s), _p(self.camera), _p(self.config), int(pass_index), int(pixel_begin), int(pixel_end), _p(out))
        return out, int(seg)

    def render(self, n_passes, first_pass=1, stride=1):
        """Float accumulation image after passes first, first+stride, ... exactly like the reference's running sum."""
        image = np.zeros((self.height, self.width, 3), np.float32)
        segs = 0
        for k in range(n_passes):
            rad, s = self.render_pass(first_pass + k * stride)
            segs += s
            lib().ptbo_accumulate(_p(image), _p(rad), image.size, 1 if k == 0 else 2, self.max_depth)
        return image, segs

    def tonemap(self, image_sum, passes, gamma=True):
        out = np.zeros(image_sum.shape, np.uint8)
        img = np.ascontiguousarray(image_sum, np.float32)
        lib().ptbo_tonemap(_p(img), _p(out), img.size, int(passes), 1 if gamma else 0)
        return out


def num_threads():
    return lib().ptbo_num_threads()
