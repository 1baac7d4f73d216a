"""ctypes binding of libptb200.so (the C ABI in include/ptb200.h) plus a host-side mirror of the
reference's `path_tracer` class (Core/path_tracer.h:54-61: init / init_scene_device_data / render /
clear / release_scene_device_data).

There is no CPU or PyTorch fallback: if the CUDA library is missing the import of the binding fails,
and every compute call fails on a host-only handle.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PTB200_LIB selects an experimental build variant of the same library (tools/ only)
_LIB_PATH = os.environ.get("PTB200_LIB") or os.path.join(_HERE, "libptb200.so")


class PtbError(RuntimeError):
    pass


class Camera(ctypes.Structure):
    """Reference `render_camera` layout (Core/camera.h:14-23), 64 bytes."""
    _fields_ = [("eye", ctypes.c_float * 3), ("view", ctypes.c_float * 3), ("up", ctypes.c_float * 3), ("_pad", ctypes.c_float),
                ("resolution", ctypes.c_float * 2), ("fov", ctypes.c_float * 2), ("aperture_radius", ctypes.c_float),
                ("focal_distance", ctypes.c_float)]

    def as_array(self):
        return np.frombuffer(bytes(self), dtype=np.float32).copy()

    @staticmethod
    def from_array(a):
        a = np.ascontiguousarray(a, np.float32)
        assert a.size == 16
        return Camera.from_buffer_copy(a.tobytes())


class Stats(ctypes.Structure):
    _fields_ = [("passes", ctypes.c_int64), ("ray_segments", ctypes.c_int64), ("kernel_launches", ctypes.c_int64),
                ("gpu_ms_total", ctypes.c_double), ("gpu_ms_extend", ctypes.c_double), ("bvh_nodes", ctypes.c_int64),
                ("bvh_bytes", ctypes.c_int64), ("nodes_visited", ctypes.c_int64), ("tris_tested", ctypes.c_int64),
                ("wide_nodes_visited", ctypes.c_int64)]


MATERIAL_DTYPE = np.dtype([("diffuse_color", np.float32, 3), ("emission_color", np.float32, 3), ("specular_color", np.float32, 3),
                           ("is_transparent", np.uint8), ("_pad", np.uint8, 3), ("roughness", np.float32),
                           ("refraction_index", np.float32), ("extinction_coefficient", np.float32),
                           ("absorption_coefficient", np.float32, 3), ("reduced_scattering_coefficient", np.float32, 3),
                           ("diffuse_texture_id", np.int32), ("specular_texture_id", np.int32)])
assert MATERIAL_DTYPE.itemsize == 84
SPHERE_DTYPE = np.dtype([("center", np.float32, 3), ("radius", np.float32), ("mat", MATERIAL_DTYPE)])
assert SPHERE_DTYPE.itemsize == 100
CONFIG_DTYPE = np.dtype([("width", np.int32), ("height", np.int32), ("use_fullscreen", np.uint8), ("_p0", np.uint8, 3),
                         ("block_size", np.int32), ("max_block_size", np.int32), ("max_tracer_depth", np.int32),
                         ("vector_bias_length", np.float32), ("energy_exist_threshold", np.float32), ("sss_threshold", np.float32),
                         ("use_sky_box", np.uint8), ("use_sky", np.uint8), ("use_bilinear", np.uint8), ("gamma_correction", np.uint8),
                         ("use_anti_alias", np.uint8), ("_p1", np.uint8, 3), ("fov", np.float32),
                         ("bvh_leaf_node_triangle_num", np.int32), ("bvh_bucket_max_divide_internal_num", np.int32),
                         ("bvh_build_block_size", np.int32), ("bvh_build", np.int32), ("air_refraction_index", np.float32),
                         ("air_absorption_coef", np.float32, 3), ("air_reduced_scattering_coef", np.float32, 3),
                         ("cuda_acceleration", np.uint8), ("_p2", np.uint8, 3)])
assert CONFIG_DTYPE.itemsize == 96

_lib = None


def load_library():
    """Loads libptb200.so, building it first if the sources are newer. Fails loudly if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        from . import build as _build
        _build.build()
    if "PTB200_NCCL_LIB" not in os.environ:
        # multi-GPU entry points dlopen NCCL lazily (csrc/multi.inc).  In a Python process torch may be imported later and needs ITS libnccl.so.2
        # (pip package nvidia-nccl); a different copy loaded first under the same soname would break that import, so point the library at it.
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            for loc in (spec.submodule_search_locations if spec else []):
                cand = os.path.join(loc, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["PTB200_NCCL_LIB"] = cand
                    break
        except Exception:
            pass
    L = ctypes.CDLL(_LIB_PATH)
    vp, ci, cf, cp = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_char_p
    L.ptb_last_error.restype = cp
    L.ptb_create.restype = vp
    L.ptb_create.argtypes = [cp, ci]
    L.ptb_destroy.argtypes = [vp]
    L.ptb_list_scenes.argtypes = [cp, cp, ci]
    L.ptb_load_scene.argtypes = [vp, cp, cp]
    L.ptb_release_scene.argtypes = [vp]
    L.ptb_default_camera.argtypes = [cf, cf, cf, cf, ctypes.POINTER(Camera)]
    L.ptb_set_camera.argtypes = [vp, ctypes.POINTER(Camera)]
    L.ptb_get_camera.argtypes = [vp, ctypes.POINTER(Camera)]
    L.ptb_render.argtypes = [vp, ci]
    L.ptb_render_async.argtypes = [vp, ci]
    L.ptb_render_strided.argtypes = [vp, ci, ci, ci]
    L.ptb_synchronize.argtypes = [vp]
    L.ptb_stream.restype = vp
    L.ptb_stream.argtypes = [vp]
    L.ptb_clear.argtypes = [vp]
    L.ptb_pass_counter.argtypes = [vp]
    L.ptb_width.argtypes = [vp]
    L.ptb_height.argtypes = [vp]
    L.ptb_image_f32.argtypes = [vp, vp, ctypes.POINTER(ci)]
    L.ptb_image_u8.argtypes = [vp, vp]
    L.ptb_last_pass_f32.argtypes = [vp, vp]
    L.ptb_image_device_ptr.restype = vp
    L.ptb_image_device_ptr.argtypes = [vp]
    L.ptb_finalize.argtypes = [vp, ci]
    L.ptb_trace_batch.argtypes = [vp, vp, ci, vp, vp, vp]
    L.ptb_trace_batch_bruteforce.argtypes = [vp, vp, ci, vp, vp]
    L.ptb_generate_rays.argtypes = [vp, ci, vp]
    L.ptb_capture_rays.argtypes = [vp, ci, ci, vp, vp, ci]
    L.ptb_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.ptb_get_depth_profile.argtypes = [vp, ci, vp, vp]
    L.ptb_set_option.argtypes = [vp, cp, cp]
    L.ptb_get_traversal_histogram.argtypes = [vp, vp]
    L.ptb_save_png.argtypes = [vp, cp]
    L.ptb_save_pfm.argtypes = [vp, cp]
    L.ptb_save_checkpoint.argtypes = [vp, cp]
    L.ptb_load_checkpoint.argtypes = [vp, cp, ci]
    L.ptb_write_png_rgb8.argtypes = [cp, vp, ci, ci]
    L.ptb_write_pfm.argtypes = [cp, vp, ci, ci, cf]
    L.ptb_set_sphere.argtypes = [vp, ci, vp]
    L.ptb_set_mesh_material.argtypes = [vp, ci, vp, ci]
    L.ptb_set_mesh_transform.argtypes = [vp, ci, vp, vp]
    L.ptb_apply_mesh_rotate.argtypes = [vp, ci, vp]
    L.ptb_get_mesh_placement.argtypes = [vp, ci, vp, vp, vp] + [ctypes.POINTER(ci)] * 4
    L.ptb_bvh_info.argtypes = [vp, vp, vp]
    L.ptb_bvh_leaf_labels.argtypes = [vp, vp]
    L.ptb_bvh_download.argtypes = [vp, vp, vp]
    L.ptb_scene_counts.argtypes = [vp] + [ctypes.POINTER(ci)] * 6
    L.ptb_scene_triangles.argtypes = [vp, vp, vp]
    L.ptb_scene_materials.argtypes = [vp, vp]
    L.ptb_scene_spheres.argtypes = [vp, vp]
    L.ptb_scene_texture.argtypes = [vp, ci, ctypes.POINTER(ci), ctypes.POINTER(ci), vp]
    L.ptb_scene_cubemap_face.argtypes = [vp, ci, vp]
    L.ptb_get_config.argtypes = [vp, vp]
    L.ptb_set_config.argtypes = [vp, vp]
    L.ptb_decode_image.argtypes = [cp, ctypes.POINTER(ci), ctypes.POINTER(ci), vp]
    L.ptb_set_jpeg_decode.argtypes = [cp]
    # multi-GPU (csrc/multi.inc)
    L.ptb_multi_create.restype = vp
    L.ptb_multi_create.argtypes = [cp, ci, vp]
    L.ptb_multi_destroy.argtypes = [vp]
    L.ptb_multi_device_count.argtypes = [vp]
    L.ptb_multi_renderer.restype = vp
    L.ptb_multi_renderer.argtypes = [vp, ci]
    L.ptb_multi_set_option.argtypes = [vp, cp, cp]
    L.ptb_multi_load_scene.argtypes = [vp, cp, cp]
    L.ptb_multi_set_camera.argtypes = [vp, ctypes.POINTER(Camera)]
    L.ptb_multi_render.argtypes = [vp, ci]
    L.ptb_multi_clear.argtypes = [vp]
    L.ptb_multi_pass_counter.argtypes = [vp]
    L.ptb_multi_image_f32.argtypes = [vp, vp, ctypes.POINTER(ci)]
    L.ptb_multi_image_u8.argtypes = [vp, vp]
    L.ptb_dist_unique_id.argtypes = [vp]
    L.ptb_dist_init.argtypes = [vp, ci, ci, vp]
    L.ptb_dist_shutdown.argtypes = [vp]
    L.ptb_dist_broadcast_scene.argtypes = [vp, ci]
    L.ptb_dist_render.argtypes = [vp, ci]
    L.ptb_dist_load_scene.argtypes = [vp, cp, cp, ci]
    L.ptb_dist_broadcast_timing.argtypes = [vp, vp]
    L.ptb_dist_reduce.argtypes = [vp, ci]
    L.ptb_dist_clear.argtypes = [vp]
    L.ptb_merged_image_f32.argtypes = [vp, vp, ctypes.POINTER(ci)]
    L.ptb_merged_image_u8.argtypes = [vp, vp]
    L.ptb_test_scene_blob_roundtrip.argtypes = [vp]
    _lib = L
    return L


def last_error():
    return load_library().ptb_last_error().decode(errors="replace")


def device_count():
    return load_library().ptb_device_count()


def default_camera(width, height, aperture_radius=-1.0, focal_distance=-1.0):
    cam = Camera()
    load_library().ptb_default_camera(float(width), float(height), float(aperture_radius), float(focal_distance), ctypes.byref(cam))
    return cam


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def decode_image(path):
    """RGBA8 (H, W, 4) the scene front-end decodes from an image file (Others/image_loader.cpp:31-95)."""
    L = load_library()
    w, h = ctypes.c_int(), ctypes.c_int()
    if L.ptb_decode_image(os.fsencode(path), ctypes.byref(w), ctypes.byref(h), None) != 0:
        raise PtbError(last_error())
    out = np.zeros((h.value, w.value, 4), np.uint8)
    if L.ptb_decode_image(os.fsencode(path), ctypes.byref(w), ctypes.byref(h), _ptr(out)) != 0:
        raise PtbError(last_error())
    return out


def set_jpeg_decode(mode):
    """'reference' (default: FreeImage as the reference calls it — ifast IDCT, replicated chroma), 'fast' (libjpeg-turbo fast
    decode) or 'accurate' (libjpeg-turbo default decode == PIL); process-wide, applies to later loads."""
    if load_library().ptb_set_jpeg_decode(mode.encode()) != 0:
        raise PtbError(last_error())


def write_png(path, rgb_u8):
    a = np.ascontiguousarray(rgb_u8, np.uint8)
    if load_library().ptb_write_png_rgb8(os.fsencode(path), _ptr(a), a.shape[1], a.shape[0]) != 0:
        raise PtbError(last_error())


def write_pfm(path, rgb_f32, scale=1.0):
    a = np.ascontiguousarray(rgb_f32, np.float32)
    if load_library().ptb_write_pfm(os.fsencode(path), _ptr(a), a.shape[1], a.shape[0], float(scale)) != 0:
        raise PtbError(last_error())


class Renderer:
    """Thin object wrapper over the ptb_* handle API."""

    def __init__(self, config_json_path, device=0):
        self.lib = load_library()
        self.handle = self.lib.ptb_create(os.fsencode(config_json_path), int(device))
        if not self.handle:
            raise PtbError(last_error())
        self.device = device
        self.width = self.lib.ptb_width(self.handle)
        self.height = self.lib.ptb_height(self.handle)

    @classmethod
    def _borrowed(cls, handle, device):
        """View of a handle owned by someone else (MultiRenderer.renderer(i)): never destroyed from here."""
        self = cls.__new__(cls)
        self.lib = load_library()
        self.handle = handle
        self.device = device
        self.width = self.lib.ptb_width(handle)
        self.height = self.lib.ptb_height(handle)
        self._owned = False
        return self

    def close(self):
        if self.handle:
            if getattr(self, "_owned", True):
                self.lib.ptb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PtbError(last_error())

    # -- one process per GPU: NCCL communicator, scene broadcast, sharded render, reduce (csrc/multi.inc) ---------
    def dist_init(self, rank, world_size, unique_id):
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._check(self.lib.ptb_dist_init(self.handle, int(rank), int(world_size), buf))

    def dist_shutdown(self):
        self.lib.ptb_dist_shutdown(self.handle)

    def dist_broadcast_scene(self, root=0):
        self._check(self.lib.ptb_dist_broadcast_scene(self.handle, int(root)))

    def dist_load_scene(self, scene_json_path, asset_root="", root=0):
        self._check(self.lib.ptb_dist_load_scene(self.handle, os.fsencode(scene_json_path or ""), os.fsencode(asset_root or ""), int(root)))

    def dist_broadcast_timing(self):
        out = np.zeros(6, np.float64)
        self._check(self.lib.ptb_dist_broadcast_timing(self.handle, _ptr(out)))
        return dict(zip(["stage_ms", "broadcast_ms", "unpack_ms", "upload_build_ms", "total_ms", "bytes"], out.tolist()))

    def dist_render(self, total_passes):
        self._check(self.lib.ptb_dist_render(self.handle, int(total_passes)))

    def dist_reduce(self, root=0):
        self._check(self.lib.ptb_dist_reduce(self.handle, int(root)))

    def scene_blob_roundtrip(self):
        self._check(self.lib.ptb_test_scene_blob_roundtrip(self.handle))

    def dist_clear(self):
        self._check(self.lib.ptb_dist_clear(self.handle))

    def merged_image_f32(self):
        out = np.zeros((self.height, self.width, 3), np.float32)
        n = ctypes.c_int()
        self._check(self.lib.ptb_merged_image_f32(self.handle, _ptr(out), ctypes.byref(n)))
        return out, n.value

    def merged_image_u8(self, out=None):
        if out is None:
            out = np.zeros((self.height, self.width, 3), np.uint8)
        self._check(self.lib.ptb_merged_image_u8(self.handle, _ptr(out)))
        return out

    # -- scene ---------------------------------------------------------------------------------
    def load_scene(self, scene_json_path, asset_root=""):
        self._check(self.lib.ptb_load_scene(self.handle, os.fsencode(scene_json_path), os.fsencode(asset_root)))

    def release_scene(self):
        self._check(self.lib.ptb_release_scene(self.handle))

    def set_option(self, key, value):
        self._check(self.lib.ptb_set_option(self.handle, key.encode(), str(value).encode()))

    def config(self):
        out = np.zeros(1, CONFIG_DTYPE)
        self._check(self.lib.ptb_get_config(self.handle, _ptr(out)))
        return out[0]

    def set_config(self, config):
        a = np.ascontiguousarray(np.asarray(config, CONFIG_DTYPE).reshape(1))
        self._check(self.lib.ptb_set_config(self.handle, _ptr(a)))

    def scene_counts(self):
        vals = [ctypes.c_int() for _ in range(6)]
        self._check(self.lib.ptb_scene_counts(self.handle, *[ctypes.byref(v) for v in vals]))
        keys = ["triangles", "materials", "spheres", "textures", "cube_length", "meshes"]
        return dict(zip(keys, [v.value for v in vals]))

    def scene_triangles(self):
        n = self.scene_counts()["triangles"]
        tri = np.zeros((n, 24), np.float32)
        mat = np.zeros(n, np.int32)
        self._check(self.lib.ptb_scene_triangles(self.handle, _ptr(tri), _ptr(mat)))
        return tri, mat

    def scene_materials(self):
        out = np.zeros(self.scene_counts()["materials"], MATERIAL_DTYPE)
        if out.size:
            self._check(self.lib.ptb_scene_materials(self.handle, _ptr(out)))
        return out

    def scene_spheres(self):
        out = np.zeros(self.scene_counts()["spheres"], SPHERE_DTYPE)
        if out.size:
            self._check(self.lib.ptb_scene_spheres(self.handle, _ptr(out)))
        return out

    def scene_texture(self, index):
        w, h = ctypes.c_int(), ctypes.c_int()
        self._check(self.lib.ptb_scene_texture(self.handle, index, ctypes.byref(w), ctypes.byref(h), None))
        out = np.zeros((h.value, w.value, 4), np.uint8)
        self._check(self.lib.ptb_scene_texture(self.handle, index, ctypes.byref(w), ctypes.byref(h), _ptr(out)))
        return out

    def scene_cubemap(self):
        n = self.scene_counts()["cube_length"]
        out = np.zeros((6, n, n, 4), np.uint8)
        for f in range(6):
            self._check(self.lib.ptb_scene_cubemap_face(self.handle, f, _ptr(out[f])))
        return out

    # -- output side (Main/window.cpp:712-740) ---------------------------------------------------
    def save_png(self, path):
        self._check(self.lib.ptb_save_png(self.handle, os.fsencode(path)))

    def save_pfm(self, path):
        self._check(self.lib.ptb_save_pfm(self.handle, os.fsencode(path)))

    def save_checkpoint(self, path):
        self._check(self.lib.ptb_save_checkpoint(self.handle, os.fsencode(path)))

    def load_checkpoint(self, path, restore_camera=True):
        self._check(self.lib.ptb_load_checkpoint(self.handle, os.fsencode(path), 1 if restore_camera else 0))

    # -- live edits (Core/path_tracer.cpp:109-369) ----------------------------------------------
    def set_sphere(self, index, sphere):
        a = np.ascontiguousarray(np.asarray(sphere, SPHERE_DTYPE).reshape(1))
        self._check(self.lib.ptb_set_sphere(self.handle, int(index), _ptr(a)))

    def set_mesh_material(self, mesh, mats):
        a = np.ascontiguousarray(np.asarray(mats, MATERIAL_DTYPE).reshape(-1))
        self._check(self.lib.ptb_set_mesh_material(self.handle, int(mesh), _ptr(a), int(a.size)))

    def set_mesh_transform(self, mesh, position, scale):
        p, s = np.ascontiguousarray(position, np.float32), np.ascontiguousarray(scale, np.float32)
        self._check(self.lib.ptb_set_mesh_transform(self.handle, int(mesh), _ptr(p), _ptr(s)))

    def apply_mesh_rotate(self, mesh, rotate_degrees):
        a = np.ascontiguousarray(rotate_degrees, np.float32)
        self._check(self.lib.ptb_apply_mesh_rotate(self.handle, int(mesh), _ptr(a)))

    def mesh_placement(self, mesh):
        p, s, ro = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
        first, count, mfirst, mcount = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self.lib.ptb_get_mesh_placement(self.handle, int(mesh), _ptr(p), _ptr(s), _ptr(ro), ctypes.byref(first), ctypes.byref(count),
                                                    ctypes.byref(mfirst), ctypes.byref(mcount)))
        return {"position": p, "scale": s, "rotate": ro, "first_triangle": first.value, "triangle_count": count.value,
                "first_material": mfirst.value, "material_count": mcount.value}

    def mesh_material_counts(self):
        return [self.mesh_placement(i)["material_count"] for i in range(self.scene_counts()["meshes"])]

    # -- camera --------------------------------------------------------------------------------
    def camera(self):
        cam = Camera()
        self._check(self.lib.ptb_get_camera(self.handle, ctypes.byref(cam)))
        return cam

    def set_camera(self, cam):
        if not isinstance(cam, Camera):
            cam = Camera.from_array(cam)
        self._check(self.lib.ptb_set_camera(self.handle, ctypes.byref(cam)))

    # -- rendering -----------------------------------------------------------------------------
    def render(self, n_passes=1):
        self._check(self.lib.ptb_render(self.handle, int(n_passes)))

    def render_async(self, n_passes=1):
        self._check(self.lib.ptb_render_async(self.handle, int(n_passes)))

    def render_strided(self, first_pass, stride, n_passes):
        self._check(self.lib.ptb_render_strided(self.handle, int(first_pass), int(stride), int(n_passes)))

    def synchronize(self):
        self._check(self.lib.ptb_synchronize(self.handle))

    def stream(self):
        return self.lib.ptb_stream(self.handle)

    def clear(self):
        self._check(self.lib.ptb_clear(self.handle))

    def pass_counter(self):
        return self.lib.ptb_pass_counter(self.handle)

    def image_f32(self, out=None):
        if out is None:
            out = np.zeros((self.height, self.width, 3), np.float32)
        passes = ctypes.c_int()
        self._check(self.lib.ptb_image_f32(self.handle, _ptr(out), ctypes.byref(passes)))
        return out

    def image_u8(self, out=None):
        if out is None:
            out = np.zeros((self.height, self.width, 3), np.uint8)
        self._check(self.lib.ptb_image_u8(self.handle, _ptr(out)))
        return out

    def last_pass_f32(self):
        out = np.zeros((self.height, self.width, 3), np.float32)
        self._check(self.lib.ptb_last_pass_f32(self.handle, _ptr(out)))
        return out

    def image_device_ptr(self):
        return self.lib.ptb_image_device_ptr(self.handle)

    def finalize(self, total_passes):
        self._check(self.lib.ptb_finalize(self.handle, int(total_passes)))

    def trace_batch(self, rays6, bruteforce=False, with_bary=False):
        rays = np.ascontiguousarray(rays6, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        if bruteforce:
            self._check(self.lib.ptb_trace_batch_bruteforce(self.handle, _ptr(rays), n, _ptr(prim), _ptr(t)))
            return prim, t
        bary = np.zeros((n, 2), np.float32) if with_bary else None
        self._check(self.lib.ptb_trace_batch(self.handle, _ptr(rays), n, _ptr(prim), _ptr(t), _ptr(bary) if with_bary else None))
        return (prim, t, bary) if with_bary else (prim, t)

    def generate_rays(self, pass_index):
        out = np.zeros((self.width * self.height, 6), np.float32)
        self._check(self.lib.ptb_generate_rays(self.handle, int(pass_index), _ptr(out)))
        return out

    def capture_rays(self, pass_index, depth):
        n_max = self.width * self.height
        pix = np.zeros(n_max, np.int32)
        rays = np.zeros((n_max, 6), np.float32)
        n = self.lib.ptb_capture_rays(self.handle, int(pass_index), int(depth), _ptr(pix), _ptr(rays), n_max)
        if n < 0:
            raise PtbError(last_error())
        return pix[:n].copy(), rays[:n].copy()

    def depth_profile(self):
        seg = np.zeros(256, np.int64)
        ms = np.zeros(256, np.float64)
        n = self.lib.ptb_get_depth_profile(self.handle, 256, _ptr(seg), _ptr(ms))
        return seg[:max(n, 0)].copy(), ms[:max(n, 0)].copy()

    def bvh_info(self):
        oi = np.zeros(10, np.int64)
        od = np.zeros(4, np.float64)
        self._check(self.lib.ptb_bvh_info(self.handle, _ptr(oi), _ptr(od)))
        return {"node_records": int(oi[0]), "inner_nodes": int(oi[1]), "leaves": int(oi[2]), "depth": int(oi[3]), "valid": bool(oi[4]),
                "built_on_gpu": bool(oi[5]), "levels": int(oi[6]), "small_subtrees": int(oi[7]), "build_ms": float(od[0]),
                "sah_cost": float(od[1]), "upload_ms": float(od[2]), "violations": int(od[3]),
                "wide_nodes": int(oi[8]), "wide_collapsed_on_gpu": bool(oi[9])}

    def bvh_download(self):
        nodes = np.zeros((self.bvh_info()["node_records"], 16), np.float32)
        order = np.zeros(self.scene_counts()["triangles"], np.int32)
        self._check(self.lib.ptb_bvh_download(self.handle, _ptr(nodes), _ptr(order)))
        return nodes, order

    def bvh_leaf_labels(self):
        out = np.zeros(self.scene_counts()["triangles"], np.int32)
        self._check(self.lib.ptb_bvh_leaf_labels(self.handle, _ptr(out)))
        return out

    def traversal_histogram(self):
        out = np.zeros(25, np.int64)
        self._check(self.lib.ptb_get_traversal_histogram(self.handle, _ptr(out)))
        return int(out[0]), out[1:].copy()

    def stats(self):
        s = Stats()
        self._check(self.lib.ptb_get_stats(self.handle, ctypes.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}


def dist_unique_id():
    """128-byte NCCL id for Renderer.dist_init: created on one rank, handed to all ranks by the host."""
    buf = ctypes.create_string_buffer(128)
    if load_library().ptb_dist_unique_id(buf) != 0:
        raise PtbError(last_error())
    return buf.raw


def nccl_version():
    return load_library().ptb_nccl_version()


class MultiRenderer:
    """One process, every GPU of the box (ptb_multi_*): the scene is parsed once, every device builds its own BVH, passes are
    sharded by index, one NCCL reduce per render call merges the accumulation buffers on device 0."""

    def __init__(self, config_json_path, n_devices, devices=None):
        self.lib = load_library()
        dev = None
        if devices is not None:
            dev = (ctypes.c_int * n_devices)(*[int(d) for d in devices])
        self.handle = self.lib.ptb_multi_create(os.fsencode(config_json_path), int(n_devices), dev)
        if not self.handle:
            raise PtbError(last_error())
        self.n_devices = n_devices
        first = self.renderer(0)
        self.width, self.height = first.width, first.height

    def _check(self, rc):
        if rc != 0:
            raise PtbError(last_error())

    def close(self):
        if self.handle:
            self.lib.ptb_multi_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def renderer(self, index):
        h = self.lib.ptb_multi_renderer(self.handle, int(index))
        if not h:
            raise PtbError("device index out of range")
        return Renderer._borrowed(h, index)

    def set_option(self, key, value):
        self._check(self.lib.ptb_multi_set_option(self.handle, key.encode(), str(value).encode()))

    def load_scene(self, scene_json_path, asset_root=""):
        self._check(self.lib.ptb_multi_load_scene(self.handle, os.fsencode(scene_json_path), os.fsencode(asset_root)))

    def set_camera(self, cam):
        if not isinstance(cam, Camera):
            c = Camera()
            ctypes.memmove(ctypes.byref(c), np.ascontiguousarray(cam).ctypes.data, 64)
            cam = c
        self._check(self.lib.ptb_multi_set_camera(self.handle, ctypes.byref(cam)))

    def render(self, total_passes):
        self._check(self.lib.ptb_multi_render(self.handle, int(total_passes)))

    def clear(self):
        self._check(self.lib.ptb_multi_clear(self.handle))

    def pass_counter(self):
        return self.lib.ptb_multi_pass_counter(self.handle)

    def image_f32(self):
        out = np.zeros((self.height, self.width, 3), np.float32)
        n = ctypes.c_int()
        self._check(self.lib.ptb_multi_image_f32(self.handle, _ptr(out), ctypes.byref(n)))
        return out

    def image_u8(self):
        out = np.zeros((self.height, self.width, 3), np.uint8)
        self._check(self.lib.ptb_multi_image_u8(self.handle, _ptr(out)))
        return out


class PathTracer:
    """Mirror of the reference's `class path_tracer` (Core/path_tracer.h:31-61, path_tracer.cpp:18-106,371-406).

    init(render_camera, config_path, scene_dir) -> list of scene files; init_scene_device_data(index) -> bool;
    render() -> image dict or None when no scene is initiated; clear(); release_scene_device_data().
    """

    def __init__(self, device=0):
        self._device = device
        self._r = None
        self._scene_files = []
        self._asset_root = ""
        self._is_initiated = False
        self._render_camera = None

    def init(self, render_camera, config_path, scene_file_directory, asset_root=None):
        self._r = Renderer(config_path, self._device)
        self._render_camera = render_camera
        # the reference resolves paths against its CWD: "res\\scene" lives directly under it
        self._asset_root = asset_root if asset_root is not None else os.getcwd()
        buf = ctypes.create_string_buffer(1 << 16)
        scene_dir = scene_file_directory.replace("\\", "/")
        if not os.path.isabs(scene_dir):
            scene_dir = os.path.join(self._asset_root, scene_dir)
        n = self._r.lib.ptb_list_scenes(os.fsencode(scene_dir), buf, len(buf))
        self._scene_files = [s for s in buf.value.decode().split("\n") if s] if n > 0 else []
        if not self._scene_files:
            print("[Warn]There exists no scene file!")
        return list(self._scene_files)

    def init_scene_device_data(self, index):
        if self._r is None or index < 0 or index >= len(self._scene_files):
            self._is_initiated = False
            return False
        try:
            self._r.load_scene(self._scene_files[index], self._asset_root)
        except PtbError as e:
            print(str(e))
            print("[Error]Load scene failed!")
            self._is_initiated = False
            return False
        self._is_initiated = True
        return True

    def render(self):
        if not self._is_initiated:
            return None
        if self._render_camera is not None:
            self._r.set_camera(self._render_camera)
        self._r.render(1)
        return {"width": self._r.width, "height": self._r.height, "pixel_count": self._r.width * self._r.height,
                "pass_counter": self._r.pass_counter(), "pixels": self._r.image_f32(), "pixels_256": self._r.image_u8()}

    def clear(self):
        if self._is_initiated:
            self._r.clear()

    def release_scene_device_data(self):
        if not self._is_initiated:
            return
        self._r.release_scene()
        self._is_initiated = False

    @property
    def renderer(self):
        return self._r
