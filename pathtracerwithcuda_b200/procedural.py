"""Deterministic procedural assets for the BASELINE.json workloads (there is no network, and the
reference checkout lacks its large meshes — SURVEY.md Appendix E).  Everything is written in the
reference's own input formats (OBJ with `vn`, `g` groups; scene JSON; config JSON; BMP images) so
that this library and the headless reference read THE SAME files.

Workloads (SURVEY.md §8d / BASELINE.md §3.3):
  c1  cornell-class box: 2 spheres (iron, glass) + 5-group box + light      512x512,  depth 5
  c2  two displaced blobs (~150k tris: silver conductor + red dielectric)   1920x1080, depth 8
  c3  textured multi-group interior (~300k tris, 8 textures) + DOF camera   1920x1080, depth 8
  c4  ~1M-tri blob with a random-walk subsurface material                   1920x1080, depth 16
  c5  ~5M tris in two meshes                                                3840x2160, depth 8
"""
import ctypes
import json
import os
import struct

import numpy as np

DEFAULT_CONFIG = {
    "Width": "1440", "Height": "900", "FullScreen": "false", "BlockSize": "64", "MaxBlockSize": "608",
    "MaxDepth": "20", "BiasLength": "0.0002", "EnergyThreshold": "0.000001", "SSSThreshold": "0.000001",
    "Skybox": "true", "BilinearSample": "true", "Sky": "false", "GammaCorrection": "true", "AntiAlias": "true",
    "FOV": "45.0", "BvhLeafNodeTriangleNum": "1", "BvhBucketMaxDivideInternalNum": "12", "BvhBuildBlockSize": "32",
    "BvhBuildMethod": "MortonCodeCUDA", "AirRefractionIndex": "1.000293", "AirAbsorptionCoef": "0.0 0.0 0.0",
    "AirReducedScatteringCoef": "0.0 0.0 0.0", "CUDAAcceleration": "true",
}


def write_config(path, **overrides):
    cfg = dict(DEFAULT_CONFIG)
    for k, v in overrides.items():
        cfg[k] = ("true" if v else "false") if isinstance(v, bool) else str(v)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(cfg, f, indent=1)
    return path


def write_bmp24(path, rgb):
    """rgb (H, W, 3) uint8, row 0 = top -> bottom-up uncompressed 24-bit BMP."""
    h, w, _ = rgb.shape
    pitch = (w * 3 + 3) & ~3
    rows = np.zeros((h, pitch), dtype=np.uint8)
    rows[:, : w * 3] = rgb[::-1, :, ::-1].reshape(h, w * 3)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"BM" + struct.pack("<IHHI", 54 + pitch * h, 0, 0, 54))
        f.write(struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, pitch * h, 2835, 2835, 0, 0))
        f.write(rows.tobytes())


def _hash2(ix, iy, seed):
    m = np.uint64(0xFFFFFFFF)
    h = (ix.astype(np.uint64) * np.uint64(374761393) + iy.astype(np.uint64) * np.uint64(668265263) + np.uint64(seed * 2246822519 & 0xFFFFFFFF)) & m
    h = ((h ^ (h >> np.uint64(13))) * np.uint64(1274126177)) & m
    h = h ^ (h >> np.uint64(16))
    return (h & np.uint64(0xFFFFFF)).astype(np.float32) / np.float32(0xFFFFFF)


def value_noise(u, v, cells, seed):
    """Periodic (in u) bilinear value noise on a cells x cells lattice."""
    x, y = u * cells, v * cells
    ix, iy = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64)
    fx, fy = (x - ix).astype(np.float32), (y - iy).astype(np.float32)
    fx, fy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
    ix0, ix1 = ix % cells, (ix + 1) % cells
    a, b = _hash2(ix0, iy, seed), _hash2(ix1, iy, seed)
    c, d = _hash2(ix0, iy + 1, seed), _hash2(ix1, iy + 1, seed)
    return (a + (b - a) * fx) + ((c + (d - c) * fx) - (a + (b - a) * fx)) * fy


def synth_cubemap(dir_path, n=512, seed=0):
    """Six procedural faces (xpos..zneg): sky gradient, sun-ish blob, ground checker."""
    names = ["xpos", "xneg", "ypos", "yneg", "zpos", "zneg"]
    y, x = np.mgrid[0:n, 0:n].astype(np.float32)
    a, b = 2 * (x + 0.5) / n - 1, 1 - 2 * (y + 0.5) / n
    one = np.ones_like(a)
    dirs = [(one, b, -a), (-one, b, a), (a, one, -b), (a, -one, b), (a, b, one), (-a, b, -one)]
    for name, (dx, dy, dz) in zip(names, dirs):
        ln = np.sqrt(dx * dx + dy * dy + dz * dz)
        dx, dy, dz = dx / ln, dy / ln, dz / ln
        t = 0.5 * (dy + 1)
        sky = np.stack([0.35 + 0.35 * (1 - t), 0.5 + 0.3 * (1 - t), 0.75 + 0.2 * t], -1)
        sun = np.clip((dx * 0.35 + dy * 0.8 + dz * 0.48 - 0.95) * 20, 0, 1)[..., None]
        chk = ((np.floor(dx / np.maximum(-dy, 1e-3) * 2 + seed) + np.floor(dz / np.maximum(-dy, 1e-3) * 2)) % 2)[..., None]
        ground = np.where(chk > 0, 0.55, 0.3) * np.ones(3, np.float32)
        img = np.where((dy < 0)[..., None], ground, sky + sun * 0.6)
        write_bmp24(os.path.join(dir_path, name + ".bmp"), (np.clip(img, 0, 1) * 255 + 0.5).astype(np.uint8))


def synth_texture(path, n, seed):
    """checker / gradient / noise blend, deterministic per seed."""
    y, x = np.mgrid[0:n, 0:n].astype(np.float32)
    u, v = x / n, y / n
    cells = 4 << (seed % 3)
    chk = ((np.floor(u * cells) + np.floor(v * cells)) % 2)
    nz = value_noise(u, v, 16, seed) * 0.6 + value_noise(u, v, 64, seed + 100) * 0.4
    base = np.array([[0.9, 0.4, 0.3], [0.3, 0.7, 0.9], [0.8, 0.8, 0.3], [0.5, 0.9, 0.5], [0.9, 0.6, 0.9], [0.7, 0.7, 0.7], [0.95, 0.75, 0.5], [0.4, 0.5, 0.9]],
                    np.float32)[seed % 8]
    img = (0.35 + 0.45 * chk[..., None] + 0.2 * nz[..., None]) * base[None, None, :] * (0.7 + 0.3 * u[..., None])
    write_bmp24(path, (np.clip(img, 0, 1) * 255 + 0.5).astype(np.uint8))


# ---------------------------------------------------------------------------------------------
# meshes
# ---------------------------------------------------------------------------------------------

def displaced_sphere(n_lat, n_lon, seed=1, amplitude=0.18, octaves=3):
    """Closed lat-long sphere displaced radially by value noise. ~2*n_lat*n_lon triangles."""
    lat = np.linspace(0.0, np.pi, n_lat + 1, dtype=np.float64)[1:-1]
    lon = np.linspace(0.0, 2 * np.pi, n_lon, endpoint=False, dtype=np.float64)
    LA, LO = np.meshgrid(lat, lon, indexing="ij")
    u, v = (LO / (2 * np.pi)).astype(np.float32), (LA / np.pi).astype(np.float32)
    r = np.ones_like(u)
    for o in range(octaves):
        r = r + amplitude * (0.5 ** o) * (value_noise(u, v, 6 << o, seed + 17 * o) - 0.5) * 2 * np.sin(LA).astype(np.float32)
    x = r * np.sin(LA) * np.cos(LO)
    y = r * np.cos(LA)
    z = r * np.sin(LA) * np.sin(LO)
    ring = np.stack([x, y, z], -1).reshape(-1, 3)
    verts = np.concatenate([[[0, 1, 0]], ring, [[0, -1, 0]]], 0).astype(np.float32)
    uvs = np.concatenate([[[0.5, 1.0]], np.stack([u, 1 - v], -1).reshape(-1, 2), [[0.5, 0.0]]], 0).astype(np.float32)
    rows = n_lat - 1
    idx = (1 + np.arange(rows * n_lon).reshape(rows, n_lon))
    nxt = np.roll(idx, -1, axis=1)
    faces = []
    faces.append(np.stack([np.zeros(n_lon, np.int64), nxt[0], idx[0]], -1))
    a, b, c, d = idx[:-1], nxt[:-1], idx[1:], nxt[1:]
    faces.append(np.stack([a, b, c], -1).reshape(-1, 3))
    faces.append(np.stack([b, d, c], -1).reshape(-1, 3))
    south = verts.shape[0] - 1
    faces.append(np.stack([np.full(n_lon, south), idx[-1], nxt[-1]], -1))
    faces = np.concatenate(faces, 0).astype(np.int64)
    return verts, uvs, faces


def vertex_normals(verts, faces):
    v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    fn = np.cross(v1 - v0, v2 - v0).astype(np.float64)
    n = np.zeros(verts.shape, np.float64)
    for k in range(3):
        np.add.at(n, faces[:, k], fn)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    ln[ln == 0] = 1
    return (n / ln).astype(np.float32)


def box_mesh(half=(0.8, 0.8, 0.8)):
    """Axis-aligned box, 12 triangles, per-face normals, outward winding."""
    hx, hy, hz = half
    corners = np.array([[-hx, -hy, hz], [hx, -hy, hz], [-hx, hy, hz], [hx, hy, hz], [-hx, hy, -hz], [hx, hy, -hz], [-hx, -hy, -hz], [hx, -hy, -hz]], np.float32)
    quads = [([0, 1, 3, 2], [0, 0, 1]), ([2, 3, 5, 4], [0, 1, 0]), ([4, 5, 7, 6], [0, 0, -1]), ([6, 7, 1, 0], [0, -1, 0]), ([1, 7, 5, 3], [1, 0, 0]), ([6, 0, 2, 4], [-1, 0, 0])]
    return corners, quads


_FASTOBJ = [False]


def _fastobj():
    """tools/fastobj/libfastobj.so if it has been built (pathtracerwithcuda_b200.build.build_fastobj), else None (numpy writer)."""
    if _FASTOBJ[0] is False:
        lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fastobj", "libfastobj.so")
        _FASTOBJ[0] = None
        if os.path.exists(lib) and not os.environ.get("PTB_NO_FASTOBJ"):
            try:
                L = ctypes.CDLL(lib)
                dp = ctypes.POINTER(ctypes.c_double)
                L.fastobj_vertices.argtypes = [ctypes.c_char_p, dp, dp, dp, ctypes.c_int64, ctypes.c_int64]
                L.fastobj_group.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int64, ctypes.c_int]
                _FASTOBJ[0] = L
            except OSError:
                pass
    return _FASTOBJ[0]


def write_obj(path, groups, verts, normals=None, uvs=None):
    """groups: list of (name, faces[int, 3]) with 0-based indices shared by v / vn / vt."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if normals is None:
        allf = np.concatenate([f for _, f in groups], 0)
        normals = vertex_normals(verts, allf)
    fast = _fastobj()
    if fast is not None:
        # same bytes as the numpy writer below, written by C stdio (tools/fastobj/fastobj.c)
        with open(path, "w") as f:
            f.write("# generated by pathtracerwithcuda_b200.procedural\n")
        v = np.ascontiguousarray(verts, np.float64)
        nrm = np.ascontiguousarray(normals, np.float64)
        uv = np.ascontiguousarray(uvs, np.float64) if uvs is not None else None
        dp = ctypes.POINTER(ctypes.c_double)
        rc = fast.fastobj_vertices(os.fsencode(path), v.ctypes.data_as(dp), uv.ctypes.data_as(dp) if uv is not None else None, nrm.ctypes.data_as(dp),
                                   v.shape[0], uv.shape[0] if uv is not None else 0)
        for name, faces in groups:
            i = np.ascontiguousarray(faces + 1, np.int64)
            rc |= fast.fastobj_group(os.fsencode(path), name.encode(), i.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), i.shape[0], 1 if uvs is not None else 0)
        if rc:
            raise IOError("fastobj: cannot write " + path)
        return path
    with open(path, "w") as f:
        f.write("# generated by pathtracerwithcuda_b200.procedural\n")
        np.savetxt(f, verts, fmt="v %.6f %.6f %.6f")
        if uvs is not None:
            np.savetxt(f, uvs, fmt="vt %.6f %.6f")
        np.savetxt(f, normals, fmt="vn %.6f %.6f %.6f")
        for name, faces in groups:
            f.write("g %s\n" % name)
            i = faces + 1
            if uvs is not None:
                cols = np.stack([i[:, 0], i[:, 0], i[:, 0], i[:, 1], i[:, 1], i[:, 1], i[:, 2], i[:, 2], i[:, 2]], -1)
                np.savetxt(f, cols, fmt="f %d/%d/%d %d/%d/%d %d/%d/%d")
            else:
                cols = np.stack([i[:, 0], i[:, 0], i[:, 1], i[:, 1], i[:, 2], i[:, 2]], -1)
                np.savetxt(f, cols, fmt="f %d//%d %d//%d %d//%d")
    return path


def write_box_obj(path, half=(0.8, 0.8, 0.8), group_per_face=False):
    corners, quads = box_mesh(half)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write("# generated box\n")
        for c in corners:
            f.write("v %.6f %.6f %.6f\n" % tuple(c))
        for _, n in quads:
            f.write("vn %.6f %.6f %.6f\n" % tuple(n))
        if not group_per_face:
            f.write("g box\n")
        for qi, (q, _) in enumerate(quads):
            if group_per_face:
                f.write("g face%d\n" % qi)
            a, b, c, d = [k + 1 for k in q]
            n = qi + 1
            f.write("f %d//%d %d//%d %d//%d\n" % (a, n, b, n, c, n))
            f.write("f %d//%d %d//%d %d//%d\n" % (a, n, c, n, d, n))
    return path


def write_room_obj(path, half=(2.8, 2.0, 3.0)):
    """Open-front cornell-class room: back, floor, ceiling, left, right as 5 groups of 2 triangles,
    normals facing inward."""
    hx, hy, hz = half
    P = {"lbf": (-hx, -hy, hz), "rbf": (hx, -hy, hz), "ltf": (-hx, hy, hz), "rtf": (hx, hy, hz),
         "lbb": (-hx, -hy, -hz), "rbb": (hx, -hy, -hz), "ltb": (-hx, hy, -hz), "rtb": (hx, hy, -hz)}
    walls = [("back", ["lbb", "rbb", "rtb", "ltb"], (0, 0, 1)), ("left", ["lbf", "lbb", "ltb", "ltf"], (1, 0, 0)),
             ("right", ["rbb", "rbf", "rtf", "rtb"], (-1, 0, 0)), ("floor", ["lbf", "rbf", "rbb", "lbb"], (0, 1, 0)),
             ("ceiling", ["ltb", "rtb", "rtf", "ltf"], (0, -1, 0))]
    names = list(P.keys())
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write("# generated room\n")
        for k in names:
            f.write("v %.6f %.6f %.6f\n" % P[k])
        for _, _, n in walls:
            f.write("vn %.6f %.6f %.6f\n" % n)
        for wi, (g, q, _) in enumerate(walls):
            f.write("g %s\n" % g)
            a, b, c, d = [names.index(k) + 1 for k in q]
            n = wi + 1
            f.write("f %d//%d %d//%d %d//%d\n" % (a, n, b, n, c, n))
            f.write("f %d//%d %d//%d %d//%d\n" % (a, n, c, n, d, n))
    return path


def _grid_for_triangles(n_tris):
    n = int(round((n_tris / 2.0) ** 0.5))
    return max(n, 8), max(n, 8)


def write_blob_obj(path, n_tris, seed=1, with_uv=False, n_groups=1, amplitude=0.18):
    n_lat, n_lon = _grid_for_triangles(n_tris)
    verts, uvs, faces = displaced_sphere(n_lat, n_lon, seed=seed, amplitude=amplitude)
    normals = vertex_normals(verts, faces)
    if n_groups <= 1:
        groups = [("blob", faces)]
    else:
        # split by latitude band so each group is a contiguous strip (exercises per-group materials)
        bands = np.array_split(np.arange(faces.shape[0]), n_groups)
        groups = [("band%d" % i, faces[b]) for i, b in enumerate(bands)]
    write_obj(path, groups, verts, normals, uvs if with_uv else None)
    return int(faces.shape[0])


# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------

WORKLOADS = {
    "c1": dict(width=512, height=512, depth=5, spp=16),
    "c2": dict(width=1920, height=1080, depth=8, spp=256),
    "c3": dict(width=1920, height=1080, depth=8, spp=512),
    "c4": dict(width=1920, height=1080, depth=16, spp=1024),
    "c5": dict(width=3840, height=2160, depth=8, spp=4096),
    # small all-features scene for parity tests: textures (diffuse+specular, 4 groups), SSS medium,
    # glass + metal spheres, emissive box, thin-lens camera
    "mix": dict(width=96, height=72, depth=8, spp=4),
}


def _bs(p):
    return p.replace("/", "\\")


def make_workload(root, name, width=None, height=None, depth=None, tri_scale=1.0, cube_size=None):
    """Writes res/{obj,texture,scene,configuration} for workload `name` under `root`.
    Returns dict(scene=<scene json path>, config=<config json path>, triangles=..., camera=(aperture, focal), ...)."""
    w = dict(WORKLOADS[name])
    if width:
        w["width"] = width
    if height:
        w["height"] = height
    if depth:
        w["depth"] = depth
    res = os.path.join(root, "res")
    cube_n = cube_size or {"c1": 512, "mix": 64}.get(name, 1024)
    tex_dir = os.path.join(res, "texture", "ptbsky%d" % cube_n)
    if not os.path.exists(os.path.join(tex_dir, "zneg.bmp")):
        synth_cubemap(tex_dir, cube_n, seed=0)
    background = {"Name": "ptbsky%d" % cube_n, "Path": "res\\texture\\", "Format": "bmp"}
    light_obj = os.path.join(res, "obj", "ptb_light.obj")
    if not os.path.exists(light_obj):
        write_box_obj(light_obj)
    scene = {"Background": background}
    tris = 12
    aperture, focal = -1.0, -1.0

    if name == "c1":
        room = os.path.join(res, "obj", "ptb_room.obj")
        write_room_obj(room)
        tris += 10
        scene["Sphere"] = [{"Material": "iron", "Center": "-0.9 0.0 -0.9", "Radius": "0.8"},
                           {"Material": "glass", "Center": "1.3 0.0 -0.4", "Radius": "0.8"}]
        scene["Mesh"] = [
            {"Material": ["wall_white", "wall_green", "wall_red", "wall_white", "wall_white"], "Path": _bs("res/obj/ptb_room.obj"),
             "Position": "0.0 1.2 -2.0", "Scale": "1.0 1.0 1.0", "Rotate": "0.0 0.0 0.0"},
            {"Material": ["light"], "Path": _bs("res/obj/ptb_light.obj"), "Position": "0.0 3.1 0.0", "Scale": "0.85 0.05 0.85", "Rotate": "0.0 0.0 0.0"}]
    elif name == "c2":
        n = int(75000 * tri_scale)
        t1 = write_blob_obj(os.path.join(res, "obj", "ptb_blob_a.obj"), n, seed=1)
        t2 = write_blob_obj(os.path.join(res, "obj", "ptb_blob_b.obj"), n, seed=2)
        tris += t1 + t2
        scene["Mesh"] = [
            {"Material": ["light"], "Path": _bs("res/obj/ptb_light.obj"), "Position": "0.0 10.0 0.0", "Scale": "2.0 0.05 2.0", "Rotate": "0.0 0.0 0.0"},
            {"Material": ["silver"], "Path": _bs("res/obj/ptb_blob_a.obj"), "Position": "-2.6 0.0 0.0", "Scale": "2.4 2.4 2.4", "Rotate": "0.0 -45.0 0.0"},
            {"Material": ["red"], "Path": _bs("res/obj/ptb_blob_b.obj"), "Position": "2.6 0.0 0.0", "Scale": "2.4 2.4 2.4", "Rotate": "0.0 30.0 0.0"}]
    elif name == "c3":
        n = int(150000 * tri_scale)
        for i in range(8):
            p = os.path.join(res, "texture", "ptbtex", "tex%d.bmp" % i)
            if not os.path.exists(p):
                synth_texture(p, 1024, i + 1)
        scene["Texture"] = [_bs("res/texture/ptbtex/tex%d.bmp" % i) for i in range(8)]
        mats = []
        for i in range(12):
            mats.append({"Name": "ptbmat%d" % i, "Diffuse": "1.0 1.0 1.0", "Emission": "0.0 0.0 0.0",
                         "Specular": "1.0 1.0 1.0" if i % 3 == 0 else "0.0 0.0 0.0", "Transparent": "false", "Roughness": "0.3",
                         "RefractionIndex": "1.491", "ExtinctionCoef": "0.0", "AbsorptionCoef": "0.0 0.0 0.0",
                         "ReducedScatteringCoef": "0.0 0.0 0.0", "DiffuseTextureId": str(i % 8), "SpecularTextureId": str((i + 3) % 8) if i % 3 == 0 else "-1"})
        scene["Material"] = mats
        t1 = write_blob_obj(os.path.join(res, "obj", "ptb_tex_a.obj"), n, seed=3, with_uv=True, n_groups=12)
        t2 = write_blob_obj(os.path.join(res, "obj", "ptb_tex_b.obj"), n, seed=4, with_uv=True, n_groups=12)
        tris += t1 + t2
        names = ["ptbmat%d" % i for i in range(12)]
        scene["Mesh"] = [
            {"Material": ["light"], "Path": _bs("res/obj/ptb_light.obj"), "Position": "0.0 10.0 0.0", "Scale": "2.0 0.05 2.0", "Rotate": "0.0 0.0 0.0"},
            {"Material": names, "Path": _bs("res/obj/ptb_tex_a.obj"), "Position": "-2.4 0.0 -1.0", "Scale": "2.5 2.5 2.5", "Rotate": "10.0 20.0 0.0"},
            {"Material": names[::-1], "Path": _bs("res/obj/ptb_tex_b.obj"), "Position": "2.4 0.0 1.0", "Scale": "2.2 2.2 2.2", "Rotate": "0.0 -30.0 15.0"}]
        aperture, focal = 0.05, 14.0
    elif name == "c4":
        n = int(1000000 * tri_scale)
        scene["Material"] = [{"Name": "ptb_sss", "Diffuse": "0.0 0.0 0.0", "Emission": "0.0 0.0 0.0", "Specular": "1.0 1.0 1.0",
                              "Transparent": "true", "Roughness": "0.01", "RefractionIndex": "1.33", "ExtinctionCoef": "0.0",
                              "AbsorptionCoef": "20.0 3.0 13.0", "ReducedScatteringCoef": "10.0 10.0 10.0"}]
        t1 = write_blob_obj(os.path.join(res, "obj", "ptb_sss.obj"), n, seed=5)
        tris += t1
        scene["Mesh"] = [
            {"Material": ["light"], "Path": _bs("res/obj/ptb_light.obj"), "Position": "0.0 10.0 0.0", "Scale": "2.0 0.05 2.0", "Rotate": "0.0 0.0 0.0"},
            {"Material": ["ptb_sss"], "Path": _bs("res/obj/ptb_sss.obj"), "Position": "0.0 0.0 0.0", "Scale": "3.5 3.5 3.5", "Rotate": "0.0 0.0 0.0"}]
    elif name == "c5":
        n = int(2500000 * tri_scale)
        t1 = write_blob_obj(os.path.join(res, "obj", "ptb_big_a.obj"), n, seed=6, amplitude=0.25)
        t2 = write_blob_obj(os.path.join(res, "obj", "ptb_big_b.obj"), n, seed=7, amplitude=0.25)
        tris += t1 + t2
        scene["Mesh"] = [
            {"Material": ["light"], "Path": _bs("res/obj/ptb_light.obj"), "Position": "0.0 10.0 0.0", "Scale": "2.0 0.05 2.0", "Rotate": "0.0 0.0 0.0"},
            {"Material": ["gold"], "Path": _bs("res/obj/ptb_big_a.obj"), "Position": "-2.6 0.0 0.0", "Scale": "2.4 2.4 2.4", "Rotate": "0.0 -45.0 0.0"},
            {"Material": ["wall_blue"], "Path": _bs("res/obj/ptb_big_b.obj"), "Position": "2.6 0.0 0.0", "Scale": "2.4 2.4 2.4", "Rotate": "0.0 30.0 0.0"}]
    elif name == "mix":
        for i in range(2):
            p = os.path.join(res, "texture", "ptbtex", "small%d.bmp" % i)
            if not os.path.exists(p):
                synth_texture(p, 32, i + 1)
        scene["Texture"] = [_bs("res/texture/ptbtex/small%d.bmp" % i) for i in range(2)]
        base = {"Emission": "0.0 0.0 0.0", "ExtinctionCoef": "0.0", "AbsorptionCoef": "0.0 0.0 0.0", "ReducedScatteringCoef": "0.0 0.0 0.0"}
        scene["Material"] = [
            dict(base, Name="mix_tex", Diffuse="1.0 1.0 1.0", Specular="1.0 1.0 1.0", Transparent="false", Roughness="0.4", RefractionIndex="1.491", DiffuseTextureId="0", SpecularTextureId="1"),
            dict(base, Name="mix_matte", Diffuse="0.8 0.7 0.3", Specular="0.0 0.0 0.0", Transparent="false", Roughness="0.01", RefractionIndex="1.491", DiffuseTextureId="1"),
            dict(base, Name="mix_sss", Diffuse="0.0 0.0 0.0", Specular="1.0 1.0 1.0", Transparent="true", Roughness="0.01", RefractionIndex="1.33", AbsorptionCoef="2.0 0.3 1.3", ReducedScatteringCoef="4.0 4.0 4.0"),
            dict(base, Name="mix_absorb", Diffuse="1.0 1.0 1.0", Specular="0.045 0.045 0.045", Transparent="true", Roughness="0.1", RefractionIndex="1.5319", AbsorptionCoef="0.8 0.01 0.8")]
        t1 = write_blob_obj(os.path.join(res, "obj", "ptb_mix_a.obj"), int(1800 * tri_scale), seed=8, with_uv=True, n_groups=4)
        t2 = write_blob_obj(os.path.join(res, "obj", "ptb_mix_b.obj"), int(800 * tri_scale), seed=9)
        write_room_obj(os.path.join(res, "obj", "ptb_room.obj"))
        tris += t1 + t2 + 10
        scene["Sphere"] = [{"Material": "gold", "Center": "-3.2 -0.2 1.5", "Radius": "0.9"}, {"Material": "glass", "Center": "3.0 -0.1 1.8", "Radius": "1.0"},
                           {"Material": "mix_absorb", "Center": "0.0 -0.4 3.0", "Radius": "0.7"}]
        scene["Mesh"] = [
            {"Material": ["light"], "Path": _bs("res/obj/ptb_light.obj"), "Position": "0.0 4.5 0.0", "Scale": "1.5 0.05 1.5", "Rotate": "0.0 0.0 0.0"},
            {"Material": ["mix_tex", "mix_matte", "mix_tex", "red"], "Path": _bs("res/obj/ptb_mix_a.obj"), "Position": "-1.4 0.3 0.0", "Scale": "1.6 1.6 1.6", "Rotate": "15.0 -40.0 5.0"},
            {"Material": ["mix_sss"], "Path": _bs("res/obj/ptb_mix_b.obj"), "Position": "1.7 0.2 0.2", "Scale": "1.4 1.5 1.3", "Rotate": "0.0 25.0 0.0"},
            {"Material": ["wall_white", "wall_green", "wall_red", "wall_white"], "Path": _bs("res/obj/ptb_room.obj"), "Position": "0.0 1.0 -1.0", "Scale": "2.0 1.2 1.5", "Rotate": "0.0 0.0 0.0"}]
        aperture, focal = 0.08, 13.0
    else:
        raise ValueError(name)

    scene_path = os.path.join(res, "scene", "ptb_%s.json" % name)
    os.makedirs(os.path.dirname(scene_path), exist_ok=True)
    with open(scene_path, "w") as f:
        json.dump(scene, f, indent=1)
    config_path = write_config(os.path.join(res, "configuration", "ptb_%s.json" % name), Width=w["width"], Height=w["height"], MaxDepth=w["depth"])
    return {"name": name, "scene": scene_path, "scene_name": "ptb_%s" % name, "config": config_path,
            "config_rel": _bs("res/configuration/ptb_%s.json" % name), "triangles": tris, "width": w["width"], "height": w["height"],
            "depth": w["depth"], "spp": w["spp"], "aperture": aperture, "focal": focal}
