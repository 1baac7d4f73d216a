"""Product scene front-end (csrc/scene_io.cpp, host code above the C ABI) against golden output of
the reference's own loaders (scene_*.npz / kat_host.json from oracle/_ref/libptref_host.so).
Runs through a host-only handle (device=-1): no compute calls, no GPU."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN
import pathtracerwithcuda_b200 as ptb
from pathtracerwithcuda_b200 import procedural as pr


def load_host(w, root, scene=None):
    r = ptb.Renderer(w["config"], device=-1)
    r.load_scene(scene or w["scene"], root)
    return r


def mats_u32(r):
    m = r.scene_materials()
    return m.view(np.uint32).reshape(-1, 21) if m.size else np.zeros((0, 21), np.uint32)


def assert_materials_equal(a, b):
    # byte 36 is the bool; bytes 37..39 are padding (uninitialised in the reference struct)
    assert a.shape == b.shape
    assert np.array_equal(a[:, :9], b[:, :9]) and np.array_equal(a[:, 10:], b[:, 10:])
    assert np.array_equal(a[:, 9] & 0xFF, b[:, 9] & 0xFF)


@pytest.mark.parametrize("name,kw", [("mix", dict(width=96, height=72)), ("c1", dict(width=64, height=64))])
def test_scene_bit_exact_vs_reference_loader(workload_root, name, kw):
    root, w = workload_root(name, **kw)
    g = np.load(os.path.join(GOLDEN, "scene_%s.npz" % name))
    r = load_host(w, root)
    tri, mat = r.scene_triangles()
    assert np.array_equal(tri.view(np.uint32), g["triangles"])      # world-space v / n / uv, bit for bit
    assert np.array_equal(mat, g["triangle_material"])
    assert_materials_equal(mats_u32(r), g["materials"])
    s = r.scene_spheres().view(np.uint32).reshape(-1, 25)
    assert np.array_equal(s[:, :13], g["spheres"][:, :13]) and np.array_equal(s[:, 14:], g["spheres"][:, 14:])
    cam = r.camera().as_array().view(np.uint32)
    keep = [i for i in range(16) if i != 9]                          # index 9 is struct padding
    assert np.array_equal(cam[keep], g["camera"][keep])


def test_awkward_obj_matches_reference(workload_root):
    sys.path.insert(0, GOLDEN)
    import objedge
    root, w = workload_root("mix", width=96, height=72)
    scene = objedge.write(root)
    g = np.load(os.path.join(GOLDEN, "scene_objedge.npz"))
    r = load_host(w, root, scene)
    tri, mat = r.scene_triangles()
    assert tri.shape == (24, 24)
    assert np.array_equal(tri.view(np.uint32), g["triangles"])
    assert np.array_equal(mat, g["triangle_material"])
    assert_materials_equal(mats_u32(r), g["materials"])
    s = r.scene_spheres()
    assert s["radius"][0] == 0.0                                      # Radius clamped to >= 0 (scene_parser.cpp:263)
    m = r.scene_materials()
    assert m["roughness"][0] == 1.0                                   # Roughness clamped to [0,1] (:187)
    assert np.all(m["diffuse_texture_id"][:3] == -1)


def _random_obj(seed, n_blocks):
    """Relative and absolute v/vt/vn indices, polygons up to hexagons, g / o / usemtl / blank / comment lines, mixed line ends."""
    import random
    rnd = random.Random(seed)
    lines, nv, nvt, nvn = ["# sliced"], 0, 0, 0
    for b in range(n_blocks):
        for i in range(rnd.randint(3, 12)):
            lines.append("v %.6f %.6g %e" % (rnd.uniform(-3, 3), rnd.uniform(-3, 3), rnd.uniform(-3, 3))); nv += 1
            if i == 0 or rnd.random() < 0.5:
                lines.append("vt %.4f %.4f" % (rnd.random(), rnd.random())); nvt += 1
            if i == 0 or rnd.random() < 0.5:
                lines.append("  vn %.4f %.4f %.4f" % (rnd.uniform(-1, 1), rnd.uniform(-1, 1), rnd.uniform(-1, 1))); nvn += 1
        r = rnd.random()
        if r < 0.2: lines.append("g grp%d" % b)
        elif r < 0.3: lines.append("o obj%d" % b)
        elif r < 0.35: lines.append("usemtl m%d" % b)
        elif r < 0.4: lines.append("")
        elif r < 0.45: lines.append("\t# comment f 1 2 3")
        for f in range(rnd.randint(1, 6)):
            toks = []
            for j in range(rnd.choice([3, 3, 3, 4, 4, 5, 6])):
                vi = -rnd.randint(1, min(nv, 30)) if rnd.random() < 0.5 else rnd.randint(max(1, nv - 40), nv)
                if rnd.random() < 0.5: toks.append("%d/%d/%d" % (vi, rnd.randint(max(1, nvt - 9), nvt), -rnd.randint(1, min(nvn, 9))))
                else: toks.append("%d/%d/%d" % (vi, -rnd.randint(1, min(nvt, 9)), rnd.randint(max(1, nvn - 9), nvn)))
            lines.append("f " + " ".join(toks))
    return "".join(l + rnd.choice(["\n", "\r\n", "\n", "\r"]) for l in lines)


def test_world_triangles_per_vertex_equal_per_corner(workload_root):
    """append_mesh transforms every position / normal of a large mesh once per VERTEX instead of once per triangle corner (scene_io.cpp,
    loader_per_vertex); the world-space triangles — and the object-space copies a later transform edit starts from — must be the same bits.
    Checked on the golden scenes' own files (mode 2 forces the per-vertex path for small meshes) and on a 75 k-triangle blob (mode 1)."""
    cases = [workload_root("mix", width=96, height=72), workload_root("c2", width=64, height=36, tri_scale=1.0)]
    r = ptb.Renderer(cases[0][1]["config"], device=-1)
    try:
        for root, w in cases:
            got = {}
            for mode in (0, 1, 2):
                r.set_option("loader_per_vertex", mode)
                r.load_scene(w["scene"], root)
                tri, mat = r.scene_triangles()
                r.set_mesh_transform(1, position=(0.3, -0.2, 0.1), scale=(1.1, 0.9, 1.3))      # re-transforms the object-space copies
                r.apply_mesh_rotate(1, (10.0, 20.0, 5.0))
                tri2, _ = r.scene_triangles()
                got[mode] = (tri.copy(), mat.copy(), tri2.copy())
            for mode in (1, 2):
                for a, b in zip(got[0], got[mode]):
                    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (w["scene"], mode)
            assert got[0][0].shape[0] == w["triangles"]
    finally:
        r.set_option("loader_per_vertex", 1)
        r.close()


def test_obj_parsed_in_slices_is_identical(workload_root):
    """Large OBJ files are cut into slices of whole lines parsed by host threads (scene_io.cpp parse_obj); relative indices,
    groups and the triangulation's vertex count must come out as in one sequential pass (tiny_obj_loader.h:1700-1830)."""
    sys.path.insert(0, GOLDEN)
    import objedge
    root, w = workload_root("mix", width=96, height=72)
    scene = objedge.write(root)
    g = np.load(os.path.join(GOLDEN, "scene_objedge.npz"))
    r = ptb.Renderer(w["config"], device=-1)
    try:
        for n in (1, 2, 3, 7, 64):                                    # 64 slices of a 44-line file: most hold one line or none
            r.set_option("loader_threads", n)
            r.load_scene(scene, root)
            tri, mat = r.scene_triangles()
            assert np.array_equal(tri.view(np.uint32), g["triangles"]) and np.array_equal(mat, g["triangle_material"]), n
        obj = os.path.join(root, "res", "obj", "objedge.obj")
        for seed in (1, 2):
            with open(obj, "w", newline="") as f:
                f.write(_random_obj(seed, 300))
            want = None
            for n in (1, 2, 5, 16, 61):
                r.set_option("loader_threads", n)
                r.load_scene(scene, root)
                tri, mat = r.scene_triangles()
                if want is None:
                    want = (tri.copy(), mat.copy())
                    assert tri.shape[0] > 2000
                assert np.array_equal(tri.view(np.uint32), want[0].view(np.uint32)) and np.array_equal(mat, want[1]), (seed, n)
        # the first error in file order wins, whichever slice a later one falls in
        with open(obj, "w", newline="") as f:
            f.write("v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvn 0 0 1\nf 1/1/1 2/1/1 3/1/1\nf 1/1/1 2/1/1\n" + "v 1 1 1\n" * 50 + "f 0/1/1 1/1/1 2/1/1\n")
        msgs = set()
        for n in (1, 4, 32):
            r.set_option("loader_threads", n)
            with pytest.raises(RuntimeError) as e:
                r.load_scene(scene, root)
            msgs.add(str(e.value))
        assert len(msgs) == 1 and "fewer than 3 vertices" in msgs.pop()
    finally:
        r.set_option("loader_threads", 0)
        r.close()


def test_builtin_materials_and_default_camera():
    with open(os.path.join(GOLDEN, "kat_host.json")) as f:
        kat = json.load(f)
    # every built-in is exercised through a scene that uses it on a sphere
    import tempfile
    root = tempfile.mkdtemp()
    w = pr.make_workload(root, "mix", width=32, height=24)
    names = sorted(kat["builtin_materials"].keys())
    assert len(names) == 27
    scene = {"Background": {"Name": "ptbsky64", "Path": "res\\texture\\", "Format": "bmp"},
             "Sphere": [{"Material": n, "Center": "0 0 0", "Radius": "1"} for n in names]}
    p = os.path.join(root, "res", "scene", "builtins.json")
    with open(p, "w") as f:
        json.dump(scene, f)
    r = load_host(w, root, p)
    s = r.scene_spheres().view(np.uint32).reshape(-1, 25)
    for i, n in enumerate(names):
        ref = np.array(kat["builtin_materials"][n], np.uint32)
        got = s[i, 4:]
        assert np.array_equal(got[:9], ref[:9]) and np.array_equal(got[10:], ref[10:]) and (got[9] & 0xFF) == (ref[9] & 0xFF), n
    for row in kat["default_camera"]:
        w_, h_, ap, fo = row["args"]
        cam = ptb.default_camera(w_, h_, ap, fo).as_array().view(np.uint32)
        keep = [i for i in range(16) if i != 9]
        assert cam[keep].tolist() == [row["cam"][i] for i in keep], row["args"]


def test_config_semantics(tmp_path):
    p = str(tmp_path / "c.json")
    pr.write_config(p, Width=320, Height=200, MaxDepth=7, FOV="true", BvhBuildMethod="mOrToNcOdEcPu", AirAbsorptionCoef="0.1 0.2 0.3", Sky=True)
    c = ptb.Renderer(p, device=-1).config()
    assert (c["width"], c["height"], c["max_tracer_depth"], c["block_size"]) == (320, 200, 7, 64)
    assert c["fov"] == 1.0                      # FOV goes through parse_bool (config_parser.cpp:111)
    assert c["bvh_build"] == 1 and c["use_sky"] == 1 and c["use_sky_box"] == 1
    assert np.allclose(c["air_absorption_coef"], [0.1, 0.2, 0.3])
    assert c["vector_bias_length"] == np.float32(0.0002)
    # all 23 keys are mandatory and every value must be a string
    cfg = dict(pr.DEFAULT_CONFIG)
    del cfg["SSSThreshold"]
    with open(p, "w") as f:
        json.dump(cfg, f)
    with pytest.raises(ptb.PtbError, match="SSSThreshold"):
        ptb.Renderer(p, device=-1)
    cfg = dict(pr.DEFAULT_CONFIG)
    cfg["Width"] = 640
    with open(p, "w") as f:
        json.dump(cfg, f)
    with pytest.raises(ptb.PtbError, match="Width"):
        ptb.Renderer(p, device=-1)
    with open(p, "w") as f:
        f.write("{ not json")
    with pytest.raises(ptb.PtbError):
        ptb.Renderer(p, device=-1)
    with pytest.raises(ptb.PtbError):
        ptb.Renderer(str(tmp_path / "missing.json"), device=-1)


def test_scene_errors(workload_root, tmp_path):
    root, w = workload_root("mix", width=96, height=72)
    r = ptb.Renderer(w["config"], device=-1)
    bg = {"Name": "ptbsky64", "Path": "res\\texture\\", "Format": "bmp"}

    def load(obj):
        p = str(tmp_path / "s.json")
        with open(p, "w") as f:
            json.dump(obj, f)
        r.load_scene(p, root)

    with pytest.raises(ptb.PtbError, match="Background not defined"):
        load({"Sphere": []})
    with pytest.raises(ptb.PtbError, match="not found"):
        load({"Background": bg, "Sphere": [{"Material": "unobtainium", "Center": "0 0 0", "Radius": "1"}]})
    with pytest.raises(ptb.PtbError, match="Background load fail"):
        load({"Background": dict(bg, Name="nope")})
    with pytest.raises(ptb.PtbError, match="Texture index out of range"):
        load({"Background": bg, "Material": [dict(Name="m", Diffuse="1 1 1", Emission="0 0 0", Specular="0 0 0", Transparent="false", Roughness="0.1",
                                                   RefractionIndex="1.5", ExtinctionCoef="0", AbsorptionCoef="0 0 0", ReducedScatteringCoef="0 0 0",
                                                   DiffuseTextureId="3")]})
    with pytest.raises(ptb.PtbError, match="Extinction coefficient of transparent"):
        load({"Background": bg, "Material": [dict(Name="m", Diffuse="1 1 1", Emission="0 0 0", Specular="0 0 0", Transparent="true", Roughness="0.1",
                                                   RefractionIndex="1.5", ExtinctionCoef="2", AbsorptionCoef="0 0 0", ReducedScatteringCoef="0 0 0")]})
    with pytest.raises(ptb.PtbError, match="must be array"):
        load({"Background": bg, "Mesh": [{"Material": "light", "Path": "res\\obj\\ptb_light.obj", "Position": "0 0 0", "Scale": "1 1 1", "Rotate": "0 0 0"}]})
    # a mesh without vn is rejected like the reference does (triangle_mesh.cpp:46-50)
    with open(os.path.join(root, "res", "obj", "novn.obj"), "w") as f:
        f.write("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    with pytest.raises(ptb.PtbError, match="does not have normal"):
        load({"Background": bg, "Mesh": [{"Material": ["light"], "Path": "res\\obj\\novn.obj", "Position": "0 0 0", "Scale": "1 1 1", "Rotate": "0 0 0"}]})
    # faces pointing outside the vertices read so far (corrupt file, forward reference): tinyobj would read wild memory while
    # triangulating; here it is a load error (found by fuzzing the loader under AddressSanitizer)
    for k, body in enumerate(("f 1//1 2//1 9999//1\n", "f 1//1 2//1 3//1 77//1\n", "f -1//1 -2//1 -50//1\n")):
        with open(os.path.join(root, "res", "obj", "wild%d.obj" % k), "w") as f:
            f.write("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\n" + body)
        with pytest.raises(ptb.PtbError):
            load({"Background": bg, "Mesh": [{"Material": ["light"], "Path": "res\\obj\\wild%d.obj" % k, "Position": "0 0 0", "Scale": "1 1 1", "Rotate": "0 0 0"}]})
    # several meshes: all files are parsed before any is appended (one allocation of the triangle arrays), but the error reported
    # is still the first in mesh order — a mesh-level error of mesh 1 wins over a parse error of mesh 2, and the other way round
    mesh = lambda path: {"Material": ["light"], "Path": path, "Position": "0 0 0", "Scale": "1 1 1", "Rotate": "0 0 0"}
    with open(os.path.join(root, "res", "obj", "badface.obj"), "w") as f:
        f.write("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1\n")
    with pytest.raises(ptb.PtbError, match="does not have normal"):
        load({"Background": bg, "Mesh": [mesh("res\\obj\\novn.obj"), mesh("res\\obj\\badface.obj")]})
    with pytest.raises(ptb.PtbError, match="fewer than 3 vertices"):
        load({"Background": bg, "Mesh": [mesh("res\\obj\\badface.obj"), mesh("res\\obj\\novn.obj")]})
    with pytest.raises(ptb.PtbError, match="fewer than 3 vertices"):
        load({"Background": bg, "Mesh": [mesh("res\\obj\\ptb_light.obj"), mesh("res\\obj\\badface.obj"), mesh("res\\obj\\missing.obj")]})
    with pytest.raises(ptb.PtbError, match="Cannot open file"):
        load({"Background": bg, "Mesh": [mesh("res\\obj\\ptb_light.obj"), mesh("res\\obj\\missing.obj"), mesh("res\\obj\\badface.obj")]})
    # non-finite or astronomically large vertices: the tree builders define no result for them; a load error, as is an edit
    # that would move a mesh there (the mesh keeps its previous placement)
    with open(os.path.join(root, "res", "obj", "inf.obj"), "w") as f:
        f.write("v 0 0 0\nv 1e39 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\n")
    with pytest.raises(ptb.PtbError, match="outside the supported range"):
        load({"Background": bg, "Mesh": [{"Material": ["light"], "Path": "res\\obj\\inf.obj", "Position": "0 0 0", "Scale": "1 1 1", "Rotate": "0 0 0"}]})
    r.load_scene(w["scene"], root)
    before = r.scene_triangles()[0].copy()
    with pytest.raises(ptb.PtbError, match="outside the supported range"):
        r.set_mesh_transform(0, (0.0, 0.0, 0.0), (1e30, 1e30, 1e30))
    with pytest.raises(ptb.PtbError):
        r.set_mesh_transform(0, (float("nan"), 0.0, 0.0), (1.0, 1.0, 1.0))
    assert np.array_equal(r.scene_triangles()[0].view(np.uint32), before.view(np.uint32))
    # empty scene (background only) is valid
    load({"Background": bg})
    assert r.scene_counts()["triangles"] == 0 and r.scene_counts()["cube_length"] == 64


def test_mesh_files_parsed_side_by_side(workload_root, tmp_path):
    """A host with more cores than one file's parser threads reads several mesh FILES at the same time (scene_io.cpp load_scene; option
    loader_mesh_lanes, 0 = by host cores).  Forced to 3 lanes here (and to a single slice per file, so that the lanes are the only
    concurrency): the golden scenes load bit-identically and a failing mesh is reported exactly as when the files are read in order."""
    root, w = workload_root("mix", width=96, height=72)
    r = ptb.Renderer(w["config"], device=-1)
    try:
        for lanes, threads in ((3, 0), (8, 1), (2, 4)):
            r.set_option("loader_mesh_lanes", lanes)
            r.set_option("loader_threads", threads)
            test_scene_bit_exact_vs_reference_loader(workload_root, "mix", dict(width=96, height=72))
            test_scene_bit_exact_vs_reference_loader(workload_root, "c1", dict(width=64, height=64))
            test_scene_errors(workload_root, tmp_path)
    finally:
        r.set_option("loader_mesh_lanes", 0)
        r.set_option("loader_threads", 0)
        r.close()


def test_images_bmp_and_tga(tmp_path):
    # BMP written bottom-up and a TGA (top-left origin, RLE) decode to the same top-down RGBA8
    rgb = (np.arange(5 * 7 * 3) * 7 % 256).astype(np.uint8).reshape(5, 7, 3)
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "res", "texture", "t"), exist_ok=True)
    for n in ["xpos", "xneg", "ypos", "yneg", "zpos", "zneg"]:
        pr.write_bmp24(os.path.join(root, "res", "texture", "t", n + ".bmp"), np.zeros((4, 4, 3), np.uint8) + 7)
    pr.write_bmp24(os.path.join(root, "res", "texture", "a.bmp"), rgb)
    # hand-written uncompressed TGA, bottom-left origin
    hdr = bytes([0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 7, 0, 5, 0, 24, 0])
    with open(os.path.join(root, "res", "texture", "b.tga"), "wb") as f:
        f.write(hdr + rgb[::-1, :, ::-1].tobytes())
    cfg = pr.write_config(os.path.join(root, "c.json"), Width=8, Height=8)
    scene = {"Background": {"Name": "t", "Path": "res\\texture\\", "Format": "bmp"}, "Texture": ["res\\texture\\a.bmp", "res/texture/b.tga"]}
    p = os.path.join(root, "s.json")
    with open(p, "w") as f:
        json.dump(scene, f)
    r = ptb.Renderer(cfg, device=-1)
    r.load_scene(p, root)
    for i in range(2):
        t = r.scene_texture(i)
        assert t.shape == (5, 7, 4)
        assert np.array_equal(t[..., :3], rgb) and np.all(t[..., 3] == 255)
    assert np.all(r.scene_cubemap()[..., :3] == 7)


def test_reference_scene_files_load_unchanged_from_its_checkout():
    """The reference's OWN scene + config files whose assets ship in its checkout (dinosaur, micro_surface),
    read in place: Windows-spelled paths, JPEG cube maps decoded natively (no side-cars exist there)."""
    ref = "/root/reference/gpu_path_tracer"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present (GPU box)")
    expect = {"dinosaur.json": dict(triangles=4012, spheres=0), "micro_surface.json": dict(triangles=12, spheres=6)}
    for name, want in expect.items():
        r = ptb.Renderer(os.path.join(ref, "res", "configuration", "config.json"), device=-1)
        r.load_scene(os.path.join(ref, "res", "scene", name), ref)
        c = r.scene_counts()
        assert c["triangles"] == want["triangles"] and c["spheres"] == want["spheres"] and c["cube_length"] == 2048, (name, c)
        cube = r.scene_cubemap()
        assert cube.shape == (6, 2048, 2048, 4) and cube[..., 3].min() == 255 and cube[..., :3].std() > 10
    # a scene whose assets are missing from the checkout fails with the reference's own message
    r = ptb.Renderer(os.path.join(ref, "res", "configuration", "config.json"), device=-1)
    with pytest.raises(ptb.PtbError) as e:
        r.load_scene(os.path.join(ref, "res", "scene", "cornell_box_simple.json"), ref)
    assert "Background load fail" in str(e.value)
