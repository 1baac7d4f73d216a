"""Live scene edits: ptb_set_sphere / ptb_set_mesh_material / ptb_set_mesh_transform /
ptb_apply_mesh_rotate against the reference's own setters (Core/scene_parser.cpp:645-673,
Core/triangle_mesh.cpp:252-426) driven by the same edit script (tests/golden/make_golden.py EDIT_SCRIPT);
the golden world-space state after every step comes from oracle/_ref/libptref_host.so (scene_mix_edits.npz).
Host-only handle: no GPU needed."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN
import pathtracerwithcuda_b200 as ptb

sys.path.insert(0, GOLDEN)
import make_golden as mg  # noqa: E402


def _mats(r):
    return r.scene_materials().view(np.uint32).reshape(-1, 21)


def test_edits_bit_exact_vs_reference_setters(workload_root):
    root, w = workload_root("mix", width=96, height=72)
    g = np.load(os.path.join(GOLDEN, "scene_mix_edits.npz"))
    r = ptb.Renderer(w["config"], device=-1)
    r.load_scene(w["scene"], root)
    before = r.scene_triangles()[0].copy()
    for k, (op, args) in enumerate(mg.EDIT_SCRIPT):
        mg.apply_edit(r, op, args)
        tri, _ = r.scene_triangles()
        assert np.array_equal(tri.view(np.uint32), g["step%d_triangles" % k]), (k, op)
        a, b = _mats(r), g["step%d_materials" % k]
        assert np.array_equal(a[:, :9], b[:, :9]) and np.array_equal(a[:, 10:], b[:, 10:]) and np.array_equal(a[:, 9] & 0xFF, b[:, 9] & 0xFF), (k, op)
        s, gs = r.scene_spheres().view(np.uint32).reshape(-1, 25), g["step%d_spheres" % k]
        assert np.array_equal(s[:, :13], gs[:, :13]) and np.array_equal(s[:, 14:], gs[:, 14:]) and np.array_equal(s[:, 13] & 0xFF, gs[:, 13] & 0xFF), (k, op)
    assert not np.array_equal(before, r.scene_triangles()[0])
    # placement bookkeeping mirrors m_mesh_position / m_mesh_scale / m_mesh_rotate_applied
    pl = r.mesh_placement(1)
    assert np.allclose(pl["position"], [-1.6, 0.1, -0.4]) and np.allclose(pl["scale"], [1, 1, 1]) and np.allclose(pl["rotate"], [40, -25, 10])
    assert r.mesh_placement(0)["scale"][1] == np.float32(0.000001)        # clamped like the UI does
    assert r.pass_counter() == 0


def test_edit_errors_and_ignored_calls(workload_root):
    root, w = workload_root("mix", width=96, height=72)
    r = ptb.Renderer(w["config"], device=-1)
    with pytest.raises(ptb.PtbError):
        r.set_mesh_transform(0, [0, 0, 0], [1, 1, 1])                     # no scene loaded
    r.load_scene(w["scene"], root)
    with pytest.raises(ptb.PtbError):
        r.set_mesh_transform(99, [0, 0, 0], [1, 1, 1])
    with pytest.raises(ptb.PtbError):
        r.set_sphere(17, r.scene_spheres()[0:1])
    # a material list of the wrong length is ignored, like triangle_mesh::set_material_device (triangle_mesh.cpp:254-257)
    mats = r.scene_materials().copy()
    r.set_mesh_material(1, mats[:1])
    assert np.array_equal(_mats(r), mats.view(np.uint32).reshape(-1, 21))
