"""Entry cuts of the camera rays (csrc/kernels_entry.cuh, k_extend_entry): starting every camera ray at the sub-trees its 8x4 pixel
tile's shaft touches must find exactly the hits a search from the root finds — the accumulated image and the per-depth segment counts
are compared BIT FOR BIT against the same render with entry_cuts=0, over cameras that stress the shaft construction (thin lens, eye
inside the geometry's bounds, grazing views, resolutions that are no multiple of the tile, anti-aliasing jitter on and off)."""
import numpy as np
import pytest

import pathtracerwithcuda_b200 as ptb

pytestmark = pytest.mark.gpu


def render(w, root, cam, passes, **options):
    r = ptb.Renderer(w["config"], device=0)
    r.set_option("entry_min_passes", 1)      # build the lists even for these 2-3 pass renders (default: from 4 passes per batch on, or for a camera seen before)
    for k, v in options.items():
        r.set_option(k, v)
    r.load_scene(w["scene"], root)
    r.set_camera(cam)
    r.render(passes)
    img = r.image_f32().copy()
    seg, _ = r.depth_profile()
    r.close()
    return img, list(seg)


def camera(w, eye=None, view=None, up=None, aperture=None, focal=None, fov_scale=1.0):
    cam = ptb.default_camera(w["width"], w["height"], w["aperture"] if aperture is None else aperture, w["focal"] if focal is None else focal)
    if eye is not None:
        for k in range(3):
            cam.eye[k] = eye[k]
    if view is not None:
        for k in range(3):
            cam.view[k] = view[k]
    if up is not None:
        for k in range(3):
            cam.up[k] = up[k]
    cam.fov[0] *= fov_scale
    cam.fov[1] *= fov_scale
    return cam


CAMERAS = [
    ("default", dict()),
    ("pinhole", dict(aperture=0.0)),
    ("wide_lens_near_focus", dict(aperture=0.6, focal=3.0)),
    ("lens_far_focus", dict(aperture=0.3, focal=27.0)),
    ("inside_looking_out", dict(eye=(0.3, 0.4, 0.2), view=(0.7, -0.1, -0.7))),
    ("grazing_from_below", dict(eye=(6.0, -1.5, 6.0), view=(-3.0, 0.45, -3.0), aperture=0.0)),
    ("long_view_vector_rolled", dict(eye=(-9.0, 5.0, 4.0), view=(18.0, -9.0, -8.5), up=(0.3, 1.0, 0.1))),
    ("narrow_fov", dict(fov_scale=0.2, aperture=0.0)),
    ("wide_fov", dict(fov_scale=2.6, aperture=0.05)),
]


@pytest.mark.parametrize("label,kw", CAMERAS)
def test_entry_cuts_find_the_same_hits_mix(workload_root, label, kw):
    root, w = workload_root("mix", width=100, height=70)      # neither a multiple of 8 nor of 4: partial tiles on both edges
    cam = camera(w, **kw)
    ref, seg0 = render(w, root, cam, 3, entry_cuts=0)
    for k, tile in ((1, "8x4"), (4, "8x4"), (15, "8x4"), (31, "8x4"), (8, "4x4"), (31, "2x2"), (6, "16x8"), (15, "64x1")):
        img, seg = render(w, root, cam, 3, entry_cuts=1, entry_k=k, entry_tile=tile)
        assert np.array_equal(ref.view(np.uint32), img.view(np.uint32)), (label, k, tile)
        assert seg == seg0, (label, k, tile)


@pytest.mark.parametrize("name,size,scale", [("c2", (320, 180), 0.1), ("c2", (300, 170), 0.1), ("c3", (256, 144), 0.05), ("c4", (192, 108), 0.03), ("c1", (96, 96), 1.0)])
def test_entry_cuts_find_the_same_hits_baseline_configs(workload_root, name, size, scale):
    root, w = workload_root(name, width=size[0], height=size[1], tri_scale=scale)
    cam = camera(w)
    ref, seg0 = render(w, root, cam, 2, entry_cuts=0)
    img, seg = render(w, root, cam, 2, entry_cuts=1)
    assert np.array_equal(ref.view(np.uint32), img.view(np.uint32))
    assert seg == seg0
    # sky_fast (default on with entry cuts): camera rays of tiles with an empty entry cut are finished by k_generate with the background
    # colour instead of being queued, searched and shaded — same image, same per-depth segment counts, fewer launches' worth of work
    img, seg = render(w, root, cam, 2, entry_cuts=1, sky_fast=0)
    assert np.array_equal(ref.view(np.uint32), img.view(np.uint32))
    assert seg == seg0
    # anti-aliasing off: every ray goes through its pixel centre (the shaft's slack still covers it)
    r0 = ptb.Renderer(w["config"], device=0); r1 = ptb.Renderer(w["config"], device=0)
    for r, on in ((r0, 0), (r1, 1)):
        r.set_option("entry_cuts", on)
        r.set_option("entry_min_passes", 1)
        r.load_scene(w["scene"], root)
        r.set_camera(cam)
        c = r.config().copy()
        c["use_anti_alias"] = 0
        r.set_config(c)
        r.render(2)
    assert np.array_equal(r0.image_f32().view(np.uint32), r1.image_f32().view(np.uint32))
    r0.close(); r1.close()


@pytest.mark.parametrize("name,size,scale", [("c2", (320, 180), 0.1), ("c3", (256, 144), 0.05), ("c1", (96, 96), 1.0), ("mix", (100, 70), 1.0), ("c4", (192, 108), 0.03)])
def test_leaf_starts_of_bounce_rays_find_the_same_hits(workload_root, name, size, scale):
    """upwalk (k_extend_upwalk): a bounce ray that leaves a triangle starts its search at that triangle's leaf and walks UP through the
    siblings of the leaf's ancestors; images and per-depth segment counts must equal a search from the root bit for bit — with the binary
    tree used for the first bounce only (the default hybrid) and for every bounce (hybrid_from_depth=99), and, in scattering media (mix,
    c4), with every search of a subsurface walk started at the leaf of the triangle the path entered through (fused_upwalk)."""
    kw = dict(width=size[0], height=size[1])
    if scale != 1.0:
        kw["tri_scale"] = scale
    root, w = workload_root(name, **kw)
    cam = camera(w)
    ref, seg0 = render(w, root, cam, 3, upwalk=0, entry_cuts=0)
    for opts in (dict(upwalk=1, entry_cuts=0), dict(upwalk=1), dict(upwalk=1, upwalk_min_nodes=1), dict(upwalk=1, upwalk_min_nodes=1, hybrid_from_depth=99), dict(upwalk=1, hybrid_from_depth=99), dict(upwalk=1, hybrid_from_depth=99, bvh_builder="host_sah"),
                 dict(upwalk=1, bvh_max_leaf=1), dict(upwalk=1, hybrid_from_depth=99, tune_refill_u=1, tune_leaf_u=1),
                 dict(upwalk=1, fused_upwalk=0), dict(upwalk=1, fused_upwalk=1, fused_from_depth=0), dict(upwalk=1, inline_scatter=0)):
        img, seg = render(w, root, cam, 3, **opts)
        assert np.array_equal(ref.view(np.uint32), img.view(np.uint32)), opts
        assert seg == seg0, opts


def test_entry_cuts_follow_camera_and_geometry_edits(workload_root):
    """The lists are rebuilt when the camera or the geometry changes between calls on ONE renderer."""
    root, w = workload_root("mix", width=96, height=72)
    rs = []
    for on in (0, 1):
        r = ptb.Renderer(w["config"], device=0)
        r.set_option("entry_cuts", on)
        r.set_option("entry_min_passes", 1)
        r.load_scene(w["scene"], root)
        rs.append(r)
    steps = [camera(w), camera(w, eye=(5.0, 2.0, -9.0), view=(-0.45, -0.15, 0.88)), camera(w, aperture=0.0, fov_scale=0.5)]
    for i, cam in enumerate(steps):
        imgs = []
        for r in rs:
            r.set_camera(cam)
            r.clear()
            r.render(2)
            imgs.append(r.image_f32().copy())
        assert np.array_equal(imgs[0].view(np.uint32), imgs[1].view(np.uint32)), i
        if i == 1:
            for r in rs:      # move a mesh: the tree is rebuilt, the old lists refer to nodes that no longer exist
                r.set_mesh_transform(1, position=(-0.6, 0.9, 0.8), scale=(1.2, 1.0, 1.4))
    for r in rs:
        r.close()


@pytest.mark.parametrize("name,kw", [("c2", dict(width=128, height=72, tri_scale=0.05)), ("c3", dict(width=96, height=64, tri_scale=0.02)), ("mix", dict(width=100, height=70))])
def test_entry_cuts_random_cameras(workload_root, name, kw):
    """40 seeded random cameras per scene — eye anywhere in and around the geometry, any view direction and roll, field of view 3..150 degrees,
    pinhole or a lens of up to 0.8 focused anywhere between 0.3 and 40 — rendered with and without entry cuts / sky fast path on ONE pair of
    renderers: the images must agree bit for bit for every camera (the lists are rebuilt at every camera change)."""
    root, w = workload_root(name, **kw)
    rs = []
    for on in (0, 1):
        r = ptb.Renderer(w["config"], device=0)
        r.set_option("entry_cuts", on)
        r.set_option("entry_min_passes", 1)
        r.set_option("passes_in_flight", 2)
        r.load_scene(w["scene"], root)
        rs.append(r)
    rng = np.random.default_rng(20261019)
    for i in range(40):
        eye = rng.uniform(-1.0, 1.0, 3) * rng.choice([0.5, 4.0, 12.0, 30.0])
        view = rng.normal(size=3)
        if i % 3 == 0:
            view = -eye + rng.normal(size=3) * 0.5          # towards the scene
        view = view / max(np.linalg.norm(view), 1e-6) * rng.choice([0.2, 1.0, 7.0])
        up = rng.normal(size=3) if i % 4 == 0 else np.array([0.0, 1.0, 0.0]) + rng.normal(size=3) * 0.05
        cam = camera(w, eye=eye, view=view, up=up, aperture=float(rng.choice([0.0, 0.0, 0.03, 0.8])), focal=float(rng.choice([0.3, 5.0, 14.0, 27.9])))
        scale = float(rng.choice([0.07, 0.5, 1.0, 2.2, 3.3]))
        cam.fov[0] = min(cam.fov[0] * scale, 150.0)
        cam.fov[1] = min(cam.fov[1] * scale, 150.0)
        imgs = []
        for r in rs:
            r.set_camera(cam)
            r.clear()
            r.render(2)
            imgs.append(r.image_f32().copy())
        same = np.array_equal(imgs[0].view(np.uint32), imgs[1].view(np.uint32))
        # NaN pixels (a degenerate random frame) compare by bits as well; report the camera on failure
        assert same, (i, eye.tolist(), view.tolist(), up.tolist(), cam.aperture_radius, cam.focal_distance, cam.fov[0])
    for r in rs:
        r.close()


def test_lists_are_not_rebuilt_for_a_camera_that_moves_every_pass(workload_root):
    """One pass per call with a new camera each time (a host dragging the view): the lists are not built (k_entry_cut would cost more than it
    saves), the search starts at the root, the image is the same; the second call with an unchanged camera builds them."""
    root, w = workload_root("c2", width=320, height=180, tri_scale=0.1)
    r = ptb.Renderer(w["config"], device=0)
    r.set_option("passes_in_flight", 1)
    r.load_scene(w["scene"], root)
    ref = ptb.Renderer(w["config"], device=0)
    ref.set_option("passes_in_flight", 1); ref.set_option("entry_cuts", 0)
    ref.load_scene(w["scene"], root)
    launches = []
    for i in range(6):
        cam = camera(w, eye=(14.0 * np.sin(0.1 * i) , 4.0, 14.0 * np.cos(0.1 * i)), view=(-np.sin(0.1 * i), -0.28, -np.cos(0.1 * i))) if i < 4 else cam
        for x in (r, ref):
            x.set_camera(cam); x.clear(); x.render(1)
        assert np.array_equal(r.image_f32().view(np.uint32), ref.image_f32().view(np.uint32)), i
        launches.append(r.stats()["kernel_launches"])
    # calls 0-3 (camera moved): no list build; call 4 is the second with the camera of call 3: + k_entry_cut + k_tile_rank; call 5: none again
    assert launches[1] == launches[0] and launches[4] == launches[0] + 2 and launches[5] == launches[0], launches


def test_unusable_cameras_fall_back_to_the_root(workload_root):
    """A camera the shaft construction does not cover (non-positive focal distance: the generator flips the rays) is traced from the root."""
    root, w = workload_root("mix", width=64, height=48)
    cam = camera(w, aperture=0.0)
    cam.focal_distance = -2.0
    a, _ = render(w, root, cam, 2, entry_cuts=0)
    b, _ = render(w, root, cam, 2, entry_cuts=1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # a camera so far from the origin that the generator's binary32 sums lose direction bits beyond the shafts' slack: also from the root
    far = camera(w, eye=(800.0, 300.0, 600.0), view=(-0.7619, -0.2857, -0.5714), aperture=0.0, fov_scale=0.02)
    a, _ = render(w, root, far, 2, entry_cuts=0)
    b, _ = render(w, root, far, 2, entry_cuts=1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
