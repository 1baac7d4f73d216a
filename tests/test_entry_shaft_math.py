"""The geometric claim behind the entry cuts of a camera WITH a lens (csrc/kernels_entry.cuh: entry_box_overlaps): a ray from lens point
eye + delta (|delta| <= aperture, delta in the lens plane) through focal point F is P(u) = [eye + u (F - eye)] + (1 - u) delta, i.e. the
pinhole ray's point at depth fraction u displaced by at most |1 - u| * aperture — so P lies at most |1 - u| * aperture outside any side plane
of the tile's pinhole pyramid, with u = dot(P - eye, w) / dot(F - eye, w).  Checked here in float64 on random cameras, tiles, lens points
and ray parameters, following the generator's own construction (pt_device.cuh: generate_camera_ray; path_tracer_kernel.cu:299-379).
CPU only: this pins the derivation; the kernels themselves are compared bit for bit on the GPU (tests/test_gpu_entry.py)."""
import numpy as np


def frame(eye, view, up, fov_x, fov_y):
    distance = np.linalg.norm(view)
    horizontal = np.cross(view, up); horizontal /= np.linalg.norm(horizontal)
    vertical = np.cross(horizontal, view); vertical /= np.linalg.norm(vertical)
    x_axis = horizontal * distance * np.tan(np.radians(fov_x) * 0.5)
    y_axis = vertical * distance * np.tan(-np.radians(fov_y) * 0.5)
    return horizontal, vertical, x_axis, y_axis


def test_lens_rays_stay_within_the_widened_pyramid():
    rng = np.random.default_rng(7)
    worst = 0.0
    for _ in range(300):
        W, H = int(rng.integers(16, 400)), int(rng.integers(16, 300))
        eye = rng.normal(size=3) * 10
        view = rng.normal(size=3); view /= np.linalg.norm(view); view *= rng.choice([0.3, 1.0, 5.0])
        up = np.array([0.0, 1.0, 0.0]) + rng.normal(size=3) * 0.2
        fov_x = rng.uniform(5, 120); fov_y = rng.uniform(5, 120)
        aperture = rng.choice([0.0, 0.05, 0.7]); focal = rng.uniform(0.5, 30.0)
        horizontal, vertical, x_axis, y_axis = frame(eye, view, up, fov_x, fov_y)
        w = view / np.linalg.norm(view)
        tw, th = 8, 4
        tx, ty = int(rng.integers(0, (W + tw - 1) // tw)), int(rng.integers(0, (H + th - 1) // th))
        px0, px1 = tx * tw - 0.5, min(tx * tw + tw - 1, W - 1) + 0.5
        py0, py1 = ty * th - 0.5, min(ty * th + th - 1, H - 1) + 0.5
        nx = lambda p: p / (W - 1) * 2 - 1
        ny = lambda p: p / (H - 1) * 2 - 1
        corners = [view + nx(a) * x_axis + ny(b) * y_axis for a, b in ((px0, py0), (px1, py0), (px1, py1), (px0, py1))]
        centre = view + nx(0.5 * (px0 + px1)) * x_axis + ny(0.5 * (py0 + py1)) * y_axis
        normals = []
        for k in range(4):
            n = np.cross(corners[k], corners[(k + 1) % 4]); n /= np.linalg.norm(n)
            normals.append(n if n @ centre > 0 else -n)
        for _ in range(200):
            sx, sy = rng.uniform(px0, px1), rng.uniform(py0, py1)
            canvas = eye + view + nx(sx) * x_axis + ny(sy) * y_axis
            direction0 = (canvas - eye) / np.linalg.norm(canvas - eye)
            F = eye + direction0 * focal
            ang, rad = rng.uniform(0, 2 * np.pi), aperture * np.sqrt(rng.uniform())
            o = eye + np.cos(ang) * rad * horizontal + np.sin(ang) * rad * vertical
            d = (F - o) / np.linalg.norm(F - o)
            t = rng.uniform(0, 4 * focal)
            P = o + t * d
            u = ((P - eye) @ w) / ((F - eye) @ w)
            assert u >= -1e-9
            for n in normals:
                outside = -(n @ (P - eye)) - aperture * abs(1 - u)
                worst = max(worst, outside)
                assert outside <= 1e-9 * (1 + np.linalg.norm(P - eye)), (aperture, focal, u, outside)
            # the distance bound: |P - o| >= |P - eye| - aperture
            assert t >= np.linalg.norm(P - eye) - aperture - 1e-9
    assert worst <= 1e-6
