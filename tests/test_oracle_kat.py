"""Oracle (oracle/pt_oracle.c) against the golden vectors generated from the UNMODIFIED reference:
host-evaluated reference header functions + thrust RNG (kat_host.json) and the reference's device
hash() run on a B200 (ref_gpu_hash.npz).  CPU only."""
import ctypes
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as orc


def f32(bits):
    return np.array(bits, np.uint32).view(np.float32)


def p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLDEN, "kat_host.json")) as f:
        return json.load(f)


def close(a, b, rel=2e-6, abs_=1e-7):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= abs_ + rel * np.abs(b))


def test_rng_bit_exact(kat):
    # thrust::default_random_engine + uniform_real_distribution<float>
    L = orc.lib()
    for row in kat["rng"]:
        for key, lo, hi in (("m05_05", -0.5, 0.5), ("u01", 0.0, 1.0)):
            out = np.zeros(6, np.float32)
            L.ptbo_rng(ctypes.c_uint32(row["seed"]), lo, hi, 6, p(out))
            assert out.view(np.uint32).tolist() == row[key], (row["seed"], key)


def test_rng_appendix_b_vectors():
    # the first vectors transcribed in SURVEY.md Appendix B
    L = orc.lib()
    out = np.zeros(2, np.float32)
    L.ptbo_rng(ctypes.c_uint32(1765928508), -0.5, 0.5, 2, p(out))
    assert np.allclose(out, [-0.071998775, -0.45245406], rtol=0, atol=1e-9)
    out3 = np.zeros(3, np.float32)
    L.ptbo_rng(ctypes.c_uint32(517078614), 0.0, 1.0, 3, p(out3))
    assert np.allclose(out3, [0.85999769, 0.948140323, 0.682819426], rtol=0, atol=1e-8)


def test_hash_bit_exact_vs_device():
    d = np.load(os.path.join(GOLDEN, "ref_gpu_hash.npz"))
    L = orc.lib()
    got = np.array([L.ptbo_hash(int(x)) for x in d["inputs"]], np.int32)
    assert np.array_equal(got, d["outputs"])
    # seeds of Appendix B: hash(1)*hash(1)*hash(0) etc. are products of these
    h = lambda a: np.uint32(L.ptbo_hash(a) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        assert int(h(1) * h(1) * h(0)) == 1765928508
        assert int(h(1) * h(0) * h(0)) == 517078614


def test_triangle(kat):
    L = orc.lib()
    n_hit = 0
    for row in kat["triangle"]:
        v, ray = f32(row["v"]), f32(row["ray"])
        out = np.zeros(3, np.float32)
        hit = L.ptbo_triangle(p(v), p(ray), p(out))
        ref = f32(row["out"])
        if row["hit"]:
            n_hit += 1
            # the oracle uses the device's FMA placement; the golden values are un-fused host
            # evaluations of the same header, so agreement is to rounding, not bit-exact
            if hit:
                assert close(out, ref, rel=5e-5, abs_=5e-6), (out, ref)
        else:
            # a miss may flip only if the barycentrics sit on an edge to rounding
            if hit:
                assert min(out[1], out[2], 1 - out[1] - out[2]) < 1e-4
    assert n_hit > 10


def test_sphere(kat):
    L = orc.lib()
    hits = 0
    for row in kat["sphere"]:
        cr, ray = f32(row["cr"]), f32(row["ray"])
        out = np.zeros(7, np.float32)
        hit = L.ptbo_sphere(p(cr), p(ray), p(out))
        assert hit == row["hit"]
        if hit:
            hits += 1
            assert close(out, f32(row["out"]), rel=2e-5, abs_=2e-5)
    assert hits > 10


def test_box(kat):
    L = orc.lib()
    for row in kat["box"]:
        box, ray = f32(row["box"]), f32(row["ray"])
        t = np.array([np.inf], np.float32)
        hit = L.ptbo_box(p(box), p(ray), p(t))
        assert hit == row["hit"]
        assert t.view(np.uint32).tolist() == row["t"]  # same un-fused arithmetic: bit-exact


def test_fresnel(kat):
    L = orc.lib()
    for row in kat["fresnel"]:
        n, d, refr = f32(row["n"]), f32(row["d"]), f32(row["refr"])
        fd = L.ptbo_fresnel_dielectric(p(n), p(d), row["n_in"], row["n_out"], p(refr))
        fc = L.ptbo_fresnel_conductor(p(n), p(d), row["nk"][0], row["nk"][1])
        assert close([fd], f32(row["F_dielectric"]), rel=2e-5, abs_=1e-6)
        assert close([fc], f32(row["F_conductor"]), rel=2e-5, abs_=1e-6)


def test_cube_uv(kat):
    L = orc.lib()
    for row in kat["cube_uv"]:
        d = f32(row["d"])
        uv = np.zeros(2, np.float32)
        idx = L.ptbo_cube_uv(float(d[0]), float(d[1]), float(d[2]), p(uv))
        assert idx == row["index"]
        assert uv.view(np.uint32).tolist() == row["uv"]


def test_texture(kat):
    L = orc.lib()
    t = kat["texture"]
    tex = np.array(t["rgba"], np.uint8)
    for s in t["samples"]:
        uv = f32(s["uv"])
        out = np.zeros(3, np.float32)
        L.ptbo_texture(t["width"], t["height"], p(tex), float(uv[0]), float(uv[1]), s["bilinear"], p(out))
        assert close(out, f32(s["rgb"]), rel=1e-6, abs_=1e-6)


def test_background(kat):
    L = orc.lib()
    b = kat["background"]
    faces = np.array(b["faces"], np.uint8).reshape(6, 4, 4, 4)
    scene = orc._Scene()
    scene.cube_length = b["length"]
    for f in range(6):
        scene.cube_faces[f] = faces[f].ctypes.data
    for s in b["samples"]:
        cfg = np.zeros(96, np.uint8)
        cfg[36], cfg[37], cfg[38] = s["sky_box"], s["sky"], s["bilinear"]
        d = f32(s["d"])
        out = np.zeros(3, np.float32)
        L.ptbo_background(ctypes.byref(scene), p(cfg), p(d), p(out))
        assert close(out, f32(s["rgb"]), rel=1e-6, abs_=1e-6)


def test_byte_over_255_is_the_ieee_division():
    """csrc/pt_device.cuh: byte_over_255 replaces the texel normalisation v / 255.0f (Core/texture.h) by v * RN(1/255) plus one Newton
    correction in fused multiply-adds; it must be the correctly rounded quotient for all 256 bytes (libm's fmaf is correctly rounded)."""
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.fmaf.restype = ctypes.c_float
    libm.fmaf.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_float]
    r = np.float32(float.fromhex("0x1.010102p-8"))
    assert r == np.float32(1.0) / np.float32(255.0)
    for v in range(256):
        fv = np.float32(v)
        q = np.float32(fv * r)
        got = libm.fmaf(libm.fmaf(-255.0, float(q), float(fv)), float(r), float(q))
        assert np.float32(got) == fv / np.float32(255.0), v
