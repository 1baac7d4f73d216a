"""Output writers of the C ABI on host buffers (csrc/image_out.cpp; replaces Main/window.cpp:712-740's lodepng
screenshot): the PNG must decode (zlib + CRC verified here by hand, and by PIL when present) to the RGBA8 image the
reference would save; the PFM carries the scaled floats bottom-up."""
import os
import struct
import zlib

import numpy as np
import pytest

import pathtracerwithcuda_b200 as ptb


def decode_png(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + body) & 0xFFFFFFFF == crc, typ
        chunks.append((typ, body))
        pos += 12 + n
    assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    w, h, depth, ctype, comp, filt, inter = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, ctype, comp, filt, inter) == (8, 6, 0, 0, 0)
    raw = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(h, w * 4 + 1)
    assert np.all(raw[:, 0] == 0)
    return raw[:, 1:].reshape(h, w, 4)


@pytest.mark.parametrize("shape", [(1, 1), (7, 5), (270, 480), (300, 333)])
def test_png_roundtrip(tmp_path, shape):
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
    p = str(tmp_path / "a.png")
    ptb.write_png(p, img)
    got = decode_png(p)
    assert np.array_equal(got[..., :3], img) and np.all(got[..., 3] == 255)
    try:
        from PIL import Image
    except ImportError:
        return
    assert np.array_equal(np.asarray(Image.open(p).convert("RGB")), img)


def test_pfm(tmp_path):
    rng = np.random.default_rng(2)
    img = rng.uniform(0, 4, size=(9, 13, 3)).astype(np.float32)
    p = str(tmp_path / "a.pfm")
    ptb.write_pfm(p, img, 0.25)
    data = open(p, "rb").read()
    head, rest = data.split(b"\n", 3)[:3], data.split(b"\n", 3)[3]
    assert head == [b"PF", b"13 9", b"-1.0"]
    got = np.frombuffer(rest, "<f4").reshape(9, 13, 3)[::-1]
    assert np.array_equal(got, img * np.float32(0.25))


def test_writer_errors(tmp_path):
    with pytest.raises(ptb.PtbError):
        ptb.write_png(str(tmp_path / "no_such_dir" / "a.png"), np.zeros((2, 2, 3), np.uint8))


# ---- native PNG decode of the scene front-end (Others/image_loader.cpp:31-95: FreeImage_Load + ConvertTo24Bits) ----
def _png_bytes(w, h, depth, ctype, rows, palette=None, level=6, interlace=0):
    def chunk(t, b):
        return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b) & 0xFFFFFFFF)
    raw = b"".join(bytes([f]) + bytes(r) for f, r in rows)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
    if palette is not None:
        out += chunk(b"PLTE", bytes(palette))
    co = zlib.compress(raw, level)
    half = len(co) // 2
    return out + chunk(b"IDAT", co[:half]) + chunk(b"IDAT", co[half:]) + chunk(b"IEND", b"")


def test_png_decode_matches_pil(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(4)
    smooth = (np.add.outer(np.arange(67), np.arange(45)) * 3 % 256).astype(np.uint8)
    cases = {"rgb": rng.integers(0, 256, (37, 53, 3), dtype=np.uint8), "rgba": rng.integers(0, 256, (20, 31, 4), dtype=np.uint8),
             "grey": smooth, "smooth_rgb": np.stack([smooth, smooth.T[:45, :45].repeat(2, 0)[:67, :45], 255 - smooth], -1),
             "big": rng.integers(0, 4, (300, 400, 3), dtype=np.uint8) * 60}
    for name, a in cases.items():
        for level in (0, 1, 9):
            p = str(tmp_path / ("%s_%d.png" % (name, level)))
            Image.fromarray(a).save(p, compress_level=level)
            want = np.asarray(Image.open(p).convert("RGB"))
            got = ptb.decode_image(p)
            assert got.shape == want.shape[:2] + (4,), name
            assert np.array_equal(got[..., :3], want), (name, level)
            assert np.all(got[..., 3] == 255)
    # palette image
    pal = Image.fromarray(cases["rgb"]).quantize(17)
    p = str(tmp_path / "pal.png")
    pal.save(p)
    assert np.array_equal(ptb.decode_image(p)[..., :3], np.asarray(Image.open(p).convert("RGB")))
    # our own writer's files decode too
    p = str(tmp_path / "own.png")
    ptb.write_png(p, cases["rgb"])
    assert np.array_equal(ptb.decode_image(p)[..., :3], cases["rgb"])


def test_png_decode_filters_depths_and_errors(tmp_path):
    rng = np.random.default_rng(5)
    w, h = 19, 11
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    # every filter type, applied by hand
    rows, prev = [], np.zeros(w * 3, np.int32)
    for y in range(h):
        cur = img[y].reshape(-1).astype(np.int32)
        f = y % 5
        left = np.concatenate([np.zeros(3, np.int32), cur[:-3]])
        upleft = np.concatenate([np.zeros(3, np.int32), prev[:-3]])
        if f == 0:
            enc = cur
        elif f == 1:
            enc = cur - left
        elif f == 2:
            enc = cur - prev
        elif f == 3:
            enc = cur - ((left + prev) >> 1)
        else:
            p = left + prev - upleft
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
            enc = cur - pred
        rows.append((f, (enc & 255).astype(np.uint8)))
        prev = cur
    p = str(tmp_path / "filters.png")
    open(p, "wb").write(_png_bytes(w, h, 8, 2, rows))
    assert np.array_equal(ptb.decode_image(p)[..., :3], img)
    # 16-bit grey -> high byte; 4-bit grey -> scaled; grey + alpha -> alpha dropped
    g16 = rng.integers(0, 65536, (h, w), dtype=np.uint16)
    open(p, "wb").write(_png_bytes(w, h, 16, 0, [(0, g16[y].astype(">u2").tobytes()) for y in range(h)]))
    assert np.array_equal(ptb.decode_image(p)[..., 0], (g16 >> 8).astype(np.uint8))
    g4 = rng.integers(0, 16, (h, 20), dtype=np.uint8)
    open(p, "wb").write(_png_bytes(20, h, 4, 0, [(0, ((g4[y, 0::2] << 4) | g4[y, 1::2]).astype(np.uint8)) for y in range(h)]))
    assert np.array_equal(ptb.decode_image(p)[..., 1], g4 * 17)
    ga = rng.integers(0, 256, (h, w, 2), dtype=np.uint8)
    open(p, "wb").write(_png_bytes(w, h, 8, 4, [(0, ga[y].reshape(-1)) for y in range(h)]))
    assert np.array_equal(ptb.decode_image(p)[..., 2], ga[..., 0])
    # corrupt stream / truncated file: loud failure, no crash
    good = _png_bytes(w, h, 8, 2, rows)
    for bad in (good[:60], good[:40] + bytes([good[40] ^ 0xFF]) + good[41:], b"\x89PNG\r\n\x1a\n" + b"\0" * 40):
        open(p, "wb").write(bad)
        with pytest.raises(ptb.PtbError):
            ptb.decode_image(p)


def test_png_adam7_interlaced(tmp_path):
    """Interlaced files (FreeImage reads them; PIL reads but cannot write them): the seven reduced passes built by hand, with the
    sub filter in odd passes, for sizes that leave some passes empty; checked against the source pixels and against PIL."""
    rng = np.random.default_rng(6)
    passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]
    for (h, w, ch) in ((11, 19, 3), (1, 1, 3), (2, 3, 3), (8, 8, 1), (33, 5, 3), (16, 40, 1)):
        img = rng.integers(0, 256, (h, w, ch), dtype=np.uint8)
        rows = []
        for k, (x0, y0, dx, dy) in enumerate(passes):
            sub = img[y0::dy, x0::dx]
            if sub.shape[0] == 0 or sub.shape[1] == 0:
                continue
            for r in sub:
                cur = r.reshape(-1).astype(np.int32)
                if k % 2:
                    left = np.concatenate([np.zeros(ch, np.int32), cur[:-ch]])
                    rows.append((1, ((cur - left) & 255).astype(np.uint8)))
                else:
                    rows.append((0, cur.astype(np.uint8)))
        p = str(tmp_path / ("adam7_%d_%d_%d.png" % (h, w, ch)))
        open(p, "wb").write(_png_bytes(w, h, 8, 2 if ch == 3 else 0, rows, interlace=1))
        got = ptb.decode_image(p)
        want = img if ch == 3 else np.repeat(img, 3, axis=2)
        assert got.shape == (h, w, 4) and np.array_equal(got[..., :3], want), (h, w, ch)
        try:
            from PIL import Image
            assert np.array_equal(np.asarray(Image.open(p).convert("RGB")), want)
        except ImportError:
            pass
    # 4-bit interlaced grey: sub-byte packing restarts in every pass row
    g4 = rng.integers(0, 16, (9, 13), dtype=np.uint8)
    rows = []
    for (x0, y0, dx, dy) in passes:
        sub = g4[y0::dy, x0::dx]
        if sub.size == 0:
            continue
        for r in sub:
            r = np.concatenate([r, np.zeros(len(r) % 2, np.uint8)])
            rows.append((0, ((r[0::2] << 4) | r[1::2]).astype(np.uint8)))
    p = str(tmp_path / "adam7_g4.png")
    open(p, "wb").write(_png_bytes(13, 9, 4, 0, rows, interlace=1))
    assert np.array_equal(ptb.decode_image(p)[..., 0], g4 * 17)
    open(p, "wb").write(_png_bytes(13, 9, 4, 0, rows[:-2], interlace=1))     # truncated pass data fails loudly
    with pytest.raises(ptb.PtbError):
        ptb.decode_image(p)


# ---- native baseline JPEG decode (csrc/jpeg_decode.cpp) ----
# "accurate" mode == libjpeg-turbo's default decode (PIL) bit for bit; "fast" mode == libjpeg-turbo with the parameters FreeImage
# sets for the reference's loads (ifast IDCT, replicated chroma); "reference" (default) = "fast" with libjpeg 9a's colour constant.
@pytest.fixture
def jpeg_mode():
    yield ptb.set_jpeg_decode
    ptb.set_jpeg_decode("reference")


def _test_image(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 1.3 + yy * 0.4) % 256, (yy * 2.1 + 30 * np.sin(xx / 9.0)) % 256, (xx * yy / 50.0) % 256], -1).astype(np.uint8)
    img[h // 4:h // 2, w // 4:w // 2] = rng.integers(0, 256, (h // 2 - h // 4, w // 2 - w // 4, 3))
    return img


@pytest.mark.parametrize("shape", [(203, 157), (16, 16), (8, 8), (1, 1), (3, 2), (17, 33), (64, 250)])
def test_jpeg_decode_matches_pil(tmp_path, shape, jpeg_mode):
    Image = pytest.importorskip("PIL.Image")
    jpeg_mode("accurate")
    img = _test_image(shape[0], shape[1], shape[0] + shape[1])
    for sub in (0, 1, 2):                       # 4:4:4, 4:2:2, 4:2:0
        for kw in (dict(quality=30), dict(quality=75), dict(quality=95, optimize=True), dict(quality=60, restart_marker_blocks=3)):
            p = str(tmp_path / ("j_%d_%d.jpg" % (sub, kw["quality"])))
            try:
                Image.fromarray(img).save(p, subsampling=sub, **kw)
            except TypeError:
                continue                        # older Pillow without restart_marker_blocks
            want = np.asarray(Image.open(p).convert("RGB"))
            got = ptb.decode_image(p)
            assert got.shape == want.shape[:2] + (4,)
            assert np.array_equal(got[..., :3], want), (shape, sub, kw)
            assert np.all(got[..., 3] == 255)
    p = str(tmp_path / "grey.jpg")
    Image.fromarray(img[..., 1]).save(p, quality=85)
    assert np.array_equal(ptb.decode_image(p)[..., :3], np.asarray(Image.open(p).convert("RGB")))


@pytest.mark.parametrize("shape", [(203, 157), (16, 16), (1, 1), (17, 33), (64, 250)])
def test_jpeg_progressive_matches_libjpeg(tmp_path, shape, jpeg_mode):
    """Progressive files (spectral selection + successive approximation, ITU T.81 Annex G): PIL writes libjpeg's standard
    10-scan script; decoded here they must equal PIL (accurate mode) and libjpeg-turbo's fast decode (fast mode)."""
    Image = pytest.importorskip("PIL.Image")
    decode = _libjpeg_shim()
    img = _test_image(shape[0], shape[1], 7 * shape[0] + shape[1])
    for sub in (0, 1, 2):
        for kw in (dict(quality=30), dict(quality=85, optimize=True), dict(quality=100), dict(quality=60, restart_marker_blocks=2)):
            p = str(tmp_path / ("p_%d_%d.jpg" % (sub, kw["quality"])))
            try:
                Image.fromarray(img).save(p, subsampling=sub, progressive=True, **kw)
            except TypeError:
                continue
            assert Image.open(p).info.get("progressive")
            jpeg_mode("accurate")
            assert np.array_equal(ptb.decode_image(p)[..., :3], np.asarray(Image.open(p).convert("RGB"))), (shape, sub, kw)
            if decode is not None:
                jpeg_mode("fast")
                assert np.array_equal(ptb.decode_image(p)[..., :3], decode(open(p, "rb").read(), 1, 0)), (shape, sub, kw)
    p = str(tmp_path / "pgrey.jpg")
    Image.fromarray(img[..., 0]).save(p, quality=70, progressive=True)
    jpeg_mode("accurate")
    assert np.array_equal(ptb.decode_image(p)[..., :3], np.asarray(Image.open(p).convert("RGB")))


def test_jpeg_sidecar_and_corrupt_files(tmp_path, jpeg_mode):
    Image = pytest.importorskip("PIL.Image")
    img = _test_image(40, 56, 3)
    p = str(tmp_path / "prog.jpg")
    Image.fromarray(img).save(p, quality=80, progressive=True)
    data = open(p, "rb").read()
    # a progressive file cut before its last scans is incomplete (libjpeg would smooth it): rejected, not half-decoded
    last_sos = data.rfind(b"\xff\xda")
    open(p, "wb").write(data[:last_sos] + b"\xff\xd9")
    with pytest.raises(ptb.PtbError):
        ptb.decode_image(p)
    with open(p + ".rgba8", "wb") as f:         # ... the documented side-car is picked up instead
        rgba = np.concatenate([img, np.full(img.shape[:2] + (1,), 255, np.uint8)], 2)
        f.write(struct.pack("<II", img.shape[1], img.shape[0]) + rgba.tobytes())
    assert np.array_equal(ptb.decode_image(p)[..., :3], img)
    q = str(tmp_path / "base.jpg")
    Image.fromarray(img).save(q, quality=80)
    data = open(q, "rb").read()
    for bad in (data[:200], data[:2] + b"\xff\xc0\x00\x05abc", b"\xff\xd8\xff\xd9"):
        open(q, "wb").write(bad)
        with pytest.raises(ptb.PtbError):
            ptb.decode_image(q)


def test_oversized_headers_fail_before_allocating(tmp_path):
    """A header announcing more pixels than any decoder should allocate (corrupt or hostile file) is an error, not a crash."""
    Image = pytest.importorskip("PIL.Image")
    p = str(tmp_path / "huge.png")
    open(p, "wb").write(_png_bytes(60000, 60000, 8, 2, [(0, bytes(30))]))
    with pytest.raises(ptb.PtbError):
        ptb.decode_image(p)
    q = str(tmp_path / "huge.jpg")
    Image.fromarray(_test_image(16, 16, 1)).save(q, quality=80)
    data = bytearray(open(q, "rb").read())
    sof = data.find(b"\xff\xc0")
    data[sof + 5:sof + 9] = b"\xff\xff\xff\xff"           # 65535 x 65535
    open(q, "wb").write(bytes(data))
    with pytest.raises(ptb.PtbError):
        ptb.decode_image(q)
    t = str(tmp_path / "huge.tga")
    open(t, "wb").write(bytes([0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0xff, 0xff, 0xff, 0xff, 24, 0]) + bytes(64))
    with pytest.raises(ptb.PtbError):
        ptb.decode_image(t)


def test_jpeg_fast_mode_matches_golden(tmp_path, jpeg_mode):
    """Committed vectors (tests/golden/make_jpeg_golden.py): libjpeg-turbo run with dct_method=JDCT_IFAST and
    do_fancy_upsampling=FALSE — what FreeImage_Load(..., 0) selects for the reference (Others/image_loader.cpp:45)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg_fast.npz"))
    n = int(g["count"])
    assert n >= 20
    worst_ref = 0.0
    for k in range(n):
        p = str(tmp_path / ("g%02d.jpg" % k))
        open(p, "wb").write(g["jpeg_%02d" % k].tobytes())
        want = g["rgb_%02d" % k]
        jpeg_mode("fast")
        got = ptb.decode_image(p)
        assert np.array_equal(got[..., :3], want), k
        assert np.all(got[..., 3] == 255)
        # the default mode differs only through one colour-table entry (libjpeg 9a's 0.344136286 vs 6b's 0.34414)
        jpeg_mode("reference")
        ref = ptb.decode_image(p)[..., :3].astype(int)
        d = np.abs(ref - want.astype(int))
        assert d.max() <= 1 and np.all(d[..., 0] == 0) and np.all(d[..., 2] == 0)
        worst_ref = max(worst_ref, float((d > 0).mean()))
    assert worst_ref < 0.01
    with pytest.raises(ptb.PtbError):
        ptb.set_jpeg_decode("best")


def _libjpeg_shim():
    """libjpeg-turbo from Pillow's wheel driven through oracle/jpeg_lib_shim.c; None where either is missing."""
    import ctypes
    import glob
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(repo, "oracle", "_build", "libjpegshim.so")
    try:
        import PIL
        found = glob.glob(os.path.join(os.path.dirname(PIL.__file__), "..", "pillow.libs", "libjpeg*.so.62*"))
        lib = ctypes.CDLL(so)
    except (ImportError, OSError):
        return None
    if not found:
        return None
    lib.jpegshim_decode.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]
    path = os.path.realpath(found[0]).encode()

    def decode(data, dct_method, fancy):
        w, h = ctypes.c_int(), ctypes.c_int()
        assert lib.jpegshim_decode(path, 62, data, len(data), dct_method, fancy, ctypes.byref(w), ctypes.byref(h), None) == 0
        out = np.zeros((h.value, w.value, 3), np.uint8)
        assert lib.jpegshim_decode(path, 62, data, len(data), dct_method, fancy, ctypes.byref(w), ctypes.byref(h), out.ctypes.data) == 0
        return out
    return decode


@pytest.mark.parametrize("shape", [(203, 157), (16, 16), (3, 2), (17, 33), (64, 250)])
def test_jpeg_fast_mode_matches_libjpeg_live(tmp_path, shape, jpeg_mode):
    Image = pytest.importorskip("PIL.Image")
    decode = _libjpeg_shim()
    if decode is None:
        pytest.skip("libjpeg shim or Pillow's libjpeg not available")
    img = _test_image(shape[0], shape[1], shape[0] * 3 + shape[1])
    jpeg_mode("fast")
    for sub in (0, 1, 2):
        for kw in (dict(quality=20), dict(quality=75), dict(quality=100), dict(quality=60, restart_marker_blocks=3)):
            p = str(tmp_path / ("f_%d_%d.jpg" % (sub, kw["quality"])))
            try:
                Image.fromarray(img).save(p, subsampling=sub, **kw)
            except TypeError:
                continue
            data = open(p, "rb").read()
            assert np.array_equal(decode(data, 0, 1), np.asarray(Image.open(p).convert("RGB")))      # the shim, with defaults, is PIL
            assert np.array_equal(ptb.decode_image(p)[..., :3], decode(data, 1, 0)), (shape, sub, kw)


def test_reference_cube_maps_decode_like_libjpeg(jpeg_mode):
    """The reference's own JPEG assets (two 2048^2 4:2:0 cube maps) — only where its checkout exists (this container)."""
    import glob
    Image = pytest.importorskip("PIL.Image")
    files = sorted(glob.glob("/root/reference/gpu_path_tracer/res/texture/*/*.jpg"))
    if not files:
        pytest.skip("reference checkout not present")
    decode = _libjpeg_shim()
    for f in files[:4]:
        jpeg_mode("accurate")
        assert np.array_equal(ptb.decode_image(f)[..., :3], np.asarray(Image.open(f).convert("RGB"))), f
        if decode is None:
            continue
        jpeg_mode("fast")
        fast = ptb.decode_image(f)[..., :3]
        assert np.array_equal(fast, decode(open(f, "rb").read(), 1, 0)), f
        jpeg_mode("reference")
        d = np.abs(ptb.decode_image(f)[..., :3].astype(int) - fast.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3
