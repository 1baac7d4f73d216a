"""Output writers of the C ABI on host buffers (csrc/image_out.cpp; replaces Main/window.cpp:712-740's lodepng
screenshot): the PNG must decode (zlib + CRC verified here by hand, and by PIL when present) to the RGBA8 image the
reference would save; the PFM carries the scaled floats bottom-up."""
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
