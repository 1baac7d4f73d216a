"""Golden vectors for the JPEG "fast" decode mode (csrc/jpeg_decode.cpp), i.e. the libjpeg parameters FreeImage uses for
the reference's loads (Others/image_loader.cpp:45, flags 0 = JPEG_FAST): dct_method = JDCT_IFAST, do_fancy_upsampling = FALSE.

Generated with the libjpeg-turbo inside Pillow's wheel, driven through oracle/jpeg_lib_shim.c (PIL cannot select those
parameters).  Run in the build container:  make -C oracle && python tests/golden/make_jpeg_golden.py
Writes tests/golden/jpeg_fast.npz: for each case the JPEG file bytes and the RGB8 image libjpeg-turbo decodes from it."""
import ctypes
import glob
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))


def shim():
    import PIL
    lib = ctypes.CDLL(os.path.join(REPO, "oracle", "_build", "libjpegshim.so"))
    lib.jpegshim_decode.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]
    found = glob.glob(os.path.join(os.path.dirname(PIL.__file__), "..", "pillow.libs", "libjpeg*.so.62*"))
    if not found:
        raise RuntimeError("Pillow's bundled libjpeg not found")
    so = os.path.realpath(found[0]).encode()

    def decode(data, dct_method, fancy):
        w, h = ctypes.c_int(), ctypes.c_int()
        rc = lib.jpegshim_decode(so, 62, data, len(data), dct_method, fancy, ctypes.byref(w), ctypes.byref(h), None)
        if rc:
            raise RuntimeError("jpegshim_decode rc=%d" % rc)
        out = np.zeros((h.value, w.value, 3), np.uint8)
        rc = lib.jpegshim_decode(so, 62, data, len(data), dct_method, fancy, ctypes.byref(w), ctypes.byref(h), out.ctypes.data)
        if rc:
            raise RuntimeError("jpegshim_decode rc=%d" % rc)
        return out
    return decode


def test_image(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 1.3 + yy * 0.4) % 256, (yy * 2.1 + 30 * np.sin(xx / 9.0)) % 256, (xx * yy / 50.0) % 256], -1).astype(np.uint8)
    img[h // 4:h // 2, w // 4:w // 2] = rng.integers(0, 256, (h // 2 - h // 4, w // 2 - w // 4, 3))
    return img


def main():
    from PIL import Image
    decode = shim()
    out = {}
    k = 0
    for (h, w) in ((48, 64), (37, 29), (8, 8), (1, 1)):
        img = test_image(h, w, h + w)
        for sub in (0, 1, 2):
            for q in (35, 90):
                b = io.BytesIO()
                Image.fromarray(img).save(b, "JPEG", subsampling=sub, quality=q)
                data = b.getvalue()
                # the shim itself is checked first: library defaults through it == PIL
                assert np.array_equal(decode(data, 0, 1), np.asarray(Image.open(io.BytesIO(data)).convert("RGB")))
                out["jpeg_%02d" % k] = np.frombuffer(data, np.uint8)
                out["rgb_%02d" % k] = decode(data, 1, 0)
                k += 1
    b = io.BytesIO()
    Image.fromarray(test_image(40, 56, 9)[..., 1]).save(b, "JPEG", quality=80)
    out["jpeg_%02d" % k] = np.frombuffer(b.getvalue(), np.uint8)
    out["rgb_%02d" % k] = decode(b.getvalue(), 1, 0)
    k += 1
    out["count"] = np.int32(k)
    np.savez_compressed(os.path.join(HERE, "jpeg_fast.npz"), **out)
    print("wrote", k, "cases")


if __name__ == "__main__":
    sys.exit(main())
