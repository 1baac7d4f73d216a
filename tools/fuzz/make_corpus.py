#!/usr/bin/env python3
"""Mutated inputs for the loader fuzz harnesses (tools/fuzz/run.sh): images of every natively decoded format, scene JSON, OBJ, config JSON.
    python tools/fuzz/make_corpus.py <out_dir> [seed] [count]"""
import io, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from pathtracerwithcuda_b200 import procedural as pr

out = sys.argv[1]
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
count = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
TOKENS = [b"-1", b"99999999999999999999", b"1e400", b"nan", b'""', b"[", b"]", b"{", b"}", b"\\", b'"', b"\nf 1 2 3 4 5 6 7\n", b"\nf -1 -2 -3\n",
          b"\nf 99999999 1 2\n", b"\nf -99999 1 2\n", b"\ng x\n", b"\no y\n", b"\nv 1e400 nan inf\n", b"\nf 2147483647 2147483648 -2147483648\n"]


def mutate(data, printable):
    a = bytearray(data)
    for _ in range(rng.integers(1, 6)):
        m, p = rng.integers(0, 4), rng.integers(0, len(a))
        if m == 0:
            a[p] = rng.integers(32, 127) if printable else rng.integers(0, 256)
        elif m == 1:
            del a[p:p + rng.integers(1, 24)]
        elif m == 2:
            a[p:p] = bytes(rng.integers(32 if printable else 0, 127 if printable else 256, rng.integers(1, 8)).astype(np.uint8))
        elif printable:
            a[p:p] = TOKENS[rng.integers(0, len(TOKENS))]
        if not a:
            a = bytearray(b"\0")
    return bytes(a)


def images():
    from PIL import Image
    noise = rng.integers(0, 256, (37, 53, 3)).astype(np.uint8)
    smooth = np.stack([np.add.outer(np.arange(40), np.arange(56)) * 3 % 256] * 3, -1).astype(np.uint8)
    seeds = []
    for im in (noise, smooth):
        for kw in (dict(format="JPEG", quality=80), dict(format="JPEG", quality=60, progressive=True, subsampling=2),
                   dict(format="JPEG", quality=90, subsampling=1, restart_marker_blocks=2), dict(format="JPEG", quality=70, progressive=True, subsampling=0),
                   dict(format="PNG"), dict(format="BMP"), dict(format="TGA")):
            b = io.BytesIO()
            Image.fromarray(im).save(b, **kw)
            seeds.append(("jpg" if kw["format"] == "JPEG" else kw["format"].lower(), b.getvalue()))
    d = os.path.join(out, "images")
    os.makedirs(d, exist_ok=True)
    for i in range(count):
        ext, data = seeds[i % len(seeds)]
        open(os.path.join(d, "f%05d.%s" % (i, ext)), "wb").write(mutate(data, False))


def scenes():
    root = os.path.join(out, "scene_root")
    w = pr.make_workload(root, "mix", width=32, height=24)
    scene = open(w["scene"], "rb").read()
    d = os.path.join(root, "fz")
    os.makedirs(d, exist_ok=True)
    for i in range(count // 3):
        open(os.path.join(d, "s%04d.json" % i), "wb").write(mutate(scene, True))
    sc = json.loads(scene)
    for i in range(count // 2):
        m = sc["Mesh"][i % len(sc["Mesh"])]
        rel = "res/obj/fz%04d.obj" % i
        open(os.path.join(root, rel), "wb").write(mutate(open(os.path.join(root, m["Path"].replace("\\", "/")), "rb").read(), True))
        json.dump(dict(sc, Mesh=[dict(m, Path=rel.replace("/", "\\"))]), open(os.path.join(d, "o%04d.json" % i), "w"))
    cfg = open(w["config"], "rb").read()
    c = os.path.join(out, "configs")
    os.makedirs(c, exist_ok=True)
    for i in range(count):
        open(os.path.join(c, "c%04d.json" % i), "wb").write(mutate(cfg, True))


if __name__ == "__main__":
    images()
    scenes()
