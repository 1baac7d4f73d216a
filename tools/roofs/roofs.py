#!/usr/bin/env python3
"""Measures the roofs of the closest-hit kernel on the GPU it runs on (tools/roofs/roofs.cu through ctypes).

    python tools/roofs/roofs.py [device]          # prints one JSON object

bench.py imports measure() and runs it BEFORE its timed region; `roofline.frac` is reported against these
numbers.  Not product code: libptb200.so neither links nor loads libptbroofs.so.
"""
import ctypes
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libptbroofs.so")


def _load():
    if not os.path.exists(LIB):
        sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
        from pathtracerwithcuda_b200 import build
        build.build_roofs()
    L = ctypes.CDLL(LIB)
    L.ptbroofs_fp32.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    L.ptbroofs_gather.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    L.ptbroofs_l1_stream.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    return L


def measure(device=0, tree_bytes=10 << 20, blocks_per_sm=8, full=False):
    """tree_bytes: working set of the L2 gather (the workload's nodes + leaf-order triangles).
    Returns {"fp32_tflops", "l2_gather_gbs" (64 B records, the binary node), "l1_gather_gbs", "l1_writeback_gbs", ...}."""
    L = _load()
    out = (ctypes.c_double * 4)()
    res = {"blocks_per_sm": blocks_per_sm, "tree_bytes": int(tree_bytes)}
    if L.ptbroofs_fp32(device, out):
        raise RuntimeError("ptbroofs_fp32 failed")
    res["fp32_tflops"] = out[0]
    res["fp32_sm_count"] = int(out[2])
    # implied clock if the FMA pipes retire 128 lanes per SM per cycle
    res["fp32_implied_mhz"] = out[0] * 1e12 / (2.0 * 128 * out[2]) / 1e6

    def gather(ws, rec):
        if L.ptbroofs_gather(device, int(ws), rec, blocks_per_sm, out):
            raise RuntimeError("ptbroofs_gather(%d, %d) failed" % (ws, rec))
        return out[0]

    res["l2_gather_gbs"] = gather(tree_bytes, 64)       # 64-byte records = one binary node per lane
    res["l1_gather_gbs"] = gather(64 << 10, 64)
    # L1 -> register write-back bandwidth (what load instructions cost the L1 data pipe): the roof of a kernel whose tree sits in cache
    if L.ptbroofs_l1_stream(device, blocks_per_sm, out):
        raise RuntimeError("ptbroofs_l1_stream failed")
    res["l1_writeback_gbs"] = out[0]
    res["l1_writeback_bytes_per_clk_per_sm"] = out[0] * 1e9 / (res["fp32_sm_count"] * res["fp32_implied_mhz"] * 1e6)
    if full:
        for rec in (32, 64, 128):
            res["l2_gather_%dB_gbs" % rec] = gather(tree_bytes, rec)
            res["l1_gather_%dB_gbs" % rec] = gather(64 << 10, rec)
        res["l2_gather_64B_96MB_gbs"] = gather(96 << 20, 64)
        res["dram_gather_64B_4GB_gbs"] = gather(4 << 30, 64)
        for b in (4, 12, 16):
            if L.ptbroofs_gather(device, int(tree_bytes), 64, b, out) == 0:
                res["l2_gather_64B_%dblocks_gbs" % b] = out[0]
    return res


if __name__ == "__main__":
    dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    print(json.dumps(measure(dev, full=True)))
