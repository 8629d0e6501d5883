#!/usr/bin/env python
"""tcgen05.mma issue/execute cost per instruction on this GPU, measured with the UMMA lab (GPU only).

Runs n back-to-back MMAs of a given flavour (same operands, accumulate) issued by one thread and reports
cycles(issue loop) and cycles(until the commit is observed) -> per-instruction slope between n = 16 and n = 64."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from diffpose_nw_b200 import _lib
from _umma import CC, idesc, run_lab, run_lab_ts


def cycles():
    out = (ctypes.c_longlong * 2)()
    _lib.check(_lib.load().dp_selftest_cycles(out), "dp_selftest_cycles")
    return out[0], out[1]


def main():
    img = np.zeros(160 * 1024, dtype=np.uint8)
    timg = np.zeros((128, 64), dtype=np.uint32)
    flavours = {
        "SS K-major N=32": lambda i: (0, CC, 128, 40960, 1536, 128, idesc(32), 0, 1),
        "SS K-major N=48": lambda i: (0, CC, 128, 40960, 1536, 128, idesc(48), 0, 1),
        "SS K-major N=96": lambda i: (0, CC, 128, 40960, 1536, 128, idesc(96), 0, 1),
        "SS K-major N=128": lambda i: (0, CC, 128, 40960, 1536, 128, idesc(128), 0, 1),
        "SS K-major N=256": lambda i: (0, CC, 128, 40960, 1536, 128, idesc(256), 0, 1),
        "SS B MN-major N=96": lambda i: (0, CC, 128, 40960, 128, CC, idesc(96, b_mn=True), 0, 1),
        "SS B MN-major N=48": lambda i: (0, CC, 128, 40960, 128, CC, idesc(48, b_mn=True), 0, 1),
        "TS B MN-major N=32": lambda i: (256, 0, 0, 40960, 128, CC, idesc(32, b_mn=True), 0, 3),
    }
    for name, f in flavours.items():
        res = {}
        for n in (1, 16, 64, 128):
            o = list(f(0))
            o[7] |= n << 16
            ops = [tuple(o)]
            best = None
            for _ in range(3):
                if "TS" in name:
                    run_lab_ts(img, timg, 256, ops, 16)
                else:
                    run_lab(img, ops, 16)
                c = cycles()
                best = c if best is None or c[1] < best[1] else best
            res[n] = best
        slope_issue = (res[128][0] - res[16][0]) / 112.0
        slope_done = (res[128][1] - res[16][1]) / 112.0
        print(f"{name:24s} n=1: issue {res[1][0]:5d} done {res[1][1]:5d} | n=16: {res[16][0]:5d} {res[16][1]:5d} | n=128: {res[128][0]:6d} {res[128][1]:6d} | per MMA: issue {slope_issue:6.1f} done {slope_done:6.1f}")


if __name__ == "__main__":
    main()
