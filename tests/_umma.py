"""numpy builders for tcgen05 shared-memory operand images (SWIZZLE_NONE canonical layouts) used by the UMMA lab tests.
A 'chunk column' is the engine's activation layout: 8 channels (16 B) of all 128 rows, rows 16 B apart, chunk columns
CC = 2064 B apart (2048 + 16 B bank skew)."""
import ctypes

import numpy as np
import torch

from diffpose_nw_b200 import _lib

CC = 2064


def idesc(n, a_mn=False, b_mn=False, m=128):
    return (1 << 4) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def put_chunkcols(img, off, mat):
    """mat [rows<=128][C] (float) -> fp16 at off + (c//8)*CC + r*16 + (c%8)*2"""
    h = mat.astype(np.float16)
    rows, C = h.shape
    v = img.view(np.float16)
    for c in range(C):
        base = (off + (c // 8) * CC + (c % 8) * 2) // 2
        v[base + np.arange(rows) * 8] = h[:, c]


def put_kmajor(img, off, mat, lbo, sbo):
    """mat [rows][K] -> element (r,k) at off + (k//8)*lbo + (r//8)*sbo + (r%8)*16 + (k%8)*2"""
    h = mat.astype(np.float16)
    v = img.view(np.float16)
    for r in range(h.shape[0]):
        for k in range(h.shape[1]):
            v[(off + (k // 8) * lbo + (r // 8) * sbo + (r % 8) * 16 + (k % 8) * 2) // 2] = h[r, k]


def run_lab(img, ops, ncols):
    dev = torch.device("cuda:0")
    image = torch.from_numpy(img.copy()).to(dev)
    arr = (_lib.DpMmaOp * len(ops))(*[_lib.DpMmaOp(*o) for o in ops])
    out = torch.zeros(128, ncols, device=dev)
    _lib.check(_lib.load().dp_selftest_umma(image.data_ptr(), image.numel(), arr, len(ops), out.data_ptr(), ncols, None), "dp_selftest_umma")
    return out.cpu().numpy().astype(np.float64)


def run_lab_ts(img, timg, tcol0, ops, ncols):
    """timg: [128][ncols_t] float16 pairs packed as uint32 (lane = row) preloaded at TMEM column tcol0."""
    dev = torch.device("cuda:0")
    image = torch.from_numpy(img.copy()).to(dev)
    t = torch.from_numpy(np.ascontiguousarray(timg).view(np.int32).copy()).to(dev)
    arr = (_lib.DpMmaOp * len(ops))(*[_lib.DpMmaOp(*o) for o in ops])
    out = torch.zeros(128, ncols, device=dev)
    _lib.check(_lib.load().dp_selftest_umma_ts(image.data_ptr(), image.numel(), t.data_ptr(), tcol0, t.shape[1], arr, len(ops),
                                               out.data_ptr(), ncols, None), "dp_selftest_umma_ts")
    return out.cpu().numpy().astype(np.float64)


def f16(x):
    return x.astype(np.float16).astype(np.float64)
