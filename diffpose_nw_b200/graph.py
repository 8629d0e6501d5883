"""Skeleton graph of the frame-based model (host side, evaluated once).

Mirrors `adj_mx_from_edges(num_pts, edges, sparse=False)` (reference models/ChebConv.py:36-48) and the
16-bone / 17-joint Human3.6M edge list of runners/diffpose_frame.py:120-124.
"""
from __future__ import annotations

import numpy as np
import torch

H36M_EDGES = ((0, 1), (1, 2), (2, 3), (0, 4), (4, 5), (5, 6), (0, 7), (7, 8), (8, 9), (9, 10),
              (8, 11), (11, 12), (12, 13), (8, 14), (14, 15), (15, 16))


def adj_mx_from_edges(num_pts=17, edges=H36M_EDGES, sparse=False):
    """Symmetric adjacency + self loops, row-normalised, dense fp32 [num_pts, num_pts]."""
    if sparse:
        raise NotImplementedError("the frame-based path uses the dense adjacency (sparse=False)")
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    a = np.zeros((num_pts, num_pts), dtype=np.float64)   # the reference normalises in float64 (scipy), then casts
    a[e[:, 0], e[:, 1]] = 1.0
    a = np.maximum(a, a.T)                      # symmetrise a 0/1 matrix
    a += np.eye(num_pts)
    deg = a.sum(axis=1)
    inv = np.zeros_like(deg)
    np.divide(1.0, deg, out=inv, where=deg != 0)
    return torch.from_numpy((inv[:, None] * a).astype(np.float32))
