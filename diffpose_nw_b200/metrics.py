"""Evaluation metrics that follow the sampler, computed on the GPU by `dp_metrics`.

Replaces `mpjpe` (reference common/loss.py:7-13) and `p_mpjpe` (common/loss.py:25-64, per-pose form
common/utils.py:155-187) as the runner uses them (runners/diffpose_frame.py:382-387), without the
`.cpu().numpy()` SVD round trip.  Per-rank partial sums `[sum mpjpe, sum p_mpjpe, count]` (fp64) are what the
multi-GPU driver all-reduces.
"""
from __future__ import annotations

import torch

from . import _lib


def pose_error_sums(pred, target, sums=None, per_pose=False):
    """pred [n,17,3] (xyz) or [n,17,5] (uvxyz: xyz = columns 2:5); target [n,17,3]; metres.

    Both are root-centred out of place (intended semantics of diffpose_frame.py:384-385).  Returns
    (sums, per_pose) where sums is a CUDA fp64 tensor [sum_pose mpjpe, sum_pose p_mpjpe, n] (accumulated into
    the tensor passed as `sums` if any) and per_pose is [n,2] fp32 or None.
    """
    if not (pred.is_cuda and target.is_cuda):
        raise RuntimeError("diffpose_nw_b200.metrics runs on CUDA only (no CPU fallback)")
    if pred.dim() != 3 or pred.shape[1] != 17 or pred.shape[2] not in (3, 5):
        raise RuntimeError(f"pred must be [n,17,3] or [n,17,5], got {tuple(pred.shape)}")
    n = pred.shape[0]
    if tuple(target.shape) != (n, 17, 3):
        raise RuntimeError(f"target must be [{n},17,3], got {tuple(target.shape)}")
    dev = pred.device
    p = pred.detach().to(torch.float32).contiguous()
    g = target.detach().to(device=dev, dtype=torch.float32).contiguous()
    stride = p.shape[2]
    if sums is None:
        sums = torch.zeros(3, device=dev, dtype=torch.float64)
    pp = torch.empty(n, 2, device=dev, dtype=torch.float32) if per_pose else None
    if n:
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(_lib.load().dp_metrics(p.data_ptr(), stride, stride - 3, g.data_ptr(), n, 17, sums.data_ptr(),
                                              pp.data_ptr() if pp is not None else None, stream), "dp_metrics")
    return sums, pp


def mpjpe(pred, target):
    """Mean per-joint position error of root-centred poses (scalar CUDA tensor, same unit as the inputs)."""
    s, _ = pose_error_sums(pred, target)
    return (s[0] / s[2]).to(torch.float32)


def p_mpjpe(pred, target):
    """Procrustes-aligned MPJPE, mean over poses (scalar CUDA tensor)."""
    s, _ = pose_error_sums(pred, target)
    return (s[1] / s[2]).to(torch.float32)
