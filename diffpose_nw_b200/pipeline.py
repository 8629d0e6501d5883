"""Evaluation pipeline around the sampler: the device-side equivalent of the body of `test_hyber`
(reference runners/diffpose_frame.py:330-391) and its multi-GPU sharding.

    uv --GCNpose--> xyz --root-centre, cat--> uvxyz --H hypotheses x T DDIM steps--> mean over H --> MPJPE / P-MPJPE

Poses are independent, so ranks take contiguous pose ranges, keep all hypotheses of a pose on the same rank and
exchange nothing inside the loop; one all-reduce of three fp64 partial sums ends the evaluation
(SURVEY.md section 8e).
"""
from __future__ import annotations

import torch

from .metrics import pose_error_sums
from .sampler import sample


def shard_range(n, rank, world):
    """Contiguous, balanced pose range [lo, hi) of `rank` out of `world`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def lift_and_refine(model_diff, x_uvxyz=None, *, model_pose=None, input_2d=None, src_mask=None, seq, betas, eta=0.0,
                    test_times=1, noise=None):
    """Two-stage inference for one batch.  Give either `x_uvxyz` [B,17,5] or (`model_pose`, `input_2d` [B,17,2]).
    Returns the hypothesis-averaged uvxyz [B,17,5]."""
    if x_uvxyz is None:
        xyz = model_pose(input_2d, src_mask)
        xyz = xyz - xyz[:, :1, :]                       # out-of-place root-centring
        x_uvxyz = torch.cat([input_2d.to(xyz.dtype), xyz], dim=2)
    return sample(model_diff, x_uvxyz, src_mask, seq, betas, eta=eta, noise=noise, n_hyp=test_times,
                  repeat_input=True, mean_over_hyp=True)


def evaluate_shard(model_diff, x_uvxyz, targets_3d, *, src_mask=None, seq, betas, eta=0.0, test_times=1, noise=None,
                   batch_size=None, model_pose=None, input_2d=None):
    """Sample this rank's poses and return fp64 partial sums [sum mpjpe, sum p_mpjpe, n] (metres) on the device."""
    n = targets_3d.shape[0]
    dev = targets_3d.device
    sums = torch.zeros(3, device=dev, dtype=torch.float64)
    bs = batch_size or max(n, 1)
    for lo in range(0, n, bs):
        hi = min(n, lo + bs)
        nz = None
        if noise is not None:   # noise is [T, H*n, 17, 5] hypothesis-major over this shard
            T = noise.shape[0]
            nz = noise.reshape(T, test_times, n, 17, -1)[:, :, lo:hi].reshape(T, test_times * (hi - lo), 17, -1)
        out = lift_and_refine(model_diff, None if x_uvxyz is None else x_uvxyz[lo:hi], model_pose=model_pose,
                              input_2d=None if input_2d is None else input_2d[lo:hi], src_mask=src_mask, seq=seq,
                              betas=betas, eta=eta, test_times=test_times, noise=nz)
        pose_error_sums(out, targets_3d[lo:hi], sums=sums)
    return sums


def reduce_metrics(sums, group=None):
    """All-reduce the per-rank partial sums (NCCL on GPUs, gloo in CPU tests) and return
    (mpjpe_mm, p_mpjpe_mm, n) as python floats -- the AverageMeter values of diffpose_frame.py:386-387."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    s = sums.detach().cpu()
    n = float(s[2])
    if n == 0:
        return float("nan"), float("nan"), 0.0
    return float(s[0]) / n * 1000.0, float(s[1]) / n * 1000.0, n
