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
        # lift + out-of-place root-centring + concat in ONE launch (dp_lift); the sampler below is
        # the second and last launch of the batch (kernel-side repeat, hypothesis mean fused into its final store)
        x_uvxyz = getattr(model_pose, "module", model_pose).lift(input_2d, src_mask)
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


class HostStream:
    """Batches that live in HOST memory, sampled at device speed: the H2D copy of batch i+1 and the D2H copy of
    batch i-1 run on their own CUDA streams while batch i is in the sampler kernel (the reference's loop,
    runners/diffpose_frame.py:330-366, copies, computes and reads back strictly one after the other).

        hs = HostStream(model_diff, batch=1024, seq=seq, betas=betas, test_times=1)
        for uvxyz_cpu in loader:            # pinned [B,17,5] fp32 tensors
            done = hs.submit(uvxyz_cpu)     # -> the result of an EARLIER batch (pinned host tensor) or None
        for out in hs.drain(): ...

    Results come back in submission order.  `depth` batches are in flight; every buffer is allocated once.
    """

    def __init__(self, model_diff, batch, seq, betas, eta=0.0, test_times=1, src_mask=None, depth=3, device=None):
        from .sampler import ddim_steps
        self.model = getattr(model_diff, "module", model_diff)
        self.dev = torch.device(device) if device is not None else self.model._device()
        if self.dev.type != "cuda":
            raise RuntimeError("HostStream needs the model on a CUDA device (diffpose_nw_b200 has no CPU path)")
        self.B, self.H, self.seq, self.betas, self.eta, self.mask = batch, test_times, seq, betas, eta, src_mask
        self.steps = ddim_steps(betas, seq, eta)
        self.depth = depth
        c = self.model._c_in
        self.x_dev = [torch.empty(batch, 17, c, device=self.dev) for _ in range(depth)]
        self.out_host = [torch.empty(batch, 17, c).pin_memory() for _ in range(depth)]
        self.out_dev = [None] * depth
        self.s_in, self.s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.n_rows = [0] * depth
        self.head = 0          # next slot to fill
        self.inflight = 0

    def _collect(self, slot):
        self.ev_out[slot].synchronize()
        return self.out_host[slot][: self.n_rows[slot]]

    def submit(self, x_host):
        """Queue one pinned host batch [n<=B,17,c].  Returns the oldest finished result when the ring is full, else None."""
        ret = None
        slot = self.head
        if self.inflight == self.depth:
            ret = self._collect(slot).clone()
            self.inflight -= 1
        n = x_host.shape[0]
        if n > self.B:
            raise RuntimeError(f"batch of {n} poses exceeds the {self.B} this HostStream was built for")
        self.n_rows[slot] = n
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_done[slot])          # the kernel that last read this input buffer is finished
            self.x_dev[slot][:n].copy_(x_host, non_blocking=True)
            self.ev_in[slot].record(self.s_in)
        cur.wait_event(self.ev_in[slot])
        cur.wait_event(self.ev_out[slot])                     # the previous result of this slot has left the device
        out = sample(self.model, self.x_dev[slot][:n], self.mask, self.seq, self.betas, eta=self.eta, n_hyp=self.H,
                     repeat_input=True, mean_over_hyp=self.H > 1, steps=self.steps)
        self.out_dev[slot] = out                              # keep the tensor alive until its copy has run
        self.ev_done[slot].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_done[slot])
            self.out_host[slot][:n].copy_(out, non_blocking=True)
            self.ev_out[slot].record(self.s_out)
        self.head = (slot + 1) % self.depth
        self.inflight += 1
        return ret

    def drain(self):
        """Wait for and return (in order) the results still in flight; the returned tensors alias the ring buffers."""
        outs = []
        slot = (self.head - self.inflight) % self.depth
        while self.inflight:
            outs.append(self._collect(slot))
            slot = (slot + 1) % self.depth
            self.inflight -= 1
        return outs
