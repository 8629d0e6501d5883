#!/usr/bin/env python
"""Per-phase timeline of the tcg engine (GPU only): where one layer's cycles go.

    python tools/phase_trace.py [--batch 1024]

Uses dp_set_trace (include/diffpose_b200.h): thread 0 of CTA 0 stamps clock64() just before every "operands ready" signal
(S) and just after every "accumulator ready" wait (A).  Between A_k and S_{k+1} the compute warps run an epilogue; between
S_k and A_k the MMA group runs (plus hand-over latency and the slowest warp's lag).  Prints mean cycles per phase of a layer.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffpose_nw_b200 as D
from diffpose_nw_b200 import _lib
from oracle import diffpose_oracle as O


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--names", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = O.default_config()
    torch.manual_seed(0)
    model = D.FusedGCNdiff(D.adj_mx_from_edges(), cfg).to(dev).set_engine("tcg")
    betas = torch.from_numpy(O.beta_schedule("linear", 1e-4, 1e-3, 51)).float()
    x = O.synthetic_poses(args.batch, seed=1).to(dev)
    for _ in range(3):
        D.generalized_steps(x, None, [0, 12], model, betas)
    buf = torch.zeros(4096, dtype=torch.int64, device=dev)
    _lib.check(_lib.load().dp_set_trace(model._handle, buf.data_ptr(), buf.numel()), "dp_set_trace")
    D.generalized_steps(x, None, [0, 12], model, betas)
    torch.cuda.synchronize()
    _lib.check(_lib.load().dp_set_trace(model._handle, None, 0), "dp_set_trace")
    full = buf.cpu().numpy()
    t = full[: buf.numel() // 2]
    t = t[t > 0]
    it = full[buf.numel() // 2:]
    it = it[it > 0]          # issuer: (before wait_rdy = all MMAs of the previous group issued, after wait_rdy) pairs
    n_layers, n_steps = 5, 2
    per_step = len(t) // n_steps                      # 2 (input conv) + 26 per layer + 2 (output conv)
    per_layer = (per_step - 4) // n_layers
    print(f"{len(t)} stamps, {per_step} per step, {per_layer} per layer; total {t[-1] - t[0]} cycles for {n_steps} steps x {n_layers} layers")
    steps = t[: per_step * n_steps].reshape(n_steps, per_step)
    lay = steps[:, 2:2 + per_layer * n_layers].reshape(n_steps * n_layers, per_layer)
    seg = np.diff(lay, axis=1)                     # [layers][per_layer-1]
    mean = seg[1:].mean(axis=0)                    # skip the very first layer (cold)
    tot_mma = tot_epi = 0.0
    for i, v in enumerate(mean):
        kind = "MMA+handover" if i % 2 == 0 else "epilogue    "
        if i % 2 == 0:
            tot_mma += v
        else:
            tot_epi += v
        print(f"  {i:2d} {kind} {v:8.0f} cyc  min {seg[1:, i].min():6d} max {seg[1:, i].max():6d}")
    cross = [lay[i + 1, 0] - lay[i, -1] for i in range(len(lay) - 1) if (i + 1) % n_layers != 0]
    print(f"  layer-crossing epilogue (residual update + LN0 of next layer): {np.mean(cross):8.0f} cyc")
    print(f"  per layer: MMA+handover {tot_mma:.0f}, epilogues {tot_epi + np.mean(cross):.0f}, sum {tot_mma + tot_epi + np.mean(cross):.0f}")
    for s_ in range(n_steps):
        st = steps[s_]
        print(f"  step {s_}: in-conv MMA {st[1] - st[0]}, X load + LN0 {st[2] - st[1]}, residual+split {st[-2] - st[-3]}, out-conv MMA {st[-1] - st[-2]}"
              + (f", eps + DDIM + gather of next step {steps[s_ + 1][0] - st[-1]}" if s_ + 1 < n_steps else ""))
    # issuer view of each hand-over k (same order as the compute stamps): S_k -> woke -> issued -> A_k
    npair = min(len(it) // 2, len(t) // 2)
    woke = it[1:2 * npair:2]
    issued = np.r_[it[2:2 * npair:2], it[2 * npair - 1]]
    S, A = t[0:2 * npair:2], t[1:2 * npair:2]
    k0 = per_step // 2                                # second step
    print("  hand-over breakdown per MMA group (second step): signal->issuer awake | issue loop | last issue->accumulator observed")
    for k in range(k0, min(k0 + 1 + per_layer // 2 + 1, npair)):
        print(f"    group {k - k0:2d}: {woke[k] - S[k]:6d} | {issued[k] - woke[k]:6d} | {A[k] - issued[k]:6d}")


if __name__ == "__main__":
    main()
