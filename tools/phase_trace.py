#!/usr/bin/env python
"""Per-phase timeline of the tcg engine (GPU only): where one layer's cycles go.

    python tools/phase_trace.py [--batch 1024]

Uses dp_set_trace (include/diffpose_b200.h): thread 0 of CTA 0 stamps clock64() just before every "operands ready" signal
(S) and just after every "accumulator ready" wait (A).  A span that ends in S is compute-warp work (epilogue, LayerNorm,
softmax); a span that ends in A is time spent waiting for an MMA group (its execution, the hand-over latency and the
slowest warp's lag).  Prints the mean cycles of every span of a layer (layers 2..5 of the second DDIM step).
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
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = O.default_config()
    torch.manual_seed(0)
    model = D.FusedGCNdiff(D.adj_mx_from_edges(), cfg).to(dev).set_engine("tcg")
    betas = torch.from_numpy(O.beta_schedule("linear", 1e-4, 1e-3, 51)).float()
    x = O.synthetic_poses(args.batch, seed=1).to(dev)
    for _ in range(3):
        D.generalized_steps(x, None, [0, 12], model, betas)
    buf = torch.zeros(8192, dtype=torch.int64, device=dev)
    _lib.check(_lib.load().dp_set_trace(model._handle, buf.data_ptr(), buf.numel()), "dp_set_trace")
    D.generalized_steps(x, None, [0, 12], model, betas)
    torch.cuda.synchronize()
    _lib.check(_lib.load().dp_set_trace(model._handle, None, 0), "dp_set_trace")
    full = buf.cpu().numpy()
    k = full[-4:]
    raw = full[: buf.numel() // 2]
    raw = raw[raw > 0]
    t, kind = raw >> 1, raw & 1
    n_layers, n_steps = 5, 2
    per_step = len(t) // n_steps
    per_layer = (per_step - 4) // n_layers         # 2 stamps for the input convolution, 2 for the output convolution
    print(f"{len(t)} stamps, {per_step} per step, {per_layer} per layer; {t[-1] - t[0]} cycles for {n_steps} steps x {n_layers} layers")
    st = t[per_step:2 * per_step]                  # second step
    kd = kind[per_step:2 * per_step]
    lay = st[2:2 + per_layer * n_layers].reshape(n_layers, per_layer)
    lk = kd[2:2 + per_layer].tolist()
    d = np.diff(lay, axis=1)[1:].mean(axis=0)
    cross = np.mean([lay[i + 1, 0] - lay[i, -1] for i in range(n_layers - 1)])
    comp = wait = 0.0
    for i, v in enumerate(d):
        what = "compute" if lk[i + 1] == 0 else "wait MMA"
        if lk[i + 1] == 0:
            comp += v
        else:
            wait += v
        print(f"  span {i:2d} {'SA'[lk[i]]}->{'SA'[lk[i + 1]]} {what:9s} {v:7.0f}")
    print(f"  layer crossing (last A -> first S of the next layer: residual + LN0): {cross:7.0f}")
    print(f"  per layer: compute {comp + cross:.0f}, waiting for MMA groups {wait:.0f}, sum {comp + cross + wait:.0f}")
    print(f"  step: in-conv MMA {st[1] - st[0]}, first LN0 {st[2] - st[1]}, residual+split {st[-2] - st[-3]}, out-conv MMA {st[-1] - st[-2]}, "
          f"eps + DDIM + panel of the next step {t[per_step] - t[per_step - 1]}")


    print(f"  kernel (CTA 0): setup {k[1] - k[0]} cycles, setup end -> first signal {t[0] - k[1]}, main loop {t[-1] - t[0]}, last wait -> tiles done {k[2] - t[-1]}, "
          f"teardown {k[3] - k[2]}; entry -> exit {k[3] - k[0]} cycles = {(k[3] - k[0]) / 1.965e3:.1f} us at 1965 MHz")
    # device time of one call, measured from outside
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(50):
        D.generalized_steps(x, None, [0, 12], model, betas)
    ev1.record()
    torch.cuda.synchronize()
    print(f"  50 back-to-back calls: {ev0.elapsed_time(ev1) / 50 * 1e3:.1f} us per call (CUDA events)")


if __name__ == "__main__":
    main()
