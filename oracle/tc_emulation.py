"""CPU emulation of the tensor-core engine's ROUNDING POINTS (test infrastructure, like the rest of oracle/).

The fast tcgen05 engine ("tcg", csrc/dp_tc2.cu) feeds fp16 operands (11-bit significand = TF32's) to the tensor cores
and accumulates in fp32.  This file restates GCNdiff.forward with a round-to-fp16 at exactly the places where that
kernel stores an fp16 operand, so that (a) the precision plan can be checked against the north_star tolerance before/without a GPU and
(b) kernel bugs (indexing, layouts) can be told apart from precision effects: the kernel must agree with this
emulation much more tightly (~1e-5) than with the fp32 oracle (~1e-4).
"""
from __future__ import annotations

import math

import torch

from . import diffpose_oracle as O


def r16(x):
    return x.to(torch.float16).to(torch.float32)


def split16(b):
    """bias travels through the MMA as hi + lo fp16 (two columns of the constant-one K slab)."""
    hi = r16(b)
    return hi + r16(b - hi)


def gcnpose_forward_tcg(sd, adj, n_layer, n_head, x, mask, p16=True):
    """GCNpose (models/gcnpose.py:101-113) on the second-generation engine: the same rounding points, no time embedding."""
    return gcndiff_forward_tcg(sd, adj, n_layer, n_head, x, mask, None, p16=p16)


def gcndiff_forward_tcg(sd, adj, n_layer, n_head, x, mask, t, p16=False, temb_in_gc2=False):
    """Rounding points of the second-generation tcgen05 engine (csrc/dp_tc2.cu): every tensor-core operand is fp16 --
    activations, weights and the learnable 17x17 matrix L^, which the kernel applies as per-pose MMAs; the Chebyshev
    matrices T1, T2 are applied exactly (integer rows in fp16 times an fp32 row scale, csrc/dp_api.cu integerise_rows);
    accumulation, the residual stream (TMEM), LayerNorm statistics and softmax are fp32.  temb_in_gc2: the sampler's form --
    the batch-uniform time embedding is not added to the hidden activation before its fp16 rounding but enters GC2 as
    the bias sum_k rowsum(T_k) (temb^T W_k) (dp_forward with per-sample timesteps keeps the addition).  Differences from the
    reference order: fc2 is commuted in front of the second L^ aggregation (L^(h W2) + b2), and the Chebyshev input
    panel is [x16 | r16(T1 x16) | r16(T2 x16)].  p16: attention probabilities are fp16 operands too."""
    hid = sd["gconv_input.weight"].shape[-1]
    basis = O.cheb_basis(adj)
    t1, t2 = basis[1], basis[2]
    temb = None
    if t is not None:     # GCNpose (t = None) has no time embedding
        temb = O.timestep_embedding(t, hid)
        temb = torch.nn.functional.linear(temb, sd["temb.dense.0.weight"], sd["temb.dense.0.bias"])
        temb = torch.nn.functional.linear(O.swish(temb), sd["temb.dense.1.weight"], sd["temb.dense.1.bias"])

    def cheb_tc(v16, w, b):
        a = torch.cat([v16, r16(torch.matmul(t1, v16)), r16(torch.matmul(t2, v16))], dim=-1)
        return a @ r16(w.reshape(3 * w.shape[2], w.shape[3])) + split16(b.reshape(-1))

    X = O.cheb_conv(x, adj, sd["gconv_input.weight"], sd["gconv_input.bias"])      # hi/lo split MMAs: fp32-level accuracy
    d_k = hid // n_head
    for l in range(n_layer):
        p, g = f"atten_layers.{l}", f"gconv_layers.{l}"
        # LayerNorm gains folded into the consumers (csrc/dp_tc2.cu tc2_pack_block_kernel): the operand is the plain
        # normalised row, the weights are a_2-scaled before the fp16 rounding, the shift b_2 W joins the bias
        one, zero = torch.ones(hid), torch.zeros(hid)
        a0, b0 = sd[f"{p}.sublayer.0.norm.a_2"], sd[f"{p}.sublayer.0.norm.b_2"]
        y = r16(O.layer_norm(X, one, zero))
        q, k, v = (r16(y @ r16(a0[:, None] * sd[f"{p}.self_attn.linears.{i}.weight"].T)
                       + split16(sd[f"{p}.self_attn.linears.{i}.bias"] + b0 @ sd[f"{p}.self_attn.linears.{i}.weight"].T)) for i in range(3))
        nb = X.shape[0]
        qh, kh, vh = (u.view(nb, -1, n_head, d_k).transpose(1, 2) for u in (q, k, v))
        sc = torch.matmul(qh, kh.transpose(-2, -1)) / math.sqrt(d_k)
        if mask is not None:
            sc = sc.masked_fill(mask.unsqueeze(1) == 0, -1e9)
        pr = torch.softmax(sc, dim=-1)
        if p16:
            pr = r16(pr)
        o = torch.matmul(pr, vh).transpose(1, 2).contiguous().view(nb, -1, hid)
        X = X + (r16(o) @ r16(sd[f"{p}.self_attn.linears.3.weight"]).T + split16(sd[f"{p}.self_attn.linears.3.bias"]))
        # GraphNet sublayer: aggregate, fc1, relu, fc2, aggregate (+ b2 straight onto the residual stream)
        a_hat = sd[f"{p}.feed_forward.A_hat"]
        dd = (a_hat.sum(0) + 1e-5) ** (-0.5)
        lhat32 = dd.view(-1, 1) * a_hat * dd.view(1, -1)
        lhat = r16(lhat32)
        # L^ (a_2 n + b_2) W1 = (L^ n)(a_2 W1) + r (b_2 W1), r = row sums of L^ (fp32; enters through the joint slab)
        a1, b1 = sd[f"{p}.sublayer.1.norm.a_2"], sd[f"{p}.sublayer.1.norm.b_2"]
        w1 = sd[f"{p}.feed_forward.gconv1.fc.weight"]
        y = r16(O.layer_norm(X, one, zero))
        g1 = r16(torch.matmul(lhat, y))
        h = r16(torch.relu(g1 @ r16(a1[:, None] * w1.T) + split16(sd[f"{p}.feed_forward.gconv1.fc.bias"])
                           + lhat32.sum(1).view(1, -1, 1) * (b1 @ w1.T).view(1, 1, -1)))
        z = r16(h @ r16(sd[f"{p}.feed_forward.gconv2.fc.weight"]).T)
        X = X + torch.matmul(lhat, z) + split16(sd[f"{p}.feed_forward.gconv2.fc.bias"])
        # residual Chebyshev block
        h1 = torch.relu(cheb_tc(r16(X), sd[f"{g}.gconv1.gconv.weight"], sd[f"{g}.gconv1.gconv.bias"]))
        tau = 0.0
        if temb is not None:
            tp = torch.nn.functional.linear(O.swish(temb), sd[f"{g}.temb_proj.weight"], sd[f"{g}.temb_proj.bias"])   # [n, hid]
            if temb_in_gc2:
                w2 = sd[f"{g}.gconv2.gconv.weight"][:, 0]                                                      # [3, hid, hid]
                rows = torch.stack([torch.ones(adj.shape[0]), t1.sum(1), t2.sum(1)])                             # [3, joints]
                tau = torch.einsum("kj,nc,kcd->njd", rows, tp, w2)
            else:
                h1 = h1 + tp[:, None, :]
        h2 = torch.relu(cheb_tc(r16(h1), sd[f"{g}.gconv2.gconv.weight"], sd[f"{g}.gconv2.gconv.bias"]) + tau)
        X = X + h2
    return O.cheb_conv(X, adj, sd["gconv_output.weight"], sd["gconv_output.bias"])
