"""CPU oracle for the DiffPose frame-based DDIM sampling path.

TEST INFRASTRUCTURE ONLY.  This module is a CPU restatement (PyTorch CPU tensors,
fp32 by default, fp64 on request) of the reference algorithm.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import it; the
product (`diffpose_nw_b200`) never does and has no CPU fallback.

Parity status: PINNED.  `oracle/gen_golden.py` runs the real reference
(imported from /root/reference in the build container) next to this restatement on
the same weights/inputs/noise and stores the reference outputs under
`tests/golden/`; `tests/test_oracle_golden.py` re-checks this file against those
vectors on every run (no reference needed), and against the survey-time known
answers of SURVEY.md section 8c.

Every function names the reference lines it restates (paths relative to the
reference repository root).  The restatement is functional (a flat state_dict of
tensors goes in) instead of the reference's nn.Module tree, but executes the same
sequence of tensor operations so that its CPU timing is representative of the
reference's CPU path (`cpu_baseline.kind = "port"`).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

# 16 bones of the 17-joint Human3.6M skeleton -- runners/diffpose_frame.py:120-124
H36M_EDGES = [(0, 1), (1, 2), (2, 3), (0, 4), (4, 5), (5, 6), (0, 7), (7, 8), (8, 9), (9, 10),
              (8, 11), (11, 12), (12, 13), (8, 14), (14, 15), (15, 16)]


# --------------------------------------------------------------------------- schedules
def beta_schedule(kind, beta_start, beta_end, n):
    """common/utils_diff.py:7-37 -- float64 numpy beta schedules."""
    if kind == "quad":
        b = np.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=np.float64) ** 2
    elif kind == "linear":
        b = np.linspace(beta_start, beta_end, n, dtype=np.float64)
    elif kind == "const":
        b = beta_end * np.ones(n, dtype=np.float64)
    elif kind == "jsd":
        b = 1.0 / np.linspace(n, 1, n, dtype=np.float64)
    elif kind == "sigmoid":
        s = np.linspace(-6, 6, n)
        b = 1.0 / (np.exp(-s) + 1.0) * (beta_end - beta_start) + beta_start
    else:
        raise NotImplementedError(kind)
    assert b.shape == (n,)
    return b


def alpha_bar(betas, t):
    """common/utils_diff.py:40-43 -- cumprod of (1-[0,beta]) gathered at t+1, shape [n,1,1]."""
    padded = torch.cat([torch.zeros(1, dtype=betas.dtype), betas], dim=0)
    return (1 - padded).cumprod(dim=0).index_select(0, t + 1).view(-1, 1, 1)


def eval_sequence(skip_type, n_diffusion, n_steps):
    """runners/diffpose_frame.py:310-317 -- the timestep subsequence used at test time."""
    if skip_type == "uniform":
        return list(range(0, n_diffusion, n_diffusion // n_steps))
    if skip_type == "quad":
        s = np.linspace(0, np.sqrt(n_diffusion * 0.8), n_steps) ** 2
        return [int(v) for v in list(s)]
    raise NotImplementedError(skip_type)


# --------------------------------------------------------------------------- graph
def adjacency(n_pts=17, edges=H36M_EDGES):
    """models/ChebConv.py:36-48 (+normalize :17-24) -- symmetric adjacency + I, row-normalised, fp32."""
    a = np.zeros((n_pts, n_pts), dtype=np.float64)   # scipy promotes to float64 (sp.eye), cast to fp32 last
    for i, j in edges:
        a[i, j] = 1.0
    up = a.T > a
    a = a + a.T * up - a * up
    a = a + np.eye(n_pts)
    rs = a.sum(1)
    inv = np.where(rs == 0, 0.0, 1.0 / rs)
    a = np.diag(inv) @ a
    return torch.tensor(a, dtype=torch.float32)


def cheb_basis(adj):
    """models/ChebConv.py:90-130 -- [T0,T1,T2] = [I, L, 2L^2-I], L = I - D^-1/2 A D^-1/2."""
    n = adj.size(0)
    d = torch.diag(torch.sum(adj, dim=-1) ** (-1 / 2))
    lap = torch.eye(n, dtype=adj.dtype) - torch.mm(torch.mm(d, adj), d)
    basis = torch.zeros(3, n, n, dtype=adj.dtype)
    basis[0] = torch.eye(n, dtype=adj.dtype)
    basis[1] = lap
    basis[2] = 2 * torch.mm(lap, basis[1]) - basis[0]
    return basis


# --------------------------------------------------------------------------- operators
def cheb_conv(x, adj, weight, bias):
    """models/ChebConv.py:74-88 -- sum_k (T_k x) W_k + b ; x [B,N,Cin], W [3,1,Cin,Cout]."""
    basis = cheb_basis(adj).unsqueeze(1)          # [3,1,N,N]
    r = torch.matmul(basis, x)                    # [3,B,N,Cin]
    r = torch.matmul(r, weight)                   # [3,B,N,Cout]
    return torch.sum(r, dim=0) + bias


def graph_conv(x, adj, weight, bias):
    """models/ChebConv.py:145-151 -- relu(dropout(relu(cheb))) == relu(cheb) in eval mode."""
    return torch.relu(cheb_conv(x, adj, weight, bias))


def layer_norm(x, a_2, b_2, eps=1e-6):
    """models/GraFormer.py:67-70 -- unbiased std, eps added to std (not to the variance)."""
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return a_2 * (x - mean) / (std + eps) + b_2


def multi_head_attention(x, mask, sd, pre, n_head):
    """models/GraFormer.py:127-140 + :99-113 -- self-attention over the joints of one pose."""
    nb, n_pts, c = x.shape
    d_k = c // n_head
    if mask is not None:
        mask = mask.unsqueeze(1)
    q, k, v = [torch.nn.functional.linear(x, sd[f"{pre}.linears.{i}.weight"], sd[f"{pre}.linears.{i}.bias"])
               .view(nb, -1, n_head, d_k).transpose(1, 2) for i in range(3)]
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(d_k)
    if mask is not None:
        scores = scores.masked_fill(mask == 0, -1e9)
    p = torch.softmax(scores, dim=-1)
    o = torch.matmul(p, v).transpose(1, 2).contiguous().view(nb, -1, n_head * d_k)
    return torch.nn.functional.linear(o, sd[f"{pre}.linears.3.weight"], sd[f"{pre}.linears.3.bias"])


def lam_gconv(x, a_hat, w, b, relu):
    """models/GraFormer.py:174-186 -- fc(L^ x), L^ = D A_hat D with D = (colsum+1e-5)^-1/2."""
    nb, n = x.size(0), a_hat.size(0)
    a = a_hat.unsqueeze(0).repeat(nb, 1, 1)
    d = (torch.sum(a, 1) + 1e-5) ** (-0.5)
    lap = d.view(nb, n, 1) * a * d.view(nb, 1, n)
    y = torch.nn.functional.linear(torch.bmm(lap, x), w, b)
    return torch.relu(y) if relu else y


def graph_net(x, sd, pre):
    """models/GraFormer.py:198-201 -- the 'feed-forward' of a GraAttenLayer."""
    a_hat = sd[f"{pre}.A_hat"]
    h = lam_gconv(x, a_hat, sd[f"{pre}.gconv1.fc.weight"], sd[f"{pre}.gconv1.fc.bias"], True)
    return lam_gconv(h, a_hat, sd[f"{pre}.gconv2.fc.weight"], sd[f"{pre}.gconv2.fc.bias"], False)


def gra_atten_layer(x, mask, sd, l, n_head):
    """models/GraFormer.py:94-96 with SublayerConnection :80-81 (pre-norm residual, dropout = id)."""
    p = f"atten_layers.{l}"
    x = x + multi_head_attention(layer_norm(x, sd[f"{p}.sublayer.0.norm.a_2"], sd[f"{p}.sublayer.0.norm.b_2"]),
                                 mask, sd, f"{p}.self_attn", n_head)
    return x + graph_net(layer_norm(x, sd[f"{p}.sublayer.1.norm.a_2"], sd[f"{p}.sublayer.1.norm.b_2"]),
                         sd, f"{p}.feed_forward")


def swish(x):
    """models/gcndiff.py:35-37."""
    return x * torch.sigmoid(x)


def timestep_embedding(t, dim):
    """models/gcndiff.py:15-33 -- sinusoidal embedding [n, dim] (sin | cos)."""
    half = dim // 2
    f = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    e = t.float()[:, None] * f[None, :]
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    if dim % 2 == 1:
        e = torch.nn.functional.pad(e, (0, 1, 0, 0))
    return e


def strip_module_prefix(sd):
    """DataParallel checkpoints prefix every key with 'module.' (runners/diffpose_frame.py:127,131-132)."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def gcndiff_forward(sd, adj, n_layer, n_head, x, mask, t):
    """models/gcndiff.py:101-113 (+ _ResChebGC_diff :48-53) -- eps = GCNdiff(x, mask, t, cemd)."""
    hid = sd["gconv_input.weight"].shape[-1]
    temb = timestep_embedding(t, hid).to(x.dtype)
    temb = torch.nn.functional.linear(temb, sd["temb.dense.0.weight"], sd["temb.dense.0.bias"])
    temb = torch.nn.functional.linear(swish(temb), sd["temb.dense.1.weight"], sd["temb.dense.1.bias"])
    out = cheb_conv(x, adj, sd["gconv_input.weight"], sd["gconv_input.bias"])
    for l in range(n_layer):
        out = gra_atten_layer(out, mask, sd, l, n_head)
        g = f"gconv_layers.{l}"
        h = graph_conv(out, adj, sd[f"{g}.gconv1.gconv.weight"], sd[f"{g}.gconv1.gconv.bias"])
        h = h + torch.nn.functional.linear(swish(temb), sd[f"{g}.temb_proj.weight"], sd[f"{g}.temb_proj.bias"])[:, None, :]
        h = graph_conv(h, adj, sd[f"{g}.gconv2.gconv.weight"], sd[f"{g}.gconv2.gconv.bias"])
        out = out + h
    return cheb_conv(out, adj, sd["gconv_output.weight"], sd["gconv_output.bias"])


def gcnpose_forward(sd, adj, n_layer, n_head, x, mask):
    """models/gcnpose.py:101-113 (+ _ResChebGC models/ChebConv.py:154-165) -- xyz = GCNpose(uv, mask)."""
    out = cheb_conv(x, adj, sd["gconv_input.weight"], sd["gconv_input.bias"])
    for l in range(n_layer):
        out = gra_atten_layer(out, mask, sd, l, n_head)
        g = f"gconv_layers.{l}"
        h = graph_conv(out, adj, sd[f"{g}.gconv1.gconv.weight"], sd[f"{g}.gconv1.gconv.bias"])
        h = graph_conv(h, adj, sd[f"{g}.gconv2.gconv.weight"], sd[f"{g}.gconv2.gconv.bias"])
        out = out + h
    return cheb_conv(out, adj, sd["gconv_output.weight"], sd["gconv_output.bias"])


# --------------------------------------------------------------------------- sampler
def ddim_scalars(betas, seq, eta):
    """common/utils_diff.py:50-64 -- the per-step scalars, evaluated with the same fp32 tensor ops.

    Returns a list (in execution order) of dicts {t, next_t, at, at_next, c1, c2} of python floats
    that are exact fp32 values.
    """
    seq = list(seq)
    seq_next = [-1] + seq[:-1]
    out = []
    for i, j in zip(reversed(seq), reversed(seq_next)):
        at = alpha_bar(betas, torch.tensor([i], dtype=torch.long))
        an = alpha_bar(betas, torch.tensor([j], dtype=torch.long))
        c1 = eta * ((1 - at / an) * (1 - an) / (1 - at)).sqrt()
        c2 = ((1 - an) - c1 ** 2).sqrt()
        out.append(dict(t=float(i), next_t=float(j), at=at.item(), at_next=an.item(), c1=c1.item(), c2=c2.item()))
    return out


def ddim_sample(x, mask, seq, denoiser, betas, eta=0.0, noise=None):
    """common/utils_diff.py:46-67 -- generalized_steps.

    `denoiser(xt, mask, t)` returns eps.  `noise` (optional, [T, n, 17, C]) replaces the
    per-step `torch.randn_like(x)` draw so that runs are comparable across implementations;
    when omitted the draw is made exactly where the reference makes it (every step, even for eta=0).
    Returns (xs, x0_preds) like the reference.
    """
    with torch.no_grad():
        n = x.size(0)
        seq = list(seq)
        seq_next = [-1] + seq[:-1]
        xs, x0_preds = [x], []
        for s, (i, j) in enumerate(zip(reversed(seq), reversed(seq_next))):
            t = torch.ones(n) * i
            nt = torch.ones(n) * j
            at = alpha_bar(betas, t.long()).to(x.dtype)
            an = alpha_bar(betas, nt.long()).to(x.dtype)
            xt = xs[-1]
            et = denoiser(xt, mask, t.float())
            x0 = (xt - et * (1 - at).sqrt()) / at.sqrt()
            x0_preds.append(x0)
            c1 = eta * ((1 - at / an) * (1 - an) / (1 - at)).sqrt()
            c2 = ((1 - an) - c1 ** 2).sqrt()
            z = torch.randn_like(x) if noise is None else noise[s]
            xs.append(an.sqrt() * x0 + c1 * z + c2 * et)
    return xs, x0_preds


def hypothesis_mean(x, n_hyp):
    """runners/diffpose_frame.py:382 -- mean over the hypothesis-major leading factor of [H*B,17,C]."""
    return torch.mean(x.reshape(n_hyp, -1, x.shape[-2], x.shape[-1]), 0)


def root_centre(x):
    """Intended semantics of runners/diffpose_frame.py:338,384,385 (x - x[:, :1]), done out of place.
    The reference's aliased in-place form is a data race on CUDA and a no-op for joints 1..16 on CPU
    (SURVEY.md 8a quirk 4); the pipeline restates the intent, not the bug."""
    return x - x[:, :1, :]


# --------------------------------------------------------------------------- metrics
def mpjpe(pred, target):
    """common/loss.py:7-13 -- mean Euclidean joint distance (scalar)."""
    return torch.mean(torch.norm(pred - target, dim=-1))


def p_mpjpe_per_pose(pred, target):
    """common/utils.py:155-187 / common/loss.py:25-64 -- Procrustes-aligned MPJPE, one value per pose
    (numpy float64 here; the scalar form of loss.py is the mean of these)."""
    pred = np.asarray(pred, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    mu_x = target.mean(axis=1, keepdims=True)
    mu_y = pred.mean(axis=1, keepdims=True)
    x0, y0 = target - mu_x, pred - mu_y
    nx = np.sqrt((x0 ** 2).sum(axis=(1, 2), keepdims=True))
    ny = np.sqrt((y0 ** 2).sum(axis=(1, 2), keepdims=True))
    x0, y0 = x0 / nx, y0 / ny
    h = np.matmul(x0.transpose(0, 2, 1), y0)
    u, s, vt = np.linalg.svd(h)
    v = vt.transpose(0, 2, 1)
    r = np.matmul(v, u.transpose(0, 2, 1))
    sign = np.sign(np.expand_dims(np.linalg.det(r), axis=1))
    v[:, :, -1] *= sign
    s[:, -1] *= sign.flatten()
    r = np.matmul(v, u.transpose(0, 2, 1))
    tr = np.expand_dims(s.sum(axis=1, keepdims=True), axis=2)
    a = tr * nx / ny
    t = mu_x - a * np.matmul(mu_y, r)
    aligned = a * np.matmul(pred, r) + t
    return np.linalg.norm(aligned - target, axis=-1).mean(axis=-1)


# --------------------------------------------------------------------------- synthetic workload
def default_config(**model_over):
    """configs/human36m_diffpose_uvxyz_cpn.yml:9-36 as a nested namespace (model/diffusion/testing)."""
    model = dict(hid_dim=96, emd_dim=96, coords_dim=[5, 5], num_layer=5, n_head=4, dropout=0.25, n_pts=17,
                 var_type="fixedsmall")
    model.update(model_over)
    return SimpleNamespace(
        model=SimpleNamespace(**model),
        diffusion=SimpleNamespace(beta_schedule="linear", beta_start=0.0001, beta_end=0.001, num_diffusion_timesteps=51),
        testing=SimpleNamespace(test_times=1, test_timesteps=2, test_num_diffusion_timesteps=24),
        training=SimpleNamespace(batch_size=1024))


def synthetic_poses(n, seed=1):
    """u,v ~ U(-1,1); xyz ~ N(0,0.3^2) root-relative (root xyz = 0)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.cat([torch.rand(n, 17, 2, generator=g) * 2 - 1, torch.randn(n, 17, 3, generator=g) * 0.3], -1)
    x[:, 0, 2:] = 0
    return x


def synthetic_targets(x_uvxyz, seed=2):
    """targets_3d = xyz + N(0,0.05^2), root = 0."""
    g = torch.Generator().manual_seed(seed)
    t = x_uvxyz[:, :, 2:] + torch.randn(x_uvxyz.shape[0], 17, 3, generator=g) * 0.05
    t[:, 0] = 0
    return t


def perturb_state_dict(sd, seed=3, scale=0.05):
    """Hard part 6 of SURVEY.md 7: default init leaves A_hat = I, LayerNorm = (1,0), Cheb bias = 0, which
    hides kernels that ignore those tensors.  Adds seeded noise to EVERY tensor (A_hat stays
    positive-column-sum so D^-1/2 is finite)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sd:
        v = sd[k].clone()
        if k.endswith("A_hat"):
            v = v + torch.rand(v.shape, generator=g) * 0.2
        else:
            v = v + torch.randn(v.shape, generator=g) * scale
        out[k] = v
    return out
