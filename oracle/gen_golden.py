"""Generate tests/golden/reference_vectors.npz by running the REAL reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py

What it does
  1. imports the reference modules with the two shims of SURVEY.md section 8c (stub `lib2to3.refactor`, make
     `Tensor.cuda` a no-op on CPU) -- no reference source is copied;
  2. builds reference models under fixed seeds, checks that `diffpose_nw_b200`'s own initialisation reproduces
     the same weights bit for bit (so fixtures carry seeds, not 4 MB of weights);
  3. runs reference forward / generalized_steps / metrics on seeded inputs and stores inputs + outputs;
  4. asserts that oracle/diffpose_oracle.py agrees with the reference on every case (this is what pins the oracle).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

# --- shims
_stub = types.ModuleType("lib2to3.refactor")
_stub.get_fixers_from_package = lambda *a, **k: []
sys.modules.setdefault("lib2to3", types.ModuleType("lib2to3"))
sys.modules["lib2to3.refactor"] = _stub
torch.Tensor.cuda = lambda self, *a, **k: self
sys.path.insert(0, REF)

from models.gcndiff import GCNdiff  # noqa: E402  (reference)
from models.gcnpose import GCNpose  # noqa: E402
from models.ChebConv import adj_mx_from_edges as ref_adj  # noqa: E402
from common.utils_diff import generalized_steps as ref_steps, get_beta_schedule as ref_betas, compute_alpha as ref_alpha  # noqa: E402
from common.loss import mpjpe as ref_mpjpe, p_mpjpe as ref_p_mpjpe  # noqa: E402

from oracle import diffpose_oracle as O  # noqa: E402
import diffpose_nw_b200 as D  # noqa: E402

torch.set_grad_enabled(False)
EDGES = torch.tensor(O.H36M_EDGES, dtype=torch.long)
out = {}


def ns(cfg):
    return cfg


def check(name, a, b, tol):
    err = (a.double() - b.double()).abs().max().item()
    print(f"  oracle vs reference  {name:<34s} max|diff| = {err:.3e}")
    assert err <= tol, (name, err)


def same_weights(ref_model, mine):
    rs, ms = ref_model.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys()), "state_dict key order differs"
    for k in rs:
        assert rs[k].shape == ms[k].shape and torch.equal(rs[k], ms[k]), k


adj = ref_adj(num_pts=17, edges=EDGES, sparse=False)
assert torch.equal(adj, O.adjacency()) and torch.equal(adj, D.adj_mx_from_edges())
out["adj"] = adj.numpy()
mask_all = torch.ones(1, 1, 17, dtype=torch.bool)
mask_part = mask_all.clone()
mask_part[0, 0, 3] = False
mask_part[0, 0, 10] = False
out["mask_part"] = mask_part.numpy()

# ---------------------------------------------------------------- schedules
for kind in ["linear", "quad", "const", "jsd", "sigmoid"]:
    b = ref_betas(kind, beta_start=1e-4, beta_end=1e-3, num_diffusion_timesteps=51)
    assert np.array_equal(b, O.beta_schedule(kind, 1e-4, 1e-3, 51))
    assert np.array_equal(b, D.get_beta_schedule(kind, beta_start=1e-4, beta_end=1e-3, num_diffusion_timesteps=51))
    out[f"betas_{kind}"] = b
betas = torch.from_numpy(out["betas_linear"]).float()
tt = torch.tensor([-1, 0, 6, 12, 23, 49], dtype=torch.long)
out["alpha_bar_t"] = tt.numpy()
out["alpha_bar"] = ref_alpha(betas, tt).numpy()
assert torch.equal(ref_alpha(betas, tt), O.alpha_bar(betas, tt)) and torch.equal(ref_alpha(betas, tt), D.compute_alpha(betas, tt))


def run_case(tag, model_over, seed_w, perturb, n, t_vals, seq, eta, mask, seed_x=1, seed_noise=7):
    cfg = O.default_config(**model_over)
    torch.manual_seed(seed_w)
    ref = GCNdiff(adj, cfg).eval()
    torch.manual_seed(seed_w)
    mine = D.FusedGCNdiff(adj, cfg)
    same_weights(ref, mine)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    if perturb:
        sd = O.perturb_state_dict(sd, seed=perturb)
        ref.load_state_dict(sd)
    out[f"{tag}.param_sum"] = np.float64(sum(v.double().sum().item() for v in sd.values()))
    x = O.synthetic_poses(n, seed=seed_x)
    out[f"{tag}.x"] = x.numpy()
    # forward
    t = torch.tensor(t_vals, dtype=torch.float32)
    out[f"{tag}.t"] = t.numpy()
    eps = ref(x, mask, t, 0)
    out[f"{tag}.eps"] = eps.numpy()
    L, nh = cfg.model.num_layer, cfg.model.n_head
    check(f"{tag}.eps", O.gcndiff_forward(sd, adj, L, nh, x, mask, t), eps, 2e-5)
    # sampling with host-supplied noise: pre-draw exactly the stream randn_like would consume
    seq = list(seq)
    out[f"{tag}.seq"] = np.asarray(seq, dtype=np.int64)
    out[f"{tag}.eta"] = np.float64(eta)
    torch.manual_seed(seed_noise)
    noise = torch.stack([torch.randn_like(x) for _ in seq])
    torch.manual_seed(seed_noise)
    xs, x0s = ref_steps(x, mask, seq, ref, betas, eta=eta)
    out[f"{tag}.noise"] = noise.numpy()
    out[f"{tag}.x_final"] = xs[-1].numpy()
    out[f"{tag}.x0_last"] = x0s[-1].numpy()
    den = lambda xt, m, tt_: O.gcndiff_forward(sd, adj, L, nh, xt, m, tt_)
    oxs, ox0 = O.ddim_sample(x, mask, seq, den, betas, eta=eta, noise=noise)
    check(f"{tag}.x_final", oxs[-1], xs[-1], 2e-5)
    check(f"{tag}.x0_last", ox0[-1], x0s[-1], 2e-5)
    sc = O.ddim_scalars(betas, seq, eta)
    out[f"{tag}.scalars"] = np.asarray([[s["t"], s["at"], s["at_next"], s["c1"], s["c2"]] for s in sc], dtype=np.float64)
    return ref, sd, x


print("GCNdiff cases")
# A: cpn.yml shape, default init, SURVEY 8c known answers (B=4, t=12, seq [0,12])
refA, sdA, xA = run_case("A", {}, 0, 0, 4, [12.0] * 4, range(0, 24, 12), 0.0, mask_all)
assert abs(out["A.param_sum"] - 1124.303041338549) < 1e-9
assert abs(out["A.eps"].astype(np.float64).sum() - 215.32830626517534) < 1e-6
assert abs(out["A.x_final"].astype(np.float64).sum() - (-5.624957477208227)) < 1e-6
# A1: same model, eta = 1 with seed-7 noise
run_case("A1", {}, 0, 0, 4, [12.0] * 4, range(0, 24, 12), 1.0, mask_all)
assert abs(out["A1.x_final"].astype(np.float64).sum() - (-7.469549811212346)) < 1e-6
# B: every parameter perturbed, per-sample t, partial key mask, 5 steps, eta 0.5
run_case("B", {}, 0, 3, 8, [0, 3, 7, 12, 23, 31, 49, 50], [0, 5, 11, 17, 23], 0.5, mask_part)
# C: gt.yml sequence [0,6], perturbed, all-true mask, eta 0
run_case("C", {}, 0, 5, 6, [6.0] * 6, range(0, 12, 6), 0.0, mask_all)
# D: 50 steps (throughput-sweep schedule), perturbed, eta 1
run_case("D", {}, 0, 3, 4, [49.0, 25.0, 1.0, 0.0], range(0, 50), 1.0, mask_all)
# E: a different architecture (hid 64, 2 layers, 2 heads) exercising the generic fp32 engine
run_case("E", dict(hid_dim=64, num_layer=2, n_head=2), 11, 4, 5, [1, 2, 3, 4, 5], [0, 10, 20], 1.0, mask_part)
# F: hid 128, 8 heads, 1 layer
run_case("F", dict(hid_dim=128, num_layer=1, n_head=8), 12, 6, 3, [9, 19, 29], [0, 25], 0.0, mask_all)

print("GCNpose cases")
for tag, perturb in [("P0", 0), ("P1", 8)]:
    cfg = O.default_config(coords_dim=[2, 3])
    torch.manual_seed(0)
    ref = GCNpose(adj, cfg).eval()
    torch.manual_seed(0)
    mine = D.FusedGCNpose(adj, cfg)
    same_weights(ref, mine)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    if perturb:
        sd = O.perturb_state_dict(sd, seed=perturb)
        ref.load_state_dict(sd)
    out[f"{tag}.param_sum"] = np.float64(sum(v.double().sum().item() for v in sd.values()))
    uv = O.synthetic_poses(7, seed=21)[:, :, :2].contiguous()
    out[f"{tag}.uv"] = uv.numpy()
    xyz = ref(uv, mask_all)
    out[f"{tag}.xyz"] = xyz.numpy()
    check(f"{tag}.xyz", O.gcnpose_forward(sd, adj, 5, 4, uv, mask_all), xyz, 2e-5)

print("metrics")
g = torch.Generator().manual_seed(31)
gt = torch.randn(16, 17, 3, generator=g) * 0.3
pred = gt + torch.randn(16, 17, 3, generator=g) * 0.05
pred[3] = gt[3] * 1.7 + 0.2            # pure similarity transform -> P-MPJPE ~ 0
pred[5] = -gt[5]                       # reflection: exercises the det(R) < 0 branch
gt_c, pred_c = gt - gt[:, :1], pred - pred[:, :1]
out["M.gt"], out["M.pred"] = gt.numpy(), pred.numpy()
out["M.mpjpe"] = np.float64(ref_mpjpe(pred_c, gt_c).item())
out["M.p_mpjpe"] = np.float64(ref_p_mpjpe(pred_c.numpy().copy(), gt_c.numpy().copy()))
pp = O.p_mpjpe_per_pose(pred_c.numpy(), gt_c.numpy())
out["M.p_mpjpe_per_pose"] = pp
assert abs(pp.mean() - out["M.p_mpjpe"]) < 1e-6 and abs(O.mpjpe(pred_c, gt_c).item() - out["M.mpjpe"]) < 1e-7
print(f"  mpjpe {out['M.mpjpe']:.6f}  p_mpjpe {out['M.p_mpjpe']:.6f}")

dst = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst), "bytes,", len(out), "arrays")
