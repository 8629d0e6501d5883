"""CPU: the oracle restatement against the reference's own outputs (tests/golden, made by oracle/gen_golden.py)
and against the survey-time known answers."""
import numpy as np
import pytest
import torch

from oracle import diffpose_oracle as O
from _cases import DIFF_CASES, POSE_CASES, betas, build_diff, build_pose, mask_for, t

torch.set_grad_enabled(False)


def test_known_answers(golden):
    # SURVEY.md 8c pins, reproduced by the reference at fixture-generation time
    assert abs(float(golden["A.param_sum"]) - 1124.303041338549) < 1e-9
    assert abs(golden["A.x"].astype(np.float64).sum() - 5.800098619190976) < 1e-9
    assert abs(golden["A.eps"].astype(np.float64).sum() - 215.32830626517534) < 1e-6
    np.testing.assert_allclose(golden["A.eps"][0, 0], [1.10350752, 1.83766949, 1.43554246, -1.34550977, 1.96559608], atol=2e-6)
    assert abs(golden["A.x_final"].astype(np.float64).sum() + 5.624957477208227) < 1e-6
    np.testing.assert_allclose(golden["A.x_final"][0, 1], [-0.23668964, 0.41739053, -0.07979971, 0.08068986, 0.11864171], atol=1e-6)
    assert abs(golden["A1.x_final"].astype(np.float64).sum() + 7.469549811212346) < 1e-6
    sc = golden["A1.scalars"]      # rows: t, at, at_next, c1, c2
    np.testing.assert_allclose(sc[0], [12, 0.99729908, 0.99989998, 0.0098143648, 0.0019221968], rtol=2e-6)
    np.testing.assert_allclose(sc[1][[0, 2, 3, 4]], [0, 1.0, 0.0, 0.0], atol=1e-12)


def test_schedules_and_alpha(golden):
    for kind in ["linear", "quad", "const", "jsd", "sigmoid"]:
        assert np.array_equal(O.beta_schedule(kind, 1e-4, 1e-3, 51), golden[f"betas_{kind}"])
    a = O.alpha_bar(betas(), t(golden, "alpha_bar_t"))
    assert np.array_equal(a.numpy(), golden["alpha_bar"])
    np.testing.assert_allclose(a.flatten().numpy()[:5], [1.0, 0.99989998, 0.99892235, 0.99729908, 0.99265730], rtol=1e-7)
    assert np.array_equal(O.adjacency().numpy(), golden["adj"])
    assert O.eval_sequence("uniform", 24, 2) == [0, 12] and O.eval_sequence("uniform", 12, 2) == [0, 6]
    assert O.eval_sequence("quad", 24, 2) == [0, 19]


@pytest.mark.parametrize("tag", sorted(DIFF_CASES))
def test_gcndiff_forward_and_sampler(golden, tag):
    cfg, adj, _, sd = build_diff(tag, golden)
    L, nh = cfg.model.num_layer, cfg.model.n_head
    x, mask = t(golden, f"{tag}.x"), mask_for(tag, golden)
    eps = O.gcndiff_forward(sd, adj, L, nh, x, mask, t(golden, f"{tag}.t"))
    assert torch.equal(eps, t(golden, f"{tag}.eps")), "oracle forward is not bit-identical to the reference"
    den = lambda xt, m, tt: O.gcndiff_forward(sd, adj, L, nh, xt, m, tt)
    xs, x0 = O.ddim_sample(x, mask, golden[f"{tag}.seq"].tolist(), den, betas(), eta=float(golden[f"{tag}.eta"]),
                           noise=t(golden, f"{tag}.noise"))
    assert torch.equal(xs[-1], t(golden, f"{tag}.x_final"))
    assert torch.equal(x0[-1], t(golden, f"{tag}.x0_last"))
    assert len(xs) == len(golden[f"{tag}.seq"]) + 1 and len(x0) == len(golden[f"{tag}.seq"])


@pytest.mark.parametrize("tag", sorted(POSE_CASES))
def test_gcnpose_forward(golden, tag):
    cfg, adj, _, sd = build_pose(tag, golden)
    xyz = O.gcnpose_forward(sd, adj, 5, 4, t(golden, f"{tag}.uv"), torch.ones(1, 1, 17, dtype=torch.bool))
    assert torch.equal(xyz, t(golden, f"{tag}.xyz"))


def test_metrics(golden):
    gt, pred = t(golden, "M.gt"), t(golden, "M.pred")
    gt, pred = O.root_centre(gt), O.root_centre(pred)
    assert abs(O.mpjpe(pred, gt).item() - float(golden["M.mpjpe"])) < 1e-7
    pp = O.p_mpjpe_per_pose(pred.numpy(), gt.numpy())
    np.testing.assert_allclose(pp, golden["M.p_mpjpe_per_pose"], atol=1e-12)
    assert abs(pp.mean() - float(golden["M.p_mpjpe"])) < 1e-6
    assert pp[3] < 1e-6          # a pure similarity transform aligns exactly


def test_sampler_properties():
    # eta = 0 makes every hypothesis identical; the mean over hypotheses is then the value
    torch.manual_seed(0)
    cfg = O.default_config(hid_dim=32, num_layer=1, n_head=2)
    import diffpose_nw_b200 as D
    adj = D.adj_mx_from_edges()
    sd = O.perturb_state_dict({k: v.detach() for k, v in D.FusedGCNdiff(adj, cfg).state_dict().items()})
    x = O.synthetic_poses(3)
    den = lambda xt, m, tt: O.gcndiff_forward(sd, adj, 1, 2, xt, m, tt)
    xr = x.repeat(4, 1, 1)
    out = O.ddim_sample(xr, None, [0, 12], den, betas(), eta=0.0)[0][-1]
    assert torch.equal(out[:3], out[3:6]) and torch.allclose(O.hypothesis_mean(out, 4), out[:3], atol=1e-7)
    # empty batch: the reference's attention `.view(nbatches, -1, h, d_k)` cannot infer -1 for 0 rows and raises;
    # the restatement keeps that behaviour (the product returns an empty tensor instead, see test_gpu_parity)
    with pytest.raises(RuntimeError):
        O.ddim_sample(x[:0], None, [0, 12], den, betas())


def test_tensor_core_emulation_is_an_fp16_perturbation_of_the_oracle():
    """oracle/tc_emulation.py restates the tensor-core engine's rounding points (fp16 operands, folded LayerNorm gains,
    time embedding inside GC2).  On CPU: it must be the
    oracle up to fp16-operand noise (1e-3 relative) -- both forms of the time embedding, masked keys and GCNpose; the two
    forms differ from each other by the same kind of noise (one extra rounding point), not more."""
    from oracle import tc_emulation as E
    import diffpose_nw_b200 as D
    adj = D.adj_mx_from_edges()
    torch.manual_seed(0)
    sd = O.perturb_state_dict({k: v.detach().clone() for k, v in D.FusedGCNdiff(adj, O.default_config()).state_dict().items()}, seed=5, scale=0.05)
    x = O.synthetic_poses(6, seed=4)
    tt = torch.tensor([0.0, 3.0, 12.0, 12.0, 37.0, 49.0])
    mask = torch.ones(1, 1, 17, dtype=torch.bool)
    mask[0, 0, 5] = mask[0, 0, 11] = False
    for m in (None, mask):
        ref = O.gcndiff_forward(sd, adj, 5, 4, x, m, tt)
        add = E.gcndiff_forward_tcg(sd, adj, 5, 4, x, m, tt, p16=True)
        fold = E.gcndiff_forward_tcg(sd, adj, 5, 4, x, m, tt, p16=True, temb_in_gc2=True)
        scale = ref.abs().max().item()
        assert (add - ref).abs().max().item() < 2e-3 * scale and (fold - ref).abs().max().item() < 2e-3 * scale
        assert (add - fold).abs().max().item() < 2e-3 * scale
    torch.manual_seed(1)
    sdp = O.perturb_state_dict({k: v.detach().clone() for k, v in D.FusedGCNpose(adj, O.default_config(coords_dim=[2, 3])).state_dict().items()}, seed=6, scale=0.05)
    uv = x[:, :, :2].contiguous()
    ref = O.gcnpose_forward(sdp, adj, 5, 4, uv, None)
    assert (E.gcnpose_forward_tcg(sdp, adj, 5, 4, uv, None) - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
