"""GPU (-m gpu): the tcgen05 engine.  First the isolated 128x96x96 tensor-core product (layouts, descriptors, TMA
staging, TMEM read-back), then the whole sampler against (a) the CPU emulation of its fp16 rounding points -- tight
tolerance, catches indexing bugs -- and (b) the fp32 oracle at the north_star tolerance."""
import numpy as np
import pytest
import torch

import diffpose_nw_b200 as D
from diffpose_nw_b200 import _lib
from oracle import diffpose_oracle as O
from oracle import tc_emulation as E
from _cases import betas, build_diff, mask_for, t

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def dev():
    return torch.device("cuda:0")


EMU = {"tc": E.gcndiff_forward_tc, "tcg": lambda *a: E.gcndiff_forward_tcg(*a, p16=True, temb_in_gc2=True)}   # the sampler's form
ENGINE_ID = {"tc": 2, "tcg": 3}


@pytest.mark.parametrize("engine", ["tc", "tcg"])
@pytest.mark.parametrize("tag", ["A", "A1", "B", "C", "D"])
def test_tc_engine_vs_emulation_and_oracle(golden, tag, engine):
    cfg, adj, model, sd = build_diff(tag, golden)
    model = model.to(dev()).set_engine(engine)
    assert model.engine() == engine
    x, mask = t(golden, f"{tag}.x"), mask_for(tag, golden)
    seq, eta = golden[f"{tag}.seq"].tolist(), float(golden[f"{tag}.eta"])
    noise = t(golden, f"{tag}.noise")
    out = D.generalized_steps(x.to(dev()), mask.to(dev()), seq, model, betas(), eta=eta, noise=noise.to(dev()))[0][-1].cpu()
    assert model.last_launch()[4] == ENGINE_ID[engine]
    emu = O.ddim_sample(x, mask, seq, lambda a, m, tt: EMU[engine](sd, adj, 5, 4, a, m, tt), betas(), eta=eta, noise=noise)[0][-1]
    ref = t(golden, f"{tag}.x_final")
    e_emu = (out - emu).abs().max().item()
    e_ref = (out - ref).abs().max().item()
    amp = (emu - ref).abs().max().item()
    print(f"{engine} {tag}: T={len(seq)} |tc-emu|={e_emu:.2e} |tc-ref|={e_ref:.2e} |emu-ref|={amp:.2e}")
    # vs the emulation: only accumulation order, reciprocal-vs-division and fp16 double rounding differ; those are
    # amplified by the sampler dynamics exactly like the operand rounding itself (amp), an indexing bug is not
    assert e_emu < max(1e-4, 1.0 * amp), f"{tag}: kernel disagrees with its rounding-point emulation ({e_emu:.3e})"
    # vs the reference: north_star tolerance; case D (50 steps on perturbed, strongly amplifying weights) is reported
    # against a 1e-2 bound because the fp16/TF32 operand precision itself (amp) exceeds 1e-3 there
    assert e_ref < (1e-2 if tag == "D" else 1e-3)


@pytest.mark.parametrize("engine", ["tc", "tcg"])
def test_tc_default_init_50_steps(engine):
    """BASELINE configs[3] schedule (T = 50, H = 2, eta = 1) on default-init weights: within 1e-3 of the fp32 oracle."""
    cfg = O.default_config()
    adj = D.adj_mx_from_edges()
    torch.manual_seed(0)
    model = D.FusedGCNdiff(adj, cfg)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(dev()).set_engine(engine)
    B, Hh, seq = 16, 2, list(range(50))
    x = O.synthetic_poses(B, seed=3)
    g = torch.Generator().manual_seed(11)
    noise = torch.randn(50, Hh * B, 17, 5, generator=g)
    xr = x.repeat(Hh, 1, 1)
    ref = O.ddim_sample(xr, None, seq, lambda a, m, tt: O.gcndiff_forward(sd, adj, 5, 4, a, m, tt), betas(), eta=1.0, noise=noise)[0][-1]
    out = D.sample(model, xr.to(dev()), None, seq, betas(), eta=1.0, noise=noise.to(dev()), n_hyp=Hh).cpu()
    err = (out - ref).abs().max().item()
    tgt = O.synthetic_targets(x)
    m_ref = O.mpjpe(O.root_centre(O.hypothesis_mean(ref, Hh)[:, :, 2:]), tgt).item() * 1000
    m_out = O.mpjpe(O.root_centre(O.hypothesis_mean(out, Hh)[:, :, 2:]), tgt).item() * 1000
    print(f"T=50 default-init: max|dx|={err:.2e} dMPJPE={abs(m_ref - m_out):.4f} mm")
    assert err < 1e-3 and abs(m_ref - m_out) < 0.05


@pytest.mark.parametrize("engine", ["tc", "tcg"])
def test_tc_matches_fp32_engine_on_many_tiles(engine):
    """Multi-tile / ragged last tile / persistent loop: 1500 poses on both engines."""
    cfg = O.default_config()
    torch.manual_seed(0)
    model = D.FusedGCNdiff(D.adj_mx_from_edges(), cfg).to(dev())
    x = O.synthetic_poses(1500, seed=12).to(dev())
    a = D.generalized_steps(x, None, [0, 12], model.set_engine("fp32"), betas())[0][-1]
    b = D.generalized_steps(x, None, [0, 12], model.set_engine(engine), betas())[0][-1]
    assert (a - b).abs().max().item() < 1e-3
    assert torch.equal(b, D.generalized_steps(x, None, [0, 12], model, betas())[0][-1])


def test_tcg_repeatable_across_ring_wraps():
    """Race hunting without a sanitizer: several tiles per CTA (every mbarrier ring wraps many times), three steps, noise and
    a key mask; 12 repetitions must be bit-identical and match the fp32 engine within the operand-precision tolerance."""
    cfg = O.default_config()
    torch.manual_seed(0)
    model = D.FusedGCNdiff(D.adj_mx_from_edges(), cfg).to(dev()).eval()
    n, Hh, seq = 3100, 2, [0, 6, 12]
    x = O.synthetic_poses(n, seed=21).to(dev())
    g = torch.Generator().manual_seed(22)
    noise = torch.randn(len(seq), Hh * n, 17, 5, generator=g).to(dev())
    mask = torch.ones(1, 1, 17, dtype=torch.bool, device=dev())
    mask[0, 0, 3] = False
    mask[0, 0, 11] = False
    run = lambda: D.sample(model, x, mask, seq, betas(), eta=1.0, noise=noise, n_hyp=Hh, repeat_input=True)
    first = run()
    assert torch.isfinite(first).all()
    for _ in range(11):
        assert torch.equal(run(), first)
    ref = D.sample(model.set_engine("fp32"), x, mask, seq, betas(), eta=1.0, noise=noise, n_hyp=Hh, repeat_input=True)
    assert (first - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("tag", ["A", "A1", "B", "C"])
def test_tcg_forward_per_sample_t(golden, tag):
    """GCNdiff.forward (per-sample timesteps, partial key mask in B) on the tensor-core engine: close to its rounding-point
    emulation, and to the fp32 reference within the operand precision."""
    from _cases import build_diff
    cfg, adj, model, sd = build_diff(tag, golden)
    model = model.to(dev()).set_engine("tcg")
    x, mask, tt = t(golden, f"{tag}.x"), mask_for(tag, golden), t(golden, f"{tag}.t")
    eps = model(x.to(dev()), mask.to(dev()), tt.to(dev()), 0).cpu()
    assert model.last_launch()[4] == 3
    emu = E.gcndiff_forward_tcg(sd, adj, 5, 4, x, mask, tt, p16=True)
    ref = t(golden, f"{tag}.eps")
    scale = max(1.0, ref.abs().max().item())
    e_emu, e_ref, amp = (eps - emu).abs().max().item(), (eps - ref).abs().max().item(), (emu - ref).abs().max().item()
    print(f"forward {tag}: |tcg-emu|={e_emu:.2e} |tcg-ref|={e_ref:.2e} |emu-ref|={amp:.2e} scale={scale:.2f}")
    assert e_emu < max(2e-3 * scale, amp)
    assert e_ref < (3e-3 if tag in ("A", "A1") else 3e-2) * scale


@pytest.mark.parametrize("tag", ["P0", "P1"])
def test_tcg_gcnpose(golden, tag):
    """GCNpose (uv -> xyz, no time embedding) on the tensor-core engine."""
    from _cases import build_pose
    cfg, adj, model, sd = build_pose(tag, golden)
    model = model.to(dev())
    uv = t(golden, f"{tag}.uv")
    mask = torch.ones(1, 1, 17, dtype=torch.bool)
    xyz = model(uv.to(dev()), mask.to(dev())).cpu()
    assert model.engine() == "tcg" and model.last_launch()[4] == 3
    emu = E.gcnpose_forward_tcg(sd, adj, 5, 4, uv, mask)
    ref = t(golden, f"{tag}.xyz")
    scale = max(1.0, ref.abs().max().item())
    e_emu, e_ref, amp = (xyz - emu).abs().max().item(), (xyz - ref).abs().max().item(), (emu - ref).abs().max().item()
    print(f"gcnpose {tag}: |tcg-emu|={e_emu:.2e} |tcg-ref|={e_ref:.2e} |emu-ref|={amp:.2e} scale={scale:.2f}")
    assert e_emu < max(2e-3 * scale, amp)
    assert e_ref < (3e-3 if tag == "P0" else 3e-2) * scale
    # the fp32 engine stays available and exact
    xyz32 = model.set_engine("fp32")(uv.to(dev()), mask.to(dev())).cpu()
    assert (xyz32 - ref).abs().max().item() < 2e-5 * scale


@pytest.mark.parametrize("n_layer", [1, 3])
def test_tcg_other_depths(n_layer):
    """config.model.num_layer is a runtime parameter of the tensor-core engine (weight ring, parameter ring and the issuer's
    program all loop over it): 1 and 3 layers, every parameter perturbed, against the fp32 oracle and the emulation."""
    cfg = O.default_config(num_layer=n_layer)
    adj = D.adj_mx_from_edges()
    torch.manual_seed(5)
    model = D.FusedGCNdiff(adj, cfg)
    sd = O.perturb_state_dict({k: v.detach().clone() for k, v in model.state_dict().items()}, seed=13)
    model.load_state_dict(sd)
    model = model.to(dev())
    assert model.engine() == "tcg"
    x = O.synthetic_poses(40, seed=17)
    seq = [0, 8, 16]
    g = torch.Generator().manual_seed(19)
    noise = torch.randn(3, 40, 17, 5, generator=g)
    ref = O.ddim_sample(x, None, seq, lambda a, m, tt: O.gcndiff_forward(sd, adj, n_layer, 4, a, m, tt), betas(), eta=1.0, noise=noise)[0][-1]
    emu = O.ddim_sample(x, None, seq, lambda a, m, tt: E.gcndiff_forward_tcg(sd, adj, n_layer, 4, a, m, tt, p16=True, temb_in_gc2=True), betas(), eta=1.0, noise=noise)[0][-1]
    out = D.generalized_steps(x.to(dev()), None, seq, model, betas(), eta=1.0, noise=noise.to(dev()))[0][-1].cpu()
    amp = (emu - ref).abs().max().item()
    assert (out - emu).abs().max().item() < max(1e-4, amp)
    assert (out - ref).abs().max().item() < 1e-3


def test_tcg_denser_graph():
    """Nothing in the engine is specific to the H36M tree: a hub joined to every joint makes every row of T2 dense (the
    5-wide input / output convolutions read a dense (T1, T2) table, the Chebyshev slab carries whatever the row sums
    are, the integerised rows fall back to fp16 where no common denominator exists).  Against the fp32 oracle and the
    fp32 engine."""
    edges = list(D.H36M_EDGES) + [(0, j) for j in range(1, 17)]
    adj = D.adj_mx_from_edges(17, edges)
    assert ((O.cheb_basis(adj)[2] != 0).sum(1) > 9).any()
    torch.manual_seed(3)
    model = D.FusedGCNdiff(adj, O.default_config())
    sd = O.perturb_state_dict({k: v.detach().clone() for k, v in model.state_dict().items()}, seed=14)
    model.load_state_dict(sd)
    model = model.to(dev())
    x = O.synthetic_poses(30, seed=18)
    seq = [0, 12]
    g = torch.Generator().manual_seed(20)
    noise = torch.randn(2, 30, 17, 5, generator=g)
    ref = O.ddim_sample(x, None, seq, lambda a, m, tt: O.gcndiff_forward(sd, adj, 5, 4, a, m, tt), betas(), eta=1.0, noise=noise)[0][-1]
    out = D.generalized_steps(x.to(dev()), None, seq, model, betas(), eta=1.0, noise=noise.to(dev()))[0][-1].cpu()
    assert model.last_launch()[4] == ENGINE_ID["tcg"]
    assert (out - ref).abs().max().item() < 1e-3
    out32 = D.generalized_steps(x.to(dev()), None, seq, model.set_engine("fp32"), betas(), eta=1.0, noise=noise.to(dev()))[0][-1].cpu()
    assert (out32 - ref).abs().max().item() < 2e-5


def test_tcg_long_schedule_steps_on_device():
    """More than 64 DDIM steps: the step scalars no longer travel by value but through a device array."""
    cfg = O.default_config()
    torch.manual_seed(0)
    model = D.FusedGCNdiff(D.adj_mx_from_edges(), cfg).to(dev()).eval()
    b = torch.from_numpy(O.beta_schedule("linear", 1e-4, 1e-3, 100)).float()
    seq = list(range(0, 70))
    x = O.synthetic_poses(20, seed=33).to(dev())
    g = torch.Generator().manual_seed(34)
    noise = torch.randn(len(seq), 20, 17, 5, generator=g).to(dev())
    out = D.generalized_steps(x, None, seq, model, b, eta=1.0, noise=noise)[0][-1]
    ref = D.generalized_steps(x, None, seq, model.set_engine("fp32"), b, eta=1.0, noise=noise)[0][-1]
    assert torch.isfinite(out).all() and (out - ref).abs().max().item() < 1e-3


def test_trace_diagnostic_is_monotonic_and_optional():
    """dp_set_trace: the traced kernel variant stamps every hand-over in program order; results are unchanged and the
    production variant is used again once the buffer is cleared."""
    cfg = O.default_config()
    torch.manual_seed(0)
    model = D.FusedGCNdiff(D.adj_mx_from_edges(), cfg).to(dev()).eval()
    x = O.synthetic_poses(14, seed=50).to(dev())
    plain = D.generalized_steps(x, None, [0, 12], model, betas())[0][-1]
    buf = torch.zeros(4096, dtype=torch.int64, device=dev())
    _lib.check(_lib.load().dp_set_trace(model._handle, buf.data_ptr(), buf.numel()), "dp_set_trace")
    traced = D.generalized_steps(x, None, [0, 12], model, betas())[0][-1]
    torch.cuda.synchronize()
    _lib.check(_lib.load().dp_set_trace(model._handle, None, 0), "dp_set_trace")
    assert torch.equal(plain, traced)
    raw = buf.cpu().numpy()[:2048]
    raw = raw[raw > 0]
    t, kind = raw >> 1, raw & 1
    assert len(t) == 2 * (2 + 5 * 35 + 2)               # per step: input conv, 5 layers x 35 hand-over stamps, output conv
    assert (np.diff(t) > 0).all() and kind[0] == 0 and kind[1] == 1
    again = D.generalized_steps(x, None, [0, 12], model, betas())[0][-1]
    assert torch.equal(plain, again)
