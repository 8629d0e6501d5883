"""GPU (-m gpu): the two tcgen05 engines.
  tcg (sampler default, fp16 operands): against (a) the CPU emulation of its fp16 rounding points -- tight tolerance,
      catches indexing bugs -- and (b) the fp32 oracle at the north_star tolerance (1e-3 abs, 0.05 mm), every case.
  tcx (forward / lifter default, split-precision operands): against the fp32 oracle at fp32-level tolerance."""
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


# tcg: the sampler's form of the rounding-point emulation; tcx has no rounding points above fp32 level: its emulation IS the oracle
EMU = {"tcx": lambda sd, adj, nl, nh, a, m, tt: O.gcndiff_forward(sd, adj, nl, nh, a, m, tt),
       "tcg": lambda *a: E.gcndiff_forward_tcg(*a, p16=True, temb_in_gc2=True)}
ENGINE_ID = {"tcx": 2, "tcg": 3}
TCX_TOL = 5e-5      # split-precision engine vs the fp32 oracle (|x| ~ 1; accumulation-order level)


@pytest.mark.parametrize("engine", ["tcx", "tcg"])
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
    # vs the reference: the north_star tolerance (1e-3 abs) on every case, the 50-step perturbed-weights case D included;
    # the split-precision engine at fp32 level (case D: 50 amplifying steps)
    assert e_ref < (1e-3 if engine == "tcg" else (4 * TCX_TOL if tag == "D" else TCX_TOL))


@pytest.mark.parametrize("engine", ["tcx", "tcg"])
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
    print(f"T=50 default-init / {engine}: max|dx|={err:.2e} dMPJPE={abs(m_ref - m_out):.4f} mm")
    assert err < (1e-3 if engine == "tcg" else TCX_TOL) and abs(m_ref - m_out) < 0.05


@pytest.mark.parametrize("engine", ["tcx", "tcg"])
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


@pytest.mark.parametrize("tag", ["A", "A1", "B", "C", "D"])
def test_forward_default_engine_is_split_precision(golden, tag):
    """GCNdiff.forward (per-sample timesteps, partial key mask in B) with the DEFAULT engine setting: forward calls run on
    the split-precision tensor-core engine and match the reference's eps at fp32 level."""
    from _cases import build_diff
    cfg, adj, model, sd = build_diff(tag, golden)
    model = model.to(dev())
    assert model.forward_engine() == "tcx" and model.engine() == "tcg"
    x, mask, tt = t(golden, f"{tag}.x"), mask_for(tag, golden), t(golden, f"{tag}.t")
    eps = model(x.to(dev()), mask.to(dev()), tt.to(dev()), 0).cpu()
    assert model.last_launch()[4] == 2
    ref = t(golden, f"{tag}.eps")
    scale = max(1.0, ref.abs().max().item())
    err = (eps - ref).abs().max().item()
    print(f"forward {tag} / tcx: |eps - ref|={err:.2e} scale={scale:.2f}")
    assert err < 2e-5 * scale


@pytest.mark.parametrize("tag", ["A", "A1", "B", "C"])
def test_tcg_forward_per_sample_t(golden, tag):
    """GCNdiff.forward (per-sample timesteps, partial key mask in B) on the fp16-operand engine when it is selected
    explicitly (the default for forward calls is tcx, above): close to its rounding-point emulation; against the fp32
    reference only within the undamped operand precision (that is why it is not the default for forward calls)."""
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
def test_gcnpose_default_engine(golden, tag):
    """GCNpose (uv -> xyz) with the default engine setting: the split-precision tensor-core engine, within 1e-3 abs of the
    reference (north_star; measured: fp32 level) -- and the fused lift (root-centre + concat, one launch)."""
    from _cases import build_pose
    cfg, adj, model, sd = build_pose(tag, golden)
    model = model.to(dev())
    uv = t(golden, f"{tag}.uv")
    mask = torch.ones(1, 1, 17, dtype=torch.bool)
    xyz = model(uv.to(dev()), mask.to(dev())).cpu()
    assert model.forward_engine() == "tcx" and model.last_launch()[4] == 2
    ref = t(golden, f"{tag}.xyz")
    scale = max(1.0, ref.abs().max().item())
    err = (xyz - ref).abs().max().item()
    print(f"gcnpose {tag} / tcx: |xyz - ref|={err:.2e} scale={scale:.2f}")
    assert err < 1e-3 and err < 2e-5 * scale
    l0 = D._lib.launch_count()
    u5 = model.lift(uv.to(dev()), mask.to(dev())).cpu()
    assert D._lib.launch_count() - l0 == 1
    want = torch.cat([uv, O.root_centre(ref)], dim=2)
    assert u5.shape == want.shape and (u5 - want).abs().max().item() < 4e-5 * scale
    assert torch.equal(u5[:, :, :2], uv) and (u5[:, 0, 2:] == 0).all()
    for eng in ("fp32", "tcg"):       # the other engines reach the same result through a glue kernel
        v5 = model.set_engine(eng).lift(uv.to(dev()), mask.to(dev())).cpu()
        assert (v5 - want).abs().max().item() < (4e-5 if eng == "fp32" else 3e-2) * scale


@pytest.mark.parametrize("tag", ["P0", "P1"])
def test_tcg_gcnpose(golden, tag):
    """GCNpose on the fp16-operand engine, selected explicitly (not the default for the lifter: its output is the xyz
    handed to the sampler, undamped -- oracle/tc_emulation.py)."""
    from _cases import build_pose
    cfg, adj, model, sd = build_pose(tag, golden)
    model = model.to(dev()).set_engine("tcg")
    uv = t(golden, f"{tag}.uv")
    mask = torch.ones(1, 1, 17, dtype=torch.bool)
    xyz = model(uv.to(dev()), mask.to(dev())).cpu()
    assert model.forward_engine() == "tcg" and model.last_launch()[4] == 3
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


def test_tcx_denser_graph():
    """The split-precision engine walks per-joint neighbour lists for T1 / T2: they must cover ANY adjacency, not only the
    <= 9 two-hop neighbours of the H36M tree.  A hub joined to every joint (T2 rows with 17 entries): forward (GCNdiff and
    the GCNpose lifter) and sampler against the oracle."""
    edges = list(D.H36M_EDGES) + [(0, j) for j in range(1, 17)]
    adj = D.adj_mx_from_edges(17, edges)
    assert ((O.cheb_basis(adj)[2] != 0).sum(1) > 9).any()
    torch.manual_seed(3)
    model = D.FusedGCNdiff(adj, O.default_config())
    sd = O.perturb_state_dict({k: v.detach().clone() for k, v in model.state_dict().items()}, seed=14)
    model.load_state_dict(sd)
    model = model.to(dev())
    x = O.synthetic_poses(30, seed=18)
    tt = torch.linspace(0, 40, 30)
    eps = model(x.to(dev()), None, tt.to(dev()), 0).cpu()
    assert model.last_launch()[4] == ENGINE_ID["tcx"]
    ref = O.gcndiff_forward(sd, adj, 5, 4, x, None, tt)
    assert (eps - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    seq = [0, 12]
    ref_s = O.ddim_sample(x, None, seq, lambda a, m, t_: O.gcndiff_forward(sd, adj, 5, 4, a, m, t_), betas())[0][-1]
    out = D.generalized_steps(x.to(dev()), None, seq, model.set_engine("tcx"), betas())[0][-1].cpu()
    assert (out - ref_s).abs().max().item() < TCX_TOL
    torch.manual_seed(4)
    pose = D.FusedGCNpose(adj, O.default_config(coords_dim=[2, 3]))
    sdp = {k: v.detach().clone() for k, v in pose.state_dict().items()}
    uv = x[:, :, :2].contiguous()
    xyz = pose.to(dev())(uv.to(dev()), None).cpu()
    refp = O.gcnpose_forward(sdp, adj, 5, 4, uv, None)
    assert (xyz - refp).abs().max().item() < 2e-5 * max(1.0, refp.abs().max().item())


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


def test_fp16_operand_range_saturates_instead_of_nan():
    """fp16 operand range of the tensor-core engines (documented limit): an activation beyond +-65504 (tcg)
    or +-131008 (tcx: hi + lo) -- possible with a trained checkpoint, impossible to rule out without one -- SATURATES in
    the fp32->fp16 conversion (cvt.rn.satfinite.f16x2.f32) instead of becoming inf and then inf - inf = NaN inside an MMA.
    Below the limit the engines keep their relative accuracy; above it the result is clipped but finite (the fp32 engine
    has no such limit)."""
    adj = D.adj_mx_from_edges()
    mask = torch.ones(1, 1, 17, dtype=torch.bool, device=dev())
    x = O.synthetic_poses(40, seed=5).to(dev())
    tt = torch.full((40,), 3.0, device=dev())
    for gain, overflow in ((2.0e4, False), (4.0e5, True)):
        torch.manual_seed(0)
        model = D.FusedGCNdiff(adj, O.default_config())
        with torch.no_grad():
            model.gconv_input.weight.mul_(gain)          # residual stream ~ gain: the Chebyshev blocks read it as an fp16 operand
        model = model.to(dev())
        ref = model.set_engine("fp32")(x, mask, tt, 0)
        big = model.set_engine("tcx")(x, mask, tt, 0)
        eps = model.set_engine("tcg")(x, mask, tt, 0)
        assert model.last_launch()[4] == 3
        scale = ref.abs().max().item()
        print(f"gain {gain:g}: |eps|max={scale:.3g} tcg rel err={(eps - ref).abs().max().item() / scale:.2e} tcx rel err={(big - ref).abs().max().item() / scale:.2e}")
        assert torch.isfinite(ref).all() and torch.isfinite(eps).all() and torch.isfinite(big).all()
        if not overflow:
            assert (eps - ref).abs().max().item() < 5e-3 * scale
            assert (big - ref).abs().max().item() < 2e-5 * scale
        out = D.generalized_steps(x, mask, [0, 12], model, betas())[0][-1]      # the sampler on the saturating engine: finite
        assert torch.isfinite(out).all()


def test_check_weights_sees_data_writes():
    """`param.data.copy_()` (the reference's EMAHelper.ema idiom, models/ema.py:27-29) bumps no version counter: the
    packed copy goes stale silently unless repack() is called.  check_weights() detects it; the package's EMAHelper
    repacks by itself."""
    torch.manual_seed(0)
    model = D.FusedGCNdiff(D.adj_mx_from_edges(), O.default_config()).to(dev()).eval()
    x = O.synthetic_poses(32, seed=2).to(dev())
    a = D.generalized_steps(x, None, [0, 12], model, betas())[0][-1]
    assert model.check_weights()
    ema = D.EMAHelper(mu=0.9)
    ema.register(model)
    for k in ema.shadow:
        ema.shadow[k] = ema.shadow[k] * 0.9
    w = model.atten_layers[2].self_attn.linears[1].weight
    w.data.copy_(w.data * 0.5)                          # invisible to the fingerprint ...
    b = D.generalized_steps(x, None, [0, 12], model, betas())[0][-1]
    assert torch.equal(a, b)                            # ... so the stale copy is still in use (documented behaviour)
    assert not model.check_weights()                    # detected, and repacked
    c = D.generalized_steps(x, None, [0, 12], model, betas())[0][-1]
    assert (c - a).abs().max().item() > 1e-6 and model.check_weights()
    ema.ema(model)                                      # the drop-in helper invalidates the packed copy itself
    d = D.generalized_steps(x, None, [0, 12], model, betas())[0][-1]
    assert (d - c).abs().max().item() > 1e-6 and model.check_weights()
    # a deep copy owns its own native handle
    import copy
    m2 = copy.deepcopy(model)
    e = D.generalized_steps(x, None, [0, 12], m2, betas())[0][-1]
    assert torch.equal(d, e) and m2._handle is not model._handle and m2._handle.value != model._handle.value
