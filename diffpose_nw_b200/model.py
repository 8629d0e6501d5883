"""Drop-in modules for the reference's frame-based denoiser and stage-1 lifter.

`FusedGCNdiff(adj, config)` replaces `GCNdiff(adj, config)` (reference models/gcndiff.py:55-113) and
`FusedGCNpose(adj, config)` replaces `GCNpose(adj, config)` (reference models/gcnpose.py:55-113):

* same constructor arguments and config keys (`config.model.{hid_dim, emd_dim, coords_dim, num_layer, n_head,
  dropout, n_pts}`), `emd_dim` overridden by `4*hid_dim` exactly as the reference does (gcndiff.py:68);
* same `state_dict` keys and shapes, so `load_state_dict(states[0])` of a reference
  checkpoint works unchanged, with or without DataParallel's `module.` prefix;
* same default initialisation *and the same RNG draw order* as the reference constructor, so
  `torch.manual_seed(s); FusedGCNdiff(adj, cfg)` holds bit-identical weights to `torch.manual_seed(s);
  GCNdiff(adj, cfg)` (the reference deep-copies one attention block and one GraphNet into every layer,
  gcndiff.py:78-85; that is reproduced);
* same forward signatures `model(x, mask, t, cemd)` / `model(x, mask)`.

The modules only hold parameters; all arithmetic happens in libdiffpose_b200.so on the GPU.  There is no CPU or
eager-PyTorch fallback: calling forward with CPU tensors raises.
"""
from __future__ import annotations

import ctypes
import math

import torch
import torch.nn as nn

from . import _lib


class _ChebParams(nn.Module):
    """Parameters of one ChebConv(in_c, out_c, K=2) (reference models/ChebConv.py:59-70)."""

    def __init__(self, in_c, out_c):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(3, 1, in_c, out_c))
        nn.init.xavier_normal_(self.weight)
        self.bias = nn.Parameter(torch.zeros(1, 1, out_c))


class _GraphConvParams(nn.Module):
    def __init__(self, in_c, out_c):
        super().__init__()
        self.gconv = _ChebParams(in_c, out_c)


class _ResChebParams(nn.Module):
    """gconv_layers.l: two graph convolutions (+ temb_proj for the diffusion denoiser)."""

    def __init__(self, hid, emd, with_temb):
        super().__init__()
        self.gconv1 = _GraphConvParams(hid, hid)
        self.gconv2 = _GraphConvParams(hid, hid)
        if with_temb:
            self.temb_proj = nn.Linear(emd, hid)


class _AttnParams(nn.Module):
    def __init__(self, hid):
        super().__init__()
        first = nn.Linear(hid, hid)  # the reference clones ONE Linear four times (GraFormer.py:123)
        self.linears = nn.ModuleList([first] + [_clone(first) for _ in range(3)])


class _LamParams(nn.Module):
    def __init__(self, in_f, out_f):
        super().__init__()
        self.fc = nn.Linear(in_f, out_f)


class _GraphNetParams(nn.Module):
    def __init__(self, hid, n_pts):
        super().__init__()
        self.A_hat = nn.Parameter(torch.eye(n_pts))
        self.gconv1 = _LamParams(hid, 2 * hid)
        self.gconv2 = _LamParams(2 * hid, hid)


class _NormParams(nn.Module):
    def __init__(self, hid):
        super().__init__()
        self.a_2 = nn.Parameter(torch.ones(hid))
        self.b_2 = nn.Parameter(torch.zeros(hid))


class _SublayerParams(nn.Module):
    def __init__(self, hid):
        super().__init__()
        self.norm = _NormParams(hid)


class _AttenLayerParams(nn.Module):
    def __init__(self, hid, attn, ffn):
        super().__init__()
        self.self_attn = attn
        self.feed_forward = ffn
        self.sublayer = nn.ModuleList([_SublayerParams(hid), _SublayerParams(hid)])


def _clone(module):
    import copy
    return copy.deepcopy(module)


def _param_order(n_layer, has_temb):
    """Canonical flattening order documented in include/diffpose_b200.h (dp_pack)."""
    keys = ["gconv_input.weight", "gconv_input.bias"]
    for l in range(n_layer):
        g, a = f"gconv_layers.{l}", f"atten_layers.{l}"
        keys += [f"{g}.gconv1.gconv.weight", f"{g}.gconv1.gconv.bias", f"{g}.gconv2.gconv.weight", f"{g}.gconv2.gconv.bias"]
        if has_temb:
            keys += [f"{g}.temb_proj.weight", f"{g}.temb_proj.bias"]
        for i in range(4):
            keys += [f"{a}.self_attn.linears.{i}.weight", f"{a}.self_attn.linears.{i}.bias"]
        keys += [f"{a}.feed_forward.A_hat", f"{a}.feed_forward.gconv1.fc.weight", f"{a}.feed_forward.gconv1.fc.bias",
                 f"{a}.feed_forward.gconv2.fc.weight", f"{a}.feed_forward.gconv2.fc.bias",
                 f"{a}.sublayer.0.norm.a_2", f"{a}.sublayer.0.norm.b_2", f"{a}.sublayer.1.norm.a_2", f"{a}.sublayer.1.norm.b_2"]
    keys += ["gconv_output.weight", "gconv_output.bias"]
    if has_temb:
        keys += ["temb.dense.0.weight", "temb.dense.0.bias", "temb.dense.1.weight", "temb.dense.1.bias"]
    return keys


class _FusedBase(nn.Module):
    _has_temb = True

    def __init__(self, adj, config):
        super().__init__()
        self.adj = adj                      # plain attribute, not a buffer (reference gcndiff.py:59)
        self.config = config
        m = config.model
        self.hid_dim, self.coords_dim = m.hid_dim, list(m.coords_dim)
        self.emd_dim = self.hid_dim * 4     # YAML emd_dim is ignored by the reference (gcndiff.py:68)
        self.n_layers, self.n_head, self.n_pts = m.num_layer, m.n_head, m.n_pts
        hid, c_in, c_out = self.hid_dim, self.coords_dim[0], self._out_dim()

        # --- parameters, created in the reference's order so the RNG stream matches (gcndiff.py:73-98)
        gconv_input = _ChebParams(c_in, hid)
        attn = _AttnParams(hid)
        ffn = _GraphNetParams(hid, self.n_pts)
        g_layers, a_layers = [], []
        for _ in range(self.n_layers):
            g_layers.append(_ResChebParams(hid, self.emd_dim, self._has_temb))
            a_layers.append(_AttenLayerParams(hid, _clone(attn), _clone(ffn)))
        self.gconv_input = gconv_input
        self.gconv_layers = nn.ModuleList(g_layers)
        self.atten_layers = nn.ModuleList(a_layers)
        self.gconv_output = _ChebParams(hid, c_out)
        self.temb = nn.Module()             # GCNpose carries an unused temb.dense as well (gcnpose.py:94-98)
        self.temb.dense = nn.ModuleList([nn.Linear(hid, self.emd_dim), nn.Linear(self.emd_dim, self.emd_dim)])

        self._c_in, self._c_out = c_in, c_out
        self._handle = None
        self._packed_version = None
        self._param_list = None
        self._packed_sum = None
        self._calls = 0
        self._engine = _lib.ENGINE_AUTO

    def _out_dim(self):
        return self.coords_dim[1]

    # ---------------------------------------------------------------- checkpoint compatibility
    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts reference checkpoints saved from a DataParallel wrapper (`module.` prefix)."""
        if any(k.startswith("module.") for k in state_dict):
            state_dict = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._packed_version = None
        self._param_list = None
        return out

    def load_checkpoint(self, path_or_states, use_ema=False, map_location="cpu"):
        """Load a reference checkpoint: the list `[model_state_dict, optimizer_state, epoch, step, (ema_shadow)]` that
        `runners/diffpose_frame.py:247-258` saves with `torch.save` (keys carry a `module.` prefix because the runner wraps
        the model in DataParallel).  The reference's evaluation loads `states[0]` (`:131-132`); `use_ema=True` loads the EMA
        shadow `states[4]` (`models/ema.py:45-49`: name -> tensor without the prefix) on top of it instead.
        Returns (epoch, step)."""
        states = path_or_states
        if isinstance(states, (str, bytes)) or hasattr(states, "__fspath__"):
            states = torch.load(states, map_location=map_location)
        if isinstance(states, dict):          # a bare state_dict is accepted too
            self.load_state_dict(states)
            return None, None
        if not isinstance(states, (list, tuple)) or len(states) < 1:
            raise RuntimeError("load_checkpoint: expected the reference's list [state_dict, optimizer, epoch, step, (ema)]")
        self.load_state_dict(states[0])
        if use_ema:
            if len(states) < 5:
                raise RuntimeError("load_checkpoint: this checkpoint carries no EMA shadow (config.model.ema was off)")
            shadow = {(k[7:] if k.startswith("module.") else k): v for k, v in states[4].items()}
            own = dict(self.named_parameters())
            missing = [k for k in shadow if k not in own]
            if missing:
                raise RuntimeError(f"load_checkpoint: EMA shadow has unknown parameters {missing[:3]}")
            with torch.no_grad():
                for k, v in shadow.items():
                    own[k].copy_(v)
            self.repack()
        epoch = states[2] if len(states) > 2 else None
        step = states[3] if len(states) > 3 else None
        return epoch, step

    def set_engine(self, engine):
        """'auto' | 'fp32' | 'tcx' | 'tcg' -- which kernel family runs the model (see include/diffpose_b200.h).
        'auto': the sampler runs on 'tcg' (fp16-operand tensor cores; its rounding is damped by the DDIM schedule) and
        plain forward calls -- GCNdiff.forward, the GCNpose lifter -- on 'tcx' (split-precision tensor cores, fp32-level)."""
        self._engine = {"auto": _lib.ENGINE_AUTO, "fp32": _lib.ENGINE_FP32, "tcx": _lib.ENGINE_TCX, "tcg": _lib.ENGINE_TCG}[engine]
        if self._handle is not None:
            _lib.check(_lib.load().dp_set_engine(self._handle, self._engine), "dp_set_engine")
        return self

    def engine(self):
        """Engine the SAMPLER (dp_sample) uses with the current setting."""
        self._ensure_packed(self._device())
        return _lib.ENGINE_NAMES[_lib.load().dp_get_engine(self._handle)]

    def forward_engine(self):
        """Engine a forward call (dp_forward / dp_lift) uses with the current setting."""
        self._ensure_packed(self._device())
        return _lib.ENGINE_NAMES[_lib.load().dp_get_forward_engine(self._handle)]

    # ---------------------------------------------------------------- device state
    def _device(self):
        return self.gconv_input.weight.device

    def _weights_version(self, full):
        """Fingerprint of (storage, version counter) of the parameters.  In training mode every parameter is checked on
        every call (an optimiser step bumps each `_version`).  In eval() mode -- the sampling path, where this runs once
        per batch and a 123-tensor walk would cost more host time than the kernel takes -- the first and last parameter
        are checked on every call and ALL of them on every 64th.  `load_state_dict`, `.to()/.cuda()`, `train()` and
        `repack()` invalidate the packed copy explicitly.

        What NO version counter sees: writes through `param.data` -- `p.data.copy_(...)`, which is exactly what the
        reference's `EMAHelper.ema()` does (models/ema.py:27-29).  After such a write call `repack()` (the `EMAHelper`
        of this package does it for you); `check_weights()` verifies the device copy against the live parameters."""
        if full:
            return tuple((p.data_ptr(), p._version) for p in self.parameters())
        ps = self._param_list
        if ps is None:
            ps = self._param_list = list(self.parameters())
        a, b = ps[0], ps[-1]
        return (len(ps), a.data_ptr(), a._version, b.data_ptr(), b._version)

    def repack(self):
        """Force the device-side packed weights to be rebuilt on the next call.  REQUIRED after writes that bypass autograd's
        version counters (`param.data.copy_()`, `EMAHelper.ema()` of the reference) and after in-place edits in eval mode."""
        self._packed_version = None
        self._param_list = None
        return self

    def check_weights(self, repack=True):
        """Debugging aid: compare a checksum of the LIVE parameters with the one taken when the device copy was packed
        (one small reduction + a device->host read, i.e. a synchronisation -- not for the per-batch path).  Returns True
        when the packed copy is current; otherwise repacks (unless `repack=False`) and returns False."""
        if self._packed_sum is None or self._packed_version is None:
            return False
        live = self._checksum(self._flat_params(self._device()))
        ok = bool(torch.equal(live, self._packed_sum))
        if not ok and repack:
            self.repack()
        return ok

    def _flat_params(self, device):
        sd = dict(self.named_parameters())
        return torch.cat([sd[k].detach().reshape(-1).to(device=device, dtype=torch.float32)
                          for k in _param_order(self.n_layers, self._has_temb)]).contiguous()

    @staticmethod
    def _checksum(flat):
        # two position-weighted fp64 sums: a changed, moved or swapped value changes at least one of them
        w = torch.arange(1, flat.numel() + 1, device=flat.device, dtype=torch.float64)
        f = flat.double()
        return torch.stack([f.sum(), (f * w).sum()])

    def _drop_handle(self):
        if getattr(self, "_handle", None) is not None and _lib._lib is not None:
            _lib._lib.dp_destroy(self._handle)
        self._handle = None
        self._packed_version = None
        self._param_list = None
        self._packed_sum = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        # .to()/.cuda() may have moved the parameters to another device: the handle (weights blob, scratch, SM count) belongs
        # to the device it was created on, so it is re-created on the next call
        self._drop_handle()
        if hasattr(self, "_mask_cache"):
            object.__delattr__(self, "_mask_cache")
        return out

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handle"] = None              # a ctypes pointer is neither picklable nor shareable between copies
        state["_packed_version"] = None
        state["_param_list"] = None
        state["_packed_sum"] = None
        state.pop("_mask_cache", None)
        return state

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def train(self, mode=True):
        out = super().train(mode)
        self._packed_version = None
        return out

    def _ensure_packed(self, device):
        if device.type != "cuda":
            raise RuntimeError("diffpose_nw_b200 runs on CUDA only (no CPU fallback): move the model and its inputs to a GPU")
        lib = _lib.load()
        if self._handle is not None and device != self._handle_dev:     # inputs on another GPU than the handle's: re-create it there
            self._drop_handle()
        if self._handle is None:
            h = ctypes.c_void_p()
            with torch.cuda.device(device):
                _lib.check(lib.dp_create(ctypes.byref(h), self.n_pts, self._c_in, self._c_out, self.hid_dim,
                                         self.n_layers, self.n_head, 1 if self._has_temb else 0), "dp_create")
            self._handle = h
            self._handle_dev = device
            _lib.check(lib.dp_set_engine(h, self._engine), "dp_set_engine")
        self._calls += 1
        full = self.training or (self._calls & 63) == 0
        version = self._weights_version(full)
        if self._packed_version is None or version != self._packed_version[1 if full else 0]:
            flat = self._flat_params(device)
            adj = torch.as_tensor(self.adj, dtype=torch.float32).detach().cpu().contiguous()
            if tuple(adj.shape) != (self.n_pts, self.n_pts):
                raise RuntimeError(f"adj must be [{self.n_pts},{self.n_pts}], got {tuple(adj.shape)}")
            with torch.cuda.device(device):
                stream = torch.cuda.current_stream(device).cuda_stream
                _lib.check(lib.dp_pack(self._handle, flat.data_ptr(), flat.numel(), adj.data_ptr(), stream), "dp_pack")
            self._packed_sum = self._checksum(flat)
            self._packed_version = (self._weights_version(False), self._weights_version(True))

    def _mask_bytes(self, mask, device):
        if mask is None:
            return None
        if mask.numel() != self.n_pts:
            raise RuntimeError(f"mask must have {self.n_pts} elements (reference passes [1,1,{self.n_pts}])")
        # the runner builds src_mask once and passes the same tensor to every call (runners/diffpose_frame.py:39-40): keep
        # its uint8 device copy instead of launching a conversion kernel per call (in-place edits bump _version)
        key = (mask.data_ptr(), mask._version, mask.device, mask.dtype, str(device))
        cached = getattr(self, "_mask_cache", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        mb = mask.reshape(-1).to(device=device, dtype=torch.uint8).contiguous()
        if mb.data_ptr() == mask.data_ptr():      # already uint8 on the device: .to() returned the caller's storage
            mb = mb.clone()
        object.__setattr__(self, "_mask_cache", (key, mb, mask))      # (the reference to `mask` keeps data_ptr from being recycled)
        return mb

    def _check_x(self, x, c):
        if not x.is_cuda:
            raise RuntimeError("diffpose_nw_b200 runs on CUDA only (no CPU fallback): got a CPU tensor")
        if x.dim() != 3 or x.shape[1] != self.n_pts or x.shape[2] != c:
            raise RuntimeError(f"expected x of shape [n,{self.n_pts},{c}], got {tuple(x.shape)}")
        if x.dtype is torch.float32 and x.is_contiguous() and not x.requires_grad:
            return x
        return x.detach().to(torch.float32).contiguous()

    def _forward(self, x, mask, t):
        x = self._check_x(x, self._c_in)
        dev = x.device
        self._ensure_packed(dev)
        n = x.shape[0]
        out = torch.empty(n, self.n_pts, self._c_out, device=dev, dtype=torch.float32)
        if n == 0:
            return out
        mb = self._mask_bytes(mask, dev)
        tt = None
        if self._has_temb:
            tt = torch.as_tensor(t, device=dev).detach().to(torch.float32).reshape(-1).contiguous()
            if tt.numel() != n:
                raise RuntimeError(f"t must have one entry per sample ({n}), got {tt.numel()}")
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(_lib.load().dp_forward(self._handle, x.data_ptr(), tt.data_ptr() if tt is not None else None,
                                              mb.data_ptr() if mb is not None else None, out.data_ptr(), n, stream),
                       "dp_forward")
        return out

    def last_launch(self):
        """(grid, block, smem_bytes, poses_per_tile, engine, n_tiles) of the last hot-path launch."""
        buf = (ctypes.c_long * 6)()
        _lib.check(_lib.load().dp_last_launch_info(self._handle, buf), "dp_last_launch_info")
        return tuple(buf)

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass


class FusedGCNdiff(_FusedBase):
    """GCNdiff(adj, config) -- the diffusion denoiser eps_theta(x_t, t)."""
    _has_temb = True

    def forward(self, x, mask, t, cemd=0):
        """x [n,17,c_in] fp32 CUDA, mask [1,1,17] bool, t [n] -> eps [n,17,c_out].  `cemd` is ignored, as in the
        reference (models/gcndiff.py:101)."""
        return self._forward(x, mask, t)


class FusedGCNpose(_FusedBase):
    """GCNpose(adj, config) -- the stage-1 2D->3D lifter; output width is hard-coded to 3 (gcnpose.py:91)."""
    _has_temb = False

    def _out_dim(self):
        return 3

    def forward(self, x, mask):
        return self._forward(x, mask, None)

    def lift(self, input_2d, mask=None):
        """The runner's glue between the two stages in ONE launch (`dp_lift`): `xyz = model_pose(input_2d, mask)`,
        root-centre (out of place), `cat([input_2d, xyz], 2)` (runners/diffpose_frame.py:337-343) -> [n,17,5]."""
        x = self._check_x(input_2d, self._c_in)
        dev = x.device
        self._ensure_packed(dev)
        n = x.shape[0]
        out = torch.empty(n, self.n_pts, self._c_in + self._c_out, device=dev, dtype=torch.float32)
        if n == 0:
            return out
        mb = self._mask_bytes(mask, dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(_lib.load().dp_lift(self._handle, x.data_ptr(), mb.data_ptr() if mb is not None else None, out.data_ptr(), n, stream),
                       "dp_lift")
        return out


class EMAHelper(object):
    """Drop-in for the reference's `models/ema.py::EMAHelper` (same methods, same arithmetic) that keeps the fused
    modules coherent: `ema()` writes the shadow into the parameters through `param.data.copy_` exactly like the reference
    (models/ema.py:27-29) -- a write no autograd version counter sees -- and then calls `repack()` on the module, so the
    device-side packed weights can never be stale after it."""

    def __init__(self, mu=0.999):
        self.mu = mu
        self.shadow = {}

    @staticmethod
    def _inner(module):
        return module.module if isinstance(module, nn.DataParallel) else module

    def register(self, module):
        for name, param in self._inner(module).named_parameters():
            if param.requires_grad:
                self.shadow[name] = param.data.clone()

    def update(self, module):
        for name, param in self._inner(module).named_parameters():
            if param.requires_grad:
                self.shadow[name].data = (1. - self.mu) * param.data + self.mu * self.shadow[name].data

    def ema(self, module):
        inner = self._inner(module)
        for name, param in inner.named_parameters():
            if param.requires_grad:
                param.data.copy_(self.shadow[name].data)
        if hasattr(inner, "repack"):
            inner.repack()

    def ema_copy(self, module):
        import copy
        module_copy = copy.deepcopy(self._inner(module))     # (the reference rebuilds from config; a deep copy is equivalent)
        self.ema(module_copy)
        return nn.DataParallel(module_copy) if isinstance(module, nn.DataParallel) else module_copy

    def state_dict(self):
        return self.shadow

    def load_state_dict(self, state_dict):
        self.shadow = state_dict
