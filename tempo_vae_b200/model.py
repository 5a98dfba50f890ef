"""TEMPO-VAE model — same module tree, constructor arguments, state_dict keys and public methods as the
reference's src/model.py, with all arithmetic running in libtvae_b200.so (sm_100a).

Reference interface mirrored here (file:line in /root/reference):
  get_conv / zero_init ................ src/model.py:13-42
  DiagonalGaussianDistribution ........ src/model.py:47-87
  AttnBlock ........................... src/model.py:92-152
  ResNetBlock / ResNetDown / ResNetUp . src/model.py:155-289
  Encoder / Decoder ................... src/model.py:294-574
  AutoencoderKL ....................... src/model.py:579-681
  SpectralVAE / get_model ............. src/model.py:684-759

How it executes. The module objects only own parameters (so state_dict()/load_state_dict()/parameters() are the
reference's). Compute is an explicit forward/backward program over NHWC tensors written against tempo_vae_b200.ops:
bf16 tensor-core operands, fp32 accumulation, an fp32 residual stream, fp32 GroupNorm statistics. Three
torch.autograd.Function nodes expose it to autograd: the whole get_loss (the training hot path), encode and
decode. Weight gradients are written straight into `param.grad` (a view of the optimiser's flat gradient buffer
when FusedAdamW owns the parameter).
"""
import os

import torch
import torch.nn as nn

from . import ops
from ._lib import TvaeError

ACT_CODES = {"identity": 0, "gelu": 1, "relu": 2, "silu": 3}


# =================================================================================================== engine state
class _Engine:
    """Process-wide knobs and counters of the execution engine."""

    def __init__(self):
        self.param_epoch = 0          # bumped whenever parameters are rewritten behind autograd's back
        self.launches = 0             # C-ABI kernel-launching calls (reported by bench.py as gpu_launches)
        self.rng_seed = 0x7E3B0         # Philox key for device-side eps
        self.rng_offset = 0           # global sample counter (keyed RNG => world-size invariant draws)
        self.grad_ready_hook = None   # set by parallel.DataParallel: called as hook(param) when param.grad is final
        self.unit_loss_grad = False   # Trainer sets this: loss.backward() is called with grad 1 (skips a sync)
        # conv outputs that only feed a GroupNorm (the h between the two convs of a ResNet block) are stored as bf16;
        # the fp32 residual stream is untouched. False restores fp32 storage for them (A/B and parity experiments).
        self.bf16_norm_inputs = os.environ.get("TVAE_BF16_NORM_INPUTS", "1") != "0"
        # Optional: weight-gradient GEMMs on a second stream, right behind the data-gradient GEMM of the same layer, so
        # that the GroupNorm backward of the next layer (HBM/ALU-bound, tensor cores idle) shares the SMs with a
        # tensor-bound kernel instead of running alone; joined before the optimiser / any collective reads gradients.
        # Measured on B200 (B=256, alternating runs): 116.0-117.1 ms vs 117.1-117.3 ms per step -- the board is
        # power-capped, so co-scheduling does not buy the time it would on an unconstrained part. Off by default.
        self.wgrad_overlap = os.environ.get("TVAE_WGRAD_OVERLAP", "0") == "1"
        # ... but the SMALL layers (32x32 and 16x16 at B=256: ~35 weight-gradient GEMMs of 30-80 us each) are bound by
        # launch latency and fixed per-kernel cost, not by power: those can hide behind the main stream's chain.
        # TVAE_WGRAD_OVERLAP_MAX_PIXELS: overlap only GEMMs over at most this many pixels (0 = off).
        self.wgrad_overlap_max_pixels = int(os.environ.get("TVAE_WGRAD_OVERLAP_MAX_PIXELS", "0"))
        self._side, self._main, self._side_busy = {}, {}, set()
        # every weight pack that an optimiser step made stale is rebuilt by ONE launch (TVAE_BATCHED_PACKING=0: one
        # launch per pack, on first use)
        self.batched_packing = os.environ.get("TVAE_BATCHED_PACKING", "1") != "0"
        # 1028-channel weight gradients as a 1024-channel tcgen05 GEMM (whole 128-row tiles: 1,520-1,570 TFLOP/s instead of
        # 1,180-1,250 with the padded 9th tile) + the 4 leftover channels by tvae_wgrad_skinny (taps on the M side, the
        # wide operand read once). Round-2 history: with the leftover as a 16-column tcgen05 GEMM the split bought nothing
        # (operand-bound, 1.7 ms per launch; profiles/wgrad_split_r2.txt). TVAE_SPLIT_WIDE_WGRAD=0 restores the single GEMM.
        self.split_wide_wgrad = os.environ.get("TVAE_SPLIT_WIDE_WGRAD", "1") != "0"
        # decoder.conv_out under get_loss(): the reconstruction loss, its sums and its gradient come out of the conv epilogue;
        # the fp32 reconstruction is never written (4 B/element) and the separate loss pass (8 B/element, 1.7 ms per B=256
        # step) disappears. TVAE_FUSE_NLL=0 restores conv -> tvae_nll_fwd.
        self.fuse_nll = os.environ.get("TVAE_FUSE_NLL", "1") != "0"
        self._packs = []
        # Weight packs are refreshed at the start of EVERY top-level forward (one batched launch, 0.34 ms): writes that
        # autograd cannot see (`p.data.copy_(ema)`, `nn.init.*_(p.data)`, weight surgery) are honoured like the
        # reference's modules, which read the live parameter on every call. `frozen_weights()` opts a sweep out.
        self.forward_serial = 0
        self._frozen = 0

    def params_changed(self):
        """Public: marks every bf16 weight pack stale (called by FusedAdamW.step, load_state_dict paths, set_precision).
        Not needed after ordinary `.data` writes between forwards -- see forward_serial."""
        self.param_epoch += 1

    def begin_forward(self):
        if not self._frozen:
            self.forward_serial += 1

    def frozen_weights(self):
        """Context manager for inference sweeps: packs are refreshed once on entry, then assumed unchanged."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            self.forward_serial += 1
            self._frozen += 1
            try:
                yield
            finally:
                self._frozen -= 1
        return cm()

    def register_pack(self, module, mode, ent):
        import weakref
        self._packs.append((weakref.ref(module), mode, ent))

    def repack_stale(self, device):
        """Rebuild, in ONE kernel launch, every registered bf16 weight pack on `device` whose parameter changed since
        it was packed (after an optimiser step: all of them -- ~80 launches become one)."""
        items, keys = [], []
        alive = []
        for ref, mode, ent in self._packs:
            mod = ref()
            if mod is None or mod.__dict__.get("_packs", {}).get(mode) is not ent:
                continue
            alive.append((ref, mode, ent))
            w = mod.weight
            if w.device != device or ent.data.device != device or not w.is_contiguous():
                continue
            key = (w.data_ptr(), w._version, self.param_epoch, self.forward_serial, False)
            if ent.version != key:
                items.append((w.detach(), mode, ent))
                keys.append((ent, key))
        self._packs = alive
        if len(items) < 2:
            return
        ops.pack_weights_batched(items)
        for ent, key in keys:
            ent.version = key

    def side_stream(self, device):
        """The weight-gradient stream of `device`, made to wait for everything enqueued so far on the current one."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        side = self._side.get(idx)
        if side is None:
            side = self._side[idx] = torch.cuda.Stream(device=idx)
        main = torch.cuda.current_stream(idx)
        self._main[idx] = main
        side.wait_stream(main)
        self._side_busy.add(idx)
        return side

    def join_side_streams(self):
        """Current streams wait for the weight-gradient streams (before gradients are consumed)."""
        for idx in list(self._side_busy):
            torch.cuda.current_stream(idx).wait_stream(self._side[idx])
        self._side_busy.clear()

    def sync_streams_for_collective(self, device):
        """Called right before a gradient bucket is handed to NCCL (which orders itself after the CURRENT stream
        only): the current stream first waits for the other one of the (compute, weight-gradient) pair."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        side, main = self._side.get(idx), self._main.get(idx)
        if side is None or main is None:
            return
        cur = torch.cuda.current_stream(idx)
        cur.wait_stream(main if cur == side else side)


ENGINE = _Engine()


def set_precision(mode: str):
    """"bf16" (default): bf16 tensor-core operands, fp32 accumulation. "fp32": forward GEMMs emulate fp32 products
    with split-bf16 operands (hi*hi + hi*lo + lo*hi, 3x the tensor work, ~2^-16 relative product error) — the
    north star's "fp32 mode" for parity runs; backward GEMMs keep plain bf16 operands."""
    if mode not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    ops.SPLIT_BF16[0] = mode == "fp32"
    ENGINE.params_changed()


def get_precision() -> str:
    return "fp32" if ops.SPLIT_BF16[0] else "bf16"


def _grad_begin(p):
    """Returns (tensor to write, accumulate?) for parameter p, installing the flat-buffer view if there is one."""
    if p.grad is None:
        view = getattr(p, "_tvae_grad", None)
        p.grad = view if view is not None else torch.empty_like(p)
        return p.grad, False
    return p.grad, True


def _grad_done(p):
    hook = ENGINE.grad_ready_hook
    if hook is not None:
        hook(p)


def _write_grad(p, fn, native_accumulate=False):
    """fn(dst) overwrites dst with the gradient of p; handles autograd's accumulate-into-.grad semantics (micro-batch
    accumulation: bench.py --global-batch). native_accumulate: fn(dst, accumulate) can add in place itself."""
    if not p.requires_grad:
        return
    g, acc = _grad_begin(p)
    if native_accumulate:
        fn(g, acc)
    elif acc:
        tmp = torch.empty_like(g)
        fn(tmp)
        g.add_(tmp)
    else:
        fn(g)
    _grad_done(p)


# =================================================================================================== helpers
@torch.no_grad()
def zero_init(module: nn.Module) -> nn.Module:
    """Sets to zero all the parameters of a module, and returns the module. (src/model.py:13-18)"""
    for p in module.parameters():
        torch.nn.init.zeros_(p.data)
    return module


class _PackedMixin:
    """bf16 GEMM-operand copies of a conv weight, rebuilt lazily when the parameter changes."""

    def packed(self, mode):
        packs = self.__dict__.setdefault("_packs", {})
        w = self.weight
        key = (w.data_ptr(), w._version, ENGINE.param_epoch, ENGINE.forward_serial, ops.SPLIT_BF16[0])
        ent = packs.get(mode)
        if ent is None or ent.data.device != w.device:
            ent = ops.pack_weight(w, mode)
            packs[mode] = ent
            ENGINE.register_pack(self, mode, ent)
        elif ent.version != key:
            if ENGINE.batched_packing and not ops.SPLIT_BF16[0]:
                ENGINE.repack_stale(w.device)          # one launch for every pack the optimiser step invalidated
            if ent.version != key:
                ops.pack_weight(w, mode, out=ent)
        ent.version = key
        return ent

    def _no_direct_call(self):
        raise TvaeError(
            f"{type(self).__name__}: layers of the B200 engine are executed by their owning Encoder/Decoder program; "
            "call AutoencoderKL.encode/decode/forward/get_loss (or Encoder/Decoder) instead of a single layer")


class Conv2d(nn.Conv2d, _PackedMixin):
    """nn.Conv2d parameters + init (identical RNG consumption); forward/backward run in conv_gemm / wgrad_gemm."""

    def conv_kind(self):
        k, s, p = self.kernel_size, self.stride, self.padding
        if k[0] != k[1] or s[0] != s[1] or self.dilation != (1, 1) or self.groups != 1 or self.padding_mode != "zeros":
            raise TvaeError(f"unsupported convolution {self}")
        if s == (1, 1) and k[0] in (1, 3) and p == (k[0] // 2, k[0] // 2):
            return 0, k[0]
        if s == (2, 2) and k == (2, 2) and p == (0, 0):
            return 1, 2
        raise TvaeError(f"unsupported convolution geometry {self} (supported: 1x1/3x3 stride 1 'same', 2x2 stride 2)")

    def forward(self, x):  # noqa: D401
        self._no_direct_call()


class ConvTranspose2d(nn.ConvTranspose2d, _PackedMixin):
    def conv_kind(self):
        if not (self.kernel_size == (2, 2) and self.stride == (2, 2) and self.padding == (0, 0)
                and self.output_padding == (0, 0) and self.groups == 1 and self.dilation == (1, 1)):
            raise TvaeError(f"unsupported transposed convolution {self} (supported: 2x2 stride 2)")
        return 2, 2

    def forward(self, x, output_size=None):  # noqa: D401
        self._no_direct_call()


def get_conv(in_channels, out_channels, **kwargs):
    """Same factory as src/model.py:21-42 (defaults k3 p1 s1, zeros padding, optional init / transposed)."""
    def_params = {
        "dim": 2,
        "kernel_size": 3,
        "padding": 1,
        "stride": 1,
        "padding_mode": "zeros",
        "dilation": 1,
        "groups": 1,
        "init": lambda x: x,
        "transposed": False,
    }
    def_params.update(kwargs)
    dim = def_params.pop("dim")
    transposed = def_params.pop("transposed")
    init = def_params.pop("init")
    if dim != 2:
        raise TvaeError("the B200 engine implements the 2-D model only (dim=3 is unused by TEMPO-VAE)")
    conv = ConvTranspose2d if transposed else Conv2d
    return init(conv(in_channels, out_channels, **def_params))


class GroupNorm(nn.GroupNorm):
    def forward(self, x):  # noqa: D401
        raise TvaeError("GroupNorm is executed fused with its activation by the owning Encoder/Decoder program")

    def affine_params(self):
        if self.affine:
            return self.weight, self.bias
        return self._ones, self._zeros


def _make_norm(num_channels, norm_params):
    gn = GroupNorm(num_channels=num_channels, **norm_params)
    if not gn.affine:  # identity scale/shift live in non-persistent buffers (state_dict keys stay the reference's)
        gn.register_buffer("_ones", torch.ones(num_channels), persistent=False)
        gn.register_buffer("_zeros", torch.zeros(num_channels), persistent=False)
    return gn


class _Act(nn.Module):
    """Parameter-free stand-in for nn.GELU/ReLU/SiLU (keeps the Sequential indices of the reference)."""

    def __init__(self, name):
        super().__init__()
        self.name = name
        self.code = ACT_CODES[name]

    def forward(self, x):  # noqa: D401
        raise TvaeError("activations are executed fused with GroupNorm by the owning Encoder/Decoder program")

    def extra_repr(self):
        return self.name


class A:
    """An activation in flight: fp32 stream tensor and/or bf16 operand tensor, NHWC, C valid channels."""

    __slots__ = ("f32", "bf16", "C", "stats")

    def __init__(self, f32=None, bf16=None, C=0, stats=None):
        self.f32, self.bf16, self.C = f32, bf16, C
        self.stats = stats      # (G, eps, [N,G,2] mean/rstd) produced by the conv epilogue that wrote f32, or None

    def as_bf16(self):
        if self.bf16 is None:
            self.bf16 = ops.f32_to_bf16(self.f32)
        return self.bf16


# =================================================================================================== conv fwd/bwd
class ConvOut(tuple):
    """(out_f32, out_bf16) with the fused GroupNorm statistics of the output attached as `.stats`."""
    stats = None


def conv_fwd(mod, x_bf16, Cin, *, residual=None, want_f32=True, want_bf16=False, out_f32=None, stats_for=None):
    """stats_for: the GroupNorm that will consume the fp32 output — its statistics are then produced by this conv's
    epilogue (when the geometry allows) instead of a separate pass over the tensor."""
    kind, R = mod.conv_kind()
    Cout = mod.out_channels
    mode = "up_fwd" if kind == 2 else "fwd"
    spec = None
    if stats_for is not None and out_f32 is None and stats_for.num_channels == Cout:
        spec = (stats_for.num_groups, stats_for.eps)
    r = ops.conv_gemm(x_bf16, Cin, mod.packed(mode), kind=kind, R=R, Cout=Cout, bias=mod.bias, residual=residual,
                      want_f32=want_f32, want_bf16=want_bf16, out_f32=out_f32, stats=spec)
    out = ConvOut(r[:2])
    if spec is not None and r[2] is not None:
        out.stats = (spec[0], spec[1], r[2])
    return out


def conv_bwd(mod, dy_bf16, x_bf16, Cin, *, dgrad=None, dgrad_residual=None, bias_grad_from=None):
    """Backward of conv_fwd. dy_bf16: gradient wrt the conv output (bf16 NHWC). Writes weight/bias grads.
    dgrad: None | "bf16" | "f32" — format of the returned input gradient."""
    kind, R = mod.conv_kind()
    Cout = mod.out_channels
    w = mod.weight
    x_bf16, dy_bf16 = ops.hi_of(x_bf16), ops.hi_of(dy_bf16)      # backward GEMMs run on plain bf16 operands
    dx = None
    if dgrad is not None:       # data gradient first: it is on the critical path, the weight gradient is not
        dx = _conv_dgrad(mod, kind, R, Cin, Cout, dy_bf16, dgrad, dgrad_residual)
    if w.requires_grad:
        def wg(dst, acc):
            if kind == 2:   # ConvTranspose2d [Cin][Cout][2][2]: P = x (coarse), Q = dy (fine)
                ops.wgrad_gemm(x_bf16, Cin, dy_bf16, Cout, kind=1, R=2, grad=dst, accumulate=acc)
            elif (kind == 0 and R == 3 and ENGINE.split_wide_wgrad and Cin > 256 and 0 < Cin % 128 <= 4
                  and Cout % 64 == 0 and Cout <= 512):
                # encoder.conv_in (1028 -> 512): 1028 = 8 x 128 + 4. A 9th 128-row tile for 4 channels is 97 % padding
                # (a full MMA sweep, 1.3-1.8 ms per launch); instead the whole tiles run with x[:, :1024] on M and the 4
                # leftover channels go through tvae_wgrad_skinny (dW[n][1024 + c][tap] = sum dY[px][n] x[px + tap][1024 + c])
                main = Cin - Cin % 128
                ops.wgrad_gemm(x_bf16[..., :main], main, dy_bf16, Cout, kind=0, R=R, grad=dst, flip=True, accumulate=acc,
                               grad_ld=Cin, grad_off=0)
                ops.wgrad_skinny(dy_bf16, Cout, x_bf16[..., main:Cin], Cin - main, sign=+1, grad=dst.view(-1)[main * 9:],
                                 stride_c=9, stride_n=9 * Cin, accumulate=acc)
            elif (kind == 0 and R == 3 and ENGINE.split_wide_wgrad and Cout > 256 and 0 < Cout % 128 <= 4
                  and Cin % 64 == 0 and Cin <= 512):
                # decoder.conv_out (512 -> 1028): the same split on the output-channel side; both pieces are contiguous row
                # blocks of the [Cout][Cin][3][3] parameter (dW[1024 + c][n][tap] = sum x[px'][n] dY[px' - tap][1024 + c])
                main = Cout - Cout % 128
                flat = dst.view(-1)
                ops.wgrad_gemm(dy_bf16[..., :main], main, x_bf16, Cin, kind=0, R=R, grad=flat[:main * Cin * 9],
                               accumulate=acc)
                ops.wgrad_skinny(x_bf16, Cin, dy_bf16[..., main:Cout], Cout - main, sign=-1, grad=flat[main * Cin * 9:],
                                 stride_c=9 * Cin, stride_n=9, accumulate=acc)
            elif kind == 0 and Cin > 256 and Cin % 256 and Cout % 128 == 0:
                # a wide, awkward channel count (1028) goes on the GEMM's M side: operand roles exchanged. (Measured:
                # the opposite choice, N = 1028 as 5 tiles of 208, pads less (1.2 % vs 12 %) but is 3 ms per launch
                # SLOWER: a 208-wide tile still stages 256 channels per K block, and the kernel is operand-bound.)
                ops.wgrad_gemm(x_bf16, Cin, dy_bf16, Cout, kind=0, R=R, grad=dst, flip=True, accumulate=acc)
            else:
                ops.wgrad_gemm(dy_bf16, Cout, x_bf16, Cin, kind=kind, R=R, grad=dst, accumulate=acc)
        if ENGINE.wgrad_overlap or dy_bf16.numel() // dy_bf16.shape[-1] <= ENGINE.wgrad_overlap_max_pixels:
            side = ENGINE.side_stream(dy_bf16.device)
            with torch.cuda.stream(side):
                _write_grad(w, wg, native_accumulate=True)
            dy_bf16.record_stream(side)       # the caching allocator must not recycle the operands early
            x_bf16.record_stream(side)
        else:
            _write_grad(w, wg, native_accumulate=True)
    if mod.bias is not None and mod.bias.requires_grad:
        if bias_grad_from is None:
            bias_grad_from = getattr(dy_bf16, "tvae_colsum", None)     # produced together with dy (norm_act_bwd)
        if bias_grad_from is not None:
            _write_grad(mod.bias, lambda dst: dst.copy_(bias_grad_from))
        else:
            def bg(dst):
                ops.colsum_bf16(dy_bf16, Cout, dst)
            _write_grad(mod.bias, bg)
    return dx


def _conv_dgrad(mod, kind, R, Cin, Cout, dy_bf16, dgrad, dgrad_residual):
    if kind == 0:
        of, ob = ops.conv_gemm(dy_bf16, Cout, mod.packed("dgrad"), kind=0, R=R, Cout=Cin, flip=True,
                               residual=dgrad_residual, want_f32=(dgrad == "f32"), want_bf16=(dgrad == "bf16"),
                               split_out=False)
    elif kind == 1:
        of, ob = ops.conv_gemm(dy_bf16, Cout, mod.packed("down_dgrad"), kind=2, R=2, Cout=Cin,
                               want_f32=(dgrad == "f32"), want_bf16=(dgrad == "bf16"), split_out=False)
    else:
        of, ob = ops.conv_gemm(dy_bf16, Cout, mod.packed("up_dgrad"), kind=1, R=2, Cout=Cin,
                               want_f32=(dgrad == "f32"), want_bf16=(dgrad == "bf16"), split_out=False)
    return of if dgrad == "f32" else ob


class NormStats:
    """(mean, rstd) statistics of a GroupNorm input, plus -- when the forward ran with save=True -- the activation
    derivative act'(y) it stored for the backward pass (bf16, or None). Travels in the saved-activation tuples where the
    bare statistics tensor used to."""

    __slots__ = ("stats", "act_grad")

    def __init__(self, stats, act_grad=None):
        self.stats, self.act_grad = stats, act_grad


def norm_act_fwd(norm, h_f32, act_code, stats=None, save=False):
    C = norm.num_channels
    gamma, beta = norm.affine_params()
    if stats is not None and stats[0] == norm.num_groups and stats[1] == norm.eps:
        st = stats[2]
    else:
        if h_f32.dtype != torch.float32:
            raise RuntimeError("a bf16 GroupNorm input needs statistics from the producing conv's epilogue")
        st = ops.gn_stats(h_f32, C, norm.num_groups, norm.eps)
    if save and act_code != 0:
        a, gp = ops.gn_act_fwd(h_f32, st, gamma, beta, norm.num_groups, act_code, want_act_grad=True)
        return a, NormStats(st, gp)
    a = ops.gn_act_fwd(h_f32, st, gamma, beta, norm.num_groups, act_code)
    return a, NormStats(st)


def norm_act_bwd(norm, h_f32, stats, da_bf16, gres_bf16, act_code):
    """Returns dx (bf16). dx is the output gradient of the conv(s) that produced h: its column sums (their bias
    gradient) come out of the same pass and ride along as `dx.tvae_colsum`, which conv_bwd picks up."""
    gamma, beta = norm.affine_params()
    C = norm.num_channels
    dev = h_f32.device
    cs = torch.empty((C,), dtype=torch.float32, device=dev)
    train_affine = norm.affine and gamma.requires_grad
    if train_affine:
        dg, acc_g = _grad_begin(gamma)
        db, acc_b = _grad_begin(beta)
    else:
        dg = torch.empty((C,), dtype=torch.float32, device=dev)
        db = torch.empty((C,), dtype=torch.float32, device=dev)
        acc_g = acc_b = False
    gp = None
    if isinstance(stats, NormStats):
        stats, gp = stats.stats, stats.act_grad
    if acc_g or acc_b:
        tg, tb = torch.empty_like(dg), torch.empty_like(db)
        dx = ops.gn_act_bwd(h_f32, stats, gamma, beta, da_bf16, gres_bf16, norm.num_groups, act_code, tg, tb, cs, gp)
        dg.add_(tg)
        db.add_(tb)
    else:
        dx = ops.gn_act_bwd(h_f32, stats, gamma, beta, da_bf16, gres_bf16, norm.num_groups, act_code, dg, db, cs, gp)
    if train_affine:
        _grad_done(gamma)
        _grad_done(beta)
    dx.tvae_colsum = cs
    return dx


# =================================================================================================== distribution
class _SampleFn(torch.autograd.Function):
    """z = mean + exp(0.5*clamp(logvar)) * eps through tvae_reparam_fwd / _bwd (moments NCHW fp32 at the API)."""

    @staticmethod
    def forward(ctx, parameters, eps):
        B, C2, h, w = parameters.shape
        Z = C2 // 2
        mom = parameters.permute(0, 2, 3, 1).contiguous()
        if eps is None:
            off = ENGINE.rng_offset
            ENGINE.rng_offset += B
            _, z, eps_used, _ = ops.reparam_fwd(mom, Z, seed=ENGINE.rng_seed, sample_offset=off, want_z_nchw=True)
        else:
            _, z, eps_used, _ = ops.reparam_fwd(mom, Z, eps=eps, want_z_nchw=True)
        ctx.save_for_backward(mom, eps_used)
        ctx.Z = Z
        return z

    @staticmethod
    def backward(ctx, dz):
        mom, eps = ctx.saved_tensors
        dz_nhwc = dz.permute(0, 2, 3, 1).contiguous().float()
        dm = ops.reparam_bwd(mom, ctx.Z, dz_nhwc, eps, None, None, 0.0)
        return dm.float().permute(0, 3, 1, 2), None


class DiagonalGaussianDistribution(object):
    """Same attributes/methods as src/model.py:47-87. `parameters` is the NCHW fp32 moments tensor."""

    def __init__(self, parameters, dim=2, deterministic=False):
        self.parameters = parameters
        self.dim = dim
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean).to(device=self.parameters.device)

    def sample(self, eps=None):
        """`eps` (optional, NCHW like mean) injects the noise; default draws it on the device (Philox), where the
        reference draws it on the CPU and copies it over (src/model.py:61-65)."""
        if self.deterministic:
            return self.mean
        return _SampleFn.apply(self.parameters, eps)

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.0])
        dims = [1, 2, 3] if self.dim == 2 else [1, 2, 3, 4]
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=dims)
        return 0.5 * torch.sum(
            torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0 - self.logvar + other.logvar,
            dim=dims)

    def mode(self):
        return self.mean


# =================================================================================================== blocks
class AttnBlock(nn.Module):
    def __init__(self, in_channels, n_heads=4, dim=2, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        assert self.in_channels % n_heads == 0, "in_channels must be divisible by n_heads"
        self.n_heads = n_heads
        self.dim = dim
        assert self.dim == 2 or self.dim == 3, "dim must be 2 or 3"
        norm_params = kwargs.get("norm_params", {})
        self.norm = _make_norm(in_channels, norm_params)
        self.q = get_conv(in_channels, in_channels, dim=self.dim, kernel_size=1, stride=1, padding=0)
        self.k = get_conv(in_channels, in_channels, dim=self.dim, kernel_size=1, stride=1, padding=0)
        self.v = get_conv(in_channels, in_channels, dim=self.dim, kernel_size=1, stride=1, padding=0)
        self.proj_out = get_conv(in_channels, in_channels, dim=self.dim, kernel_size=1, stride=1, padding=0)

    # h: A with f32 stream. returns A (f32), saved
    def fwd(self, h, save, next_norm=None):
        C = self.in_channels
        N, H, W, _ = h.f32.shape
        hn, stats = norm_act_fwd(self.norm, h.f32, 0, h.stats, save)
        qkv = torch.empty((N, H, W, 3 * C), dtype=torch.float32, device=h.f32.device)
        for i, m in enumerate((self.q, self.k, self.v)):
            conv_fwd(m, hn, C, out_f32=qkv[..., i * C:(i + 1) * C])
        o_bf16, o_f32, lse = ops.attn_fwd(qkv, C, self.n_heads, N, H * W)
        o4 = o_bf16.view(N, H, W, C)          # tensor or split-bf16 Pair
        r = conv_fwd(self.proj_out, o4, C, residual=h.f32, stats_for=next_norm)
        saved = (h.f32, stats, hn, qkv, o4, o_f32, lse) if save else None
        return A(f32=r[0], C=C, stats=r.stats), saved

    def bwd(self, g, saved):
        h_f32, stats, hn, qkv, o4, o_f32, lse = saved
        C = self.in_channels
        N, H, W, _ = h_f32.shape
        d_o = conv_bwd(self.proj_out, g, o4, C, dgrad="f32")
        dqkv = ops.attn_bwd(qkv, o_f32, d_o, lse, C, self.n_heads, N, H * W).view(N, H, W, 3 * C)
        acc = None
        mods = (self.q, self.k, self.v)
        for i, m in enumerate(mods):
            dyi = dqkv[..., i * C:(i + 1) * C]
            last = i == len(mods) - 1
            r = conv_bwd(m, dyi, hn, C, dgrad=("bf16" if last else "f32"), dgrad_residual=acc)
            acc = r
        return norm_act_bwd(self.norm, h_f32, stats, acc, g, 0)


class ResNetBlock(nn.Module):
    def __init__(self, ch_in, ch_out, dim=2, conditioning_dims=None, dropout_prob=0.0, nca_params={},
                 cond_proj_type="zerolinear"):
        super().__init__()
        self.ch_in = ch_in
        self.ch_out = ch_out
        self.dim = dim
        assert self.dim in [2, 3], "dim must be 2 or 3"
        self.conditioning_dims = conditioning_dims
        if conditioning_dims is not None:
            raise TvaeError("conditioning is unused by TEMPO-VAE (conditionings=None, src/model.py:412,556)")
        if dropout_prob > 0.0:
            raise TvaeError("dropout_prob > 0 is not implemented by the B200 engine (configs use 0.0)")
        self.nca_params = nca_params
        norm_params = self.nca_params.get("norm_params", {})
        get_act = self.nca_params.get("get_act", lambda: _Act("gelu"))
        conv_params = self.nca_params.get("conv_params", {})
        self.net1 = nn.Sequential(
            _make_norm(ch_in, norm_params),
            get_act(),
            get_conv(ch_in, ch_out, dim=self.dim, **conv_params),
        )
        self.net2 = nn.Sequential(
            _make_norm(ch_out, norm_params),
            get_act(),
            get_conv(ch_out, ch_out, dim=self.dim, init=zero_init, **conv_params),
        )
        if ch_in != ch_out:
            self.skip_conv = get_conv(ch_in, ch_out, dim=self.dim, kernel_size=1, padding=0)

    def fwd(self, h, save, want_bf16=False, next_norm=None):
        act = self.net1[1].code
        a1, st1 = norm_act_fwd(self.net1[0], h.f32, act, h.stats, save)
        # h1 only feeds net2's GroupNorm (it is not on the fp32 residual stream): when its statistics come out of the
        # conv epilogue (taken from the fp32 accumulators) it is stored as bf16 -- half the bytes for the conv
        # epilogue, the norm's forward and both passes of its backward
        n2 = self.net2[0]
        N_, H_, W_ = a1.shape[0], a1.shape[1], a1.shape[2]
        h1_bf16 = (ENGINE.bf16_norm_inputs and not ops.SPLIT_BF16[0] and ops.gn_fast_ok(self.ch_out, n2.num_groups)
                   and ops.fused_stats_ok(N_, H_, W_, self.ch_out, n2.num_groups, 0, H_, W_))
        r1 = conv_fwd(self.net1[2], a1, self.ch_in, stats_for=n2, want_f32=not h1_bf16, want_bf16=h1_bf16)
        h1 = r1[1] if h1_bf16 else r1[0]
        a2, st2 = norm_act_fwd(n2, h1, self.net2[1].code, r1.stats, save)
        if self.ch_in != self.ch_out:
            xb = h.as_bf16()
            res, _ = conv_fwd(self.skip_conv, xb, self.ch_in)
        else:
            xb = None
            res = h.f32
        r2 = conv_fwd(self.net2[-1], a2, self.ch_out, residual=res, want_bf16=want_bf16, stats_for=next_norm)
        saved = (h.f32, st1, a1, h1, st2, a2, xb) if save else None
        return A(f32=r2[0], bf16=r2[1], C=self.ch_out, stats=r2.stats), saved

    def bwd(self, g, saved):
        x_f32, st1, a1, h1, st2, a2, xb = saved
        skip = self.ch_in != self.ch_out
        d_a2 = conv_bwd(self.net2[-1], g, a2, self.ch_out, dgrad="bf16")
        d_h1 = norm_act_bwd(self.net2[0], h1, st2, d_a2, None, self.net2[1].code)
        d_a1 = conv_bwd(self.net1[2], d_h1, a1, self.ch_in, dgrad="bf16")
        if skip:
            g_res = conv_bwd(self.skip_conv, g, xb, self.ch_in, dgrad="bf16")
        else:
            g_res = g
        return norm_act_bwd(self.net1[0], x_f32, st1, d_a1, g_res, self.net1[1].code)


class ResNetDown(nn.Module):
    def __init__(self, resnet_blocks, attention_blocks=None):
        super().__init__()
        self.resnet_blocks = resnet_blocks
        self.attention_blocks = attention_blocks
        self.dim = self.resnet_blocks[-1].dim
        self.down = get_conv(self.resnet_blocks[-1].ch_out, self.resnet_blocks[-1].ch_out, dim=self.dim,
                             kernel_size=2, stride=2, padding=0)

    def fwd(self, h, save, no_down=False, want_bf16=False, next_norm=None):
        saved = []
        nb = len(self.resnet_blocks)
        for i, blk in enumerate(self.resnet_blocks):
            last = i == nb - 1
            feeds_conv = last and self.attention_blocks is None and not no_down
            after = (self.resnet_blocks[i + 1].net1[0] if not last else (next_norm if no_down else None))
            blk_next = self.attention_blocks[i].norm if self.attention_blocks is not None else after
            h, s = blk.fwd(h, save, want_bf16=feeds_conv, next_norm=blk_next)
            saved.append(s)
            if self.attention_blocks is not None:
                h, s = self.attention_blocks[i].fwd(h, save, next_norm=after)
                saved.append(s)
        if no_down:
            return h, (saved, None)
        xb = h.as_bf16()
        C = self.down.in_channels
        r = conv_fwd(self.down, xb, C, want_bf16=want_bf16, stats_for=next_norm)
        return A(f32=r[0], bf16=r[1], C=self.down.out_channels, stats=r.stats), (saved, xb if save else None)

    def bwd(self, g, saved):
        blocks, xb = saved
        if xb is not None:
            g = conv_bwd(self.down, g, xb, self.down.in_channels, dgrad="bf16")
        idx = len(blocks) - 1
        for i in reversed(range(len(self.resnet_blocks))):
            if self.attention_blocks is not None:
                g = self.attention_blocks[i].bwd(g, blocks[idx])
                idx -= 1
            g = self.resnet_blocks[i].bwd(g, blocks[idx])
            idx -= 1
        return g


class ResNetUp(nn.Module):
    def __init__(self, resnet_blocks, attention_blocks=None, ch_out=None, conv_params={}):
        super().__init__()
        self.resnet_blocks = resnet_blocks
        self.ch_out = ch_out if ch_out is not None else self.resnet_blocks[-1].ch_out
        self.attention_blocks = attention_blocks
        self.dim = self.resnet_blocks[-1].dim
        self.up = get_conv(self.resnet_blocks[-1].ch_out, self.ch_out, dim=self.dim, kernel_size=2, stride=2,
                           padding=0, transposed=True)

    def fwd(self, h, save, no_up=False, next_norm=None):
        saved = []
        nb = len(self.resnet_blocks)
        for i, blk in enumerate(self.resnet_blocks):
            last = i == nb - 1
            feeds_conv = last and self.attention_blocks is None and not no_up
            after = (self.resnet_blocks[i + 1].net1[0] if not last else (next_norm if no_up else None))
            blk_next = self.attention_blocks[i].norm if self.attention_blocks is not None else after
            h, s = blk.fwd(h, save, want_bf16=feeds_conv, next_norm=blk_next)
            saved.append(s)
            if self.attention_blocks is not None:
                h, s = self.attention_blocks[i].fwd(h, save, next_norm=after)
                saved.append(s)
        if no_up:
            return h, (saved, None)
        xb = h.as_bf16()
        r = conv_fwd(self.up, xb, self.up.in_channels, stats_for=next_norm)
        return A(f32=r[0], C=self.ch_out, stats=r.stats), (saved, xb if save else None)

    def bwd(self, g, saved):
        blocks, xb = saved
        if xb is not None:
            g = conv_bwd(self.up, g, xb, self.up.in_channels, dgrad="bf16")
        idx = len(blocks) - 1
        for i in reversed(range(len(self.resnet_blocks))):
            if self.attention_blocks is not None:
                g = self.attention_blocks[i].bwd(g, blocks[idx])
                idx -= 1
            g = self.resnet_blocks[i].bwd(g, blocks[idx])
            idx -= 1
        return g


def _enc_dec_common(self, shape, chs, attn_sizes, mid_attn, num_res_blocks, dropout_prob, z_channels, double_z,
                    n_attention_heads, norm_groups, norm_eps, norm_affine, act, conv_kernel_size, conv_padding_mode):
    self.shape = shape
    self.in_channels = self.shape[0]
    self.input_size = self.shape[1]
    self.chs = chs
    self.dim = len(self.shape) - 1
    self.attn_sizes = attn_sizes
    self.mid_attn = mid_attn
    if (len(self.attn_sizes) > 0 or self.mid_attn) and self.dim == 3:
        raise ValueError("3D attention very highly discouraged.")
    if self.dim != 2:
        raise TvaeError("the B200 engine implements the 2-D model only")
    self.num_res_blocks = num_res_blocks
    self.dropout_prob = dropout_prob
    self.z_channels = z_channels
    self.double_z = double_z
    self.n_attention_heads = n_attention_heads
    assert conv_kernel_size % 2 == 1, "conv_kernel_size must be odd"
    if conv_kernel_size not in (1, 3):
        raise TvaeError("the B200 engine implements conv_kernel_size 1 and 3")
    if conv_padding_mode != "zeros":
        raise TvaeError("the B200 engine implements conv_padding_mode='zeros' only")
    norm_params = dict(num_groups=norm_groups, eps=norm_eps, affine=norm_affine)
    assert act in ["gelu", "relu", "silu"], "act must be gelu or relu or silu"
    self.act_name = act

    def get_act():
        return _Act(act)

    padding = conv_kernel_size // 2
    conv_params = dict(kernel_size=conv_kernel_size, padding=padding, padding_mode=conv_padding_mode)
    nca_params = dict(norm_params=norm_params, get_act=get_act, conv_params=conv_params)
    resnet_params = dict(dim=self.dim, conditioning_dims=None, dropout_prob=self.dropout_prob, nca_params=nca_params)
    self.n_sizes = len(self.chs)
    return norm_params, get_act, conv_params, resnet_params


class Encoder(nn.Module):
    def __init__(self, shape, chs=[48, 96, 192], attn_sizes=[], mid_attn=False, num_res_blocks=1, dropout_prob=0.0,
                 z_channels=4, double_z=True, n_attention_heads=1, norm_groups=8, norm_eps=1e-6, norm_affine=True,
                 act="gelu", conv_kernel_size=3, conv_padding_mode="zeros"):
        super().__init__()
        norm_params, get_act, conv_params, resnet_params = _enc_dec_common(
            self, shape, chs, attn_sizes, mid_attn, num_res_blocks, dropout_prob, z_channels, double_z,
            n_attention_heads, norm_groups, norm_eps, norm_affine, act, conv_kernel_size, conv_padding_mode)
        self.conv_in = get_conv(self.in_channels, self.chs[0], dim=self.dim, **conv_params)
        curr_size = self.input_size
        self.downs = nn.ModuleList()
        for i_level in range(self.n_sizes):
            ch_in = chs[0] if i_level == 0 else chs[i_level - 1]
            ch_out = chs[i_level]
            resnets = nn.ModuleList()
            attentions = nn.ModuleList()
            for _ in range(self.num_res_blocks):
                resnets.append(ResNetBlock(ch_in, ch_out, **resnet_params))
                if curr_size in self.attn_sizes:
                    attentions.append(AttnBlock(ch_out, n_heads=self.n_attention_heads, dim=self.dim,
                                                norm_params=norm_params))
                ch_in = ch_out
            if len(attentions) == 0:
                attentions = None
            down = ResNetDown(resnets, attentions)
            curr_size = curr_size // 2
            self.downs.append(down)
        self.mid1 = ResNetBlock(ch_in, ch_in, **resnet_params)
        if self.mid_attn:
            self.mid_attn1 = AttnBlock(ch_in, n_heads=self.n_attention_heads, dim=self.dim, norm_params=norm_params)
        self.mid2 = ResNetBlock(ch_in, ch_in, **resnet_params)
        self.norm_out = _make_norm(ch_in, norm_params)
        self.act_out = get_act()
        self.conv_out = get_conv(in_channels=ch_in, out_channels=2 * z_channels if double_z else z_channels,
                                 dim=self.dim, init=zero_init, **conv_params)

    # ---- engine program: x_bf16 NHWC -> A(conv_out output); `tail_bf16` asks for a bf16 copy (feeds quant_conv)
    def fwd(self, x_bf16, save, tail_bf16=False):
        saved = {}
        r = conv_fwd(self.conv_in, x_bf16, self.in_channels, stats_for=self.downs[0].resnet_blocks[0].net1[0])
        saved["x"] = x_bf16 if save else None
        h = A(f32=r[0], C=self.chs[0], stats=r.stats)
        levels = []
        n = len(self.downs)
        for i, down in enumerate(self.downs):
            last = i == n - 1
            nxt_skip = (not last) and (self.chs[i] != self.chs[i + 1])
            nxt_norm = self.mid1.net1[0] if last else self.downs[i + 1].resnet_blocks[0].net1[0]
            h, s = down.fwd(h, save, no_down=last, want_bf16=nxt_skip, next_norm=nxt_norm)
            levels.append(s)
        saved["levels"] = levels
        h, saved["mid1"] = self.mid1.fwd(h, save, next_norm=(self.mid_attn1.norm if self.mid_attn else self.mid2.net1[0]))
        if self.mid_attn:
            h, saved["attn"] = self.mid_attn1.fwd(h, save, next_norm=self.mid2.net1[0])
        h, saved["mid2"] = self.mid2.fwd(h, save, next_norm=self.norm_out)
        a, st = norm_act_fwd(self.norm_out, h.f32, self.act_out.code, h.stats, save)
        saved["out"] = (h.f32, st, a) if save else None
        Cz = self.conv_out.out_channels
        of, ob = conv_fwd(self.conv_out, a, self.norm_out.num_channels, want_f32=not tail_bf16, want_bf16=tail_bf16)
        return A(f32=of, bf16=ob, C=Cz), (saved if save else None)

    def bwd(self, g_bf16, saved, need_input_grad=False):
        h_f32, st, a = saved["out"]
        d_a = conv_bwd(self.conv_out, g_bf16, a, self.norm_out.num_channels, dgrad="bf16")
        g = norm_act_bwd(self.norm_out, h_f32, st, d_a, None, self.act_out.code)
        g = self.mid2.bwd(g, saved["mid2"])
        if self.mid_attn:
            g = self.mid_attn1.bwd(g, saved["attn"])
        g = self.mid1.bwd(g, saved["mid1"])
        for down, s in zip(reversed(self.downs), reversed(saved["levels"])):
            g = down.bwd(g, s)
        return conv_bwd(self.conv_in, g, saved["x"], self.in_channels, dgrad=("f32" if need_input_grad else None))

    def forward(self, x):
        """NCHW fp32 in -> NCHW fp32 out (src/model.py:410-431), differentiable."""
        return _ModuleFn.apply(x, self, *[p for p in self.parameters()])


class Decoder(nn.Module):
    def __init__(self, shape, chs=[48, 96, 192], attn_sizes=[], mid_attn=False, num_res_blocks=1, dropout_prob=0.0,
                 z_channels=4, double_z=True, n_attention_heads=1, norm_groups=8, norm_eps=1e-6, norm_affine=True,
                 act="gelu", conv_kernel_size=3, conv_padding_mode="zeros"):
        super().__init__()
        norm_params, get_act, conv_params, resnet_params = _enc_dec_common(
            self, shape, chs, attn_sizes, mid_attn, num_res_blocks, dropout_prob, z_channels, double_z,
            n_attention_heads, norm_groups, norm_eps, norm_affine, act, conv_kernel_size, conv_padding_mode)
        ch_in = self.chs[-1]
        self.conv_in = get_conv(self.z_channels, ch_in, dim=self.dim, **conv_params)
        self.mid1 = ResNetBlock(ch_in, ch_in, **resnet_params)
        if self.mid_attn:
            self.mid_attn1 = AttnBlock(ch_in, n_heads=self.n_attention_heads, dim=self.dim, norm_params=norm_params)
        self.mid2 = ResNetBlock(ch_in, ch_in, **resnet_params)
        curr_size = self.input_size // 2 ** (self.n_sizes - 1)
        self.ups = nn.ModuleList()
        for i_level in reversed(range(self.n_sizes)):
            ch_in = self.chs[i_level]
            resnets = nn.ModuleList()
            attentions = nn.ModuleList()
            for _ in range(self.num_res_blocks):
                resnets.append(ResNetBlock(ch_in, ch_in, **resnet_params))
                if curr_size in self.attn_sizes:
                    attentions.append(AttnBlock(ch_in, n_heads=self.n_attention_heads, dim=self.dim,
                                                norm_params=norm_params))
            if len(attentions) == 0:
                attentions = None
            ch_out = self.chs[0] if i_level == 0 else self.chs[i_level - 1]
            up = ResNetUp(ch_out=ch_out, resnet_blocks=resnets, attention_blocks=attentions)
            curr_size = curr_size // 2
            self.ups.append(up)
        self.norm_out = _make_norm(ch_out, norm_params)
        self.act_out = get_act()
        self.conv_out = get_conv(in_channels=ch_out, out_channels=self.in_channels, dim=self.dim, init=zero_init,
                                 **conv_params)

    def fwd(self, z_bf16, save, nll=None):
        self.last_z_shape = (z_bf16.shape[0], self.z_channels, z_bf16.shape[1], z_bf16.shape[2])
        saved = {}
        r = conv_fwd(self.conv_in, z_bf16, self.z_channels, stats_for=self.mid1.net1[0])
        saved["z"] = z_bf16 if save else None
        h = A(f32=r[0], C=self.chs[-1], stats=r.stats)
        h, saved["mid1"] = self.mid1.fwd(h, save, next_norm=(self.mid_attn1.norm if self.mid_attn else self.mid2.net1[0]))
        if self.mid_attn:
            h, saved["attn"] = self.mid_attn1.fwd(h, save, next_norm=self.mid2.net1[0])
        h, saved["mid2"] = self.mid2.fwd(h, save, next_norm=self.ups[0].resnet_blocks[0].net1[0])
        levels = []
        for i, up in enumerate(self.ups):
            last = i == self.n_sizes - 1
            nxt_norm = self.norm_out if last else self.ups[i + 1].resnet_blocks[0].net1[0]
            h, s = up.fwd(h, save, no_up=last, next_norm=nxt_norm)
            levels.append(s)
        saved["levels"] = levels
        a, st = norm_act_fwd(self.norm_out, h.f32, self.act_out.code, h.stats, save)
        saved["out"] = (h.f32, st, a) if save else None
        if nll is not None:     # get_loss(): the loss and its gradient come out of this conv's epilogue (tvae_conv_args.nll_*)
            Cout = self.conv_out.out_channels
            kind, R = self.conv_out.conv_kind()
            _, g = ops.conv_gemm(a, self.norm_out.num_channels, self.conv_out.packed("fwd"), kind=kind, R=R, Cout=Cout,
                                 bias=self.conv_out.bias, want_f32=False, want_bf16=True, bf16_pitch=ops.round_up(Cout, 8),
                                 nll=nll)
            nll["dxhat"] = g
            return A(C=self.in_channels), (saved if save else None)
        of, _ = conv_fwd(self.conv_out, a, self.norm_out.num_channels)
        return A(f32=of, C=self.in_channels), (saved if save else None)

    def bwd(self, g_bf16, saved, input_grad="bf16"):
        h_f32, st, a = saved["out"]
        d_a = conv_bwd(self.conv_out, g_bf16, a, self.norm_out.num_channels, dgrad="bf16")
        g = norm_act_bwd(self.norm_out, h_f32, st, d_a, None, self.act_out.code)
        for up, s in zip(reversed(self.ups), reversed(saved["levels"])):
            g = up.bwd(g, s)
        g = self.mid2.bwd(g, saved["mid2"])
        if self.mid_attn:
            g = self.mid_attn1.bwd(g, saved["attn"])
        g = self.mid1.bwd(g, saved["mid1"])
        return conv_bwd(self.conv_in, g, saved["z"], self.z_channels, dgrad=input_grad)

    def forward(self, z):
        """NCHW fp32 in -> NCHW fp32 out (src/model.py:552-574), differentiable."""
        return _ModuleFn.apply(z, self, *[p for p in self.parameters()])


# =================================================================================================== autograd glue
def _needs_grad(mod_params):
    return torch.is_grad_enabled() and any(p.requires_grad for p in mod_params)


def _check_input(x, C):
    if not torch.is_tensor(x) or x.dim() != 4:
        raise TvaeError(f"expected a 4-D NCHW tensor, got {type(x).__name__} with shape {getattr(x, 'shape', None)}")
    ops.require_cuda(x, "input")
    if x.shape[1] != C:
        raise TvaeError(f"expected {C} input channels, got {x.shape[1]}")


class _ModuleFn(torch.autograd.Function):
    """Generic NCHW-in / NCHW-out node over an engine program (Encoder, Decoder, _EncodeTail, _DecodeHead)."""

    @staticmethod
    def forward(ctx, x, mod, *params):
        _check_input(x, mod.in_channels_api())
        ENGINE.begin_forward()
        save = any(ctx.needs_input_grad)
        xb = ops.input_nhwc_bf16(x)
        out, saved = mod.program_fwd(xb, save)
        ctx.mod, ctx.saved = mod, saved
        ctx.x_needs_grad = x.requires_grad
        ctx.nparams = len(params)
        return ops.nhwc_to_nchw_f32(out.f32, out.C)

    @staticmethod
    def backward(ctx, g):
        mod = ctx.mod
        if ctx.saved is None:
            raise TvaeError("backward through a forward that ran without saved activations")
        gb = ops.nchw_to_nhwc_bf16(g.contiguous())
        dx = mod.program_bwd(gb, ctx.saved, ctx.x_needs_grad)
        ctx.saved = None
        ENGINE.join_side_streams()
        gx = None
        if ctx.x_needs_grad:
            gx = ops.nhwc_to_nchw_f32(dx, mod.in_channels_api())
        return (gx, None) + (None,) * ctx.nparams


def _enc_in_channels(self):
    return self.in_channels


Encoder.in_channels_api = _enc_in_channels
Encoder.program_fwd = lambda self, xb, save: self.fwd(xb, save)
Encoder.program_bwd = lambda self, gb, saved, need: self.bwd(gb, saved, need_input_grad=need)
Decoder.in_channels_api = lambda self: self.z_channels
Decoder.program_fwd = lambda self, zb, save: self.fwd(zb, save)
Decoder.program_bwd = lambda self, gb, saved, need: self.bwd(gb, saved, input_grad=("f32" if need else None))


class _EncodeProgram:
    """encoder + quant_conv (AutoencoderKL.encode, src/model.py:634-638) as one engine program."""

    def __init__(self, vae):
        self.vae = vae

    def in_channels_api(self):
        return self.vae.encoder.in_channels

    def program_fwd(self, xb, save):
        v = self.vae
        h, s = v.encoder.fwd(xb, save, tail_bf16=True)
        mom, _ = conv_fwd(v.quant_conv, h.bf16, h.C)
        return A(f32=mom, C=v.quant_conv.out_channels), ((s, h.bf16) if save else None)

    def program_bwd(self, gb, saved, need):
        v = self.vae
        s, hb = saved
        d_h = conv_bwd(v.quant_conv, gb, hb, v.quant_conv.in_channels, dgrad="bf16")
        return v.encoder.bwd(d_h, s, need_input_grad=need)


class _DecodeProgram:
    """post_quant_conv + decoder (AutoencoderKL.decode, src/model.py:640-643)."""

    def __init__(self, vae):
        self.vae = vae

    def in_channels_api(self):
        return self.vae.post_quant_conv.in_channels

    def program_fwd(self, zb, save, nll=None):
        v = self.vae
        _, pq = conv_fwd(v.post_quant_conv, zb, v.post_quant_conv.in_channels, want_f32=False, want_bf16=True)
        out, s = v.decoder.fwd(pq, save, nll=nll)
        return out, ((s, zb) if save else None)

    def program_bwd(self, gb, saved, need):
        v = self.vae
        s, zb = saved
        d_pq = v.decoder.bwd(gb, s, input_grad="bf16")
        return conv_bwd(v.post_quant_conv, d_pq, zb, v.post_quant_conv.in_channels, dgrad=("f32" if need else None))


class _VAELossFn(torch.autograd.Function):
    """The training hot path: AutoencoderKL.get_loss (src/model.py:654-669) as ONE autograd node.

    forward : NCHW->NHWC bf16, encoder, quant_conv, fused reparam+KL, post_quant_conv, decoder, fused NLL (+ its
              gradient wrt the reconstruction), loss scalars.
    backward: decoder, fused reparam backward (KL gradient folded in), encoder; parameter gradients are written
              into param.grad as each layer finishes (reverse forward order, so bucketed all-reduce can overlap).
    """

    @staticmethod
    def forward(ctx, x, eps, vae, extra, *params):
        _check_input(x, vae.encoder.in_channels)
        ENGINE.begin_forward()
        B = x.shape[0]
        train = any(ctx.needs_input_grad)
        C = vae.encoder.in_channels
        Z = vae.embed_dim
        xb = ops.input_nhwc_bf16(x)
        enc, dec = _EncodeProgram(vae), _DecodeProgram(vae)
        mom, enc_saved = enc.program_fwd(xb, train)
        if eps is None:
            off = ENGINE.rng_offset + extra.get("sample_offset", 0)
            ENGINE.rng_offset += extra.get("global_batch", B)
            z_bf16, _, eps_used, kl = ops.reparam_fwd(mom.f32, Z, seed=ENGINE.rng_seed, sample_offset=off)
        else:
            z_bf16, _, eps_used, kl = ops.reparam_fwd(mom.f32, Z, eps=eps)
        loss_type = 0 if vae.nll_loss_type == "l1" else 1
        if (ENGINE.fuse_nll and train and not extra.get("keep_xhat") and not ops.SPLIT_BF16[0]
                and vae.decoder.conv_out.conv_kind()[0] == 0):
            # training: loss sums and d loss / d reconstruction straight from decoder.conv_out's epilogue
            nll = {"x": xb, "loss_type": loss_type, "logvar": vae.logvar.detach(), "batch": B}
            xhat, dec_saved = dec.program_fwd(z_bf16, train, nll=nll)
            sums, dxhat = nll["sums"], nll["dxhat"]
            # conv_out's bias gradient = column sums of the gradient just written. The l1 gradient is +-g with g =
            # exp(-logvar) / B ROUNDED TO BF16 in every element, so its column sums carry that one rounding as a common
            # factor (up to 2^-9): divide it out (tvae_nll_fwd sums the fp32 values; l2 gradients round independently).
            cs = torch.empty((C,), dtype=torch.float32, device=dxhat.device)
            ops.colsum_bf16(dxhat, C, cs)
            if loss_type == 0:
                gsc = torch.exp(-vae.logvar.detach().float().reshape(-1)[:1]) / B
                cs.mul_(gsc / gsc.to(torch.bfloat16).float())
            dxhat.tvae_colsum = cs
        else:
            xhat, dec_saved = dec.program_fwd(z_bf16, train)
            sums, dxhat = ops.nll_fwd(xb, xhat.f32, C, loss_type, vae.logvar.detach(), B, train)
        n_elem = float(x.numel())
        scal = ops.vae_loss_finalize(sums, kl, vae.logvar.detach(), n_elem, vae.kl_weight)
        # optional L2-product supervision on a SECOND posterior sample (src/model_with_l2.py:124-168)
        l2 = extra.get("l2")
        l2_state = None
        if l2 is not None:
            if l2.get("eps2") is None:
                off2 = ENGINE.rng_offset + extra.get("sample_offset", 0)
                ENGINE.rng_offset += extra.get("global_batch", B)
                z2, _, eps2, _ = ops.reparam_fwd(mom.f32, Z, seed=ENGINE.rng_seed ^ 0x5bd1e995, sample_offset=off2)
            else:
                z2, _, eps2, _ = ops.reparam_fwd(mom.f32, Z, eps=l2["eps2"])
            pred, head_saved = l2["head"].fwd(z2, train)
            h, w = pred.shape[1], pred.shape[2]
            l2sums = ops.l2head_loss_fwd(pred, l2["targets"], B, h, w)
            total = ops.l2head_finalize(l2sums, l2["weights"], scal)
            l2_state = (l2, eps2, pred, head_saved, l2sums, h, w)
            extra["l2_out"] = total          # fp32 [1 + nprod]: total loss, per-product masked MSE (NaN = skipped)
            extra["l2_pred"] = pred
            extra["z2_bf16"] = z2
            loss_out = total[0].clone()
        else:
            loss_out = scal[0].clone()
        ctx.vae = vae
        ctx.state = (enc, dec, enc_saved, dec_saved, mom.f32, eps_used, dxhat, scal, B, l2_state) if train else None
        ctx.nparams = len(params)
        ctx.x_needs_grad = x.requires_grad
        extra["scalars"] = scal
        extra["moments"] = mom.f32
        extra["eps"] = eps_used
        extra["z_bf16"] = z_bf16
        if extra.get("keep_xhat"):
            extra["xhat"] = xhat
        return loss_out

    @staticmethod
    def backward(ctx, g):
        vae = ctx.vae
        if ctx.state is None:
            raise TvaeError("backward through get_loss() that ran without grad")
        enc, dec, enc_saved, dec_saved, mom, eps_used, dxhat, scal, B, l2_state = ctx.state
        ctx.state = None
        gs = 1.0
        if not ENGINE.unit_loss_grad:
            gs = float(g.item())
            if gs != 1.0:
                dxhat.mul_(gs)
        dz = dec.program_bwd(dxhat, dec_saved, True)                     # fp32 NHWC [B,h,w,Z]
        dz2 = eps2 = None
        if l2_state is not None:
            l2, eps2, pred, head_saved, l2sums, h, w = l2_state
            dpred = ops.l2head_loss_bwd(pred, l2["targets"], B, h, w, l2sums, l2["weights"], gs)
            dz2 = l2["head"].bwd(dpred, head_saved)                      # fp32 NHWC [B,h,w,Z]
        dm = ops.reparam_bwd(mom, vae.embed_dim, dz, eps_used, dz2, eps2, vae.kl_weight / B * gs)
        enc.program_bwd(dm, enc_saved, False)
        if vae.logvar.requires_grad:
            _write_grad(vae.logvar, lambda dst: dst.copy_(scal[4] * gs if gs != 1.0 else scal[4]))
        if ctx.x_needs_grad:
            raise TvaeError("gradient with respect to the input of get_loss() is not implemented")
        ENGINE.join_side_streams()
        return (None, None, None, None) + (None,) * ctx.nparams


# =================================================================================================== main VAE
class AutoencoderKL(nn.Module):
    def __init__(self, enc_dec_params, embed_dim=8, learning_rate=1e-3, weight_decay=1.0e-5, nll_loss_type="l1",
                 kl_weight=0.000001, no2_weight=0.0, no2_mlp_hidden=None, **kwargs):
        super().__init__()
        self.enc_dec_params = enc_dec_params
        self.encoder = Encoder(**self.enc_dec_params)
        self.decoder = Decoder(**self.enc_dec_params)
        self.dim = self.encoder.dim
        self.embed_dim = embed_dim
        self.learning_rate = learning_rate
        self.weight_decay = weight_decay
        self.nll_loss_type = nll_loss_type
        assert self.nll_loss_type in ["l1", "l2"], "nll_loss_type must be l1 or l2"
        self.kl_weight = kl_weight
        self.no2_weight = no2_weight
        z_channels = self.encoder.z_channels
        self.quant_conv = get_conv(2 * z_channels, 2 * self.embed_dim, dim=self.dim, kernel_size=1, padding=0)
        self.post_quant_conv = get_conv(self.embed_dim, z_channels, dim=self.dim, kernel_size=1, padding=0)
        self.logvar = nn.Parameter(torch.ones(size=(), dtype=torch.float32) * 6.0)
        self.no2_probe = None
        if no2_mlp_hidden is not None and no2_weight > 0:
            raise TvaeError("the legacy in-model NO2 probe (no2_mlp_hidden) is unused by the shipped configs and "
                            "not implemented; use VAEWithL2Supervision")
        self._last = {}

    # -- API -----------------------------------------------------------------------------------------------
    def encode(self, x):
        moments = _ModuleFn.apply(x, _EncodeProgram(self), *self._enc_params())
        return DiagonalGaussianDistribution(moments)

    def decode(self, z):
        return _ModuleFn.apply(z, _DecodeProgram(self), *self._dec_params())

    def forward(self, input, sample_posterior=True, eps=None):
        posterior = self.encode(input)
        z = posterior.sample(eps) if sample_posterior else posterior.mode()
        dec = self.decode(z)
        return dec, posterior

    def get_loss(self, x, eps=None, **extra):
        """Returns (loss, {"kl_loss","nll_loss","loss"}) like src/model.py:654-669. `eps` optionally injects the
        reparameterisation noise (NCHW [B, embed_dim, h, w]); by default it is drawn on the device."""
        extra = dict(extra)
        loss = _VAELossFn.apply(x, eps, self, extra, *[p for p in self.parameters()])
        scal = extra["scalars"]
        self._last = extra
        metrics = {"kl_loss": scal[2], "nll_loss": scal[1], "loss": loss}
        return loss, metrics

    def last_pixel_mse(self):
        """mean((x - recon)^2) of the most recent get_loss() forward (device scalar)."""
        return self._last["scalars"][3]

    def predict_no2(self, x):
        raise ValueError("NO2 probe not initialized")

    # -- helpers -------------------------------------------------------------------------------------------
    def _enc_params(self):
        return [p for p in self.encoder.parameters()] + [p for p in self.quant_conv.parameters()]

    def _dec_params(self):
        return [p for p in self.post_quant_conv.parameters()] + [p for p in self.decoder.parameters()]


class SpectralVAE(nn.Module):
    """Simple wrapper for TEMPO spectral data (src/model.py:684-705)."""

    def __init__(self, vae):
        super(SpectralVAE, self).__init__()
        self.vae = vae

    def forward(self, x):
        x_rec, _ = self.vae(x)
        return x_rec

    def get_latent(self, x):
        """The reference runs the full decode here only to return the posterior (src/model.py:695-697);
        the posterior does not depend on it, so the decode is skipped."""
        return self.vae.encode(x)

    def get_loss(self, x, **kw):
        loss, metrics = self.vae.get_loss(x, **kw)
        return loss, metrics

    def get_metrics(self, x):
        _, metrics = self.vae.get_loss(x)
        return metrics


DEFAULT_ENC_DEC = dict(
    shape=(1028, 64, 64),
    chs=[512, 256, 128],
    attn_sizes=[],
    mid_attn=True,
    num_res_blocks=1,
    dropout_prob=0.0,
    z_channels=32,
    double_z=True,
    n_attention_heads=4,
    norm_groups=8,
    norm_eps=1e-6,
    norm_affine=True,
    act="gelu",
    conv_kernel_size=3,
    conv_padding_mode="zeros",
)


def get_model(model_params, device):
    """Same contract as src/model.py:708-759: builds SpectralVAE(AutoencoderKL) on `device` and attaches an AdamW
    optimiser as `model.optimizer` (here the fused flat-buffer FusedAdamW, state_dict-compatible with
    torch.optim.AdamW)."""
    assert model_params["architecture_type"] == "vae"
    enc_dec_params = {k: (list(v) if isinstance(v, (list, tuple)) and k != "shape" else v)
                      for k, v in DEFAULT_ENC_DEC.items()}
    config_params = model_params["architecture_params"]["enc_dec_params"]
    for key in enc_dec_params.keys():
        if key in config_params:
            enc_dec_params[key] = config_params[key]
    embed_dim = config_params.get("embed_dim", 32)
    kl_weight = config_params.get("kl_weight", 0.000001)
    nll_loss_type = config_params.get("nll_loss_type", "l1")
    no2_weight = config_params.get("no2_weight", 0.0)
    no2_mlp_hidden = config_params.get("no2_mlp_hidden", None)
    vae = AutoencoderKL(enc_dec_params=enc_dec_params, embed_dim=embed_dim, learning_rate=1e-3, weight_decay=1.0e-5,
                        nll_loss_type=nll_loss_type, kl_weight=kl_weight, no2_weight=no2_weight,
                        no2_mlp_hidden=no2_mlp_hidden)
    model = SpectralVAE(vae)
    device = torch.device(device)
    if device.type != "cuda":
        raise TvaeError(f"get_model(device={device}): the B200 engine runs on CUDA only; there is no CPU fallback")
    model = model.to(device)
    assert model_params["optimizer_type"] == "AdamW"
    from .optim import FusedAdamW
    optimizer = FusedAdamW(model.parameters(), **model_params["optimizer_params"])
    model.optimizer = optimizer
    return model
