"""Tensor-level wrappers over the C ABI (include/tvae.h).

PyTorch is used here for device memory and streams only; every arithmetic op below is a call into
libtvae_b200.so on torch's current CUDA stream. Activations are NHWC tensors `[N, H, W, pitch]` (bf16 operands,
fp32 residual stream); `C` (the number of valid channels, <= pitch) travels next to them.
"""
import ctypes as C
import functools
import os

import torch

from . import _lib
from ._lib import ConvArgs, WgradArgs, lib
from ._lib import check as _check


def _stream():
    """Current stream of the CURRENT device; every tensor-taking wrapper below runs under `_on_device`, which makes the
    device that owns its operands current for the duration of the call (the C ABI launches on the calling thread's
    current device and never switches it)."""
    return torch.cuda.current_stream().cuda_stream


def _device_of(args):
    for a in args:
        if torch.is_tensor(a):
            if a.is_cuda:
                return a.device
        elif isinstance(a, Pair):
            return a.hi.device
        elif isinstance(a, (list, tuple)) and a and isinstance(a[0], tuple) and torch.is_tensor(a[0][0]):
            return a[0][0].device                    # pack_weights_batched(items)
    return None


def _on_device(fn):
    """Run `fn` with the device of its first CUDA tensor argument current (no-op when it already is). The reference's
    torch ops work from any current device (`model.to('cuda:3')` while the current device is 0); so must these."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = _device_of(args)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def check(rc, what):
    _check(rc, what)
    KERNEL_LAUNCHES[0] += _KERNELS_PER_CALL.get(what, 0)


# bench.py hooks: PROFILE["conv"|"wgrad"] = {"match": fn(pixels, Cout|Cm, Cin|Cn, kind, R) -> bool, "events": []}
# brackets the matching launches with CUDA events on the launching stream (per-kernel roofline numbers)
PROFILE = {}

KERNEL_LAUNCHES = [0]   # kernels of libtvae_b200.so enqueued through this module (bench.py reports the delta)


def _timed(name, nbytes, call):
    """bench.py hook for the HBM-bound kernels: PROFILE["hbm"] = {"min_bytes": n, "events": {}} brackets every call
    whose ALGORITHMIC byte count (operands read once + results written once) is at least min_bytes with CUDA events
    on the launching stream and records (start, stop, bytes) under `name`."""
    prof = PROFILE.get("hbm")
    if prof is None or nbytes < prof.get("min_bytes", 0):
        return call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = call()
    e1.record()
    prof["events"].setdefault(name, []).append((e0, e1, nbytes))
    return r

# kernels launched by one call of each C-ABI entry point
_KERNELS_PER_CALL = {
    "tvae_pack_weight": 1, "tvae_pack_weights_batched": 1, "tvae_conv_gemm": 1, "tvae_wgrad_gemm": 2, "tvae_nchw_f32_to_nhwc_bf16": 1,
    "tvae_nhwc_f32_to_nchw_f32": 1, "tvae_nhwc_f32_to_nhwc_bf16": 1, "tvae_normalize_radiance": 1, "tvae_recon_metrics": 2, "tvae_nhwc_bf16_to_nchw_f32": 1, "tvae_f32_to_bf16": 1, "tvae_gn_stats": 1,
    "tvae_gn_act_fwd": 1, "tvae_gn_stats_finalize": 1, "tvae_gn_act_bwd": 4, "tvae_colsum_bf16": 2, "tvae_attn_fwd": 1, "tvae_attn_bwd": 2, "tvae_attn_fwd_tc": 1, "tvae_attn_bwd_tc": 2, "tvae_wgrad_skinny": 2,
    "tvae_reparam_fwd": 1, "tvae_reparam_bwd": 1, "tvae_nll_fwd": 2, "tvae_vae_loss_finalize": 1,
    "tvae_l2head_loss_fwd": 1, "tvae_l2head_loss_bwd": 1, "tvae_l2head_finalize": 1, "tvae_sumsq": 2, "tvae_adamw": 1,
    "tvae_gather_rows": 1, "tvae_extract_tiles": 1, "tvae_spectrum_stats_accum": 2, "tvae_spectrum_stats_finalize": 1,
    "tvae_batch_stats": 2, "tvae_act_dropout_fwd": 1, "tvae_act_dropout_bwd": 1, "tvae_probe_mse": 1,
    "tvae_nan_moments": 1, "tvae_select_hist": 1, "tvae_component_pool": 1,
}


# mirror of the library's wgrad scheduling switch (only used to count launches correctly)
WGRAD_CTA_PAIR = [os.environ.get("TVAE_WGRAD_CTA_PAIR", "1") != "0"]


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def round_up(v, m):
    return (v + m - 1) // m * m


# "fp32 mode": tensor-core operands carried as split-bf16 pairs (hi = bf16(v), lo = bf16(v - hi)); the conv GEMM then
# accumulates hi*hi + hi*lo + lo*hi in fp32 (~2^-16 relative product error, 3x the tensor work). Forward only:
# backward GEMMs use the hi halves. Toggled by tempo_vae_b200.set_precision().
SPLIT_BF16 = [False]


class Pair:
    """A split-bf16 operand: two same-shaped bf16 tensors."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape

    @property
    def device(self):
        return self.hi.device

    def view(self, *shape):
        return Pair(self.hi.view(*shape), self.lo.view(*shape))

    def __getitem__(self, idx):
        return Pair(self.hi[idx], self.lo[idx])


def hi_of(x):
    return x.hi if isinstance(x, Pair) else x


def _lo_like(t):
    return torch.empty_like(t) if SPLIT_BF16[0] else None


def _pair(hi, lo):
    return Pair(hi, lo) if lo is not None else hi


def pitch_of(t):
    """Channel pitch (elements between consecutive pixels) of an NHWC / [rows, C] tensor or channel-slice view."""
    assert t.stride(-1) == 1, "channels must be the contiguous dimension"
    return t.stride(-2)


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise _lib.TvaeError(
            f"{name} is on {t.device}: the TEMPO-VAE B200 path runs on CUDA only (there is no CPU fallback)")


# ----------------------------------------------------------------------------------------------- weight packing
class PackedWeight:
    """bf16 K-major GEMM operand [rows][k_pitch] built from a parameter in the reference's layout."""

    __slots__ = ("data", "lo", "rows", "k_pitch", "c_pad", "version")

    def __init__(self, data, rows, k_pitch, c_pad, lo=None):
        self.data, self.rows, self.k_pitch, self.c_pad = data, rows, k_pitch, c_pad
        self.lo = lo            # low-order half (split-bf16 mode) or None
        self.version = -1


# mode -> how the GEMM operand is cut out of the parameter
#  "fwd"        Conv2d [Co][Ci][R][S]     -> [Co][tap][Ci_pad]            (forward)
#  "dgrad"      Conv2d                    -> [Ci][tap][Co_pad]            (stride-1 dgrad, taps flipped by the kernel)
#  "down_dgrad" Conv2d 2x2 s2             -> [(tap, Ci)][Co_pad]          (dgrad as transposed-conv GEMM)
#  "up_fwd"     ConvT  [Ci][Co][2][2]     -> [(tap, Co)][Ci_pad]          (forward)
#  "up_dgrad"   ConvT                     -> [Ci][tap][Co_pad]            (dgrad as 2x2 s2 conv)
def pack_geometry(shape, mode):
    if len(shape) == 2:                # nn.Linear [out, in] = a 1x1 convolution (probes)
        shape = tuple(shape) + (1, 1)
    a, b, r, s = shape
    taps = r * s
    if mode == "fwd":
        return dict(Crow=a, TR=1, TK=taps, C=b, s_row=b * taps, s_col=taps, s_tap=1)
    if mode == "dgrad":
        return dict(Crow=b, TR=1, TK=taps, C=a, s_row=taps, s_col=b * taps, s_tap=1)
    if mode == "down_dgrad":
        return dict(Crow=b, TR=taps, TK=1, C=a, s_row=taps, s_col=b * taps, s_tap=1)
    if mode == "up_fwd":
        return dict(Crow=b, TR=taps, TK=1, C=a, s_row=taps, s_col=b * taps, s_tap=1)
    if mode == "up_dgrad":
        return dict(Crow=a, TR=1, TK=taps, C=b, s_row=b * taps, s_col=taps, s_tap=1)
    raise ValueError(mode)


@_on_device
def pack_weight(w, mode, out=None):
    require_cuda(w, "weight")
    g = pack_geometry(tuple(w.shape), mode)
    c_pad = round_up(g["C"], 64)
    rows = g["TR"] * g["Crow"]
    k_pitch = g["TK"] * c_pad
    if out is None:
        out = PackedWeight(torch.empty((rows, k_pitch), dtype=torch.bfloat16, device=w.device), rows, k_pitch, c_pad)
    if SPLIT_BF16[0] and out.lo is None:
        out.lo = torch.empty_like(out.data)
    wc = w.detach()
    assert wc.is_contiguous() and wc.dtype == torch.float32
    check(lib.tvae_pack_weight(wc.data_ptr(), out.data.data_ptr(), g["Crow"], g["TR"], g["TK"], g["C"], c_pad,
                               g["s_row"], g["s_col"], g["s_tap"], _ptr(out.lo) if SPLIT_BF16[0] else 0, _stream()),
          "tvae_pack_weight")
    return out


class PackDesc(C.Structure):      # tvae_pack_desc
    _fields_ = [("w", C.c_void_p), ("out_bf16", C.c_void_p), ("Crow", C.c_int32), ("TR", C.c_int32), ("TK", C.c_int32),
                ("C", C.c_int32), ("c_pad", C.c_int32), ("reserved", C.c_int32), ("s_row", C.c_int64),
                ("s_col", C.c_int64), ("s_tap", C.c_int64)]


_pack_tables = {}      # signature of (weight ptr, pack ptr, mode) triples -> (descs tensor, block_start tensor, n, total)


@_on_device
def pack_weights_batched(items):
    """items: list of (weight parameter, mode, PackedWeight). Rebuilds every pack with ONE kernel launch
    (tvae_pack_weights_batched). The descriptor table lives on the device and is reused while the pointers stay the
    same (they do: parameters sit in FusedAdamW's flat buffer, packs are persistent)."""
    dev = items[0][0].device
    sig = tuple((w.data_ptr(), ent.data.data_ptr(), mode) for w, mode, ent in items)
    tab = _pack_tables.get(dev)
    if tab is None or tab[0] != sig:
        chunk = lib.tvae_pack_chunk_elems()
        descs = (PackDesc * len(items))()
        starts = [0]
        for i, (w, mode, ent) in enumerate(items):
            g = pack_geometry(tuple(w.shape), mode)
            assert w.is_contiguous() and w.dtype == torch.float32
            d = descs[i]
            d.w, d.out_bf16 = w.data_ptr(), ent.data.data_ptr()
            d.Crow, d.TR, d.TK, d.C, d.c_pad = g["Crow"], g["TR"], g["TK"], g["C"], ent.c_pad
            d.s_row, d.s_col, d.s_tap = g["s_row"], g["s_col"], g["s_tap"]
            total = g["TR"] * g["Crow"] * g["TK"] * ent.c_pad
            starts.append(starts[-1] + (total + chunk - 1) // chunk)
        raw = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(dev)
        bs = torch.tensor(starts, dtype=torch.int64).to(dev)
        tab = _pack_tables[dev] = (sig, raw, bs, len(items), starts[-1])
    _, raw, bs, n, total_blocks = tab
    check(lib.tvae_pack_weights_batched(raw.data_ptr(), bs.data_ptr(), n, total_blocks, _stream()),
          "tvae_pack_weights_batched")


# ----------------------------------------------------------------------------------------------- conv GEMM
def fused_stats_ok(N, oH, oW, Cout, G, kind, H, W):
    """Can the conv epilogue produce the GroupNorm statistics of its output? (see tvae_conv_args.stats_part)"""
    if G <= 0 or Cout % G:
        return False
    gs = Cout // G
    bn = Cout if Cout <= 256 else 256
    grid = H * W if kind == 2 else oH * oW          # the GEMM's pixel grid
    return gs % 16 == 0 and Cout % 16 == 0 and (Cout <= 256 or Cout % 256 == 0) and bn % gs == 0 \
        and bn // gs <= 16 and grid % 128 == 0


@_on_device
def conv_gemm(x, C_in, wp, *, kind, R, Cout, flip=False, bias=None, residual=None, want_f32=True, want_bf16=False,
              bf16_pitch=None, bn=0, out_f32=None, out_bf16=None, stats=None, split_out=True, nll=None):
    """x: bf16 [N,H,W,pitch]. Returns (out_f32 or None, out_bf16 or None) as NHWC tensors; with stats=(G, eps) the
    GroupNorm statistics [N, G, 2] of the output are produced by the epilogue and returned as a third value
    (None when the geometry does not allow it). nll = {"x": target bf16 NHWC, "loss_type", "logvar", "batch"}: the fused
    reconstruction loss of tvae_conv_args.nll_* -- the bf16 output is then the loss gradient wrt the reconstruction and
    nll["sums"] receives the fp64[3] sums of tvae_nll_fwd."""
    x_lo = x.lo if isinstance(x, Pair) else None
    x = hi_of(x)
    N, H, W, _ = x.shape
    pitch = pitch_of(x)
    if kind == 1:
        oH, oW = H // 2, W // 2
    elif kind == 2:
        oH, oW = 2 * H, 2 * W
    else:
        oH, oW = H, W
    dev = x.device
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty((N, oH, oW, round_up(Cout, 4)), dtype=torch.float32, device=dev)
    if want_bf16 and out_bf16 is None:
        bp = bf16_pitch or round_up(Cout, 8)
        out_bf16 = torch.empty((N, oH, oW, bp), dtype=torch.bfloat16, device=dev)
    a = ConvArgs()
    a.x = x.data_ptr(); a.N, a.H, a.W, a.C, a.x_pitch = N, H, W, C_in, pitch
    a.kind, a.R, a.flip = kind, R, int(flip)
    a.w = wp.data.data_ptr(); a.w_rows, a.k_pitch, a.c_pad = wp.rows, wp.k_pitch, wp.c_pad
    a.Cout = Cout
    a.bias = _ptr(bias)
    a.residual = _ptr(residual); a.res_pitch = pitch_of(residual) if residual is not None else 0
    a.out_f32 = _ptr(out_f32); a.out_f32_pitch = pitch_of(out_f32) if out_f32 is not None else 0
    a.out_bf16 = _ptr(out_bf16); a.out_bf16_pitch = pitch_of(out_bf16) if out_bf16 is not None else 0
    a.bn = bn
    out_lo = None
    if x_lo is not None and wp.lo is not None and SPLIT_BF16[0]:
        assert pitch_of(x_lo) == pitch
        a.x_lo = x_lo.data_ptr()
        a.w_lo = wp.lo.data_ptr()
    if out_bf16 is not None and SPLIT_BF16[0] and want_bf16 and split_out:
        out_lo = torch.empty_like(out_bf16)
        a.out_bf16_lo = out_lo.data_ptr()
    if nll is not None:
        tgt = hi_of(nll["x"])
        assert kind == 0 and not flip and out_f32 is None and out_bf16 is not None and residual is None and stats is None
        assert tgt.shape[:3] == (N, oH, oW)
        nll["sums"] = torch.empty((3,), dtype=torch.float64, device=dev)
        nws = _workspace(lib.tvae_conv_nll_workspace_bytes(N * oH * oW, Cout), dev, "conv_nll")
        a.nll_x, a.nll_x_pitch = tgt.data_ptr(), pitch_of(tgt)
        a.nll_loss_type, a.nll_logvar, a.nll_batch = int(nll["loss_type"]), nll["logvar"].data_ptr(), int(nll["batch"])
        a.nll_workspace, a.nll_sums = nws.data_ptr(), nll["sums"].data_ptr()
        KERNEL_LAUNCHES[0] += 2       # the two-stage reduction of the loss partials
    part = None
    if stats is not None and fused_stats_ok(N, oH, oW, Cout, stats[0], kind, H, W):
        grid_px = (H * W) if kind == 2 else (oH * oW)
        spi = grid_px // 128 * (4 if kind == 2 else 1)
        part = torch.empty((N * spi, stats[0], 2), dtype=torch.float32, device=dev)
        a.stats_part = part.data_ptr()
        a.stats_groups = stats[0]
    prof = PROFILE.get("conv")
    if prof is not None and prof["match"](N * oH * oW if kind != 2 else N * H * W, Cout, C_in, kind, R):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.tvae_conv_gemm(C.byref(a), _stream()), "tvae_conv_gemm")
        e1.record()
        prof["events"].append((e0, e1))
    else:
        check(lib.tvae_conv_gemm(C.byref(a), _stream()), "tvae_conv_gemm")
    out_bf16 = _pair(out_bf16, out_lo)
    if stats is None:
        return out_f32, out_bf16
    st = None
    if part is not None:
        st = torch.empty((N, stats[0], 2), dtype=torch.float32, device=dev)
        check(lib.tvae_gn_stats_finalize(part.data_ptr(), spi, N, stats[0], float(oH * oW * (Cout // stats[0])),
                                         float(stats[1]), st.data_ptr(), _stream()), "tvae_gn_stats_finalize")
    return out_f32, out_bf16, st


_wgrad_ws = {}


def _workspace(nbytes, device, key="ws"):
    """Grow-only scratch buffer per (device, key, stream): reuse is ordered by the stream it is used on, so kernels on
    different streams (weight gradients on the side stream) must not share one. A replaced (too small) buffer may still
    be in use by kernels already enqueued: the caching allocator is told which stream that is."""
    stream = torch.cuda.current_stream(device)
    k = (device, key, stream.cuda_stream)
    buf = _wgrad_ws.get(k)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            buf.record_stream(stream)
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _wgrad_ws[k] = buf
    return buf


@_on_device
def wgrad_gemm(p, Cm, q, Cn, *, kind, R, grad, accumulate=False, splits=0, flip=False, grad_ld=0, grad_off=0):
    """grad[m][n][tap] (+)= sum_pixels p[pixel][m] * q[pixel (+) tap][n]; p: bf16 [N,H,W,pitch] (the dense grid).
    flip=True (stride-1 only): p = x, q = dY read at pixel (-) tap, grad written as [n][m][tap] (see include/tvae.h)."""
    p, q = hi_of(p), hi_of(q)
    N, H, W, _ = p.shape
    pp = pitch_of(p)
    taps = R * R if kind == 0 else 4
    if splits <= 0:
        splits = lib.tvae_wgrad_splits(Cm, Cn, taps, N * H * W)
    nbytes = lib.tvae_wgrad_workspace_bytes(Cm, Cn, taps, splits)
    ws = _workspace(nbytes, p.device, "wgrad")
    assert grad.is_contiguous() and grad.dtype == torch.float32
    if grad_ld:          # the GEMM fills a sub-block of the parameter's inner channel dimension (see tvae_wgrad_args)
        assert grad.numel() == (Cn if flip else Cm) * grad_ld * taps and grad_off + (Cm if flip else Cn) <= grad_ld
    else:
        assert grad.numel() == Cm * Cn * taps
    a = WgradArgs()
    a.p = p.data_ptr(); a.p_pitch = pp; a.Cm = Cm
    a.q = q.data_ptr(); a.q_pitch = pitch_of(q); a.Cn = Cn
    a.N, a.H, a.W = N, H, W
    a.kind, a.R, a.splits = kind, R, splits
    a.workspace = ws.data_ptr(); a.grad = grad.data_ptr(); a.accumulate = int(accumulate); a.flip = int(flip)
    a.grad_ld, a.grad_off = int(grad_ld), int(grad_off)
    m_tiles = (Cm + 127) // 128
    if WGRAD_CTA_PAIR[0] and m_tiles >= 3 and m_tiles % 2 == 1:
        KERNEL_LAUNCHES[0] += 2       # the odd last M tile runs as its own GEMM + reduce launch (see tvae_wgrad_gemm)
    prof = PROFILE.get("wgrad")
    if prof is not None and prof["match"](N * H * W, Cm, Cn, kind, R):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.tvae_wgrad_gemm(C.byref(a), _stream()), "tvae_wgrad_gemm")
        e1.record()
        prof["events"].append((e0, e1))
    else:
        check(lib.tvae_wgrad_gemm(C.byref(a), _stream()), "tvae_wgrad_gemm")


# ----------------------------------------------------------------------------------------------- layout
@_on_device
def nchw_to_nhwc_bf16(x, pitch=None):
    require_cuda(x, "input")
    N, Cc, H, W = x.shape
    pitch = pitch or round_up(Cc, 8)
    x = x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    out = torch.empty((N, H, W, pitch), dtype=torch.bfloat16, device=x.device)
    lo = _lo_like(out)
    _timed("nchw_to_nhwc_bf16", N * Cc * H * W * 6,
           lambda: check(lib.tvae_nchw_f32_to_nhwc_bf16(x.data_ptr(), out.data_ptr(), N, Cc, H * W, pitch, _ptr(lo),
                                                        _stream()), "tvae_nchw_f32_to_nhwc_bf16"))
    return _pair(out, lo)


@_on_device
def input_nhwc_bf16(x):
    """The engine's bf16 channels-last operand for an NCHW-SHAPED input tensor [N, C, H, W], whatever its memory
    layout: (a) bf16 with channels-last strides and a row pitch that is a multiple of 8 (what DeviceTileCache
    yields) is used in place, no kernel; (b) fp32 with channels-last strides (torch.channels_last, or a permuted
    view of [N, H, W, C] tiles) is cast row by row; (c) anything else goes through the NCHW transpose."""
    require_cuda(x, "input")
    N, Cc, H, W = x.shape
    pitch = round_up(Cc, 8)
    cl = x.stride(1) == 1 and x.stride(2) == W * x.stride(3) and x.stride(0) == H * x.stride(2) and x.stride(3) >= Cc
    if cl and x.dtype == torch.bfloat16 and not SPLIT_BF16[0] and x.stride(3) % 8 == 0 and x.data_ptr() % 16 == 0:
        p = x.stride(3)
        need = ((N * H * W - 1) * p + Cc) * 2                    # bytes actually addressed (pad lanes are never read)
        if x.untyped_storage().nbytes() - x.storage_offset() * 2 >= need:
            return x.permute(0, 2, 3, 1)                          # [N, H, W, C] view, pixel pitch p: zero copy
    if cl and x.dtype == torch.float32:
        out = torch.empty((N, H, W, pitch), dtype=torch.bfloat16, device=x.device)
        lo = _lo_like(out)
        check(lib.tvae_nhwc_f32_to_nhwc_bf16(x.data_ptr(), x.stride(3), N * H * W, Cc, out.data_ptr(), pitch,
                                             _ptr(lo), _stream()), "tvae_nhwc_f32_to_nhwc_bf16")
        return _pair(out, lo)
    return nchw_to_nhwc_bf16(x)


@_on_device
def normalize_radiance(rad, mean, std, min_radiance, clip_min, clip_max, want_f32=True, want_bf16=False):
    """rad [..., C] fp32 on the device -> (z fp32 [..., C] or None, z bf16 [..., pitch] operand rows or None)."""
    require_cuda(rad, "radiance")
    rad = rad.contiguous().float()
    Cc = rad.shape[-1]
    rows = rad.numel() // Cc
    mean = mean.to(rad.device, torch.float32).contiguous()
    std = std.to(rad.device, torch.float32).contiguous()
    if mean.numel() != Cc or std.numel() != Cc:
        raise _lib.TvaeError(f"mean/std spectra must have {Cc} channels")
    zf = torch.empty_like(rad) if want_f32 else None
    pitch = round_up(Cc, 8)
    zb = torch.empty(rad.shape[:-1] + (pitch,), dtype=torch.bfloat16, device=rad.device) if want_bf16 else None
    check(lib.tvae_normalize_radiance(rad.data_ptr(), mean.data_ptr(), std.data_ptr(), rows, Cc, float(min_radiance),
                                      float(clip_min), float(clip_max), _ptr(zf), _ptr(zb), pitch, _stream()),
          "tvae_normalize_radiance")
    return zf, zb


@_on_device
def nhwc_to_nchw_f32(x, Cc):
    x = hi_of(x)
    N, H, W, pitch = x.shape
    out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=x.device)
    if x.dtype == torch.float32:
        check(lib.tvae_nhwc_f32_to_nchw_f32(x.data_ptr(), out.data_ptr(), N, Cc, H * W, pitch, _stream()),
              "tvae_nhwc_f32_to_nchw_f32")
    else:
        check(lib.tvae_nhwc_bf16_to_nchw_f32(x.data_ptr(), out.data_ptr(), N, Cc, H * W, pitch, _stream()),
              "tvae_nhwc_bf16_to_nchw_f32")
    return out


@_on_device
def f32_to_bf16(x):
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = _lo_like(out)
    check(lib.tvae_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _ptr(lo), _stream()), "tvae_f32_to_bf16")
    return _pair(out, lo)


# ----------------------------------------------------------------------------------------------- GroupNorm
@_on_device
def gn_stats(x, Cc, G, eps):
    N, H, W, pitch = x.shape
    assert pitch == Cc and x.dtype == torch.float32
    stats = torch.empty((N, G, 2), dtype=torch.float32, device=x.device)
    check(lib.tvae_gn_stats(x.data_ptr(), N, H * W, Cc, G, float(eps), stats.data_ptr(), _stream()), "tvae_gn_stats")
    return stats


def gn_fast_ok(Cc, G):
    """Geometry of the vectorised GroupNorm kernels (gn_fast.cu); only these accept a bf16 input."""
    if G <= 0 or Cc % G or (Cc // G) % 8:
        return False
    U = Cc // 8
    return 1 <= U <= 256 and 256 % U == 0 and 256 % (Cc // G // 8) == 0


# Handing the activation derivative from forward to backward (tvae_gn_act_fwd2 / _bwd2) removes GELU' from the backward
# row-sum pass (16 of its 24 instructions per element) for 2 B/element written by the forward and read by the backward.
# Measured inside the live B=256 step (two A/B pairs, one box): gn_act_bwd 0.95 -> 0.61 ms per large call, gn_act_fwd
# 0.33 -> 0.43 ms, i.e. -2.5 ms of kernel time per step -- and the STEP got 1.2 ms slower (110.3 -> 111.6 ms). The board
# sits at its 1 kW power cap for the whole step: what bounds the step is energy, and 4 B/element of extra DRAM traffic
# costs about as much energy as the 16 instructions it saves, so the GEMMs that follow simply clock lower. OFF by
# default (TVAE_GN_SAVE_ACT_GRAD=1 turns it on); kept because on an uncapped part the trade goes the other way.
SAVE_ACT_GRAD = [os.environ.get("TVAE_GN_SAVE_ACT_GRAD", "0") == "1"]


@_on_device
def gn_act_fwd(x, stats, gamma, beta, G, act, want_act_grad=False):
    """x: the GroupNorm input, dense NHWC, fp32 or bf16 (see tvae_gn_act_fwd). With want_act_grad (training, an
    activation, the vectorised geometry, bf16 mode) returns (out, act'(y) as bf16), else out."""
    N, H, W, Cc = x.shape
    assert x.is_contiguous() and x.dtype in (torch.float32, torch.bfloat16)
    out = torch.empty((N, H, W, Cc), dtype=torch.bfloat16, device=x.device)
    lo = _lo_like(out)
    gp = None
    if want_act_grad and (SAVE_ACT_GRAD[0] or want_act_grad == "force") and act != 0 and lo is None and gn_fast_ok(Cc, G):
        gp = torch.empty((N, H, W, Cc), dtype=torch.bfloat16, device=x.device)
    _timed("gn_act_fwd", x.numel() * (x.element_size() + 2 + (2 if gp is not None else 0)),
           lambda: check(lib.tvae_gn_act_fwd2(x.data_ptr(), int(x.dtype == torch.bfloat16), stats.data_ptr(),
                                              gamma.data_ptr(), beta.data_ptr(), N, H * W, Cc, G, int(act),
                                              out.data_ptr(), _ptr(lo), _ptr(gp), _stream()), "tvae_gn_act_fwd"))
    if want_act_grad:
        return _pair(out, lo), gp
    return _pair(out, lo)


@_on_device
def gn_act_bwd(x, stats, gamma, beta, da, gres, G, act, dgamma, dbeta, dx_colsum=None, act_grad=None):
    """dx_colsum (optional fp32 [C]): receives the column sums of dx (= bias gradient of the conv that produced x).
    act_grad (optional bf16, from gn_act_fwd(want_act_grad=True)): the backward then never evaluates the activation."""
    N, H, W, Cc = x.shape
    da, gres = hi_of(da), hi_of(gres)
    assert da.shape[-1] == Cc and da.dtype == torch.bfloat16
    dx = torch.empty((N, H, W, Cc), dtype=torch.bfloat16, device=x.device)
    ws = _workspace(lib.tvae_gn_bwd_workspace_bytes(N, H * W, Cc, G), x.device, "gn")
    assert x.is_contiguous() and x.dtype in (torch.float32, torch.bfloat16)
    # algorithmic bytes of ONE pass: x + da (+ the residual-branch gradient, + the saved act') read, dx written
    nbytes = x.numel() * (x.element_size() + 2 + (2 if gres is not None else 0) + (2 if act_grad is not None else 0) + 2)
    _timed("gn_act_bwd", nbytes,
           lambda: check(lib.tvae_gn_act_bwd2(x.data_ptr(), int(x.dtype == torch.bfloat16), stats.data_ptr(),
                                              gamma.data_ptr(), beta.data_ptr(), da.data_ptr(), _ptr(gres),
                                              _ptr(act_grad), N, H * W, Cc, G, int(act), dx.data_ptr(),
                                              dgamma.data_ptr(), dbeta.data_ptr(), _ptr(dx_colsum), ws.data_ptr(),
                                              _stream()), "tvae_gn_act_bwd"))
    if dx_colsum is not None:
        KERNEL_LAUNCHES[0] += 2
    return dx


@_on_device
def colsum_bf16(x, Cc, out):
    x = hi_of(x)
    rows = x.numel() // x.shape[-1]
    ws = _workspace(lib.tvae_colsum_workspace_bytes(rows, Cc), x.device, "colsum")
    check(lib.tvae_colsum_bf16(x.data_ptr(), rows, Cc, pitch_of(x), out.data_ptr(), ws.data_ptr(), _stream()),
          "tvae_colsum_bf16")


@_on_device
def wgrad_skinny(wide, Cw, skinny, Cs, *, sign, grad, stride_c, stride_n, accumulate=False):
    """grad[c * stride_c + n * stride_n + tap] (+)= sum_pixels wide[pixel][n] * skinny[pixel + sign * tap][c] for a 3x3
    stride-1 convolution: the weight gradient of <= 4 tail channels (1028 = 8 x 128 + 4) with the nine taps on the M side of
    a 36 x Cw x pixels GEMM, so the wide operand is read once (tvae_wgrad_skinny). wide / skinny: bf16 [N,H,W,*] views."""
    wide, skinny = hi_of(wide), hi_of(skinny)
    N, H, W, _ = wide.shape
    assert skinny.shape[:3] == wide.shape[:3] and grad.dtype == torch.float32 and grad.is_contiguous()
    ws = _workspace(lib.tvae_wgrad_skinny_workspace_bytes(Cw), wide.device, "wgrad_skinny")
    check(lib.tvae_wgrad_skinny(wide.data_ptr(), Cw, pitch_of(wide), skinny.data_ptr(), Cs, pitch_of(skinny), N, H, W,
                                int(sign), grad.data_ptr(), int(stride_c), int(stride_n), int(accumulate), ws.data_ptr(),
                                _stream()), "tvae_wgrad_skinny")


# ----------------------------------------------------------------------------------------------- attention
ATTN_TENSOR_CORES = [True]     # TF32 mma.sync kernels for head dim 32; the exact fp32 kernels otherwise / in fp32 mode


def attn_uses_tensor_cores(Cc, heads):
    return ATTN_TENSOR_CORES[0] and not SPLIT_BF16[0] and Cc == 32 * heads


@_on_device
def attn_fwd(qkv, Cc, heads, B, T):
    """qkv: fp32 [B*T (any leading shape), 3C]. Returns (o_bf16, o_f32, lse)."""
    pitch = qkv.shape[-1]
    dev = qkv.device
    o_bf16 = torch.empty((B * T, Cc), dtype=torch.bfloat16, device=dev)
    o_f32 = torch.empty((B * T, Cc), dtype=torch.float32, device=dev)
    lse = torch.empty((B, heads, T), dtype=torch.float32, device=dev)
    base = qkv.data_ptr()
    if attn_uses_tensor_cores(Cc, heads):
        check(lib.tvae_attn_fwd_tc(base, base + 4 * Cc, base + 8 * Cc, pitch, B, T, Cc, heads, o_bf16.data_ptr(),
                                   o_f32.data_ptr(), lse.data_ptr(), _stream()), "tvae_attn_fwd_tc")
    else:
        check(lib.tvae_attn_fwd(base, base + 4 * Cc, base + 8 * Cc, pitch, B, T, Cc, heads, o_bf16.data_ptr(),
                                o_f32.data_ptr(), lse.data_ptr(), _stream()), "tvae_attn_fwd")
    if SPLIT_BF16[0]:
        o_bf16 = f32_to_bf16(o_f32)
    return o_bf16, o_f32, lse


@_on_device
def attn_bwd(qkv, o_f32, d_out, lse, Cc, heads, B, T):
    pitch = qkv.shape[-1]
    dev = qkv.device
    dqkv = torch.empty((B * T, 3 * Cc), dtype=torch.bfloat16, device=dev)
    ws = torch.empty((B * heads * T,), dtype=torch.float32, device=dev)
    base = qkv.data_ptr()
    fn, name = (lib.tvae_attn_bwd_tc, "tvae_attn_bwd_tc") if attn_uses_tensor_cores(Cc, heads) else \
        (lib.tvae_attn_bwd, "tvae_attn_bwd")
    check(fn(base, base + 4 * Cc, base + 8 * Cc, pitch, o_f32.data_ptr(), d_out.data_ptr(), lse.data_ptr(), B, T, Cc,
             heads, dqkv.data_ptr(), ws.data_ptr(), _stream()), name)
    return dqkv


# ----------------------------------------------------------------------------------------------- latent / losses
@_on_device
def reparam_fwd(moments, Z, *, eps=None, seed=0, sample_offset=0, want_z_nchw=False, z_pitch=None):
    """moments: fp32 [B,h,w,2Z]. Returns (z_bf16 [B,h,w,z_pitch], z_nchw or None, eps_nchw, kl[B])."""
    B, h, w, _ = moments.shape
    dev = moments.device
    z_pitch = z_pitch or round_up(Z, 8)
    z_bf16 = torch.zeros((B, h, w, z_pitch), dtype=torch.bfloat16, device=dev) if z_pitch != Z else \
        torch.empty((B, h, w, Z), dtype=torch.bfloat16, device=dev)
    z_nchw = torch.empty((B, Z, h, w), dtype=torch.float32, device=dev) if want_z_nchw else None
    kl = torch.empty((B,), dtype=torch.float32, device=dev)
    if eps is not None:
        eps = eps.contiguous().float()
        eps_out = None
    else:
        eps_out = torch.empty((B, Z, h, w), dtype=torch.float32, device=dev)
    z_lo = torch.zeros_like(z_bf16) if SPLIT_BF16[0] else None
    check(lib.tvae_reparam_fwd(moments.data_ptr(), _ptr(eps), seed, sample_offset, B, h * w, Z, z_bf16.data_ptr(),
                               z_pitch, _ptr(z_nchw), _ptr(eps_out), kl.data_ptr(), _ptr(z_lo), _stream()),
          "tvae_reparam_fwd")
    return _pair(z_bf16, z_lo), z_nchw, (eps if eps is not None else eps_out), kl


@_on_device
def reparam_bwd(moments, Z, dz1, eps1, dz2, eps2, kl_scale):
    B, h, w, _ = moments.shape
    dm = torch.empty((B, h, w, 2 * Z), dtype=torch.bfloat16, device=moments.device)
    check(lib.tvae_reparam_bwd(moments.data_ptr(), _ptr(dz1), _ptr(eps1), _ptr(dz2), _ptr(eps2), float(kl_scale), B,
                               h * w, Z, dm.data_ptr(), _stream()), "tvae_reparam_bwd")
    return dm


@_on_device
def nll_fwd(x_bf16, xhat, Cc, loss_type, logvar, batch, want_grad):
    """x_bf16 [N,H,W,xp] bf16; xhat [N,H,W,hp] fp32. Returns (sums fp64[3], dxhat bf16 or None); dxhat carries its
    column sums (the bias gradient of the conv that produced xhat) as `dxhat.tvae_colsum`."""
    x_bf16 = hi_of(x_bf16)
    P = x_bf16.numel() // x_bf16.shape[-1]          # pixels (the last dim may be a channel-slice view of pitched rows)
    dev = xhat.device
    sums = torch.empty((3,), dtype=torch.float64, device=dev)
    ws = _workspace(lib.tvae_nll_workspace_bytes(Cc), dev, "nll")
    dx = cs = None
    if want_grad:
        dx = torch.empty(x_bf16.shape[:-1] + (round_up(Cc, 8),), dtype=torch.bfloat16, device=dev)
        cs = torch.empty((Cc,), dtype=torch.float32, device=dev)
    _timed("nll_fwd", P * Cc * (2 + 4 + (2 if dx is not None else 0)),
           lambda: check(lib.tvae_nll_fwd(x_bf16.data_ptr(), pitch_of(x_bf16), xhat.data_ptr(), pitch_of(xhat), P, Cc,
                                          loss_type, _ptr(logvar), batch, _ptr(dx), dx.shape[-1] if dx is not None else 0,
                                          _ptr(cs), sums.data_ptr(), ws.data_ptr(), _stream()), "tvae_nll_fwd"))
    if dx is not None:
        dx.tvae_colsum = cs
        KERNEL_LAUNCHES[0] += 1
    return sums, dx


@_on_device
def recon_metrics(x_bf16, xhat, Cc):
    """x_bf16 [N,H,W,xp] bf16, xhat [N,H,W,hp] fp32 -> fp32 [N, 2] = per-sample (MAE, MSE)."""
    x_bf16 = hi_of(x_bf16)
    N, H, W = xhat.shape[0], xhat.shape[1], xhat.shape[2]
    out = torch.empty((N, 2), dtype=torch.float32, device=xhat.device)
    ws = _workspace(lib.tvae_recon_metrics_workspace_bytes(N), xhat.device, "recon_metrics")
    check(lib.tvae_recon_metrics(x_bf16.data_ptr(), pitch_of(x_bf16), xhat.data_ptr(), pitch_of(xhat), N, H * W, Cc,
                                 out.data_ptr(), ws.data_ptr(), _stream()), "tvae_recon_metrics")
    return out


@_on_device
def vae_loss_finalize(sums, kl, logvar, n_elem, kl_weight):
    """Returns fp32[5] = (loss, nll_loss, kl_loss, pixel_mse, dloss/dlogvar) on the device."""
    out = torch.empty((5,), dtype=torch.float32, device=kl.device)
    check(lib.tvae_vae_loss_finalize(sums.data_ptr(), kl.data_ptr(), kl.numel(), logvar.data_ptr(), float(n_elem),
                                     float(kl_weight), out.data_ptr(), _stream()), "tvae_vae_loss_finalize")
    return out


def _target_array(targets):
    arr = (C.c_void_p * len(targets))()
    for i, t in enumerate(targets):
        arr[i] = _ptr(t)
    return arr


@_on_device
def l2head_loss_fwd(pred, targets, B, h, w):
    sums = torch.empty((len(targets), 2), dtype=torch.float64, device=pred.device)
    check(lib.tvae_l2head_loss_fwd(pred.data_ptr(), pred.shape[-1], _target_array(targets), len(targets), B, h, w,
                                   sums.data_ptr(), _stream()), "tvae_l2head_loss_fwd")
    return sums


@_on_device
def l2head_loss_bwd(pred, targets, B, h, w, sums, weights, grad_scale, dp_pitch=8):
    dpred = torch.empty((B, h, w, dp_pitch), dtype=torch.bfloat16, device=pred.device)
    check(lib.tvae_l2head_loss_bwd(pred.data_ptr(), pred.shape[-1], _target_array(targets), len(targets), B, h, w,
                                   sums.data_ptr(), weights.data_ptr(), float(grad_scale), dpred.data_ptr(), dp_pitch,
                                   _stream()), "tvae_l2head_loss_bwd")
    return dpred


@_on_device
def l2head_finalize(sums, weights, vae_scal):
    out = torch.empty((1 + sums.shape[0],), dtype=torch.float32, device=sums.device)
    check(lib.tvae_l2head_finalize(sums.data_ptr(), weights.data_ptr(), sums.shape[0], vae_scal.data_ptr(),
                                   out.data_ptr(), _stream()), "tvae_l2head_finalize")
    return out


# ----------------------------------------------------------------------------------------------- optimiser
@_on_device
def sumsq(g, out):
    ws = _workspace(lib.tvae_sumsq_workspace_bytes(g.numel()), g.device, "sumsq")
    check(lib.tvae_sumsq(g.data_ptr(), g.numel(), out.data_ptr(), ws.data_ptr(), _stream()), "tvae_sumsq")


@_on_device
def adamw(p, g, m, v, *, lr, beta1, beta2, eps, weight_decay, step, sumsq_buf=None, max_norm=0.0, grad_scale=1.0):
    _timed("adamw", p.numel() * 28,
           lambda: check(lib.tvae_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1,
                                        beta2, eps, weight_decay, step, _ptr(sumsq_buf), max_norm, grad_scale,
                                        _stream()), "tvae_adamw"))


# ----------------------------------------------------------------------------------------------- data side
@_on_device
def gather_rows(src, idx, out):
    """out[j] = src[idx[j]] along dim 0 (contiguous rows, row bytes a multiple of 16); idx int64 on the device."""
    assert src.is_contiguous() and out.is_contiguous() and idx.dtype == torch.int64 and idx.is_cuda
    row_bytes = src[0].numel() * src.element_size() if src.dim() > 1 else src.element_size()
    if row_bytes % 16:
        raise _lib.TvaeError(f"gather_rows: rows of {row_bytes} bytes are not a multiple of 16")
    n = idx.numel()
    for i in range(0, n, 65535):
        k = min(65535, n - i)
        check(lib.tvae_gather_rows(src.data_ptr(), src.shape[0], row_bytes, idx.data_ptr() + 8 * i, k,
                                   out.data_ptr() + i * row_bytes, _stream()), "tvae_gather_rows")
    return out


@_on_device
def extract_tiles(rad, spec, T, mean=None, std=None, min_radiance=1.0, clip_min=-10.0, clip_max=10.0, want_f32=True,
                  want_bf16=False, out_bf16=None):
    """rad fp32 [M, NT, C] on the device; spec int32 [n, 4] = (row0, col0, flags, k) on the device. Returns
    (tiles fp32 [n, T, T, C] or None, tiles bf16 [n, T, T, pitch] or None) -- see tvae_extract_tiles."""
    require_cuda(rad, "radiance")
    assert rad.is_contiguous() and rad.dtype == torch.float32 and rad.dim() == 3
    assert spec.is_contiguous() and spec.dtype == torch.int32 and spec.dim() == 2 and spec.shape[1] == 4
    M, NT, Cc = rad.shape
    n = spec.shape[0]
    if mean is not None:
        mean = mean.to(rad.device, torch.float32).contiguous()
        std = std.to(rad.device, torch.float32).contiguous()
        if mean.numel() != Cc or std.numel() != Cc:
            raise _lib.TvaeError(f"mean/std spectra must have {Cc} channels")
    pitch = round_up(Cc, 8)
    of = torch.empty((n, T, T, Cc), dtype=torch.float32, device=rad.device) if want_f32 else None
    ob = out_bf16
    if ob is not None:          # e.g. a slice of DeviceTileCache.data: tiles land in the cache without a staging copy
        assert ob.is_contiguous() and ob.dtype == torch.bfloat16 and tuple(ob.shape[:3]) == (n, T, T)
        pitch = ob.shape[3]
    elif want_bf16:
        ob = torch.empty((n, T, T, pitch), dtype=torch.bfloat16, device=rad.device)
    check(lib.tvae_extract_tiles(rad.data_ptr(), M, NT, Cc, spec.data_ptr(), n, T, _ptr(mean), _ptr(std),
                                 float(min_radiance), float(clip_min), float(clip_max), _ptr(of), _ptr(ob), pitch,
                                 _stream()), "tvae_extract_tiles")
    return of, ob


@_on_device
def spectrum_stats_accum(rad, acc, min_radiance=1.0, take_log=True):
    """Adds the pixels of rad [..., C] (fp32, device) to the running fp64 sums acc [2, C]."""
    require_cuda(rad, "radiance")
    rad = rad.contiguous()
    Cc = rad.shape[-1]
    rows = rad.numel() // Cc
    assert acc.dtype == torch.float64 and tuple(acc.shape) == (2, Cc) and acc.is_contiguous()
    ws = _workspace(lib.tvae_spectrum_stats_workspace_bytes(rows, Cc), rad.device, "spectrum")
    check(lib.tvae_spectrum_stats_accum(rad.data_ptr(), rows, Cc, float(min_radiance), int(take_log), acc.data_ptr(),
                                        ws.data_ptr(), _stream()), "tvae_spectrum_stats_accum")
    return rows


@_on_device
def spectrum_stats_finalize(acc, total_rows):
    Cc = acc.shape[1]
    mean = torch.empty((Cc,), dtype=torch.float32, device=acc.device)
    std = torch.empty((Cc,), dtype=torch.float32, device=acc.device)
    check(lib.tvae_spectrum_stats_finalize(acc.data_ptr(), int(total_rows), Cc, mean.data_ptr(), std.data_ptr(),
                                           _stream()), "tvae_spectrum_stats_finalize")
    return mean, std


@_on_device
def batch_stats(x):
    """fp32[4] = (min, max, mean, unbiased std) of a device tensor: any contiguous fp32/bf16 tensor, or an NCHW-shaped
    channels-last view (the bf16 batches of the tile stores; pad lanes are skipped)."""
    require_cuda(x, "batch")
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    if x.is_contiguous():
        rows, Cc, pitch = 1, x.numel(), x.numel()
    elif x.dim() == 4 and x.stride(1) == 1 and x.stride(2) == x.shape[3] * x.stride(3) \
            and x.stride(0) == x.shape[2] * x.stride(2):
        rows, Cc, pitch = x.shape[0] * x.shape[2] * x.shape[3], x.shape[1], x.stride(3)
    else:
        x = x.contiguous()
        rows, Cc, pitch = 1, x.numel(), x.numel()
    out = torch.empty((4,), dtype=torch.float32, device=x.device)
    ws = _workspace(lib.tvae_batch_stats_workspace_bytes(), x.device, "batch_stats")
    check(lib.tvae_batch_stats(x.data_ptr(), int(x.dtype == torch.bfloat16), rows, Cc, pitch, out.data_ptr(),
                               ws.data_ptr(), _stream()), "tvae_batch_stats")
    return out


# ----------------------------------------------------------------------------------------------- probes
@_on_device
def rows_f32_to_bf16(x, out):
    """x fp32 [rows, C] (row pitch = x.stride(0)) -> out bf16 [rows, pitch] (pad lanes zeroed)."""
    require_cuda(x, "rows")
    assert x.dtype == torch.float32 and x.stride(1) == 1 and out.dtype == torch.bfloat16 and out.stride(1) == 1
    if x.shape[0] == 0:
        return out
    check(lib.tvae_nhwc_f32_to_nhwc_bf16(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], out.data_ptr(),
                                         out.stride(0), None, _stream()), "tvae_nhwc_f32_to_nhwc_bf16")
    return out


@_on_device
def act_dropout_fwd(x, Cc, act, p, seed, offset):
    """x fp32 [rows, pitch] -> bf16 [rows, round_up(C, 8)] = dropout_p(act(x)) (tvae_act_dropout_fwd)."""
    rows = x.shape[0]
    out = torch.empty((rows, round_up(Cc, 8)), dtype=torch.bfloat16, device=x.device)
    check(lib.tvae_act_dropout_fwd(x.data_ptr(), x.stride(0), rows, Cc, int(act), float(p), int(seed) & (2 ** 64 - 1),
                                   int(offset), out.data_ptr(), out.stride(0), _stream()), "tvae_act_dropout_fwd")
    return out


@_on_device
def act_dropout_bwd(x, da, Cc, act, p, seed, offset):
    rows = x.shape[0]
    dx = torch.empty((rows, round_up(Cc, 8)), dtype=torch.bfloat16, device=x.device)
    check(lib.tvae_act_dropout_bwd(x.data_ptr(), x.stride(0), da.data_ptr(), da.stride(0), rows, Cc, int(act), float(p),
                                   int(seed) & (2 ** 64 - 1), int(offset), dx.data_ptr(), dx.stride(0), _stream()),
          "tvae_act_dropout_bwd")
    return dx


@_on_device
def probe_mse(pred, y, n_valid, rows_padded, sums, dpred=None):
    """y: fp32 vector (any element stride: a column of the shuffled [X | y] matrix is read in place)."""
    assert pred.dtype == torch.float32 and y.dtype == torch.float32 and y.dim() == 1
    check(lib.tvae_probe_mse(pred.data_ptr(), pred.stride(0), y.data_ptr(), y.stride(0) if y.numel() > 1 else 1,
                             int(n_valid), int(rows_padded), sums.data_ptr(), _ptr(dpred),
                             dpred.stride(0) if dpred is not None else 0, _stream()), "tvae_probe_mse")
    return sums


# ----------------------------------------------------------------------------------------------- probe targets
@_on_device
def nan_moments(x, center=0.0):
    """fp64 [5] on the device: count, sum (x - c), sum (x - c)^2, min, max over the non-NaN elements (tvae_nan_moments)."""
    require_cuda(x, "component field")
    assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() > 0
    out = torch.empty(5, dtype=torch.float64, device=x.device)
    check(lib.tvae_nan_moments(x.data_ptr(), x.numel(), float(center), out.data_ptr(), _stream()), "tvae_nan_moments")
    return out


@_on_device
def select_hist(x, center, use_abs, prefix, prefix_mask, shift, hist=None):
    """One 8-bit pass of the exact radix select (tvae_select_hist): uint64-as-int64 [256] counts on the device."""
    require_cuda(x, "component field")
    assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() > 0
    if hist is None:
        hist = torch.empty(256, dtype=torch.int64, device=x.device)
    check(lib.tvae_select_hist(x.data_ptr(), x.numel(), float(center), int(bool(use_abs)), int(prefix), int(prefix_mask),
                               int(shift), hist.data_ptr(), _stream()), "tvae_select_hist")
    return hist


@_on_device
def component_pool(x, mode, a, b, pool=4, want_normalized=False):
    """x fp32 [H, W] (row pitch = x.stride(0)) -> (normalised [H, W] or None, pooled [H // pool, W // pool])
    (tvae_component_pool: normalisation and the nanmean pooling in one pass)."""
    require_cuda(x, "component field")
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    H, W = x.shape
    norm = torch.empty((H, W), dtype=torch.float32, device=x.device) if want_normalized else None
    pooled = torch.empty((H // pool, W // pool), dtype=torch.float32, device=x.device)
    check(lib.tvae_component_pool(x.data_ptr(), H, W, x.stride(0), int(pool), int(mode), float(a), float(b), _ptr(norm),
                                  pooled.data_ptr(), _stream()), "tvae_component_pool")
    if want_normalized and (H % pool or W % pool):
        KERNEL_LAUNCHES[0] += 1
    return norm, pooled
