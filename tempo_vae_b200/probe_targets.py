"""Probe TARGETS on the GPU (the data side of the probes): the arithmetic that
`process_file` of src/scripts/linear_probe_analysis.py applies to an L2 component field before a probe sees it --

  * `normalize_component(data, norm_type, stats=None)` (:60-110): "zscore", "minmax", "asinh" (scale = 1.4826 x the median
    absolute deviation) or "logit" (squeezed by eps = 0.01), statistics over the valid (non-NaN) pixels of the first
    file unless given; same name, arguments and `(normalized, stats)` return as the reference;
  * 4x4 nanmean pooling down to the latent grid (:180-190): `pool_component`;
  * both in ONE pass over the field: `component_targets` (what a caller that only needs the pooled target uses);
  * `sample_probe_pairs` (main(), :455-486): the valid pooled pixels of one file, sampled without replacement, paired
    with the latent vectors at the same positions.

NetCDF reading and file matching stay the reference's (out of scope); everything here takes the field as an array.
Execution: statistics by `tvae_nan_moments` (fp64 sums, min, max) and an exact radix select (`tvae_select_hist`, four
8-bit passes per order statistic; np.median's even-count average is formed from the two middle values), the
transform and the pooling by `tvae_component_pool`, the latent gather by `tvae_gather_rows`. Statistics are returned
as numpy float32 scalars (what the reference's float32 fields give), so `stats` dictionaries are interchangeable with
the reference's in both directions. There is no CPU fallback.
"""
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._lib import TvaeError

LOGIT_EPS = 0.01          # src/scripts/linear_probe_analysis.py:98
_F32 = np.float32


def _field(data, device=None) -> torch.Tensor:
    """[H, W] fp32 on the device (host arrays are copied over; NaN = invalid pixel)."""
    t = torch.as_tensor(np.asarray(data, dtype=np.float32)) if not torch.is_tensor(data) else data
    if t.dim() != 2:
        raise TvaeError(f"component field must be 2-D [H, W], got shape {tuple(t.shape)}")
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise TvaeError("probe targets are computed on CUDA only (there is no CPU fallback)")
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    t = t.to(torch.float32)
    return t if t.stride(1) == 1 else t.contiguous()


def _key_to_float(key: int) -> np.float32:
    bits = (key ^ 0x80000000) if key >> 31 else (~key & 0xFFFFFFFF)
    return np.array([bits], dtype=np.uint32).view(np.float32)[0]


def order_statistics(x: torch.Tensor, ranks_of_n, center: float = 0.0, use_abs: bool = False):
    """Exact order statistics of the valid (non-NaN) values of v = x (or |x - center|): `ranks_of_n(n)` maps the valid
    count to the 0-based ranks wanted; returns (n, {rank: np.float32 value}). Four 8-bit radix passes per distinct
    path through the key space (two ranks that share a bucket share the pass)."""
    flat = x.reshape(-1) if x.is_contiguous() else x.contiguous().reshape(-1)
    hist0 = ops.select_hist(flat, center, use_abs, 0, 0, 24).cpu().numpy()
    n = int(hist0.sum())
    if n == 0:
        return 0, {}
    out: Dict[int, np.float32] = {}

    def descend(hist, prefix, mask, shift, wanted):          # wanted: [(rank, rank inside this prefix)]
        cum = np.cumsum(hist)
        groups: Dict[int, list] = {}
        for rank, inside in wanted:
            b = int(np.searchsorted(cum, inside, side="right"))
            groups.setdefault(b, []).append((rank, inside - (int(cum[b - 1]) if b else 0)))
        for b, grp in groups.items():
            p = prefix | (b << shift)
            if shift == 0:
                for rank, _ in grp:
                    out[rank] = _key_to_float(p)
            else:
                m = mask | (0xFF << shift)
                h = ops.select_hist(flat, center, use_abs, p, m, shift - 8).cpu().numpy()
                descend(h, p, m, shift - 8, grp)

    ranks = sorted(set(int(r) for r in ranks_of_n(n)))
    descend(hist0, 0, 0, 24, [(r, r) for r in ranks])
    return n, out


def nan_median(x: torch.Tensor, center: float = 0.0, use_abs: bool = False) -> np.float32:
    """np.median over the valid values of x (or of |x - center|), as float32: the middle value, or the float32 mean of
    the two middle values for an even count; NaN when nothing is valid."""
    n, vals = order_statistics(x, lambda n: ((n - 1) // 2, n // 2), center, use_abs)
    if n == 0:
        return _F32(np.nan)
    lo, hi = vals[(n - 1) // 2], vals[n // 2]
    return lo if n % 2 else _F32((lo + hi) / _F32(2.0))


def component_stats(x: torch.Tensor, norm_type: str) -> Dict[str, np.float32]:
    """The statistics `normalize_component` computes when none are given (src/scripts/linear_probe_analysis.py:62-100)."""
    if norm_type == "logit":
        return {"eps": LOGIT_EPS}
    if norm_type == "asinh":
        median = nan_median(x)
        mad = nan_median(x, center=float(median), use_abs=True)
        return {"scale": _F32(_F32(1.4826) * mad), "median": median}
    if norm_type not in ("zscore", "minmax"):
        raise ValueError(f"Unknown normalization type: {norm_type}")
    flat = x.reshape(-1) if x.is_contiguous() else x.contiguous().reshape(-1)
    m = ops.nan_moments(flat).cpu().numpy()
    if m[0] == 0:
        nan = _F32(np.nan)
        return {"mean": nan, "std": nan} if norm_type == "zscore" else {"min": nan, "max": nan}
    if norm_type == "minmax":
        return {"min": _F32(m[3]), "max": _F32(m[4])}
    mean = _F32(m[1] / m[0])
    c = ops.nan_moments(flat, center=float(mean)).cpu().numpy()          # centred: no cancellation in the variance
    var = c[2] / c[0] - (c[1] / c[0]) ** 2
    return {"mean": mean, "std": _F32(np.sqrt(max(var, 0.0)))}


def _transform(norm_type: str, stats) -> Tuple[int, float, float]:
    """(mode, a, b) of tvae_component_pool; the float32 arithmetic of the reference's expressions."""
    if norm_type == "zscore":
        return 0, float(_F32(stats["mean"])), float(_F32(stats["std"]) + _F32(1e-8))
    if norm_type == "minmax":
        return 0, float(_F32(stats["min"])), float(_F32(stats["max"]) - _F32(stats["min"]) + _F32(1e-8))
    if norm_type == "asinh":
        return 1, 0.0, float(_F32(stats["scale"]) + _F32(1e-8))
    if norm_type == "logit":
        return 2, float(stats["eps"]), 1.0
    raise ValueError(f"Unknown normalization type: {norm_type}")


def normalize_component(data, norm_type: str, stats: Optional[dict] = None, device=None):
    """Drop-in for src/scripts/linear_probe_analysis.py:60-110: returns (normalized [H, W] fp32 on the device, stats)."""
    x = _field(data, device)
    if stats is None:
        stats = component_stats(x, norm_type)
    mode, a, b = _transform(norm_type, stats)
    norm, _ = ops.component_pool(x, mode, a, b, pool=1, want_normalized=True)
    return norm, stats


def pool_component(normalized, pool: int = 4, device=None) -> torch.Tensor:
    """[H, W] -> [H // pool, W // pool]: mean of the valid values of every pool x pool block, NaN where a block has none
    (`reshape(h//4, 4, w//4, 4)` + `np.nanmean(axis=(1, 3))`, src/scripts/linear_probe_analysis.py:183-190)."""
    x = _field(normalized, device)
    return ops.component_pool(x, 0, 0.0, 1.0, pool=pool)[1]


def component_targets(data, norm_type: str, stats: Optional[dict] = None, pool: int = 4,
                      crop: Optional[Sequence[int]] = None, device=None):
    """Field -> (pooled target on the latent grid, stats) in one pass over the field (normalisation fused into the
    pooling). `crop` = (rows, cols) restricts the field to the L1 crop first, as process_file does (:176), BEFORE the
    statistics are taken."""
    x = _field(data, device)
    if crop is not None:
        x = x[:crop[0], :crop[1]]
    if stats is None:
        stats = component_stats(x, norm_type)
    mode, a, b = _transform(norm_type, stats)
    return ops.component_pool(x, mode, a, b, pool=pool)[1], stats


def sample_probe_pairs(latent: torch.Tensor, pooled: torch.Tensor, n_pixels: int, rng=np.random):
    """One file's (X [n, Z] fp32, y [n] fp32) on the device: `n = min(n_pixels, #valid)` valid pooled pixels drawn
    without replacement by `rng.choice` exactly as the reference does (src/scripts/linear_probe_analysis.py:455-486; pass
    `np.random` after `np.random.seed(seed)` for the reference's draw), the latent rows gathered on the device.
    latent: [1, Z, h, w] or [Z, h, w]; pooled: [h, w]. Returns (None, None) when no pixel is valid."""
    lat = latent[0] if latent.dim() == 4 else latent
    Z, h, w = lat.shape
    if tuple(pooled.shape) != (h, w):
        raise TvaeError(f"pooled component {tuple(pooled.shape)} does not match the latent grid {(h, w)}")
    comp = pooled.reshape(-1).cpu().numpy()
    valid = np.where(~np.isnan(comp))[0]
    if len(valid) == 0:
        return None, None
    idx = rng.choice(valid, min(int(n_pixels), len(valid)), replace=False)
    rows = lat.reshape(Z, h * w).t().contiguous().to(torch.float32)      # [h w, Z], Z * 4 bytes per row
    X = torch.empty((len(idx), Z), dtype=torch.float32, device=lat.device)
    ops.gather_rows(rows, torch.as_tensor(idx, dtype=torch.int64, device=lat.device), X)
    return X, torch.as_tensor(comp[idx], device=lat.device)
