"""Tile preparation on the GPU: spectrum statistics, radiance normalisation and random
augmented tile extraction straight from raw granules — the arithmetic of the reference's data-preparation scripts,
fused so that neither the normalised granule nor the 51 GB intermediate fp32 tile set has to exist.

Reference call sites mirrored (file:line in /root/reference):
  extract_tiles(z_rad, tile_size, n_tiles, seed) ... src/scripts/prepare_tempo_tiles.py:21-58
      random crop (overlap allowed) -> torch.flip(dims=[0]) w.p. 1/2 -> torch.flip(dims=[1]) w.p. 1/2 ->
      torch.rot90(k in 0..3, dims=[0, 1]); the np.random draws are made here in the SAME order, so a seed picks the
      same tiles as the reference
  process_file: log -> z-score -> clip -> tiles .... src/scripts/prepare_tempo_tiles.py:61-93 (global statistics, or
      the per-file fallback mean / unbiased std over the granule's own pixels)
  spectrum statistics ............................ src/scripts/compute_tempo_stats.py:58-86 (np.log(np.clip) ->
      per-channel mean and population std over the stacked pixels of all files)

The NetCDF reading, directory walking and .pt writing of those scripts stay host-side callers (out of scope, SURVEY.md
section 2); these functions take the radiance array the scripts read (`[mirror, track, C]`, host or device).
"""
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._lib import TvaeError


def _to_device(rad: torch.Tensor, device=None) -> torch.Tensor:
    if rad.is_cuda:
        return rad.float().contiguous()
    if not torch.cuda.is_available():
        raise TvaeError("tile preparation runs on CUDA only (there is no CPU fallback)")
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    return rad.to(dev, dtype=torch.float32, non_blocking=True).contiguous()


def draw_tile_specs(n_mirror: int, n_track: int, tile_size: Sequence[int], n_tiles: int,
                    seed: Optional[int] = None) -> np.ndarray:
    """int32 [n_tiles, 4] = (row0, col0, flags, k): the random choices of the reference's extract_tiles, drawn from
    np.random in its order (randint row, randint col, rand flip0, rand flip1, randint k)."""
    tm, tt = int(tile_size[0]), int(tile_size[1])
    if seed is not None:
        np.random.seed(seed)
    spec = np.zeros((n_tiles, 4), dtype=np.int32)
    for t in range(n_tiles):
        i = np.random.randint(0, n_mirror - tm + 1)
        j = np.random.randint(0, n_track - tt + 1)
        f0 = np.random.rand() > 0.5
        f1 = np.random.rand() > 0.5
        k = np.random.randint(0, 4)
        spec[t] = (i, j, int(f0) | (int(f1) << 1), k)
    return spec


def _check_tile(shape, tile_size):
    n_mirror, n_track = shape[0], shape[1]
    tm, tt = int(tile_size[0]), int(tile_size[1])
    if n_mirror < tm or n_track < tt:
        return None
    if tm != tt:
        raise TvaeError(f"tile_size {tuple(tile_size)}: rot90 augmentation needs square tiles (the reference's "
                        "torch.stack fails on mixed shapes too)")
    return tm


def extract_tiles(z_rad: torch.Tensor, tile_size: Sequence[int], n_tiles: int, seed: Optional[int] = None):
    """Same contract as src/scripts/prepare_tempo_tiles.py:21-58: `[n_tiles, T, T, C]` fp32 augmented tiles of an
    already normalised granule (None when the granule is smaller than a tile). One kernel for all tiles."""
    T = _check_tile(z_rad.shape, tile_size)
    if T is None:
        return None
    if n_tiles <= 0:
        return None
    z = _to_device(z_rad)
    spec = torch.from_numpy(draw_tile_specs(z.shape[0], z.shape[1], tile_size, n_tiles, seed)).to(z.device)
    return ops.extract_tiles(z, spec, T, want_f32=True)[0]


def granule_statistics(rad: torch.Tensor, min_radiance: float = 1.0, unbiased: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-channel (mean, std) of log(clamp(rad, min_radiance)) over the granule's own pixels: the per-file fallback of
    process_file (unbiased=True: torch's .std) or one file's worth of compute_tempo_stats (unbiased=False: np.std)."""
    st = SpectrumStats(rad.shape[-1], device=rad.device if rad.is_cuda else None, min_radiance=min_radiance)
    st.update(rad)
    return st.finalize(unbiased=unbiased)


def process_granule(rad: torch.Tensor, params: dict, mean_spectrum: Optional[torch.Tensor] = None,
                    std_spectrum: Optional[torch.Tensor] = None, seed: Optional[int] = None, cache=None,
                    want_f32: bool = True):
    """process_file of the reference minus the NetCDF read (src/scripts/prepare_tempo_tiles.py:61-93): raw radiance
    `[mirror, track, C]` -> `[tiles_per_file, T, T, C]` fp32 tiles (the on-disk format), normalised with the global
    spectra or, without them, with the granule's own mean / unbiased std. `params` holds the script's `processing`
    keys: min_radiance, clip_min, clip_max, tile_size, tiles_per_file.

    cache (optional DeviceTileCache): the tiles are ALSO written, as the bf16 channels-last rows the conv kernels read,
    directly into the cache's storage by the same kernel (no staging tensor, no second pass); with want_f32=False that is
    the only output and the function returns the number of tiles added."""
    T = _check_tile(rad.shape, params["tile_size"])
    if T is None:
        return None
    dev = cache.device if cache is not None else None
    r = _to_device(rad, dev)
    if mean_spectrum is None or std_spectrum is None:
        mean_spectrum, std_spectrum = granule_statistics(r, params["min_radiance"], unbiased=True)
    n = int(params["tiles_per_file"])
    spec = torch.from_numpy(draw_tile_specs(r.shape[0], r.shape[1], params["tile_size"], n, seed)).to(r.device)
    out_bf16 = None
    if cache is not None:
        if (cache.H, cache.W, cache.C) != (T, T, r.shape[2]):
            raise TvaeError("cache geometry does not match the tiles")
        if cache.n + n > cache.data.shape[0]:
            raise ValueError("DeviceTileCache capacity exceeded")
        out_bf16 = cache.data[cache.n:cache.n + n]
    tiles, _ = ops.extract_tiles(r, spec, T, mean_spectrum, std_spectrum, params["min_radiance"], params["clip_min"],
                                 params["clip_max"], want_f32=want_f32, out_bf16=out_bf16)
    if cache is not None:
        cache.n += n
    return tiles if want_f32 else n


class SpectrumStats:
    """Running per-channel statistics of log-radiance over any number of granules (compute_tempo_stats.py:58-86):

        st = SpectrumStats(1028)
        for rad in granules: st.update(rad)            # raw radiance [mirror, track, 1028], host or device
        mean_spectrum, std_spectrum = st.finalize()    # fp32 [1028] each, population std like np.std

    Sums are kept in fp64 on the device; the reduction order is fixed (bit-reproducible)."""

    def __init__(self, n_channels: int, device=None, min_radiance: float = 1.0):
        if not torch.cuda.is_available():
            raise TvaeError("SpectrumStats runs on CUDA only (there is no CPU fallback)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.C = n_channels
        self.min_radiance = float(min_radiance)
        self.acc = torch.zeros((2, n_channels), dtype=torch.float64, device=self.device)
        self.rows = 0

    def update(self, rad: torch.Tensor, take_log: bool = True):
        if rad.shape[-1] != self.C:
            raise TvaeError(f"expected {self.C} spectral channels, got {rad.shape[-1]}")
        r = _to_device(rad, self.device)
        self.rows += ops.spectrum_stats_accum(r, self.acc, self.min_radiance, take_log)
        return self

    def finalize(self, unbiased: bool = False):
        if self.rows == 0:
            raise ValueError("FATAL: No files could be loaded")
        mean, std = ops.spectrum_stats_finalize(self.acc, self.rows)
        if unbiased and self.rows > 1:
            std = std * float(np.sqrt(self.rows / (self.rows - 1.0)))
        return mean, std
