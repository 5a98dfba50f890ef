"""Inference-only encode sweep: radiance granule -> normalised patches -> posterior means (latents).

Callers this serves in the reference (they stay the reference's scripts; these are the hot-path calls they make):
  * patch-batched evaluation ........ src/scripts/evaluate_reconstruction.py:70-90  (`model(batch)` on [B,1028,64,64])
  * whole-granule latent extraction .. src/scripts/linear_probe_analysis.py:113-140  (`model.get_latent(x).mean` on
                                       [1,1028,128,2048]; fully convolutional, attention over 16,384 tokens)
  * normalisation .................... src/scripts/prepare_tempo_tiles.py:69-83 / linear_probe_analysis.py:121-124:
                                       log(clamp(rad, min_radiance)) -> z-score with the per-channel spectrum
                                       statistics -> clip to [-10, 10]

BASELINE.json config 5 is the patch-batched semantics: a [>=128, 2048] granule crop is 2 x 32 = 64 non-overlapping
64x64 patches; latents are `AutoencoderKL.encode(patches).mean` per patch (encoder only: no sampling, no decoder),
sharded by patch across ranks with no collective except the optional final gather.
"""
from typing import Optional

import torch

from .model import TvaeError


def normalize_radiance(rad: torch.Tensor, mean_spectrum: torch.Tensor, std_spectrum: torch.Tensor,
                       min_radiance: float = 1.0, clip: float = 10.0, device=None) -> torch.Tensor:
    """[mirror, track, C] raw radiance -> z-scored log-radiance clipped to [-clip, clip] (the formula of the
    reference's data preparation and analysis scripts, src/scripts/prepare_tempo_tiles.py:67-79), computed in one
    fused pass on the GPU (`tvae_normalize_radiance`). A host tensor is copied to `device` (default: the current
    CUDA device) first; the result lives on the device."""
    from . import ops
    if not rad.is_cuda:
        if not torch.cuda.is_available():
            raise TvaeError("normalize_radiance runs on CUDA only (there is no CPU fallback)")
        rad = rad.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()),
                     non_blocking=True)
    return ops.normalize_radiance(rad, mean_spectrum, std_spectrum, min_radiance, -clip, clip)[0]


def granule_to_patches(z_rad: torch.Tensor, tile: int = 64) -> torch.Tensor:
    """[mirror, track, C] -> [n_patches, C, tile, tile]: largest multiple-of-tile crop, non-overlapping, row-major
    (mirror block, then cross-track block)."""
    M, T, C = z_rad.shape
    mh, tw = M // tile, T // tile
    if mh == 0 or tw == 0:
        raise TvaeError(f"granule {tuple(z_rad.shape)} is smaller than one {tile}x{tile} patch")
    x = z_rad[:mh * tile, :tw * tile, :].reshape(mh, tile, tw, tile, C)
    return x.permute(0, 2, 4, 1, 3).reshape(mh * tw, C, tile, tile)


@torch.no_grad()
def encode_patches(model, patches: torch.Tensor, batch_size: int = 256, rank: int = 0, world: int = 1,
                   device: Optional[torch.device] = None) -> torch.Tensor:
    """Posterior means [n_local, Z, h, w] for patches[rank::world] (round-robin shard). `patches` may live on the
    host (each chunk is copied over) or on the device."""
    vae = model.vae if hasattr(model, "vae") else model
    dev = device or next(vae.parameters()).device
    mine = patches[rank::world]
    outs = []
    for i in range(0, mine.shape[0], batch_size):
        chunk = mine[i:i + batch_size].to(dev, dtype=torch.float32, non_blocking=True)
        outs.append(vae.encode(chunk).mean)
    if not outs:
        Z = vae.embed_dim
        return torch.empty((0, Z, 0, 0), device=dev)
    return torch.cat(outs, 0)


@torch.no_grad()
def encode_granule_whole(model, z_rad: torch.Tensor, tile: int = 64) -> torch.Tensor:
    """The reference's whole-granule call: one fully convolutional pass over the [1, C, H, W] crop (GroupNorm
    statistics and mid-block attention span the whole granule, so this differs from patch tiling by design)."""
    vae = model.vae if hasattr(model, "vae") else model
    dev = next(vae.parameters()).device
    M, T, C = z_rad.shape
    x = z_rad[:(M // tile) * tile, :(T // tile) * tile, :].permute(2, 0, 1).unsqueeze(0)
    return vae.encode(x.to(dev, dtype=torch.float32)).mean
