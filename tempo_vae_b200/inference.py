"""Inference-only encode sweep: radiance granule -> normalised patches -> posterior means (latents).

Callers this serves in the reference (they stay the reference's scripts; these are the hot-path calls they make):
  * patch-batched evaluation ........ src/scripts/evaluate_reconstruction.py:70-90  (`model(batch)` on [B,1028,64,64])
  * whole-granule latent extraction .. src/scripts/linear_probe_analysis.py:113-140  (`model.get_latent(x).mean` on
                                       [1,1028,128,2048]; fully convolutional, attention over 16,384 tokens)
  * whole-granule reconstruction ..... src/scripts/analyze_reconstruction.py:111-127  (`model(x)` on the same crop)
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
    from .model import ENGINE
    mine = patches[rank::world]
    outs = []
    with ENGINE.frozen_weights():          # weights do not change inside the sweep: pack them once, not per batch
        for i in range(0, mine.shape[0], batch_size):
            chunk = mine[i:i + batch_size]
            if chunk.dtype != torch.bfloat16:          # bf16 channels-last views (tile stores) are consumed in place
                chunk = chunk.to(dev, dtype=torch.float32, non_blocking=True)
            elif chunk.device != dev:
                chunk = chunk.to(dev, non_blocking=True)
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


@torch.no_grad()
def reconstruct_granule_whole(model, z_rad: torch.Tensor, tile: int = 64, sample_posterior: bool = True,
                              eps: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The reference's whole-granule reconstruction (src/scripts/analyze_reconstruction.py:111-127: `recon = model(x)` on
    the [1, C, H, W] crop, converted back to [H, W, C]): one fully convolutional encode -> sample -> decode pass; the
    result stays on the device. `sample_posterior=False` decodes the posterior mode, `eps` ([1, Z, H/4, W/4]) injects the
    noise (by default it is drawn on the device, as `model(x)` draws it)."""
    vae = model.vae if hasattr(model, "vae") else model
    dev = next(vae.parameters()).device
    M, T, C = z_rad.shape
    x = z_rad[:(M // tile) * tile, :(T // tile) * tile, :].permute(2, 0, 1).unsqueeze(0)
    rec, _ = vae(x.to(dev, dtype=torch.float32), sample_posterior=sample_posterior, eps=eps)
    return rec[0].permute(1, 2, 0)


class GranuleGraph:
    """CUDA-graph replay of a whole-granule pass for ONE granule shape (opt-in).

    A [1, C, 128, 2048] pass is a chain of 50-100 short launches on an otherwise idle GPU, so the host's launch cost shows
    up as gaps between kernels (1.3-1.5 ms of a 6.7 / 12.5 ms pass, `tools/granule_timeline.py`). Here the chain is
    captured once (`torch.cuda.graph`: the C ABI launches on torch's current stream, which is the capturing stream;
    every buffer comes from torch's allocator, whose graph pool keeps the addresses -- and with them the TMA descriptors
    baked into the launches -- valid) and replayed per granule: `g = GranuleGraph(model, z.shape); lat = g(z)`.

      * `reconstruct=False`: `encode_granule_whole` -> latent means [1, Z, H/4, W/4];
      * `reconstruct=True`: `reconstruct_granule_whole` -> [H, W, C]; the noise is injected through a static buffer that
        is refilled per call (`eps=` or, by default, `torch.randn` on the device; the eager path draws Philox noise).

    The weight packs are rebuilt inside the graph from the LIVE parameters (one launch), so optimiser steps or
    `load_state_dict` between calls are picked up as long as the parameters stay where they are (capture after the
    optimiser has been built: FusedAdamW moves them into its flat buffer). The returned tensor is the graph's static
    output: it is overwritten by the next call (`.clone()` it to keep it)."""

    def __init__(self, model, shape, tile: int = 64, reconstruct: bool = False, warmup: int = 3):
        vae = model.vae if hasattr(model, "vae") else model
        dev = next(vae.parameters()).device
        M, T, C = (int(v) for v in shape)
        if M < tile or T < tile:
            raise TvaeError(f"granule {tuple(shape)} is smaller than one {tile}x{tile} patch")
        self.model, self.reconstruct, self.shape = model, reconstruct, (M, T, C)
        self.z = torch.zeros((M, T, C), dtype=torch.float32, device=dev)
        self.eps = torch.zeros((1, vae.embed_dim, (M // tile) * tile // 4, (T // tile) * tile // 4), device=dev) \
            if reconstruct else None
        run = (lambda: reconstruct_granule_whole(model, self.z, tile, eps=self.eps)) if reconstruct else \
            (lambda: encode_granule_whole(model, self.z, tile))
        with torch.cuda.device(dev):
            # warm-up AND capture on one stream of our own: the engine's scratch buffers are keyed by stream, so the ones
            # the captured launches use are the ones allocated (outside the capture) during the warm-up
            self._stream = torch.cuda.Stream()
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                for _ in range(max(1, warmup)):
                    run()
            torch.cuda.current_stream().wait_stream(self._stream)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self._stream):
                self.out = run()

    @torch.no_grad()
    def __call__(self, z_rad: torch.Tensor, eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        if tuple(z_rad.shape) != self.shape:
            raise TvaeError(f"this graph was captured for granules of shape {self.shape}, got {tuple(z_rad.shape)}")
        self.z.copy_(z_rad, non_blocking=True)
        if self.reconstruct:
            if eps is None:
                self.eps.normal_()
            else:
                self.eps.copy_(eps, non_blocking=True)
        self.graph.replay()
        return self.out


@torch.no_grad()
def evaluate_reconstruction(model, x: torch.Tensor, eps: Optional[torch.Tensor] = None,
                            sample_posterior: bool = True, max_val: float = 20.0) -> dict:
    """Per-sample reconstruction metrics of src/scripts/evaluate_reconstruction.py:23-42 for a batch x [B, C, H, W]:
    {"mse": [B], "mae": [B], "psnr": [B]} on the device (PSNR = 10 log10(max_val^2 / (mse + 1e-10)), max_val = 20 for
    data clipped to [-10, 10]). One stochastic forward like the reference's `model(x)` (or the posterior mode with
    sample_posterior=False; `eps` injects the noise); the errors are reduced by `tvae_recon_metrics` straight from the
    channels-last reconstruction, which is never converted back to NCHW."""
    from . import ops
    from .model import ENGINE, _DecodeProgram, _EncodeProgram, _check_input
    vae = model.vae if hasattr(model, "vae") else model
    C, Z = vae.encoder.in_channels, vae.embed_dim
    _check_input(x, C)
    B = x.shape[0]
    ENGINE.begin_forward()
    xb = ops.input_nhwc_bf16(x)
    mom, _ = _EncodeProgram(vae).program_fwd(xb, False)
    if not sample_posterior:
        eps = torch.zeros((B, Z, mom.f32.shape[1], mom.f32.shape[2]), device=x.device)
    if eps is None:
        off = ENGINE.rng_offset
        ENGINE.rng_offset += B
        z = ops.reparam_fwd(mom.f32, Z, seed=ENGINE.rng_seed, sample_offset=off)[0]
    else:
        z = ops.reparam_fwd(mom.f32, Z, eps=eps)[0]
    xhat, _ = _DecodeProgram(vae).program_fwd(z, False)
    m = ops.recon_metrics(xb, xhat.f32, C)
    mse = m[:, 1]
    return {"mse": mse, "mae": m[:, 0], "psnr": 10.0 * torch.log10(max_val ** 2 / (mse + 1e-10))}
