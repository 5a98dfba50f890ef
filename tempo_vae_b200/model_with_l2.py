"""VAE with L2-product supervision heads — same classes, signatures and state_dict keys as the reference's
src/model_with_l2.py, executed by the B200 engine.

Reference interface mirrored here (file:line in /root/reference):
  L2PredictionHead ..................... src/model_with_l2.py:11-42   (1x1-conv MLP 32 -> 512 -> 512 -> 4, GN(8) + GELU)
  VAEWithL2Supervision.forward ......... src/model_with_l2.py:61-93
  VAEWithL2Supervision.compute_loss .... src/model_with_l2.py:95-182

Reference behaviours kept on purpose: the head is fed a SECOND posterior sample (not the one
that was decoded); compute_loss's `kl_weight` argument is ignored in favour of `self.vae.kl_weight`; a product
with no valid (non-NaN after 4x4 average pooling) pixel contributes nothing and has no `<product>_loss` metric.
"""
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .model import (A, ENGINE, Conv2d, GroupNorm, TvaeError, _Act, _ModuleFn, _VAELossFn, conv_bwd, conv_fwd,
                    norm_act_bwd, norm_act_fwd)


class L2PredictionHead(nn.Module):
    """Conv1D MLP for predicting ALL L2 products from VAE latents."""

    def __init__(self, latent_channels: int = 32, hidden_dims: list = [512, 512], n_outputs: int = 4):
        super().__init__()
        layers = []
        in_channels = latent_channels
        for hidden_dim in hidden_dims:
            layers.extend([
                Conv2d(in_channels, hidden_dim, kernel_size=1, bias=False),
                GroupNorm(8, hidden_dim),
                _Act("gelu"),
            ])
            in_channels = hidden_dim
        layers.append(Conv2d(in_channels, n_outputs, kernel_size=1))
        self.mlp = nn.Sequential(*layers)
        self.latent_channels = latent_channels
        self.n_outputs = n_outputs

    # ---- engine program: z_bf16 NHWC [B,h,w,Z] -> fp32 NHWC [B,h,w,n_outputs]
    def fwd(self, z_bf16, save):
        mods = list(self.mlp)
        h = z_bf16
        cin = self.latent_channels
        saved = []
        for i in range(0, len(mods) - 1, 3):
            conv, norm, act = mods[i], mods[i + 1], mods[i + 2]
            r = conv_fwd(conv, h, cin, stats_for=norm)
            f32 = r[0]
            a, st = norm_act_fwd(norm, f32, act.code, r.stats, save)
            saved.append((h, f32, st))
            h, cin = a, conv.out_channels
        pred, _ = conv_fwd(mods[-1], h, cin)
        return pred, ((saved, h) if save else None)

    def bwd(self, dpred_bf16, saved_all):
        saved, last_in = saved_all
        mods = list(self.mlp)
        nhid = len(saved)
        d = conv_bwd(mods[-1], dpred_bf16, last_in, mods[-1].in_channels, dgrad="bf16")
        for j in reversed(range(nhid)):
            conv, norm, act = mods[3 * j], mods[3 * j + 1], mods[3 * j + 2]
            h_in, f32, st = saved[j]
            df = norm_act_bwd(norm, f32, st, d, None, act.code)
            d = conv_bwd(conv, df, h_in, conv.in_channels, dgrad=("f32" if j == 0 else "bf16"))
        return d

    # program interface for _ModuleFn (NCHW in/out)
    def in_channels_api(self):
        return self.latent_channels

    def program_fwd(self, zb, save):
        pred, s = self.fwd(zb, save)
        return A(f32=pred, C=self.n_outputs), s

    def program_bwd(self, gb, saved, need):
        d = self.bwd(gb, saved)
        return d if need else None

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        """z: [B, latent_channels, H/4, W/4] -> L2 predictions [B, 4, H/4, W/4]"""
        return _ModuleFn.apply(z, self, *[p for p in self.parameters()])


class VAEWithL2Supervision(nn.Module):
    """VAE with multi-task L2 product supervision."""

    def __init__(self, base_vae, latent_channels: int = 32, mlp_hidden: list = [512, 512]):
        super().__init__()
        self.vae = base_vae
        self.l2_head = L2PredictionHead(latent_channels, mlp_hidden, n_outputs=4)
        dev = next(base_vae.parameters()).device
        if dev.type == "cuda":
            self.l2_head.to(dev)
        self.l2_products = ['NO2', 'O3TOT', 'HCHO', 'CLDO4']
        self.downsample = nn.AvgPool2d(kernel_size=4, stride=4)  # kept for API parity; pooling is fused in the loss kernel
        self._last = {}

    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        posterior = self.vae.encode(x)
        z = posterior.sample()
        reconstruction = self.vae.decode(z)
        l2_all = self.l2_head(z)
        l2_predictions = {}
        for i, product in enumerate(self.l2_products):
            l2_predictions[product] = l2_all[:, i:i + 1, :, :]
        return {
            'reconstruction': reconstruction,
            'posterior': posterior,
            'z': z,
            'l2_predictions': l2_predictions
        }

    def compute_loss(
        self,
        batch: Dict[str, torch.Tensor],
        kl_weight: float = 1e-6,
        l2_weights: Optional[Dict[str, float]] = None,
        eps: Optional[torch.Tensor] = None,
        eps2: Optional[torch.Tensor] = None,
        return_device_metrics: bool = False,
    ) -> Tuple[torch.Tensor, Dict[str, float]]:
        """Combined VAE + L2 supervision loss (src/model_with_l2.py:95-182) as one fused autograd node.

        Returns (total_loss, metrics) with metrics as python floats like the reference (one host sync for all of
        them instead of the reference's 4+3 `.item()` calls); `return_device_metrics=True` returns device scalars
        and does not synchronise. `eps` / `eps2` optionally inject the two noise draws."""
        if l2_weights is None:
            l2_weights = {'NO2': 0.1, 'O3TOT': 0.1, 'HCHO': 0.1, 'CLDO4': 0.1}
        x = batch['spectral']
        present = [p for p in self.l2_products if p in batch]
        targets = []
        for p in self.l2_products:
            if p in batch:
                t = batch[p]
                ops.require_cuda(t, p)
                if t.dim() != 3 or t.shape[0] != x.shape[0] or t.shape[1] != x.shape[2] or t.shape[2] != x.shape[3]:
                    raise TvaeError(f"{p}: expected a [B, H, W] target matching the spectral tile, got {tuple(t.shape)}")
                targets.append(t.contiguous().float())
            else:
                targets.append(None)
        w = torch.tensor([float(l2_weights[p]) if p in batch else 0.0 for p in self.l2_products], dtype=torch.float32)
        w = w.to(x.device, non_blocking=True)
        extra = {"l2": {"head": self.l2_head, "targets": targets, "weights": w, "eps2": eps2}}
        params = [p for p in self.parameters()]
        total_loss = _VAELossFn.apply(x, eps, self.vae, extra, *params)
        self._last = extra
        self.vae._last = extra
        scal, l2_out = extra["scalars"], extra["l2_out"]
        if return_device_metrics:
            metrics = {'loss': total_loss.detach(), 'nll_loss': scal[1], 'kl_loss': scal[2]}
            for i, p in enumerate(self.l2_products):
                if p in present:
                    metrics[f'{p}_loss'] = l2_out[1 + i]
            return total_loss, metrics
        host = torch.cat([l2_out, scal]).tolist()        # ONE device->host sync
        n = len(self.l2_products)
        metrics = {'loss': host[0], 'nll_loss': host[1 + n + 1], 'kl_loss': host[1 + n + 2]}
        for i, p in enumerate(self.l2_products):
            v = host[1 + i]
            if p in present and v == v:                   # NaN marks "no valid pixel": reference emits no metric
                metrics[f'{p}_loss'] = v
        return total_loss, metrics
