"""Training runtime for the B200 engine — the reference's src/train_utils.py surface (seed_all, get_device,
get_sqrt_schedule, Trainer with train_step / validate / train / save_checkpoint / load_checkpoint / save_metrics)
over the fused get_loss node and the fused clip + AdamW step.

Reference interface mirrored here (file:line in /root/reference):
  seed_all / get_device / get_sqrt_schedule ... src/train_utils.py:17-63
  Trainer.__init__ ............................ src/train_utils.py:66-120
  save_checkpoint / load_checkpoint ........... src/train_utils.py:122-147  (same dict keys, same file names)
  train_step .................................. src/train_utils.py:149-183
  validate .................................... src/train_utils.py:185-212
  train ....................................... src/train_utils.py:214-301
  L2SupervisedTrainer.train_step .............. src/scripts/train_vae_l2_supervised.py:140-175

Differences that are deliberate (documented in DESIGN.md):
  * `pixel_mse` is measured on the reconstruction of the SAME forward pass that produced the loss (one fused
    reduction) instead of a second stochastic forward (src/train_utils.py:165-168): same estimator, fresh-noise
    draw skipped, 165.8 GFLOP/sample saved. `exact_pixel_mse=True` restores the extra forward.
  * one device->host synchronisation per step for all metrics instead of one `.item()` per metric.
"""
import json
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn
from tqdm import tqdm

from .model import ENGINE
from .optim import FusedAdamW


def seed_all(seed: int):
    """Set all random seeds for reproducibility (also keys the device-side Philox eps stream)."""
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    ENGINE.rng_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    ENGINE.rng_offset = 0


def get_device() -> torch.device:
    """CUDA device with the most free memory (src/train_utils.py:24-38). The engine has no CPU path, so a machine
    without CUDA raises instead of silently returning 'cpu'."""
    if not torch.cuda.is_available():
        raise RuntimeError("tempo_vae_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
    free = []
    for i in range(torch.cuda.device_count()):
        free.append((torch.cuda.mem_get_info(i)[0], -i))       # ties go to the LOWEST index
    best = -max(free)[1]
    torch.cuda.set_device(best)       # the chosen GPU becomes the current device (kernels launch where the tensors live
    return torch.device(f"cuda:{best}")   # either way -- ops._on_device -- but allocations default to it from here on)


def get_sqrt_schedule(n_steps: int, n_saves: int = 100) -> List[int]:
    """Checkpoint steps spaced as sqrt(linspace) over the run, always ending at n_steps."""
    pts = (np.sqrt(np.linspace(0, 1, n_saves)) * n_steps).astype(int)
    steps = sorted(set(int(s) for s in pts))
    if n_steps not in steps:
        steps.append(n_steps)
    return steps


def _is_skipped_product(key: str, value: float) -> bool:
    """A `<PRODUCT>_loss` that came back NaN marks an L2 product without a single valid pixel in the batch: the
    reference emits NO key for it (src/model_with_l2.py:155-166), so neither do we -- a NaN would poison the running
    means of Trainer.train and the sums of validate(), and put non-standard tokens into metrics.json."""
    return key.endswith('_loss') and key not in ('loss', 'nll_loss', 'kl_loss') and value != value


def _to_floats(metrics: Dict[str, object]) -> Dict[str, float]:
    """Device scalars -> python floats with a single synchronising copy."""
    keys = [k for k, v in metrics.items() if torch.is_tensor(v)]
    out = {k: v for k, v in metrics.items() if not torch.is_tensor(v)}
    if keys:
        vals = torch.stack([metrics[k].detach().reshape(()).float() for k in keys]).tolist()
        out.update(dict(zip(keys, vals)))
    return {k: out[k] for k in metrics if not _is_skipped_product(k, out[k])}


class Trainer:
    """Simple trainer for VAE (same constructor and methods as the reference)."""

    def __init__(self, model: nn.Module, optimizer: torch.optim.Optimizer, device: torch.device, output_dir: Path,
                 save_every: int = 1000, val_every: int = 100, log_every: int = 10, plot_every: int = 50,
                 max_grad_norm: float = 1.0, exact_pixel_mse: bool = False):
        self.model = model
        self.optimizer = optimizer
        self.device = device
        self.output_dir = Path(output_dir)
        self.save_every = save_every
        self.val_every = val_every
        self.log_every = log_every
        self.plot_every = plot_every
        self.max_grad_norm = max_grad_norm
        self.exact_pixel_mse = exact_pixel_mse

        self.ckpt_dir = self.output_dir / 'checkpoints'
        self.ckpt_dir.mkdir(parents=True, exist_ok=True)
        self.summary_dir = self.output_dir / 'summary'
        self.summary_dir.mkdir(parents=True, exist_ok=True)

        self.train_metrics = []
        self.val_metrics = []
        self.step = 0

        self.plot_steps = []
        self.plot_losses = []
        self.plot_nll = []
        self.plot_kl = []
        self.plot_val_losses = []
        self.plot_pixel_mse = []

    # ------------------------------------------------------------------------------------------------ checkpoints
    def save_checkpoint(self, step: Optional[int] = None):
        if step is None:
            step = self.step
        checkpoint = {
            'step': step,
            'model_state_dict': self.model.state_dict(),
            'optimizer_state_dict': self.optimizer.state_dict(),
            'train_metrics': self.train_metrics,
            'val_metrics': self.val_metrics,
        }
        ckpt_path = self.ckpt_dir / f'ckpt_step={step:06d}.pt'
        torch.save(checkpoint, ckpt_path)
        return ckpt_path

    def load_checkpoint(self, ckpt_path: str):
        checkpoint = torch.load(ckpt_path, map_location=self.device)
        self.model.load_state_dict(checkpoint['model_state_dict'])
        self.optimizer.load_state_dict(checkpoint['optimizer_state_dict'])
        ENGINE.params_changed()
        self.step = checkpoint['step']
        self.train_metrics = checkpoint.get('train_metrics', [])
        self.val_metrics = checkpoint.get('val_metrics', [])
        print(f"Loaded checkpoint from step {self.step}")

    # ------------------------------------------------------------------------------------------------ one step
    def _loss_and_metrics(self, batch):
        """(loss, metrics-with-device-scalars) for one batch; overridden by the L2 trainer."""
        if batch.dtype != torch.bfloat16:      # bf16 batches (DeviceTileCache views) are consumed in place by the engine
            batch = batch.to(self.device, dtype=torch.float32, non_blocking=True)
        elif batch.device != self.device:
            batch = batch.to(self.device, non_blocking=True)
        if self.step == 0 and torch.is_grad_enabled():
            from . import ops
            mn, mx, mean, std = ops.batch_stats(batch).tolist()          # one fused reduction (tvae_batch_stats)
            print(f"Batch stats - min: {mn:.3f}, max: {mx:.3f}, mean: {mean:.3f}, std: {std:.3f}")
        loss, metrics = self.model.get_loss(batch)
        return batch, loss, dict(metrics)

    def _vae(self):
        m = self.model
        return m.vae if hasattr(m, "vae") else m

    def _optimizer_step(self):
        if isinstance(self.optimizer, FusedAdamW):
            self.optimizer.step(max_grad_norm=self.max_grad_norm)
        else:  # any torch optimiser still works (parameters are ordinary nn.Parameters)
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.max_grad_norm)
            self.optimizer.step()
            ENGINE.params_changed()

    def train_step_device(self, batch) -> Dict[str, torch.Tensor]:
        """One optimisation step; metrics stay on the device (no host synchronisation)."""
        self.model.train()
        batch, loss, metrics = self._loss_and_metrics(batch)
        if self.exact_pixel_mse:
            with torch.no_grad():
                x = batch if torch.is_tensor(batch) else batch['spectral']
                recon, _ = self._vae()(x)          # fresh noise draw, like the reference's second forward
                metrics['pixel_mse'] = torch.mean((x - recon) ** 2)
        else:
            metrics['pixel_mse'] = self._vae().last_pixel_mse()
        self.optimizer.zero_grad()
        prev = ENGINE.unit_loss_grad
        ENGINE.unit_loss_grad = True
        try:
            loss.backward()
        finally:
            ENGINE.unit_loss_grad = prev
        self._optimizer_step()
        return metrics

    def train_step(self, batch) -> Dict[str, float]:
        """Single training step (src/train_utils.py:149-183): returns python floats."""
        return _to_floats(self.train_step_device(batch))

    def train_step_accumulate(self, batches) -> Dict[str, torch.Tensor]:
        """One optimiser step over several micro-batches (gradient accumulation; the mean of the micro-batch losses is
        what gets differentiated, so equal-sized micro-batches reproduce the step on their concatenation). Extension of
        the reference API for global batches that do not fit one forward. Device metrics (means over micro-batches)."""
        self.model.train()
        if not isinstance(self.optimizer, FusedAdamW):
            raise TypeError("train_step_accumulate needs the flat-buffer FusedAdamW optimiser")
        batches = list(batches)
        self.optimizer.zero_grad()
        total = None
        prev = ENGINE.unit_loss_grad
        ENGINE.unit_loss_grad = True
        try:
            for b in batches:
                _, loss, metrics = self._loss_and_metrics(b)
                metrics['pixel_mse'] = self._vae().last_pixel_mse()
                loss.backward()                       # adds into the flat gradient buffer (wgrad accumulates in place)
                metrics = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in metrics.items()}
                total = metrics if total is None else {k: total[k] + v for k, v in metrics.items()}
        finally:
            ENGINE.unit_loss_grad = prev
        self.optimizer.step(max_grad_norm=self.max_grad_norm, grad_scale=1.0 / len(batches))
        return {k: v / len(batches) for k, v in total.items()}

    def validate(self, val_loader, n_batches: int = 10) -> Dict[str, float]:
        """Sample-weighted mean of get_loss metrics over n_batches, `val_` prefixed (src/train_utils.py:185-212)."""
        self.model.eval()
        rows = []
        with torch.no_grad():
            for i, batch in enumerate(val_loader):
                if i >= n_batches:
                    break
                b, _, metrics = self._loss_and_metrics(batch)
                bs = (b if torch.is_tensor(b) else b['spectral']).shape[0]
                rows.append((bs, metrics))
        if not rows:
            return {}
        # one device->host copy for all batches; a product that had no valid pixel in a batch (NaN) is left out of
        # that batch's contribution, exactly like the reference, whose metrics dict simply lacks the key there
        keys = list(rows[0][1])
        flat = torch.stack([torch.stack([(m[k].detach().reshape(()).double() if torch.is_tensor(m[k])
                                          else torch.tensor(float(m[k]), dtype=torch.float64, device=self.device))
                                         for k in keys]) for _, m in rows]).tolist()
        n_samples = sum(bs for bs, _ in rows)
        out = {}
        for j, k in enumerate(keys):
            vals = [(bs, r[j]) for (bs, _), r in zip(rows, flat) if not _is_skipped_product(k, r[j])]
            if vals:
                out[f'val_{k}'] = sum(bs * v for bs, v in vals) / n_samples
        return out

    # ------------------------------------------------------------------------------------------------ loop
    def train(self, train_loader, val_loader=None, n_steps: int = 10000):
        pbar = tqdm(total=n_steps, desc="Training", initial=self.step)
        train_iter = iter(train_loader)
        running_metrics = {}
        batch = None
        while self.step < n_steps:
            try:
                batch = next(train_iter)
            except StopIteration:
                train_iter = iter(train_loader)
                batch = next(train_iter)
            metrics = self.train_step(batch)
            self.step += 1

            alpha = 0.99 if running_metrics else 0.0
            for k, v in metrics.items():
                running_metrics[k] = alpha * running_metrics.get(k, 0) + (1 - alpha) * v

            if self.step % self.log_every == 0:
                self.train_metrics.append({'step': self.step, **running_metrics})
                self.plot_steps.append(self.step)
                self.plot_losses.append(running_metrics.get('loss', 0))
                self.plot_nll.append(running_metrics.get('nll_loss', 0))
                self.plot_kl.append(running_metrics.get('kl_loss', 0))
                self.plot_pixel_mse.append(running_metrics.get('pixel_mse', 0))
                pbar.set_postfix(running_metrics)

            if self.step % self.plot_every == 0 and self.step > 0:
                self.update_plots()

            if val_loader is not None and self.step % self.val_every == 0:
                val_metrics = self.validate(val_loader)
                self.val_metrics.append({'step': self.step, **val_metrics})
                tqdm.write(f"Step {self.step}: " + ", ".join(f"{k}={v:.4f}" for k, v in val_metrics.items()))

            if self.step % self.save_every == 0:
                ckpt_path = self.save_checkpoint()
                tqdm.write(f"Saved checkpoint: {ckpt_path}")
                self.save_reconstructions(batch, self.step)

            pbar.update(1)
        pbar.close()

        ckpt_path = self.save_checkpoint()
        print(f"Training complete. Final checkpoint: {ckpt_path}")
        self.save_metrics()

    def save_metrics(self):
        metrics_path = self.output_dir / 'metrics.json'
        with open(metrics_path, 'w') as f:
            json.dump({'train': self.train_metrics, 'val': self.val_metrics}, f, indent=2)
        print(f"Saved metrics to {metrics_path}")

    # ------------------------------------------------------------------------------------------------ cosmetics
    # Plotting is outside the hot path; it needs matplotlib, which is optional here.
    @staticmethod
    def _pyplot():
        try:
            import matplotlib
            matplotlib.use('Agg')
            import matplotlib.pyplot as plt
            return plt
        except Exception:  # noqa: BLE001
            return None

    def update_plots(self):
        plt = self._pyplot()
        if plt is None or not self.plot_steps:
            return
        fig, axes = plt.subplots(1, 4, figsize=(20, 4))
        for ax, ys, name in zip(axes, (self.plot_losses, self.plot_nll, self.plot_kl, self.plot_pixel_mse),
                                ("loss", "nll_loss", "kl_loss", "pixel_mse")):
            ax.plot(self.plot_steps, ys)
            ax.set_title(name)
            ax.set_xlabel("step")
        fig.tight_layout()
        fig.savefig(self.summary_dir / 'training_curves.png', dpi=100)
        plt.close(fig)

    def save_reconstructions(self, batch, step):
        plt = self._pyplot()
        if plt is None or batch is None:
            return
        self.model.eval()
        x = batch if torch.is_tensor(batch) else batch['spectral']
        with torch.no_grad():
            x = x[:4].to(self.device, dtype=torch.float32)
            recon = self.model.forward(x) if torch.is_tensor(batch) else self.model.forward(x)['reconstruction']
        x, recon = x.cpu().numpy(), recon.cpu().numpy()
        fig, axes = plt.subplots(2, x.shape[0], figsize=(4 * x.shape[0], 8), squeeze=False)
        ch = x.shape[1] // 2
        for i in range(x.shape[0]):
            axes[0, i].imshow(x[i, ch]); axes[0, i].set_title(f"input ch {ch}")
            axes[1, i].imshow(recon[i, ch]); axes[1, i].set_title("reconstruction")
        fig.tight_layout()
        fig.savefig(self.summary_dir / f'reconstructions_step={step:06d}.png', dpi=100)
        plt.close(fig)


class L2SupervisedTrainer(Trainer):
    """Trainer for dict batches {'spectral', 'NO2', 'O3TOT', 'HCHO', 'CLDO4'} over VAEWithL2Supervision
    (src/scripts/train_vae_l2_supervised.py:27-292)."""

    def __init__(self, model, optimizer, device, output_dir, kl_weight: float = 1e-6, l2_weights=None, **kw):
        super().__init__(model, optimizer, device, output_dir, **kw)
        self.kl_weight = kl_weight
        self.l2_weights = l2_weights

    def _loss_and_metrics(self, batch):
        batch = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}
        loss, metrics = self.model.compute_loss(batch, kl_weight=self.kl_weight, l2_weights=self.l2_weights,
                                                return_device_metrics=True)
        return batch, loss, dict(metrics)
