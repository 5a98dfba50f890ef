"""Linear / MLP probes on VAE latents — the reference's `LinearProbe`, `MLPProbe` and
`train_probe` (src/scripts/linear_probe_analysis.py:212-353) on the B200 engine.

Same classes, constructor arguments and state_dict keys (`linear.weight [1, 32]`, `mlp.{0,3,6}.weight`, ...: the
parameters are nn.Linear's, created in the same order, so a seed gives the reference's initial weights), same
training procedure (AdamW(lr, weight_decay), MSE, one `torch.randperm` on the host per epoch, mini-batches of
`batch_size`, the last one partial, full-batch validation every epoch, best-epoch bookkeeping). Execution:
  * every nn.Linear is a 1x1 convolution over a [1, 1, rows, C] channels-last tensor: `tvae_conv_gemm` forward (bias in
    the epilogue) and data gradient, `tvae_wgrad_gemm` weight gradient -- the tcgen05 kernels of the VAE itself (rows
    are padded to a multiple of 128 with zero-gradient rows);
  * activation + dropout: `tvae_act_dropout_fwd / _bwd` (Philox keep mask regenerated in backward, never stored);
  * loss, its gradient and the R^2 sums: `tvae_probe_mse`; optimiser: FusedAdamW (`tvae_adamw`).
One device->host read per EPOCH (train and validation loss together) instead of one `.item()` per mini-batch.

Reference behaviours kept on purpose: `best_state = probe.state_dict().copy()` is a SHALLOW copy, so the weights the
reference returns are the last epoch's, not the best epoch's (`best_epoch` / `best_val_loss` are still tracked and
printed); dropout draws come from the device Philox stream (ENGINE seed) instead of torch's generator, so runs with
dropout > 0 agree with the reference statistically, runs with dropout = 0 step for step.
"""
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import TvaeError, lib
from .model import ENGINE, _PackedMixin, conv_bwd, conv_fwd
from .optim import FusedAdamW

_ACT = {"relu": 2, "gelu": 1, "tanh": 4}


class _Linear(nn.Linear, _PackedMixin):
    """nn.Linear parameters and init; executed as a 1x1 convolution by the engine."""

    @property
    def in_channels(self):
        return self.in_features

    @property
    def out_channels(self):
        return self.out_features

    def conv_kind(self):
        return 0, 1

    def forward(self, x):  # noqa: D401
        self._no_direct_call()


class _ProbeBase(nn.Module):
    input_dim: int

    # ---- engine program over rows: X fp32 [n, input_dim] on the device -> prediction fp32 [n_pad, pitch] (column 0)
    def _layers(self):
        raise NotImplementedError

    def _run_fwd(self, X, train: bool, seed: int, offset: int):
        n = X.shape[0]
        n_pad = ops.round_up(max(n, 1), 128)
        cin = self.input_dim
        ENGINE.begin_forward()
        xb = torch.zeros((1, 1, n_pad, ops.round_up(cin, 8)), dtype=torch.bfloat16, device=X.device)
        ops.rows_f32_to_bf16(X, xb.view(n_pad, -1)[:n])
        saved = []
        h = xb
        layers = self._layers()
        for li, (lin, act, p) in enumerate(layers):
            out, _ = conv_fwd(lin, h, cin)                       # fp32 [1, 1, n_pad, round_up(cout, 4)], bias included
            cout = lin.out_features
            if act is None:
                saved.append((lin, h, cin, None, None, 0.0))
                return out.view(n_pad, -1), saved, n_pad
            pd = p if train else 0.0
            a = ops.act_dropout_fwd(out.view(n_pad, -1), cout, act, pd, seed + 7919 * (li + 1), offset)
            saved.append((lin, h, cin, out, act, pd))
            h, cin = a.view(1, 1, n_pad, -1), cout
        raise TvaeError("probe without an output layer")

    def _run_bwd(self, dpred, saved, n_pad, seed: int, offset: int):
        d = dpred.view(1, 1, n_pad, -1)
        for li in reversed(range(len(saved))):
            lin, h_in, cin, pre, act, pd = saved[li]
            if act is not None:
                d = ops.act_dropout_bwd(pre.view(n_pad, -1), d.view(n_pad, -1), lin.out_features, act, pd,
                                        seed + 7919 * (li + 1), offset).view(1, 1, n_pad, -1)
            d = conv_bwd(lin, d, h_in, cin, dgrad=("bf16" if li > 0 else None))

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [n, input_dim] -> [n, 1] (inference: dropout off, no autograd graph)."""
        ops.require_cuda(x, "probe input")
        X = x.float()
        if X.dim() != 2 or X.stride(1) != 1:
            X = X.reshape(-1, self.input_dim).contiguous()
        pred, _, _ = self._run_fwd(X, False, 0, 0)
        return pred[:X.shape[0], :1].clone()


class LinearProbe(_ProbeBase):
    """Simple linear probe from latent channels to component (src/scripts/linear_probe_analysis.py:212-219)."""

    def __init__(self, input_dim=32, output_dim=1):
        super().__init__()
        if output_dim != 1:
            raise TvaeError("probes predict one component (output_dim = 1), like every call site of the reference")
        self.input_dim = input_dim
        self.linear = _Linear(input_dim, output_dim)

    def _layers(self):
        return [(self.linear, None, 0.0)]


class MLPProbe(_ProbeBase):
    """MLP probe with configurable hidden layers (src/scripts/linear_probe_analysis.py:222-252)."""

    def __init__(self, input_dim=32, hidden_dims=[512, 512], output_dim=1, dropout=0.1, activation='relu'):
        super().__init__()
        if output_dim != 1:
            raise TvaeError("probes predict one component (output_dim = 1), like every call site of the reference")
        if activation not in _ACT:
            raise ValueError(f"Unknown activation: {activation}")
        self.input_dim = input_dim
        self.dropout = float(dropout)
        self.activation = activation
        layers = []
        prev = input_dim
        for hd in hidden_dims:
            layers.append(_Linear(prev, hd))
            layers.append(nn.Identity())                 # placeholder at the index of the reference's activation module
            if dropout > 0:
                layers.append(nn.Identity())             # ... and of its nn.Dropout (keeps the state_dict indices)
            prev = hd
        layers.append(_Linear(prev, output_dim))
        self.mlp = nn.Sequential(*layers)

    def _layers(self):
        lins = [m for m in self.mlp if isinstance(m, _Linear)]
        out = [(m, _ACT[self.activation], self.dropout) for m in lins[:-1]]
        out.append((lins[-1], None, 0.0))
        return out


def _mse(pred, y, n, n_pad, want_grad):
    sums = torch.empty((3,), dtype=torch.float64, device=pred.device)
    dpred = torch.empty((n_pad, 8), dtype=torch.bfloat16, device=pred.device) if want_grad else None
    ops.probe_mse(pred, y, n, n_pad, sums, dpred)
    return sums, dpred


def probe_metrics(probe: _ProbeBase, X: torch.Tensor, y: torch.Tensor, chunk: int = 65536) -> Dict[str, float]:
    """{'mse', 'r2_score'} of the probe on (X [n, input_dim], y [n]) -- sklearn's r2_score / mean_squared_error of
    src/scripts/linear_probe_analysis.py:676-681, reduced on the device."""
    X = X.float().contiguous()
    y = y.float().reshape(-1)
    tot = torch.zeros((3,), dtype=torch.float64, device=X.device)
    with torch.no_grad():
        for i in range(0, X.shape[0], chunk):
            xb, yb = X[i:i + chunk], y[i:i + chunk]
            pred, _, n_pad = probe._run_fwd(xb, False, 0, 0)
            sums, _ = _mse(pred, yb, xb.shape[0], n_pad, False)
            tot += sums
    ss_res, sy, syy = tot.tolist()
    n = X.shape[0]
    ss_tot = syy - sy * sy / n
    return {"mse": ss_res / n, "r2_score": 1.0 - ss_res / ss_tot if ss_tot > 0 else float("nan")}


def train_probe(X_train, y_train, X_val, y_val, config, device=None, verbose: bool = True):
    """Train a probe (linear or MLP): same contract as src/scripts/linear_probe_analysis.py:255-353 -- returns
    (probe, train_losses, val_losses). Inputs may be numpy arrays or tensors."""
    if not torch.cuda.is_available():
        raise TvaeError("train_probe runs on CUDA only (there is no CPU fallback)")
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    architecture = config.get('architecture', 'linear')
    if architecture == 'mlp':
        probe = MLPProbe(input_dim=32, hidden_dims=config.get('hidden_dims', [512, 512]), output_dim=1,
                         dropout=config.get('dropout', 0.1), activation=config.get('activation', 'relu')).to(device)
    else:
        probe = LinearProbe(input_dim=32, output_dim=1).to(device)
    weight_decay = config.get('weight_decay', 0.01)
    optimizer = FusedAdamW(probe.parameters(), lr=config['learning_rate'], weight_decay=weight_decay)

    def dev(a):
        return torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a, dtype=torch.float32).to(device).contiguous()
    X_train, y_train, X_val, y_val = dev(X_train), dev(y_train).reshape(-1), dev(X_val), dev(y_val).reshape(-1)
    n_train = X_train.shape[0]
    batch_size = config.get('batch_size', 512)
    n_batches = (n_train + batch_size - 1) // batch_size
    train_losses: List[float] = []
    val_losses: List[float] = []
    best_val_loss = float('inf')
    best_state = None
    best_epoch = 0
    seed = int(ENGINE.rng_seed)
    drop_offset = 0
    # [X | y | pad]: one matrix with 16-byte-granular rows, so ONE tvae_gather_rows per epoch applies the permutation to
    # inputs and targets together; batches are row slices of it (X read with its row pitch, y as a strided column)
    D = X_train.shape[1]
    width = ops.round_up(D + 1, 4)
    XY = torch.zeros((n_train, width), dtype=torch.float32, device=device)
    XY[:, :D] = X_train
    XY[:, D] = y_train
    XYs = torch.empty_like(XY)
    for epoch in range(config['max_epochs']):
        probe.train()
        perm = torch.randperm(n_train).to(device)                 # host RNG, like the reference
        ops.gather_rows(XY, perm, XYs)
        epoch_sum = torch.zeros((), dtype=torch.float64, device=device)
        for b in range(n_batches):
            s, e = b * batch_size, min((b + 1) * batch_size, n_train)
            xb, yb = XYs[s:e, :D], XYs[s:e, D]
            optimizer.zero_grad()
            pred, saved, n_pad = probe._run_fwd(xb, True, seed, drop_offset)
            sums, dpred = _mse(pred, yb, e - s, n_pad, True)
            probe._run_bwd(dpred, saved, n_pad, seed, drop_offset)
            ENGINE.join_side_streams()
            optimizer.step()
            drop_offset += n_pad
            epoch_sum += sums[0]                                  # batch_loss * batch rows = sum of squared errors
        probe.eval()
        with torch.no_grad():
            vm = torch.zeros((3,), dtype=torch.float64, device=device)
            for i in range(0, X_val.shape[0], 65536):
                xb, yb = X_val[i:i + 65536], y_val[i:i + 65536]
                pred, _, n_pad = probe._run_fwd(xb, False, 0, 0)
                vm += _mse(pred, yb, xb.shape[0], n_pad, False)[0]
        train_sse, val_sse = torch.stack([epoch_sum, vm[0]]).tolist()          # ONE host sync per epoch
        train_loss, val_loss = train_sse / n_train, val_sse / X_val.shape[0]
        train_losses.append(train_loss)
        val_losses.append(val_loss)
        if val_loss < best_val_loss:
            best_val_loss = val_loss
            best_state = probe.state_dict().copy()                # shallow, like the reference (see module docstring)
            best_epoch = epoch
        if verbose and epoch % 100 == 0:
            print(f"Epoch {epoch}: Train Loss = {train_loss:.4f}, Val Loss = {val_loss:.4f}")
    if best_state is not None:
        probe.load_state_dict(best_state)
        ENGINE.params_changed()
    if verbose:
        print(f"Best model from epoch {best_epoch} with val loss {best_val_loss:.4f}")
    return probe, train_losses, val_losses
