"""Probe training on latents through the C ABI against tests/golden/probes.pt -- produced by the
reference's OWN LinearProbe / MLPProbe / train_probe (src/scripts/linear_probe_analysis.py:212-353, run by
oracle/make_golden_data.py) -- and against the oracle restatement.

Tolerances: the engine feeds the Linear layers bf16 operands (fp32 accumulate): per-epoch losses within 2e-2 (linear) /
5e-2 (MLP, dropout 0) of the reference's curve from the same seed (same initial weights, same torch.randperm sequence);
R^2 within 0.01 / 0.02. With dropout > 0 the keep masks come from the device Philox stream instead of torch's generator:
final losses within 20 %, R^2 within 0.05."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tempo_vae_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu
FX = os.path.join(ROOT, "tests", "golden", "probes.pt")


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("name,curve_tol,r2_tol", [("linear", 2e-2, 0.01), ("mlp_relu_nodrop", 5e-2, 0.02),
                                                    ("mlp_gelu_drop", None, 0.05)])
def test_train_probe_matches_the_reference_run(name, curve_tol, r2_tol, capsys):
    import tempo_vae_b200 as t
    fx = torch.load(FX, weights_only=False)
    c = fx["cases"][name]
    t.seed_all(3)
    torch.manual_seed(fx["seed"])                    # the reference run's seed: initial weights and permutations
    probe, tl, vl = t.train_probe(fx["X_train"].numpy(), fx["y_train"].numpy(), fx["X_val"].numpy(), fx["y_val"].numpy(),
                                  c["config"], verbose=False)
    assert list(probe.state_dict().keys()) == list(c["state_dict"].keys())
    assert all(probe.state_dict()[k].shape == v.shape for k, v in c["state_dict"].items())
    assert len(tl) == len(c["train_losses"]) == c["config"]["max_epochs"]
    dev_t = max(abs(a - b) / b for a, b in zip(tl, c["train_losses"]))
    dev_v = max(abs(a - b) / b for a, b in zip(vl, c["val_losses"]))
    m = t.probe_metrics(probe, fx["X_val"].cuda(), fx["y_val"].cuda())
    ref_r2 = orc.r2_score(fx["y_val"], c["pred_val"])
    pred = probe(fx["X_val"].cuda()).squeeze(1)
    assert abs(m["r2_score"] - orc.r2_score(fx["y_val"], pred.cpu())) < 1e-4          # device reduction = sklearn formula
    assert abs(m["mse"] - float(((pred.cpu() - fx["y_val"]) ** 2).mean())) / m["mse"] < 1e-4
    with capsys.disabled():
        print(f"\n[probe {name}] max rel deviation of the loss curves: train {dev_t:.3e} val {dev_v:.3e}; "
              f"R^2 {m['r2_score']:.4f} (reference {ref_r2:.4f})")
    if curve_tol is not None:
        assert dev_t < curve_tol and dev_v < curve_tol, (dev_t, dev_v)
        assert rel(pred, c["pred_val"]) < 5e-2
    else:
        assert abs(tl[-1] - c["train_losses"][-1]) / c["train_losses"][-1] < 0.2
        assert abs(vl[-1] - c["val_losses"][-1]) / c["val_losses"][-1] < 0.2
    assert abs(m["r2_score"] - ref_r2) < r2_tol


def test_probe_forward_with_reference_weights():
    """Loading the reference's trained state_dict into the engine's probe reproduces its predictions (bf16 operands)."""
    import tempo_vae_b200 as t
    fx = torch.load(FX, weights_only=False)
    for name in ("linear", "mlp_relu_nodrop", "mlp_gelu_drop"):
        c = fx["cases"][name]
        cfg = c["config"]
        if cfg["architecture"] == "mlp":
            probe = t.MLPProbe(32, cfg["hidden_dims"], 1, cfg["dropout"], cfg["activation"]).cuda()
        else:
            probe = t.LinearProbe(32, 1).cuda()
        probe.load_state_dict(c["state_dict"])
        probe.eval()
        pred = probe(fx["X_val"].cuda())
        assert pred.shape == (fx["X_val"].shape[0], 1)
        assert rel(pred.squeeze(1), c["pred_val"]) < 1e-2, name
        odd = probe(fx["X_val"][:77].cuda())                               # a row count that is not a multiple of 128
        assert torch.equal(odd, pred[:77])


def test_activation_dropout_kernels():
    """tvae_act_dropout_fwd / _bwd: activations against torch, keep probability, 1/(1-p) scaling, and the backward
    regenerating exactly the forward's mask."""
    from tempo_vae_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    rows, C = 1024, 64
    x = torch.randn((rows, 68), device="cuda", generator=g)            # row pitch 68 > C
    for act, fn in ((2, torch.relu), (1, torch.nn.functional.gelu), (4, torch.tanh), (0, lambda v: v)):
        out = ops.act_dropout_fwd(x, C, act, 0.0, 1, 0)
        assert out.shape == (rows, 64)
        assert torch.equal(out, fn(x[:, :C]).to(torch.bfloat16)) or rel(out, fn(x[:, :C])) < 4e-3
    p = 0.25
    a0 = ops.act_dropout_fwd(x, C, 2, 0.0, 5, 1000).float()
    a1 = ops.act_dropout_fwd(x, C, 2, p, 5, 1000).float()
    a2 = ops.act_dropout_fwd(x, C, 2, p, 5, 1000).float()
    a3 = ops.act_dropout_fwd(x, C, 2, p, 5, 1000 + rows).float()
    assert torch.equal(a1, a2) and not torch.equal(a1, a3)                  # counter-based: reproducible, offset-keyed
    pos = a0 > 0
    kept = (a1 != 0) & pos
    frac = float(kept.sum()) / float(pos.sum())
    assert abs(frac - (1 - p)) < 0.02
    assert rel(a1[kept], a0[kept] / (1 - p)) < 8e-3
    da = torch.randn((rows, 64), device="cuda", generator=g).to(torch.bfloat16)
    dx = ops.act_dropout_bwd(x, da, C, 2, p, 5, 1000).float()
    expect = da.float() * (x[:, :C] > 0) * (a1 != 0) / (1 - p)
    assert rel(dx, expect) < 8e-3
    assert torch.equal(dx != 0, (a1 != 0) & (da.float() != 0))
