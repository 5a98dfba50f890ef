"""North star: "a loss curve within 1 % over 500 steps". The CUDA path and the fp32 oracle (plain torch ops on the same
GPU, TF32 off) train the DEFAULT model (27.3 M parameters) from the same seed-42 init on the same structured synthetic
batches with the same injected noise, 500 AdamW steps each.

A third run puts the result in context: the oracle under `orc.bf16_operands()` -- an IDEAL bf16-operand engine (conv
inputs, weights and the gradient stream rounded to bf16, everything else fp32). Training is chaotic, so two runs that
differ only by rounding drift apart over hundreds of steps whoever does the rounding; the informative curves
(pixel_mse, kl) of the engine are held to the 1 % / 3 % bands AND reported next to the drift of that ideal run.

Asserted: loss within 1e-2 at EVERY step (measured 2e-6); pixel_mse within 3e-2, kl_loss within 5e-2 at every step.
The curves are written to gpurun_out/parity_500.json (copied to profiles/ by the builder)."""
import json
import os
import sys
import time

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tempo_vae_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


def test_500_step_loss_curve_default_model(capsys):
    import tempo_vae_b200 as t
    from bench import DEFAULT_MODEL
    steps = int(os.environ.get("TVAE_PARITY_STEPS", "500"))
    B = 4
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        dev = torch.device("cuda")
        cfg = orc.DEFAULT_CFG
        t.seed_all(42)
        model = t.get_model(DEFAULT_MODEL, dev)
        init = {k: v.detach().clone() for k, v in model.state_dict().items()}
        ref_init = orc.init_state_dict(cfg, seed=42)                       # the oracle's own seed-42 weights ...
        assert all(torch.equal(init[k].cpu(), ref_init[k]) for k in ref_init)   # ... are the product's, bit for bit
        runs = {"oracle": ({k: v.clone() for k, v in init.items()}, {}),
                "ideal_bf16": ({k: v.clone() for k, v in init.items()}, {})}
        g = torch.Generator().manual_seed(99)
        names = ("loss", "pixel_mse", "kl_loss", "logvar")
        curves = {f"{r}_{n}": [] for r in ("engine", "oracle", "ideal_bf16") for n in names}
        t0 = time.time()
        for step in range(1, steps + 1):
            x = orc.structured_batch(B, cfg, seed=5000 + step).to(dev)
            eps = torch.randn((B, 32, 16, 16), generator=g).to(dev)
            for name, (params, state) in runs.items():
                if name == "ideal_bf16":
                    with orc.bf16_operands():
                        grads, out = orc.grads_of(lambda leaves: orc.vae_loss(leaves, x, eps, cfg), params)
                else:
                    grads, out = orc.grads_of(lambda leaves: orc.vae_loss(leaves, x, eps, cfg), params)
                orc.clip_and_adamw(params, grads, state, step=step)
                for n in ("loss", "pixel_mse", "kl_loss"):
                    curves[f"{name}_{n}"].append(out[n].item())
                curves[f"{name}_logvar"].append(params["vae.logvar"].item())
                del grads, out
            loss, metrics = model.get_loss(x, eps=eps)
            model.optimizer.zero_grad()
            loss.backward()
            model.optimizer.step(max_grad_norm=1.0)
            curves["engine_loss"].append(loss.item())
            curves["engine_pixel_mse"].append(model.vae.last_pixel_mse().item())
            curves["engine_kl_loss"].append(metrics["kl_loss"].item())
            curves["engine_logvar"].append(model.vae.logvar.item())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32

    def worst(run, n):
        return max(abs(a - b) / abs(b) for a, b in zip(curves[f"{run}_{n}"], curves[f"oracle_{n}"]))

    summary = {"steps": steps, "batch": B, "seconds": round(time.time() - t0, 1),
               "model": "default TEMPO-VAE (27.3 M parameters), structured synthetic patches",
               "max_rel_dev_vs_fp32_oracle": {"engine": {n: worst("engine", n) for n in names},
                                              "ideal_bf16_operand_oracle": {n: worst("ideal_bf16", n) for n in names}},
               "final": {k: v[-1] for k, v in curves.items()},
               "every_50": {k: v[49::50] for k, v in curves.items()}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_500.json"), "w") as f:
        json.dump(summary, f, indent=1)
    with capsys.disabled():
        print(f"\n[{steps}-step curve, default model] max relative deviation from the fp32 oracle: "
              + json.dumps(summary["max_rel_dev_vs_fp32_oracle"]) + f" ({summary['seconds']} s)")
    e = summary["max_rel_dev_vs_fp32_oracle"]["engine"]
    assert e["loss"] < 1e-2, e
    assert e["pixel_mse"] < 3e-2 and e["kl_loss"] < 5e-2 and e["logvar"] < 1e-3, e
