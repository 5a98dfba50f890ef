"""500-step loss-curve parity (north star: "a loss curve within 1% over 500 steps"): the CUDA path and the fp32
oracle (plain torch ops on the same GPU, TF32 off) train the DEFAULT model from the same init on the same
structured synthetic batches with the same injected noise. Writes gpurun_out/parity_500.json."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tempo_vae_oracle as orc  # noqa: E402
import tempo_vae_b200 as t  # noqa: E402
from bench import DEFAULT_MODEL  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda")
cfg = orc.DEFAULT_CFG
t.seed_all(42)
model = t.get_model(DEFAULT_MODEL, dev)
params = {k: v.detach().clone() for k, v in model.state_dict().items()}
state = {}
g = torch.Generator().manual_seed(99)
curves = {k: [] for k in ("loss", "oracle_loss", "pixel_mse", "oracle_pixel_mse", "kl_loss", "oracle_kl_loss",
                          "logvar", "oracle_logvar")}
t0 = time.time()
for step in range(1, steps + 1):
    x = orc.structured_batch(B, cfg, seed=5000 + step).to(dev)
    eps = torch.randn((B, 32, 16, 16), generator=g).to(dev)
    grads, out = orc.grads_of(lambda leaves: orc.vae_loss(leaves, x, eps, cfg), params)
    orc.clip_and_adamw(params, grads, state, step=step)
    loss, metrics = model.get_loss(x, eps=eps)
    model.optimizer.zero_grad()
    loss.backward()
    model.optimizer.step(max_grad_norm=1.0)
    curves["loss"].append(loss.item()); curves["oracle_loss"].append(out["loss"].item())
    curves["pixel_mse"].append(model.vae.last_pixel_mse().item()); curves["oracle_pixel_mse"].append(out["pixel_mse"].item())
    curves["kl_loss"].append(metrics["kl_loss"].item()); curves["oracle_kl_loss"].append(out["kl_loss"].item())
    curves["logvar"].append(model.vae.logvar.item()); curves["oracle_logvar"].append(params["vae.logvar"].item())
    if step % 50 == 0:
        print(step, {k: v[-1] for k, v in curves.items()}, f"{time.time() - t0:.0f}s", flush=True)


def worst(a, b):
    return max(abs(x - y) / abs(y) for x, y in zip(curves[a], curves[b]))


summary = {"steps": steps, "batch": B, "model": "default TEMPO-VAE (27.3 M parameters), structured synthetic patches",
           "max_rel_dev": {"loss": worst("loss", "oracle_loss"), "pixel_mse": worst("pixel_mse", "oracle_pixel_mse"),
                           "kl_loss": worst("kl_loss", "oracle_kl_loss"), "logvar": worst("logvar", "oracle_logvar")},
           "final": {k: v[-1] for k, v in curves.items()},
           "every_50": {k: v[49::50] for k, v in curves.items()}}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "parity_500.json"), "w") as f:
    json.dump(summary, f, indent=1)
print(json.dumps(summary["max_rel_dev"]))
