"""How much forward error do bf16 tensor-core operands cost BY THEMSELVES on the parity fixtures?

Runs on the CPU (no GPU, no product code). Three restatements of the forward pass on the same weights/input/noise:
  fp32      : the oracle as is (== the reference's numbers, tests/golden)
  ideal     : the oracle with every convolution's input AND weight rounded to bf16, fp32 accumulation, everything
              else (GroupNorm, GELU, residual stream, attention, softmax) in fp32 -- the best ANY bf16-operand
              engine can do; this engine's design point
  autocast  : the oracle under torch.autocast(bf16) -- what the reference itself gives in PyTorch's bf16 mode
              (activations stored in bf16 as well)
and prints the relative L2 error of mean / logvar / recon of `ideal` and `autocast` against fp32. The output is
committed as profiles/bf16_floor_r2.json and quoted by tests/test_model_gpu.py's tolerance header.
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tempo_vae_oracle as orc  # noqa: E402


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def bf(x):
    return x.to(torch.bfloat16).float()


def forward(sd, x, eps, cfg, mode):
    real_conv2d, real_convT = F.conv2d, F.conv_transpose2d
    if mode == "ideal":
        F.conv2d = lambda inp, w, b=None, **kw: real_conv2d(bf(inp), bf(w), b, **kw)
        F.conv_transpose2d = lambda inp, w, b=None, **kw: real_convT(bf(inp), bf(w), b, **kw)
    try:
        with torch.no_grad():
            if mode == "autocast":
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    mean, logvar, _ = orc.encode(sd, x, cfg)
                    z = mean.float() + torch.exp(0.5 * logvar.float()) * eps
                    recon = orc.decode(sd, z, cfg)
            else:
                mean, logvar, _ = orc.encode(sd, x, cfg)
                z = mean + torch.exp(0.5 * logvar) * eps
                recon = orc.decode(sd, z, cfg)
    finally:
        F.conv2d, F.conv_transpose2d = real_conv2d, real_convT
    return mean.float(), logvar.float(), recon.float()


def main():
    out = {}
    fx = torch.load(os.path.join(ROOT, "tests/golden/tiny_train.pt"), weights_only=False)
    cfg, sd = fx["cfg"], fx["state_dict"]
    x, eps = fx["x"][0], fx["eps"][0]
    ref = forward(sd, x, eps, cfg, "fp32")
    for mode in ("ideal", "autocast"):
        got = forward(sd, x, eps, cfg, mode)
        out[f"tiny_{mode}"] = dict(zip(("mean", "logvar", "recon"), (rel(g, r) for g, r in zip(got, ref))))
    if "--default" in sys.argv:
        fxd = torch.load(os.path.join(ROOT, "tests/golden/default_train_b2.pt"), weights_only=False)
        cfg = fxd["cfg"]
        sys.path.insert(0, "/root/reference")
        torch.manual_seed(42)
        from src.model import get_model          # seed-42 constructor weights of the real reference (CPU)
        params = dict(architecture_type="vae", architecture_params=dict(enc_dec_params=dict(
            shape=list(cfg["shape"]), embed_dim=32, chs=cfg["chs"], attn_sizes=[], mid_attn=True, num_res_blocks=1,
            dropout_prob=0.0, z_channels=32, double_z=True, n_attention_heads=4, norm_groups=8, norm_eps=1e-6,
            norm_affine=True, act="gelu", conv_kernel_size=3, conv_padding_mode="zeros", kl_weight=1e-6,
            nll_loss_type="l1")), optimizer_type="AdamW", optimizer_params=dict(lr=1e-4, betas=[0.9, 0.95], weight_decay=0.05))
        import numpy as np
        np.random.seed(42)
        m = get_model(params, torch.device("cpu"))
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        orc.rerandomize_zero_init(sd, seed=1234)
        x = orc.structured_batch(fxd["B"], cfg, seed=fxd["x_seeds"][0])
        eps = fxd["eps"][0]
        ref = forward(sd, x, eps, cfg, "fp32")
        s0 = fxd["steps"][0]
        out["default_fp32_vs_golden"] = dict(mean=rel(ref[0], s0["mean"]), logvar=rel(ref[1], s0["logvar"]),
                                             recon=rel(ref[2][:, ::16, ::4, ::4], s0["recon"]))
        for mode in ("ideal", "autocast"):
            got = forward(sd, x, eps, cfg, mode)
            out[f"default_b2_{mode}"] = dict(zip(("mean", "logvar", "recon"), (rel(g, r) for g, r in zip(got, ref))))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
