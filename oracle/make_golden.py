"""Generate tests/golden/*.pt by RUNNING THE REAL REFERENCE (imported from /root/reference; CPU, fp32).

Run in the build container only:  python oracle/make_golden.py
The fixtures pin oracle/tempo_vae_oracle.py and are the known-answer vectors of the GPU parity tests (the GPU box
has no /root/reference). Every fixture stores its inputs, so nothing has to be regenerated at test time, except
the default-size weights, which are re-created from seed 42 (bit-identical constructor RNG, checked in the tests).
"""
import os
import sys

import torch

REF = os.environ.get("TVAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

import src.model as ref_model  # noqa: E402
import src.model_with_l2 as ref_l2  # noqa: E402
import tempo_vae_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(max(1, os.cpu_count() or 1))


def model_params(cfg, lr=1e-4):
    edp = {k: cfg[k] for k in ("shape", "chs", "attn_sizes", "mid_attn", "num_res_blocks", "z_channels", "double_z",
                               "n_attention_heads", "norm_groups", "norm_eps", "act")}
    edp.update(embed_dim=cfg["embed_dim"], kl_weight=cfg["kl_weight"], nll_loss_type=cfg["nll_loss_type"])
    return dict(architecture_type="vae", architecture_params=dict(enc_dec_params=edp), optimizer_type="AdamW",
                optimizer_params=dict(lr=lr, betas=[0.9, 0.95], weight_decay=0.05))


class EpsInjector:
    """Makes the reference's posterior.sample() consume a supplied list of eps tensors (in call order)."""

    def __init__(self, eps_list):
        self.eps = list(eps_list)
        self.orig = ref_model.DiagonalGaussianDistribution.sample

    def __enter__(self):
        inj = self

        def sample(dist):
            e = inj.eps.pop(0)
            return dist.mean + dist.std * e
        ref_model.DiagonalGaussianDistribution.sample = sample
        return self

    def __exit__(self, *a):
        ref_model.DiagonalGaussianDistribution.sample = self.orig


def build_ref(cfg, seed=42, rerandomize=True):
    torch.manual_seed(seed)
    model = ref_model.get_model(model_params(cfg), torch.device("cpu"))
    if rerandomize:
        sd = model.state_dict()
        orc.rerandomize_zero_init(sd, seed=1234)
        model.load_state_dict(sd)
    return model


def train_fixture(cfg, B, steps, tag, full):
    model = build_ref(cfg)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    C, H, W = cfg["shape"]
    Z, hz = cfg["embed_dim"], H // 2 ** (len(cfg["chs"]) - 1)
    g = torch.Generator().manual_seed(7)
    xs = [orc.structured_batch(B, cfg, seed=100 + s) for s in range(steps)]
    eps = [torch.randn((B, Z, hz, hz), generator=g) for _ in range(steps)]
    fx = dict(cfg=cfg, B=B, x=xs if full else None, x_seeds=[100 + s for s in range(steps)], eps=eps, steps=[])
    if full:
        fx["state_dict"] = sd0
    opt = model.optimizer
    for s in range(steps):
        with EpsInjector([eps[s]]):
            loss, metrics = model.get_loss(xs[s])
            with torch.no_grad(), EpsInjector([eps[s]]):
                recon, post = model.vae(xs[s])
        opt.zero_grad()
        loss.backward()
        grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()}
        gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        opt.step()
        rec = dict(loss=loss.item(), nll_loss=metrics["nll_loss"].item(), kl_loss=metrics["kl_loss"].item(),
                   pixel_mse=torch.mean((xs[s] - recon) ** 2).item(), grad_norm=float(gnorm),
                   logvar_after=model.vae.logvar.item(),
                   grad_norms={k: (float(v.norm()) if v is not None else None) for k, v in grads.items()})
        if full or s == 0:
            rec["mean"] = post.mean.clone()
            rec["logvar"] = post.logvar.clone()
            rec["recon"] = recon.clone() if full else recon[:, ::16, ::4, ::4].clone()
        if full:
            rec["grads"] = grads
            rec["params_after"] = {k: v.clone() for k, v in model.state_dict().items()}
        elif s == 0:
            small = {k: v for k, v in grads.items() if v is not None and v.numel() <= 2048}
            rec["grads_small"] = small
            rec["grads_sub"] = {k: v.reshape(-1)[::997].clone() for k, v in grads.items()
                                if v is not None and v.numel() > 2048}
        fx["steps"].append(rec)
        print(tag, "step", s, {k: v for k, v in rec.items() if isinstance(v, float)}, flush=True)
    torch.save(fx, os.path.join(OUT, f"{tag}.pt"))


def l2_fixture(cfg, B, tag):
    base = build_ref(cfg)
    torch.manual_seed(43)
    model = ref_l2.VAEWithL2Supervision(base.vae, latent_channels=cfg["embed_dim"], mlp_hidden=[64, 64])
    C, H, W = cfg["shape"]
    Z, hz = cfg["embed_dim"], H // 2 ** (len(cfg["chs"]) - 1)
    g = torch.Generator().manual_seed(11)
    batch = {"spectral": orc.structured_batch(B, cfg, seed=300)}
    for i, p in enumerate(("NO2", "O3TOT", "HCHO", "CLDO4")):
        t = torch.randn((B, H, W), generator=g)
        t[torch.rand((B, H, W), generator=g) < 0.03] = float("nan")
        batch[p] = t
    batch["CLDO4"][:] = float("nan")           # exercises "no valid pixel => product skipped"
    eps = torch.randn((B, Z, hz, hz), generator=g)
    eps2 = torch.randn((B, Z, hz, hz), generator=g)
    weights = {"NO2": 0.1, "O3TOT": 0.2, "HCHO": 0.3, "CLDO4": 0.4}
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    with EpsInjector([eps, eps2]):
        total, metrics = model.compute_loss(batch, l2_weights=weights)
    total.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    fx = dict(cfg=cfg, B=B, batch=batch, eps=eps, eps2=eps2, weights=weights, state_dict=sd0, total=total.item(),
              metrics=metrics, grads=grads, mlp_hidden=[64, 64])
    print(tag, metrics, flush=True)
    torch.save(fx, os.path.join(OUT, f"{tag}.pt"))


if __name__ == "__main__":
    train_fixture(orc.TINY_CFG, B=2, steps=3, tag="tiny_train", full=True)
    l2_fixture(orc.TINY_CFG, B=2, tag="tiny_l2")
    train_fixture(orc.DEFAULT_CFG, B=2, steps=2, tag="default_train_b2", full=False)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
