"""ORACLE — test infrastructure only. A functional fp32 restatement of the reference's TEMPO-VAE hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module,
and only as the checker / CPU baseline. The product package (tempo_vae_b200/) never imports it.

What it restates (file:line in /root/reference). The reference's arithmetic lives in the third-party dependency
PyTorch (torch>=2.0.0, pyproject.toml:17, unpinned; this image: torch 2.11.0+cu128), so the restatement is written
against plain `torch.nn.functional` fp32 ops on a state_dict — no nn.Module, no code shared with the reference:
  conv / transposed conv ................ src/model.py:21-42, 240-247, 270-278
  GroupNorm(8, eps) + GELU(erf) ......... src/model.py:105,179,202,333-339,400,542
  ResNetBlock ........................... src/model.py:212-231
  AttnBlock (channel-interleaved heads) . src/model.py:120-152
  Encoder / Decoder ..................... src/model.py:410-431, 552-574
  posterior (chunk, clamp, sample, kl) .. src/model.py:47-75
  encode / decode / forward / get_loss .. src/model.py:634-669
  L2 head + masked-MSE loss ............. src/model_with_l2.py:11-42, 95-182
  clip_grad_norm_ + AdamW step .......... src/train_utils.py:171-177, src/model.py:756-758

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned against outputs of
the reference itself: tests/golden/*.pt are produced by oracle/make_golden.py, which imports the real reference
from /root/reference in the build container; tests/test_oracle_cpu.py checks this file against those fixtures
(and against the live reference whenever /root/reference is present).
"""
import math

import torch
import torch.nn.functional as F

DEFAULT_CFG = dict(shape=(1028, 64, 64), chs=[512, 256, 128], attn_sizes=[], mid_attn=True, num_res_blocks=1,
                   z_channels=32, double_z=True, n_attention_heads=4, norm_groups=8, norm_eps=1e-6, act="gelu",
                   embed_dim=32, kl_weight=1e-6, nll_loss_type="l1")

TINY_CFG = dict(shape=(20, 16, 16), chs=[32, 16, 16], attn_sizes=[], mid_attn=True, num_res_blocks=1,
                z_channels=4, double_z=True, n_attention_heads=4, norm_groups=8, norm_eps=1e-6, act="gelu",
                embed_dim=4, kl_weight=1e-6, nll_loss_type="l1")

# the 12 zero-initialised convs (src/model.py:205,402-408,544-550): forward parity on fresh weights is vacuous
# unless they are re-randomised


def zero_init_keys(sd):
    return [k for k in sd if (k.endswith("net2.2.weight") or k.endswith("net2.2.bias")
                              or k.endswith("coder.conv_out.weight") or k.endswith("coder.conv_out.bias"))]


def rerandomize_zero_init(sd, seed=1234, scale=1.0):
    """Replace the zero-initialised convs by seeded kaiming-uniform-like values (in place); returns sd."""
    g = torch.Generator().manual_seed(seed)
    for k in zero_init_keys(sd):
        t = sd[k]
        if k.endswith("weight"):
            fan_in = t.shape[1] * t.shape[2] * t.shape[3]
            bound = scale / math.sqrt(fan_in)
        else:
            wk = k[:-4] + "weight"
            fan_in = sd[wk].shape[1] * sd[wk].shape[2] * sd[wk].shape[3]
            bound = scale / math.sqrt(fan_in)
        t.copy_((torch.rand(t.shape, generator=g) * 2 - 1) * bound)
    return sd


def _act(x, name):
    if name == "gelu":
        return F.gelu(x)
    if name == "relu":
        return F.relu(x)
    if name == "silu":
        return F.silu(x)
    raise ValueError(name)


def _conv(sd, name, x, stride=1, padding=0):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _gn(sd, name, x, groups, eps):
    return F.group_norm(x, groups, sd.get(name + ".weight"), sd.get(name + ".bias"), eps)


def resblock(sd, pre, x, cfg):
    """x + net2(net1(x)), 1x1 skip when the channel count changes (src/model.py:212-231)."""
    g, eps, act = cfg["norm_groups"], cfg["norm_eps"], cfg["act"]
    h = _conv(sd, pre + ".net1.2", _act(_gn(sd, pre + ".net1.0", x, g, eps), act), padding=1)
    h = _conv(sd, pre + ".net2.2", _act(_gn(sd, pre + ".net2.0", h, g, eps), act), padding=1)
    if (pre + ".skip_conv.weight") in sd:
        x = _conv(sd, pre + ".skip_conv", x)
    return x + h


def attnblock(sd, pre, x, cfg):
    """softmax(q^T k / sqrt(c_)) v with heads interleaved over channels (src/model.py:120-152)."""
    nh = cfg["n_attention_heads"]
    hn = _gn(sd, pre + ".norm", x, cfg["norm_groups"], cfg["norm_eps"])
    q, k, v = _conv(sd, pre + ".q", hn), _conv(sd, pre + ".k", hn), _conv(sd, pre + ".v", hn)
    b, c, hh, ww = q.shape
    cd = c // nh
    q = q.reshape(b, cd, nh, hh * ww)
    k = k.reshape(b, cd, nh, hh * ww)
    v = v.reshape(b, cd, nh, hh * ww)
    w = torch.einsum("bcnq,bcnk->bnqk", q, k) * (cd ** -0.5)
    w = torch.softmax(w, dim=-1)
    o = torch.einsum("bnqk,bcnk->bcnq", w, v).reshape(b, c, hh, ww)
    return x + _conv(sd, pre + ".proj_out", o)


def encoder(sd, x, cfg, pre="vae.encoder"):
    """src/model.py:410-431"""
    n = len(cfg["chs"])
    h = _conv(sd, pre + ".conv_in", x, padding=1)
    for i in range(n):
        for j in range(cfg["num_res_blocks"]):
            h = resblock(sd, f"{pre}.downs.{i}.resnet_blocks.{j}", h, cfg)
            if (f"{pre}.downs.{i}.attention_blocks.{j}.norm.weight") in sd:
                h = attnblock(sd, f"{pre}.downs.{i}.attention_blocks.{j}", h, cfg)
        if i != n - 1:
            h = _conv(sd, f"{pre}.downs.{i}.down", h, stride=2)
    h = resblock(sd, pre + ".mid1", h, cfg)
    if cfg["mid_attn"]:
        h = attnblock(sd, pre + ".mid_attn1", h, cfg)
    h = resblock(sd, pre + ".mid2", h, cfg)
    h = _act(_gn(sd, pre + ".norm_out", h, cfg["norm_groups"], cfg["norm_eps"]), cfg["act"])
    return _conv(sd, pre + ".conv_out", h, padding=1)


def decoder(sd, z, cfg, pre="vae.decoder"):
    """src/model.py:552-574"""
    n = len(cfg["chs"])
    h = _conv(sd, pre + ".conv_in", z, padding=1)
    h = resblock(sd, pre + ".mid1", h, cfg)
    if cfg["mid_attn"]:
        h = attnblock(sd, pre + ".mid_attn1", h, cfg)
    h = resblock(sd, pre + ".mid2", h, cfg)
    for i in range(n):
        for j in range(cfg["num_res_blocks"]):
            h = resblock(sd, f"{pre}.ups.{i}.resnet_blocks.{j}", h, cfg)
            if (f"{pre}.ups.{i}.attention_blocks.{j}.norm.weight") in sd:
                h = attnblock(sd, f"{pre}.ups.{i}.attention_blocks.{j}", h, cfg)
        if i != n - 1:
            up = f"{pre}.ups.{i}.up"
            h = F.conv_transpose2d(h, sd[up + ".weight"], sd.get(up + ".bias"), stride=2)
    h = _act(_gn(sd, pre + ".norm_out", h, cfg["norm_groups"], cfg["norm_eps"]), cfg["act"])
    return _conv(sd, pre + ".conv_out", h, padding=1)


def encode(sd, x, cfg):
    """moments -> (mean, clamped logvar) (src/model.py:634-638, 47-59)"""
    mom = _conv(sd, "vae.quant_conv", encoder(sd, x, cfg))
    mean, logvar = torch.chunk(mom, 2, dim=1)
    return mean, torch.clamp(logvar, -30.0, 20.0), mom


def decode(sd, z, cfg):
    """src/model.py:640-643"""
    return decoder(sd, _conv(sd, "vae.post_quant_conv", z), cfg)


def kl_per_sample(mean, logvar):
    """src/model.py:67-75"""
    return 0.5 * torch.sum(mean ** 2 + torch.exp(logvar) - 1.0 - logvar, dim=[1, 2, 3])


def vae_loss(sd, x, eps, cfg):
    """AutoencoderKL.get_loss with the noise supplied (src/model.py:645-669). Returns a dict of tensors."""
    mean, logvar, mom = encode(sd, x, cfg)
    z = mean + torch.exp(0.5 * logvar) * eps
    recon = decode(sd, z, cfg)
    rec = (x - recon).abs() if cfg["nll_loss_type"] == "l1" else (x - recon) ** 2
    lv = sd["vae.logvar"]
    nll = torch.sum(rec / torch.exp(lv) + lv) / x.shape[0]
    kl = cfg["kl_weight"] * torch.sum(kl_per_sample(mean, logvar)) / x.shape[0]
    return dict(loss=nll + kl, nll_loss=nll, kl_loss=kl, recon=recon, mean=mean, logvar=logvar, moments=mom, z=z,
                pixel_mse=torch.mean((x - recon) ** 2))


def l2_head(sd, z, pre="l2_head.mlp"):
    """Conv1x1 -> GN(8, eps 1e-5) -> GELU, twice, then Conv1x1 to 4 products (src/model_with_l2.py:11-42)."""
    idx = sorted({int(k.split(".")[2]) for k in sd if k.startswith(pre + ".")})
    h = z
    last = idx[-1]
    for i in idx:
        name = f"{pre}.{i}"
        w = sd[name + ".weight"]
        if w.dim() == 4:
            h = F.conv2d(h, w, sd.get(name + ".bias"))
            if i == last:
                break
        else:
            h = F.gelu(F.group_norm(h, 8, w, sd[name + ".bias"], 1e-5))
    return h


def l2_supervised_loss(sd, batch, eps, eps2, cfg, l2_weights, products=("NO2", "O3TOT", "HCHO", "CLDO4")):
    """VAEWithL2Supervision.compute_loss with both noise draws supplied (src/model_with_l2.py:95-182)."""
    x = batch["spectral"]
    out = vae_loss(sd, x, eps, cfg)
    z2 = out["mean"] + torch.exp(0.5 * out["logvar"]) * eps2
    pred = l2_head(sd, z2)
    total = out["loss"]
    per = {}
    for i, p in enumerate(products):
        if p not in batch:
            continue
        tgt = F.avg_pool2d(batch[p].unsqueeze(1), 4)
        m = ~torch.isnan(tgt)
        if m.sum() > 0:
            l = F.mse_loss(pred[:, i:i + 1][m], tgt[m])
            per[p] = l
            total = total + l2_weights[p] * l
    out.update(total=total, l2_losses=per, l2_pred=pred, z2=z2)
    return out


def clip_and_adamw(params, grads, state, step, lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05, max_norm=1.0):
    """clip_grad_norm_(max_norm) followed by one decoupled-weight-decay Adam step on every tensor that has a
    gradient (src/train_utils.py:175-177; torch.optim.AdamW single-tensor algorithm). In place on params/state."""
    live = [k for k in params if grads.get(k) is not None]
    total = torch.sqrt(sum((grads[k].double() ** 2).sum() for k in live)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    for k in live:
        g = grads[k] * coef
        if k not in state:
            state[k] = dict(m=torch.zeros_like(params[k]), v=torch.zeros_like(params[k]))
        m, v = state[k]["m"], state[k]["v"]
        params[k].mul_(1 - lr * weight_decay)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = v.sqrt() / math.sqrt(1 - b2 ** step) + eps
        params[k].addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
    return total


def grads_of(loss_fn, sd):
    """Gradients of loss_fn(sd_with_grad) wrt every floating tensor of sd -> {key: grad or None}."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    out = loss_fn(leaves)
    loss = out["total"] if "total" in out else out["loss"]
    gs = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    return {k: g for k, g in zip(leaves, gs)}, out


def structured_batch(B, cfg, seed=0, device="cpu"):
    """Synthetic z-scored-log-radiance-like patches: a few smooth spatial fields mixed through a
    smooth spectral basis plus white noise, clipped to [-10, 10] like src/scripts/prepare_tempo_tiles.py:69-83."""
    C, H, W = cfg["shape"]
    g = torch.Generator().manual_seed(seed)
    R = 8
    t = torch.linspace(0, 1, C)
    basis = torch.stack([torch.cos(math.pi * r * t + 0.3 * r) for r in range(R)], 1)        # [C, R]
    basis = basis / basis.norm(dim=1, keepdim=True)
    f = torch.randn((B, R, H, W), generator=g)
    k = torch.tensor([1., 4., 6., 4., 1.]); k = (k[:, None] * k[None, :]); k = (k / k.sum())[None, None]
    for _ in range(3):
        f = F.conv2d(F.pad(f.reshape(B * R, 1, H, W), (2, 2, 2, 2), mode="reflect"), k).reshape(B, R, H, W)
    f = f / f.std()
    x = torch.einsum("cr,brhw->bchw", basis, f) + 0.05 * torch.randn((B, C, H, W), generator=g)
    return torch.clamp(x, -10, 10).to(device)
