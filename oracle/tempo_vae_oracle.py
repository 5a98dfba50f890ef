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


# ------------------------------------------------------------------------------------------------ oracle utilities
class bf16_operands:
    """Context manager: every convolution of this oracle rounds its input AND weight to bf16 (fp32 accumulation);
    GroupNorm, activations, the residual stream, attention and the loss stay fp32. That is the BEST any bf16-operand
    engine can do, so the forward error it shows against plain fp32 is the floor the GPU parity tests measure the
    engine against (tools/bf16_floor.py, profiles/bf16_floor_r2.json)."""

    def __enter__(self):
        self._conv, self._convT = F.conv2d, F.conv_transpose2d
        c, ct = self._conv, self._convT

        def r(t):
            return t.to(torch.bfloat16).to(t.dtype)
        F.conv2d = lambda inp, w, b=None, **kw: c(r(inp), r(w), b, **kw)
        F.conv_transpose2d = lambda inp, w, b=None, **kw: ct(r(inp), r(w), b, **kw)
        return self

    def __exit__(self, *a):
        F.conv2d, F.conv_transpose2d = self._conv, self._convT


def init_state_dict(cfg=None, seed=42, l2_hidden=None):
    """Seed-`seed` initial weights of the reference WITHOUT importing it (or the product): plain torch.nn modules
    created in the reference's construction order (src/model.py:358-408 encoder, :502-550 decoder, :609-632
    AutoencoderKL; src/model_with_l2.py:19-40 head), so the global RNG is consumed exactly as `seed_all(seed);
    get_model(...)` consumes it. Checked bit for bit against the real reference in tests/test_oracle_cpu.py."""
    import torch.nn as nn
    cfg = cfg or DEFAULT_CFG
    torch.manual_seed(seed)
    sd = {}

    def conv(name, cin, cout, k=3, zero=False, transposed=False, bias=True):
        m = (nn.ConvTranspose2d if transposed else nn.Conv2d)(cin, cout, kernel_size=k, stride=(2 if k == 2 else 1),
                                                              padding=(1 if k == 3 else 0), bias=bias)
        sd[name + ".weight"] = torch.zeros_like(m.weight) if zero else m.weight.detach().clone()
        if bias:
            sd[name + ".bias"] = torch.zeros_like(m.bias) if zero else m.bias.detach().clone()

    def norm(name, c):
        sd[name + ".weight"], sd[name + ".bias"] = torch.ones(c), torch.zeros(c)

    def res(pre, cin, cout):
        norm(pre + ".net1.0", cin); conv(pre + ".net1.2", cin, cout)
        norm(pre + ".net2.0", cout); conv(pre + ".net2.2", cout, cout, zero=True)
        if cin != cout:
            conv(pre + ".skip_conv", cin, cout, k=1)

    def attn(pre, c):
        norm(pre + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(f"{pre}.{n}", c, c, k=1)

    C, size = cfg["shape"][0], cfg["shape"][1]
    chs, nres = cfg["chs"], cfg["num_res_blocks"]
    zc = cfg["z_channels"]
    sd["vae.logvar"] = None                                       # placeholder: keeps the reference's key order
    e = "vae.encoder"
    conv(e + ".conv_in", C, chs[0])
    cur, cin = size, chs[0]
    for i, cout in enumerate(chs):
        cin = chs[0] if i == 0 else chs[i - 1]
        for j in range(nres):
            res(f"{e}.downs.{i}.resnet_blocks.{j}", cin, cout)
            if cur in cfg["attn_sizes"]:
                attn(f"{e}.downs.{i}.attention_blocks.{j}", cout)
            cin = cout
        conv(f"{e}.downs.{i}.down", cout, cout, k=2)
        cur //= 2
    res(e + ".mid1", cin, cin)
    if cfg["mid_attn"]:
        attn(e + ".mid_attn1", cin)
    res(e + ".mid2", cin, cin)
    norm(e + ".norm_out", cin)
    conv(e + ".conv_out", cin, 2 * zc if cfg["double_z"] else zc, zero=True)
    d = "vae.decoder"
    cin = chs[-1]
    conv(d + ".conv_in", zc, cin)
    res(d + ".mid1", cin, cin)
    if cfg["mid_attn"]:
        attn(d + ".mid_attn1", cin)
    res(d + ".mid2", cin, cin)
    cur = size // 2 ** (len(chs) - 1)
    for n, i in enumerate(reversed(range(len(chs)))):
        cin = chs[i]
        for j in range(nres):
            res(f"{d}.ups.{n}.resnet_blocks.{j}", cin, cin)
            if cur in cfg["attn_sizes"]:
                attn(f"{d}.ups.{n}.attention_blocks.{j}", cin)
        cout = chs[0] if i == 0 else chs[i - 1]
        conv(f"{d}.ups.{n}.up", cin, cout, k=2, transposed=True)
        cur //= 2
    norm(d + ".norm_out", cout)
    conv(d + ".conv_out", cout, C, zero=True)
    conv("vae.quant_conv", 2 * zc, 2 * cfg["embed_dim"], k=1)
    conv("vae.post_quant_conv", cfg["embed_dim"], zc, k=1)
    sd["vae.logvar"] = torch.tensor(6.0)
    if l2_hidden is not None:
        cin = cfg["embed_dim"]
        idx = 0
        for h in l2_hidden:
            conv(f"l2_head.mlp.{idx}", cin, h, k=1, bias=False)
            norm(f"l2_head.mlp.{idx + 1}", h)
            idx += 3
            cin = h
        conv(f"l2_head.mlp.{idx}", cin, 4, k=1)
    return sd


# ------------------------------------------------------------------------------------------------ data preparation
def normalize_radiance(rad, mean_spectrum, std_spectrum, min_radiance=1.0, clip_min=-10.0, clip_max=10.0):
    """log -> z-score -> clip (src/scripts/prepare_tempo_tiles.py:69-83)."""
    log_rad = torch.log(torch.clamp(rad, min_radiance, float("inf")))
    return torch.clamp((log_rad - mean_spectrum) / (std_spectrum + 1e-8), clip_min, clip_max)


def extract_tiles(z_rad, tile_size, n_tiles, seed=None):
    """Random crop + flip(dims=[0]) + flip(dims=[1]) + rot90(k, dims=[0,1]), numpy draws in the reference's order
    (src/scripts/prepare_tempo_tiles.py:21-58). Returns (tiles [n, T, T, C], specs [n, 4])."""
    import numpy as np
    if seed is not None:
        np.random.seed(seed)
    n_mirror, n_track, _ = z_rad.shape
    tm, tt = tile_size
    if n_mirror < tm or n_track < tt:
        return None, None
    tiles, specs = [], []
    for _ in range(n_tiles):
        i = np.random.randint(0, n_mirror - tm + 1)
        j = np.random.randint(0, n_track - tt + 1)
        tile = z_rad[i:i + tm, j:j + tt].clone()
        f0 = np.random.rand() > 0.5
        if f0:
            tile = torch.flip(tile, dims=[0])
        f1 = np.random.rand() > 0.5
        if f1:
            tile = torch.flip(tile, dims=[1])
        k = np.random.randint(0, 4)
        if k > 0:
            tile = torch.rot90(tile, k, dims=[0, 1])
        tiles.append(tile)
        specs.append((i, j, int(f0) | (int(f1) << 1), k))
    return torch.stack(tiles), specs


def spectrum_statistics(rads, min_radiance=1.0):
    """Per-channel mean / population std of log(clip(rad, min_radiance)) over the stacked pixels of all granules, in
    numpy like src/scripts/compute_tempo_stats.py:58-86. Returns (mean, std) as float32 torch tensors."""
    import numpy as np
    allp = np.vstack([np.log(np.clip(r.numpy(), min_radiance, None)).reshape(-1, r.shape[-1]) for r in rads])
    return (torch.from_numpy(allp.mean(axis=0).astype(np.float32)), torch.from_numpy(allp.std(axis=0).astype(np.float32)))


# ------------------------------------------------------------------------------------------------ probes
def probe_init(input_dim=32, hidden_dims=None, seed=None):
    """nn.Linear-initialised weights of LinearProbe (hidden_dims None) / MLPProbe in the reference's construction order
    (src/scripts/linear_probe_analysis.py:212-252): [(W [out, in], b [out]), ...]."""
    import torch.nn as nn
    if seed is not None:
        torch.manual_seed(seed)
    dims = [input_dim] + list(hidden_dims or []) + [1]
    out = []
    for a, b in zip(dims[:-1], dims[1:]):
        m = nn.Linear(a, b)
        out.append((m.weight.detach().clone(), m.bias.detach().clone()))
    return out


def probe_forward(params, x, activation="relu", dropout=0.0, train=False):
    acts = {"relu": F.relu, "gelu": F.gelu, "tanh": torch.tanh}
    h = x
    for i, (w, b) in enumerate(params):
        h = F.linear(h, w, b)
        if i < len(params) - 1:
            h = acts[activation](h)
            if dropout > 0:
                h = F.dropout(h, dropout, training=train)
    return h


def train_probe(X_train, y_train, X_val, y_val, config, params):
    """The reference's probe training loop (src/scripts/linear_probe_analysis.py:255-353) on explicit weight tensors:
    torch.optim.AdamW(lr, weight_decay), nn.MSELoss, one torch.randperm per epoch, mini-batches with a partial last
    one, full-batch validation per epoch. Returns (params, train_losses, val_losses); like the reference (whose
    `best_state` is a shallow copy) the returned weights are the last epoch's."""
    leaves = [t.clone().requires_grad_(True) for wb in params for t in wb]
    pairs = [(leaves[2 * i], leaves[2 * i + 1]) for i in range(len(params))]
    opt = torch.optim.AdamW(leaves, lr=config["learning_rate"], weight_decay=config.get("weight_decay", 0.01))
    X_train, X_val = torch.as_tensor(X_train).float(), torch.as_tensor(X_val).float()
    y_train, y_val = torch.as_tensor(y_train).float().unsqueeze(1), torch.as_tensor(y_val).float().unsqueeze(1)
    bs = config.get("batch_size", 512)
    nb = (len(X_train) + bs - 1) // bs
    act, p = config.get("activation", "relu"), (config.get("dropout", 0.1) if len(params) > 1 else 0.0)
    tl, vl = [], []
    for _ in range(config["max_epochs"]):
        perm = torch.randperm(len(X_train))
        Xs, ys = X_train[perm], y_train[perm]
        tot = 0.0
        for b in range(nb):
            s, e = b * bs, min((b + 1) * bs, len(X_train))
            opt.zero_grad()
            loss = F.mse_loss(probe_forward(pairs, Xs[s:e], act, p, True), ys[s:e])
            loss.backward()
            opt.step()
            tot += loss.item() * (e - s)
        tl.append(tot / len(X_train))
        with torch.no_grad():
            vl.append(F.mse_loss(probe_forward(pairs, X_val, act, p, False), y_val).item())
    return [(w.detach(), b.detach()) for w, b in pairs], tl, vl


def r2_score(y, pred):
    """sklearn.metrics.r2_score for 1-D arrays (src/scripts/linear_probe_analysis.py:680)."""
    y, pred = torch.as_tensor(y).double().reshape(-1), torch.as_tensor(pred).double().reshape(-1)
    return float(1.0 - ((y - pred) ** 2).sum() / ((y - y.mean()) ** 2).sum())


# ------------------------------------------------------------------------------------------------ probe targets
def component_statistics(field, norm_type):
    """The statistics src/scripts/linear_probe_analysis.py:62-100 derives from the valid (non-NaN) pixels of a float32
    component field (numpy float32 arithmetic, as the reference's `np.mean / np.std / np.median` give on float32)."""
    import numpy as np
    field = np.asarray(field, dtype=np.float32)
    good = field[np.isfinite(field) | np.isinf(field)]            # everything that is not NaN
    if norm_type == "zscore":
        return {"mean": good.mean(), "std": good.std()}
    if norm_type == "minmax":
        return {"min": good.min(), "max": good.max()}
    if norm_type == "asinh":
        centre = np.median(good)
        return {"scale": 1.4826 * np.median(np.abs(good - centre)), "median": centre}
    if norm_type == "logit":
        return {"eps": 0.01}
    raise ValueError(f"Unknown normalization type: {norm_type}")


def normalize_component(field, norm_type, stats=None):
    """(normalised field, stats) of src/scripts/linear_probe_analysis.py:60-110; NaN pixels stay NaN."""
    import numpy as np
    field = np.asarray(field, dtype=np.float32)
    stats = component_statistics(field, norm_type) if stats is None else stats
    if norm_type == "zscore":
        out = (field - stats["mean"]) / (stats["std"] + 1e-8)
    elif norm_type == "minmax":
        out = (field - stats["min"]) / (stats["max"] - stats["min"] + 1e-8)
    elif norm_type == "asinh":
        out = np.arcsinh(field / (stats["scale"] + 1e-8))
    elif norm_type == "logit":
        p = stats["eps"] + (1 - 2 * stats["eps"]) * field
        with np.errstate(divide="ignore", invalid="ignore"):
            out = np.log(p / (1 - p)).astype(np.float32)          # scipy.special.logit
    else:
        raise ValueError(f"Unknown normalization type: {norm_type}")
    return out, stats


def nanmean_pool(field, pool=4):
    """[H, W] -> [H // pool, W // pool] block means over the valid pixels, NaN for an all-NaN block
    (src/scripts/linear_probe_analysis.py:183-190)."""
    import numpy as np
    field = np.asarray(field, dtype=np.float32)
    h, w = field.shape[0] // pool, field.shape[1] // pool
    blocks = field[:h * pool, :w * pool].reshape(h, pool, w, pool).transpose(0, 2, 1, 3).reshape(h, w, pool * pool)
    ok = ~np.isnan(blocks)
    total = np.where(ok, blocks, np.float32(0)).sum(axis=2, dtype=np.float32)
    count = ok.sum(axis=2)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (total / count.astype(np.float32)).astype(np.float32)
