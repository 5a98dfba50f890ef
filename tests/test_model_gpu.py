"""End-to-end parity of the CUDA path (through the reference-shaped Python API over the C ABI) against
 (a) golden fixtures produced by the real reference (tests/golden, oracle/make_golden.py) and
 (b) the fp32 oracle (oracle/tempo_vae_oracle.py) on the same seeded inputs.

Stated tolerances (bf16 tensor-core operands, fp32 accumulation / statistics / residual stream):
  forward mean, logvar, reconstruction ... relative L2 error <= 1e-2 on the default model (north star: "rel 1e-2 in
                                           bf16"). On the tiny fixture model the bound is max(1e-2, 1.1 x FLOOR), where
                                           FLOOR is the error of an IDEAL bf16-operand engine on the same fixture,
                                           computed live by the oracle under `orc.bf16_operands()` (every conv input and
                                           weight rounded to bf16, everything else fp32): 1.32e-2 / 9.7e-3 / 1.04e-2 for
                                           mean / logvar / recon -- no bf16-operand engine can meet 1e-2 there (2-4-channel
                                           GroupNorm groups, fully re-randomised residual branches), and the reference's
                                           own autocast-bf16 path measures 1.44e-2 / 1.18e-2 / 1.48e-2 on it and
                                           1.14e-2 / 1.26e-2 / 1.37e-2 on the default-model fixture
                                           (tools/bf16_floor.py -> profiles/bf16_floor_r2.json)
  loss, nll_loss ......................... relative error   <= 1e-4   (dominated by N * logvar)
  kl_loss, pixel_mse ..................... relative error   <= 2e-2
  parameter gradients .................... see check_grads: every tensor is held to max(0.15, 3 x its own FLOOR error)
                                           and to a cosine >= 0.8 with the reference gradient; the whole vector and the
                                           median to 2 x FLOOR (FLOOR = the same ideal bf16-operand oracle, whose
                                           autograd also rounds the gradient stream to bf16 at every conv); tensors whose
                                           true gradient is numerically zero are compared at an absolute floor (1e-5 x
                                           largest grad norm). Default model (no live floor: too slow on the host): gradient
                                           norms <= 5e-2, per-tensor median rel-L2 <= 5e-2, every tensor <= 3e-1
  500-step criterion (loss within 1 %) ... tests/test_parity_500_gpu.py (default model), shortened run here
"""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tempo_vae_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def gold(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def relinf(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def params_for(cfg, lr=1e-4):
    edp = {k: cfg[k] for k in ("shape", "chs", "attn_sizes", "mid_attn", "num_res_blocks", "z_channels", "double_z",
                               "n_attention_heads", "norm_groups", "norm_eps", "act")}
    edp.update(embed_dim=cfg["embed_dim"], kl_weight=cfg["kl_weight"], nll_loss_type=cfg["nll_loss_type"])
    return dict(architecture_type="vae", architecture_params=dict(enc_dec_params=edp), optimizer_type="AdamW",
                optimizer_params=dict(lr=lr, betas=[0.9, 0.95], weight_decay=0.05))


def build(cfg, state_dict=None, seed=42):
    import tempo_vae_b200 as t
    t.seed_all(seed)
    model = t.get_model(params_for(cfg), torch.device("cuda"))
    if state_dict is not None:
        model.load_state_dict(state_dict)
    return model


def rel_errors(got, ref):
    """(whole-vector rel-L2 without logvar, {tensor: rel-L2}) of a gradient dict against the reference's."""
    num = den = 0.0
    errs = {}
    for k, g in ref.items():
        if g is None or got.get(k) is None:
            continue
        a = got[k].detach().float().cpu()
        if not k.endswith("logvar"):
            num += float((a - g).double().pow(2).sum())
            den += float(g.double().pow(2).sum())
        errs[k] = rel(a, g)
    return (num / max(den, 1e-300)) ** 0.5, errs


def bf16_floor_grads(loss_fn, sd):
    """Gradients of the IDEAL bf16-operand engine (oracle under orc.bf16_operands(): conv inputs / weights and, through
    autograd, the gradient stream rounded to bf16; everything else fp32) -- the floor check_grads measures against."""
    with orc.bf16_operands():
        grads, _ = orc.grads_of(loss_fn, sd)
    return grads


def check_grads(model, ref, tol, report, global_tol=8e-2, median_tol=1e-1, floor=None):
    """Gradient parity of a bf16 gradient stream against fp32 autograd.
      * the whole gradient vector (logvar aside: its 4e6-scale entry would hide everything else) and the median
        per-tensor rel-L2: <= 2 x the ideal-bf16 floor when `floor` (bf16_floor_grads) is given, else the fixed
        global_tol / median_tol (measured: 0.8 x .. 1.55 x the floor);
      * every tensor: rel-L2 <= max(0.15, 3 x its own floor error; 6 x below 64 elements) (with a floor) or <= tol (without), AND cosine with
        the reference >= 0.8 -- a sign, layout, tap-order or missing-term bug fails both;
      * a tensor that misses its relative bound passes only if its ABSOLUTE error is below 0.1 % of the norm of the
        whole gradient vector (0.5 % below 64 elements: cancellation-dominated sums such as a 4-element bias gradient
        over 64 pixels) and its cosine is >= 0.8; every tensor that needed this is named in the report;
      * tensors whose true gradient is numerically zero (e.g. attention k-bias) are compared at an absolute floor."""
    norms = [float(g.norm()) for k, g in ref.items() if g is not None and not k.endswith("logvar")]
    zero_floor = 1e-5 * max(norms)
    named = dict(model.named_parameters())
    got = {}
    for k, p in named.items():
        g = ref[k]
        if g is None:
            assert p.grad is None, f"{k}: reference has no gradient here"
            continue
        assert p.grad is not None, k
        got[k] = p.grad
    glob, errs = rel_errors(got, ref)
    gnorm = sum(float(g.double().pow(2).sum()) for k, g in ref.items() if g is not None and not k.endswith("logvar")) ** 0.5
    live = {}
    for k, e in errs.items():
        if float(ref[k].norm()) < zero_floor:
            assert float((got[k].detach().float().cpu() - ref[k]).norm()) < zero_floor, k
        else:
            live[k] = e
    vals = sorted(live.values())
    med = vals[len(vals) // 2]
    worst = max(live.items(), key=lambda kv: kv[1])
    line = f"gradient vector rel-L2 {glob:.3e}; per-tensor median {med:.3e}, worst {worst[1]:.3e} at {worst[0]}"
    fl_err = {}
    if floor is not None:
        fglob, fl_err = rel_errors(floor, ref)
        fvals = sorted(v for k, v in fl_err.items() if k in live)
        fmed = fvals[len(fvals) // 2]
        line += f" [ideal-bf16 floor: vector {fglob:.3e}, median {fmed:.3e}]"
        global_tol, median_tol = 2.0 * fglob, 2.0 * fmed
    bad, escaped = {}, []
    for k, e in live.items():
        a, g = got[k].detach().float().cpu(), ref[k]
        # (a floor error is ONE realisation of rounding noise: for a handful of elements -- a 4-element bias of the tiny
        # model -- the ratio of two such realisations scatters widely, hence the wider factor below 64 elements)
        bound = max(0.15, (3.0 if g.numel() >= 64 else 6.0) * fl_err[k]) if floor is not None else tol
        cos = float((a * g).sum() / (a.norm() * g.norm()).clamp_min(1e-30))
        if e < bound and cos >= 0.8:
            continue
        # absolute escape: 0.1 % of the whole gradient norm; 0.5 % for tensors of fewer than 64 elements, whose own floor
        # estimate is too noisy to lean on (post_quant_conv.bias of the tiny model: 4 elements, floor 0.05 .. 0.15
        # depending on the host CPU's summation order) -- and only with the right direction (cosine >= 0.8)
        if float((a - g).norm()) <= (1e-3 if g.numel() >= 64 else 5e-3) * gnorm and cos >= 0.8:
            escaped.append(f"{k} (rel {e:.2f}, cos {cos:.2f}, {g.numel()} elements)")
            continue
        bad[k] = (round(e, 4), round(cos, 3), round(bound, 3))
    if escaped:
        line += "; passed on absolute error only: " + ", ".join(escaped)
    report.append(line)
    assert not bad, bad
    assert glob < global_tol and med < median_tol, (glob, med, global_tol, median_tol)


def bf16_floor_forward(sd, x, eps, cfg):
    """(mean, logvar, recon) rel-L2 errors of the ideal bf16-operand oracle against the fp32 oracle on this fixture."""
    with torch.no_grad():
        ref = orc.vae_loss(sd, x, eps, cfg)
        with orc.bf16_operands():
            idl = orc.vae_loss(sd, x, eps, cfg)
    return tuple(rel(idl[k], ref[k]) for k in ("mean", "logvar", "recon"))


def test_tiny_forward_loss_grads_and_three_steps_vs_reference_golden(capsys):
    import tempo_vae_b200 as t
    fx = gold("tiny_train.pt")
    cfg = fx["cfg"]
    model = build(cfg, fx["state_dict"])
    assert list(model.state_dict().keys()) == list(fx["state_dict"].keys())
    report = []
    for i, s in enumerate(fx["steps"]):
        x, eps = fx["x"][i].cuda(), fx["eps"][i].cuda()
        if i == 0:
            with torch.no_grad():
                recon, post = model.vae(x, eps=eps)
            e = (rel(post.mean, s["mean"]), rel(post.logvar, s["logvar"]), rel(recon, s["recon"]))
            fl = bf16_floor_forward(fx["state_dict"], fx["x"][0], fx["eps"][0], cfg)
            report.append(f"forward rel-L2 mean {e[0]:.3e} logvar {e[1]:.3e} recon {e[2]:.3e} "
                          f"[ideal-bf16 floor {fl[0]:.3e} {fl[1]:.3e} {fl[2]:.3e}]; "
                          f"rel-Linf recon {relinf(recon, s['recon']):.3e}")
            for got_e, floor_e in zip(e, fl):
                assert got_e < max(1e-2, 1.1 * floor_e), (e, fl)
            z = post.mode()
            assert z.shape == post.mean.shape and recon.shape == x.shape
        loss, metrics = model.get_loss(x, eps=eps)
        model.optimizer.zero_grad()
        loss.backward()
        assert abs(loss.item() - s["loss"]) / s["loss"] < 1e-4
        assert abs(metrics["nll_loss"].item() - s["nll_loss"]) / s["nll_loss"] < 1e-4
        assert abs(metrics["kl_loss"].item() - s["kl_loss"]) / s["kl_loss"] < 2e-2
        assert abs(model.vae.last_pixel_mse().item() - s["pixel_mse"]) / s["pixel_mse"] < 2e-2
        floor = None
        if i == 0:       # the golden parameters of later steps are the reference's own trajectory; the floor is per state
            floor = bf16_floor_grads(lambda leaves: orc.vae_loss(leaves, fx["x"][0], fx["eps"][0], cfg), fx["state_dict"])
        check_grads(model, s["grads"], 3e-1, report, floor=floor)
        gn = model.optimizer.grad_norm().item()
        assert abs(gn - s["grad_norm"]) / s["grad_norm"] < 1e-3
        model.optimizer.step(max_grad_norm=1.0)
        assert abs(model.vae.logvar.item() - s["logvar_after"]) < 2e-6
        sd = model.state_dict()
        worst = max(((sd[k].cpu() - v).abs().max().item(), k) for k, v in s["params_after"].items())
        report.append(f"step {i}: loss {loss.item():.4f} (ref {s['loss']:.4f}); max |param - ref| {worst[0]:.3e} ({worst[1]})")
        # AdamW moves every weight by <= lr (1e-4) per step; sign flips of ~zero gradients bound the error by 2*lr
        assert worst[0] < 2.5e-4 * (i + 1), worst
    with capsys.disabled():
        print("\n[tiny vs reference golden] " + "\n  ".join(report))


def test_default_config_b2_vs_reference_golden(capsys):
    """Config 1 of BASELINE.json (default model, reference CPU fp32) at B=2: weights from seed 42 by our constructor."""
    fx = gold("default_train_b2.pt")
    cfg = fx["cfg"]
    model = build(cfg)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    orc.rerandomize_zero_init(sd, seed=1234)
    model.load_state_dict(sd)
    s0 = fx["steps"][0]
    x = orc.structured_batch(fx["B"], cfg, seed=fx["x_seeds"][0]).cuda()
    eps = fx["eps"][0].cuda()
    with torch.no_grad():
        recon, post = model.vae(x, eps=eps)
    e = (rel(post.mean, s0["mean"]), rel(post.logvar, s0["logvar"]), rel(recon[:, ::16, ::4, ::4], s0["recon"]))
    assert max(e) < 1e-2, e
    loss, metrics = model.get_loss(x, eps=eps)
    model.optimizer.zero_grad()
    loss.backward()
    assert abs(loss.item() - s0["loss"]) / s0["loss"] < 1e-4
    assert abs(metrics["kl_loss"].item() - s0["kl_loss"]) / s0["kl_loss"] < 2e-2
    assert abs(model.vae.last_pixel_mse().item() - s0["pixel_mse"]) / s0["pixel_mse"] < 2e-2
    # the 4 never-used tensors get no gradient
    none = [k for k, p in model.named_parameters() if p.grad is None]
    assert sorted(none) == sorted(k for k, v in s0["grad_norms"].items() if v is None) and len(none) == 4
    norms = {k: float(p.grad.norm()) for k, p in model.named_parameters() if p.grad is not None}
    floor = 1e-5 * max(v for k, v in s0["grad_norms"].items() if v is not None and not k.endswith("logvar"))
    worst = (0.0, None)
    for k, v in s0["grad_norms"].items():
        if v is None or v < floor:
            continue
        d = abs(norms[k] - v) / v
        if d > worst[0]:
            worst = (d, k)
    named = dict(model.named_parameters())
    errs = {}
    for k, g in s0["grads_small"].items():
        if float(g.norm()) > floor:
            errs[k] = rel(named[k].grad, g)
    for k, g in s0["grads_sub"].items():                   # every 997th element of the large tensors
        if float(g.norm()) > floor * 0.03:
            errs[k] = rel(named[k].grad.reshape(-1)[::997], g)
    vals = sorted(errs.values())
    med, top = vals[len(vals) // 2], max(errs.items(), key=lambda kv: kv[1])
    with capsys.disabled():
        print(f"\n[default B=2 gradients] worst grad-norm rel err {worst[0]:.3e} at {worst[1]}; per-tensor rel-L2 "
              f"median {med:.3e}, worst {top[1]:.3e} at {top[0]}")
    assert worst[0] < 5e-2, worst
    assert med < 5e-2 and top[1] < 3e-1, (med, top)
    gn = model.optimizer.grad_norm().item()
    assert abs(gn - s0["grad_norm"]) / s0["grad_norm"] < 1e-3
    model.optimizer.step(max_grad_norm=1.0)
    assert abs(model.vae.logvar.item() - s0["logvar_after"]) < 2e-6
    with capsys.disabled():
        print(f"\n[default B=2 vs reference golden] forward rel-L2 mean {e[0]:.3e} logvar {e[1]:.3e} recon {e[2]:.3e}; "
              f"loss {loss.item():.1f} (ref {s0['loss']:.1f}); worst grad-norm rel err {worst[0]:.3e} at {worst[1]}")


def test_l2_variant_vs_reference_golden(capsys):
    import tempo_vae_b200 as t
    fx = gold("tiny_l2.pt")
    cfg = fx["cfg"]
    base = build(cfg)
    model = t.VAEWithL2Supervision(base.vae, latent_channels=cfg["embed_dim"], mlp_hidden=fx["mlp_hidden"]).cuda()
    model.load_state_dict(fx["state_dict"])
    opt = t.FusedAdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05)
    batch = {k: v.cuda() for k, v in fx["batch"].items()}
    total, metrics = model.compute_loss(batch, l2_weights=fx["weights"], eps=fx["eps"].cuda(), eps2=fx["eps2"].cuda())
    opt.zero_grad()
    total.backward()
    assert abs(total.item() - fx["total"]) / fx["total"] < 1e-4
    assert set(metrics) == set(fx["metrics"])
    for k, v in fx["metrics"].items():
        tol = 1e-4 if k in ("loss", "nll_loss") else 3e-2
        assert abs(metrics[k] - v) / abs(v) < tol, (k, metrics[k], v)
    report = []
    floor = bf16_floor_grads(lambda leaves: orc.l2_supervised_loss(leaves, fx["batch"], fx["eps"], fx["eps2"], cfg,
                                                                   fx["weights"]), fx["state_dict"])
    check_grads(model, fx["grads"], 3e-1, report, floor=floor)
    out = model(batch["spectral"])
    assert out["reconstruction"].shape == batch["spectral"].shape
    assert set(out["l2_predictions"]) == {"NO2", "O3TOT", "HCHO", "CLDO4"}
    assert out["l2_predictions"]["NO2"].shape == (fx["B"], 1, 4, 4)
    with capsys.disabled():
        print("\n[L2 variant vs reference golden] " + "; ".join(report), metrics)


def test_modular_api_is_differentiable_and_matches_oracle():
    """encode / sample / decode as separate autograd nodes (the non-fused API) against the fp32 oracle."""
    cfg = dict(orc.TINY_CFG, nll_loss_type="l2")
    fx = gold("tiny_train.pt")
    model = build(cfg, fx["state_dict"])
    x = fx["x"][1].cuda()
    eps = fx["eps"][1].cuda()
    post = model.vae.encode(x)
    z = post.sample(eps)
    recon = model.vae.decode(z)
    loss = ((recon - x) ** 2).mean() + 1e-3 * post.kl().mean()
    loss.backward()

    def ref_loss(leaves):
        mean, logvar, _ = orc.encode(leaves, fx["x"][1], cfg)
        zz = mean + torch.exp(0.5 * logvar) * fx["eps"][1]
        r = orc.decode(leaves, zz, cfg)
        return dict(loss=((r - fx["x"][1]) ** 2).mean() + 1e-3 * orc.kl_per_sample(mean, logvar).mean(), recon=r)
    grads, out = orc.grads_of(ref_loss, fx["state_dict"])
    with torch.no_grad(), orc.bf16_operands():
        ideal = ref_loss(fx["state_dict"])["recon"]
    assert rel(recon, out["recon"]) < max(1e-2, 1.1 * rel(ideal, out["recon"]))
    assert abs(loss.item() - out["loss"].item()) / out["loss"].item() < 2e-2
    grads["vae.logvar"] = None
    floor = bf16_floor_grads(ref_loss, fx["state_dict"])
    floor["vae.logvar"] = None
    check_grads(model, grads, 3e-1, [], floor=floor)
    # deterministic path + latent helper
    with torch.no_grad():
        r2, p2 = model.vae(x, sample_posterior=False)
        lat = model.get_latent(x)
    assert torch.equal(lat.mean, p2.mean)
    assert r2.shape == x.shape


def test_philox_sampling_statistics_and_world_size_invariance():
    cfg = orc.TINY_CFG
    model = build(cfg, gold("tiny_train.pt")["state_dict"])
    x = orc.structured_batch(8, cfg, seed=1).cuda()
    import tempo_vae_b200 as t
    t.seed_all(5)
    with torch.no_grad():
        model.vae.get_loss(x)
        eps_full = model.vae._last["eps"].clone()
    t.seed_all(5)
    with torch.no_grad():   # two "ranks" of 4 samples each, keyed by the global sample index
        model.vae.get_loss(x[:4], sample_offset=0, global_batch=8)
        e0 = model.vae._last["eps"].clone()
    t.seed_all(5)
    with torch.no_grad():
        model.vae.get_loss(x[4:], sample_offset=4, global_batch=8)
        e1 = model.vae._last["eps"].clone()
    assert torch.equal(eps_full, torch.cat([e0, e1]))


def test_trainer_step_checkpoint_roundtrip(tmp_path):
    import tempo_vae_b200 as t
    cfg = orc.TINY_CFG
    model = build(cfg)
    tr = t.Trainer(model, model.optimizer, torch.device("cuda"), tmp_path, save_every=10, val_every=5, log_every=1)
    batches = [orc.structured_batch(4, cfg, seed=s) for s in range(6)]
    m0 = tr.train_step(batches[0])
    assert set(m0) == {"kl_loss", "nll_loss", "loss", "pixel_mse"} and all(isinstance(v, float) for v in m0.values())
    tr.step = 1
    val = tr.validate(batches[1:3], n_batches=2)
    assert set(val) == {"val_kl_loss", "val_nll_loss", "val_loss"}
    path = tr.save_checkpoint()
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"step", "model_state_dict", "optimizer_state_dict", "train_metrics", "val_metrics"}
    st = ck["optimizer_state_dict"]["state"]
    assert len(st) == len(list(model.parameters())) - 4          # the 4 grad-less tensors have no AdamW state
    assert set(next(iter(st.values()))) == {"step", "exp_avg", "exp_avg_sq"}
    # a torch.optim.AdamW over same-shaped parameters accepts this state dict (format compatibility)
    clone = [torch.nn.Parameter(p.detach().cpu().clone()) for p in model.parameters()]
    ref_opt = torch.optim.AdamW(clone, lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05)
    ref_opt.load_state_dict(ck["optimizer_state_dict"])
    # resume: same next-step result as continuing
    m1 = tr.train_step(batches[3])
    model2 = build(cfg, seed=7)
    tr2 = t.Trainer(model2, model2.optimizer, torch.device("cuda"), tmp_path / "b")
    tr2.load_checkpoint(str(path))
    assert tr2.step == 1
    t.seed_all(42); t.ENGINE.rng_offset = 12  # noise-stream position of the first trainer: 4 (step) + 8 (validate)
    m1b = tr2.train_step(batches[3])
    assert abs(m1b["loss"] - m1["loss"]) / m1["loss"] < 1e-6


def test_loss_curve_tracks_oracle_over_40_steps(capsys):
    """Shortened version of the 500-step criterion: same data, same injected noise, fp32 oracle vs CUDA path."""
    cfg = orc.TINY_CFG
    fx = gold("tiny_train.pt")
    model = build(cfg, fx["state_dict"])
    params = {k: v.clone() for k, v in fx["state_dict"].items()}
    state = {}
    g = torch.Generator().manual_seed(77)
    worst = {"loss": 0.0, "pixel_mse": 0.0, "kl_loss": 0.0}
    for step in range(1, 41):
        x = orc.structured_batch(4, cfg, seed=1000 + step)
        eps = torch.randn((4, cfg["embed_dim"], 4, 4), generator=g)
        grads, out = orc.grads_of(lambda leaves: orc.vae_loss(leaves, x, eps, cfg), params)
        orc.clip_and_adamw(params, grads, state, step=step)
        loss, metrics = model.get_loss(x.cuda(), eps=eps.cuda())
        model.optimizer.zero_grad()
        loss.backward()
        model.optimizer.step(max_grad_norm=1.0)
        worst["loss"] = max(worst["loss"], abs(loss.item() - out["loss"].item()) / out["loss"].item())
        worst["pixel_mse"] = max(worst["pixel_mse"], abs(model.vae.last_pixel_mse().item() - out["pixel_mse"].item())
                                 / out["pixel_mse"].item())
        worst["kl_loss"] = max(worst["kl_loss"], abs(metrics["kl_loss"].item() - out["kl_loss"].item())
                               / out["kl_loss"].item())
    with capsys.disabled():
        print("\n[40-step curve vs oracle] worst relative deviations:", worst)
    assert worst["loss"] < 1e-2 and worst["pixel_mse"] < 2e-2 and worst["kl_loss"] < 5e-2


def test_inference_patch_sweep_and_whole_granule_vs_oracle(capsys):
    """BASELINE config 5 call pattern (patch-batched encode -> posterior mean, sharded round-robin) and the
    reference's whole-granule call (fully convolutional, global attention / GroupNorm statistics)."""
    import tempo_vae_b200 as t
    cfg = orc.TINY_CFG
    fx = gold("tiny_train.pt")
    model = build(cfg, fx["state_dict"])
    C = cfg["shape"][0]
    g = torch.Generator().manual_seed(21)
    rad = torch.exp(torch.randn((131, 256, C), generator=g) * 0.5 + 3.0)
    mean_s, std_s = torch.full((C,), 3.0), torch.full((C,), 0.5)
    z = t.normalize_radiance(rad, mean_s, std_s)             # fused CUDA kernel; host input is copied over
    z_ref = torch.clamp((torch.log(torch.clamp(rad, 1.0, float("inf"))) - mean_s) / (std_s + 1e-8), -10.0, 10.0)
    assert z.is_cuda and float((z.cpu() - z_ref).abs().max()) < 1e-5     # src/scripts/prepare_tempo_tiles.py:67-79
    z = z.cpu()
    patches = t.granule_to_patches(z)                       # 2 x 4 patches of [C, 64, 64]
    assert patches.shape == (8, C, 64, 64)
    lat = torch.cat([t.encode_patches(model, patches, batch_size=3, rank=r, world=2).cpu() for r in range(2)])
    order = list(range(0, 8, 2)) + list(range(1, 8, 2))
    with torch.no_grad():
        ref_mean, _, _ = orc.encode(fx["state_dict"], patches[order], cfg)
    assert lat.shape == ref_mean.shape == (8, cfg["embed_dim"], 16, 16)
    e_patch = rel(lat, ref_mean)
    whole = t.encode_granule_whole(model, z)
    with torch.no_grad():
        ref_whole, _, _ = orc.encode(fx["state_dict"], z[:128, :256].permute(2, 0, 1).unsqueeze(0), cfg)
    assert whole.shape == ref_whole.shape == (1, cfg["embed_dim"], 32, 64)
    e_whole = rel(whole, ref_whole)
    with torch.no_grad(), orc.bf16_operands():            # the ideal bf16-operand engine on the same two inputs
        f_patch = rel(orc.encode(fx["state_dict"], patches[order], cfg)[0], ref_mean)
        f_whole = rel(orc.encode(fx["state_dict"], z[:128, :256].permute(2, 0, 1).unsqueeze(0), cfg)[0], ref_whole)
    with capsys.disabled():
        print(f"\n[inference] patch-sweep latent rel-L2 {e_patch:.3e} (ideal-bf16 floor {f_patch:.3e}); whole-granule "
              f"(2048-token attention) {e_whole:.3e} (floor {f_whole:.3e})")
    assert e_patch < max(1e-2, 1.1 * f_patch) and e_whole < max(1e-2, 1.1 * f_whole)


def test_fp32_mode_forward_parity_1e4(capsys):
    """North star: forward recon, mu and logvar within rel 1e-4 in fp32 mode (split-bf16 emulated fp32 GEMMs)."""
    import tempo_vae_b200 as t
    t.set_precision("fp32")
    try:
        fx = gold("tiny_train.pt")
        model = build(fx["cfg"], fx["state_dict"])
        s = fx["steps"][0]
        with torch.no_grad():
            recon, post = model.vae(fx["x"][0].cuda(), eps=fx["eps"][0].cuda())
        e_tiny = (rel(post.mean, s["mean"]), rel(post.logvar, s["logvar"]), rel(recon, s["recon"]))
        loss, metrics = model.get_loss(fx["x"][0].cuda(), eps=fx["eps"][0].cuda())
        assert abs(loss.item() - s["loss"]) / s["loss"] < 1e-5
        assert abs(metrics["kl_loss"].item() - s["kl_loss"]) / s["kl_loss"] < 1e-3
        loss.backward()                                    # backward still works (bf16 operands, hi halves)
        fxd = gold("default_train_b2.pt")
        cfg = fxd["cfg"]
        model = build(cfg)
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        orc.rerandomize_zero_init(sd, seed=1234)
        model.load_state_dict(sd)
        s0 = fxd["steps"][0]
        x = orc.structured_batch(fxd["B"], cfg, seed=fxd["x_seeds"][0]).cuda()
        with torch.no_grad():
            recon, post = model.vae(x, eps=fxd["eps"][0].cuda())
        e_def = (rel(post.mean, s0["mean"]), rel(post.logvar, s0["logvar"]), rel(recon[:, ::16, ::4, ::4], s0["recon"]))
    finally:
        t.set_precision("bf16")
    with capsys.disabled():
        print(f"\n[fp32 mode] forward rel-L2 (mean, logvar, recon): tiny {tuple(f'{v:.2e}' for v in e_tiny)}, "
              f"default B=2 {tuple(f'{v:.2e}' for v in e_def)}")
    assert max(e_tiny) < 1e-4 and max(e_def) < 1e-4, (e_tiny, e_def)


VARIANTS = {
    # every knob get_model honours (src/model.py:713-742) that changes the executed program
    "relu_l2loss": dict(act="relu", nll_loss_type="l2"),
    "silu_two_blocks": dict(act="silu", num_res_blocks=2),
    "attn_at_8": dict(attn_sizes=[8]),
    "no_mid_attn_wide": dict(mid_attn=False, chs=[64, 32, 32], shape=(12, 32, 32)),
    "no_affine": dict(norm_affine=False),
    "two_levels_odd_batch": dict(chs=[32, 32], shape=(20, 16, 16)),
    # 260 = 2 x 128 + 4 spectral channels: the 1028-channel code paths of the default model at test size -- weight
    # gradients of conv_in / conv_out as whole-tile GEMM + tvae_wgrad_skinny, the fused loss epilogue with a ragged
    # last channel chunk and pad lanes
    "tail_channels_260": dict(chs=[64, 32, 32], shape=(260, 16, 16)),
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_config_variants_loss_and_grads_vs_oracle(name, capsys):
    import tempo_vae_b200 as t
    cfg = dict(orc.TINY_CFG, norm_affine=True)
    cfg.update(VARIANTS[name])
    B = 3 if "odd_batch" in name else 2
    mp = params_for({**cfg})
    mp["architecture_params"]["enc_dec_params"]["norm_affine"] = cfg["norm_affine"]
    t.seed_all(11)
    model = t.get_model(mp, torch.device("cuda"))
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    orc.rerandomize_zero_init(sd, seed=4321)
    model.load_state_dict(sd)
    x = orc.structured_batch(B, cfg, seed=31)
    hz = cfg["shape"][1] // 2 ** (len(cfg["chs"]) - 1)
    eps = torch.randn((B, cfg["embed_dim"], hz, hz), generator=torch.Generator().manual_seed(8))
    loss, metrics = model.get_loss(x.cuda(), eps=eps.cuda())
    model.optimizer.zero_grad()
    loss.backward()
    grads, out = orc.grads_of(lambda leaves: orc.vae_loss(leaves, x, eps, cfg), sd)
    assert abs(loss.item() - out["loss"].item()) / out["loss"].item() < 1e-4
    assert abs(metrics["kl_loss"].item() - out["kl_loss"].item()) / out["kl_loss"].item() < 3e-2
    assert abs(model.vae.last_pixel_mse().item() - out["pixel_mse"].item()) / out["pixel_mse"].item() < 2e-2
    report = []
    # the floor is computed per variant: e.g. ReLU's discontinuous derivative flips whole gradient terms wherever an
    # activation straddles 0 between a bf16 and an fp32 forward, in the ideal bf16-operand oracle exactly as in the engine
    floor = bf16_floor_grads(lambda leaves: orc.vae_loss(leaves, x, eps, cfg), sd)
    check_grads(model, grads, 4e-1, report, floor=floor)
    with capsys.disabled():
        print(f"\n[variant {name}] " + report[0])


def test_unsupported_geometry_and_inputs_raise():
    import tempo_vae_b200 as t
    model = build(orc.TINY_CFG)
    with pytest.raises(t.TvaeError):
        model.vae.encode(torch.zeros(2, 20, 16, 16))                        # CPU tensor
    with pytest.raises(t.TvaeError):
        model.vae.encode(torch.zeros(2, 19, 16, 16, device="cuda"))         # wrong channel count
    with pytest.raises(t.TvaeError):
        model.vae.encode(torch.zeros(2, 20, 12, 12, device="cuda"))         # 12 is not a power of two / multiple of 128
    z = torch.zeros(1, 4, 4, 4, device="cuda")
    assert model.vae.decode(z).shape == (1, 20, 16, 16)                     # batch of one works


def test_wgrad_stream_overlap_gives_identical_gradients():
    """ENGINE.wgrad_overlap moves the weight-gradient GEMMs to a second stream; the kernels and their inputs are the
    same, so every gradient must be bit-identical to the single-stream run."""
    from tempo_vae_b200.model import ENGINE
    cfg = orc.TINY_CFG
    x = orc.structured_batch(6, cfg, seed=5).cuda()
    eps = torch.randn((6, cfg["embed_dim"], cfg["shape"][1] // 4, cfg["shape"][2] // 4),
                      generator=torch.Generator().manual_seed(3)).cuda()
    grads = []
    prev = ENGINE.wgrad_overlap, ENGINE.wgrad_overlap_max_pixels
    try:
        # the last mode overlaps only the layers of at most 400 pixels (the 8x8 and 4x4 levels at B=6): main-stream and
        # side-stream weight gradients then run concurrently and must not share a split-K workspace
        for mode, max_px in ((False, 0), (True, 0), (True, 0), (False, 400)):
            ENGINE.wgrad_overlap, ENGINE.wgrad_overlap_max_pixels = mode, max_px
            model = build(cfg, seed=7)
            sd = orc.rerandomize_zero_init({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
            model.load_state_dict(sd)
            loss, _ = model.vae.get_loss(x, eps=eps)
            loss.backward()
            torch.cuda.synchronize()
            grads.append({k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})
    finally:
        ENGINE.wgrad_overlap, ENGINE.wgrad_overlap_max_pixels = prev
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 20
    for k in grads[0]:
        assert torch.equal(grads[0][k], grads[1][k]), k
        assert torch.equal(grads[1][k], grads[2][k]), k
        assert torch.equal(grads[2][k], grads[3][k]), k


def test_wide_wgrad_split_and_attention_switches_leave_the_step_unchanged():
    """260 = 2 x 128 + 4 channels (the default model's 1028 at test size). ENGINE.split_wide_wgrad computes the weight
    gradients of conv_in / conv_out as a whole-tile tcgen05 GEMM + tvae_wgrad_skinny instead of one padded GEMM: those two
    tensors agree to fp32 summation accuracy, every other gradient is bit-identical. tvae_attn_set_tcgen05 swaps the
    tcgen05 kind::tf32 attention for the mma.sync kernels: same TF32 operands, another summation order -- gradients
    agree to 2e-3 of each tensor's norm (they pass through bf16 roundings downstream)."""
    from tempo_vae_b200.model import ENGINE
    from tempo_vae_b200._lib import lib
    cfg = dict(orc.TINY_CFG, chs=[64, 32, 32], shape=(260, 16, 16))
    x = orc.structured_batch(4, cfg, seed=9).cuda()
    eps = torch.randn((4, cfg["embed_dim"], 4, 4), generator=torch.Generator().manual_seed(6)).cuda()

    def run():
        model = build(cfg, seed=7)
        sd = orc.rerandomize_zero_init({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
        model.load_state_dict(sd)
        loss, _ = model.vae.get_loss(x, eps=eps)
        loss.backward()
        torch.cuda.synchronize()
        return float(loss.detach()), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    prev = ENGINE.split_wide_wgrad
    try:
        ENGINE.split_wide_wgrad = True
        l_split, g_split = run()
        ENGINE.split_wide_wgrad = False
        l_one, g_one = run()
        ENGINE.split_wide_wgrad = True
        lib.tvae_attn_set_tcgen05(0)
        l_mma, g_mma = run()
    finally:
        ENGINE.split_wide_wgrad = prev
        lib.tvae_attn_set_tcgen05(1)
    assert l_split == l_one
    wide = ("vae.encoder.conv_in.weight", "vae.decoder.conv_out.weight")
    for k in g_split:
        if k in wide:
            assert rel(g_split[k], g_one[k]) < 1e-5, (k, rel(g_split[k], g_one[k]))
        else:
            assert torch.equal(g_split[k], g_one[k]), k
    assert abs(l_mma - l_split) <= 1e-5 * abs(l_split)
    worst = max(rel(g_mma[k], g_split[k]) for k in g_split)
    assert worst < 2e-2, worst           # per-tensor; typically 1e-3


@pytest.mark.parametrize("loss_type", ["l1", "l2"])
def test_loss_fused_into_conv_out_gives_identical_gradients(loss_type):
    """ENGINE.fuse_nll: decoder.conv_out's epilogue forms the reconstruction loss and writes its gradient instead of the
    fp32 reconstruction. The gradient it writes is bit-identical to what conv -> tvae_nll_fwd produces, so every
    parameter gradient must be bit-identical too -- except conv_out's own bias gradient, which the two-kernel path sums
    from the fp32 gradient values and the fused path from their bf16 roundings (with the common rounding factor of the
    l1 gradient divided out): equal to fp32 summation accuracy. The loss scalars differ by the summation order."""
    from tempo_vae_b200.model import ENGINE
    cfg = dict(orc.TINY_CFG, nll_loss_type=loss_type)
    x = orc.structured_batch(6, cfg, seed=8).cuda()
    eps = torch.randn((6, cfg["embed_dim"], cfg["shape"][1] // 4, cfg["shape"][2] // 4),
                      generator=torch.Generator().manual_seed(4)).cuda()
    out = []
    prev = ENGINE.fuse_nll
    try:
        for mode in (False, True):
            ENGINE.fuse_nll = mode
            model = build(cfg, seed=7)
            sd = orc.rerandomize_zero_init({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
            model.load_state_dict(sd)
            loss, metrics = model.vae.get_loss(x, eps=eps)
            loss.backward()
            torch.cuda.synchronize()
            out.append((float(loss.detach()), {k: float(v.detach() if torch.is_tensor(v) else v) for k, v in metrics.items()},
                        {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}))
    finally:
        ENGINE.fuse_nll = prev
    (l0, m0, g0), (l1, m1, g1) = out
    assert abs(l0 - l1) <= 1e-6 * abs(l0)
    for k in m0:
        assert abs(m0[k] - m1[k]) <= 1e-5 * max(abs(m0[k]), 1e-12), k
    assert g0.keys() == g1.keys() and len(g0) > 20
    for k in g0:
        if k.endswith("decoder.conv_out.bias"):
            tol = 1e-5 if loss_type == "l1" else 2e-3      # l2: independent bf16 roundings of 6 x 256 values per channel
            assert rel(g1[k], g0[k]) < tol, (k, rel(g1[k], g0[k]))
        else:
            assert torch.equal(g0[k], g1[k]), k


def test_channels_last_inputs_are_consumed_without_the_nchw_detour(tmp_path):
    """"accept channels_last strides without copying". The same values fed as (a) NCHW fp32, (b) fp32
    with channels-last strides, (c) a permuted view of [N, H, W, C] tiles and (d) the bf16 channels-last view that
    DeviceTileCache yields give bit-identical losses and gradients."""
    import tempo_vae_b200 as t
    cfg = orc.TINY_CFG
    C, H, W = cfg["shape"]
    tiles = orc.structured_batch(8, cfg, seed=11).permute(0, 2, 3, 1).contiguous()          # [8, H, W, C] as on disk
    tiles = tiles.to(torch.bfloat16).float()                                                 # bf16-representable
    torch.save(tiles[:5].clone(), tmp_path / "a.pt")
    torch.save(tiles[5:].clone(), tmp_path / "b.pt")
    cache = t.DeviceTileCache.from_dir(str(tmp_path), torch.device("cuda"))
    assert len(cache) == 8 and cache.pitch % 8 == 0
    assert torch.equal(cache.data[:8, :, :, :C].float().cpu(), tiles)
    # one epoch visits every tile exactly once
    seen = torch.cat([x.float().permute(0, 2, 3, 1).cpu() for x in cache.batches(4, seed=1, epochs=1)], 0)
    assert seen.shape[0] == 8
    key = lambda z: sorted(round(float(v), 4) for v in z.reshape(z.shape[0], -1).sum(1))      # noqa: E731
    assert key(seen) == key(tiles)
    # rank sharding: disjoint halves
    halves = [torch.cat([x.float().cpu() for x in cache.batches(2, seed=2, epochs=1, rank=r, world=2)], 0) for r in (0, 1)]
    assert halves[0].shape[0] == halves[1].shape[0] == 4
    assert key(torch.cat(halves, 0).permute(0, 2, 3, 1)) == key(tiles)

    eps = torch.randn((4, cfg["embed_dim"], H // 4, W // 4), generator=torch.Generator().manual_seed(3)).cuda()
    x_nchw = tiles[:4].permute(0, 3, 1, 2).contiguous().cuda()
    variants = {
        "nchw_f32": x_nchw,
        "channels_last_f32": x_nchw.contiguous(memory_format=torch.channels_last),
        "tile_view_f32": tiles[:4].cuda().permute(0, 3, 1, 2),
        "cache_bf16": cache.data[:4, :, :, :C].permute(0, 3, 1, 2),
    }
    assert variants["cache_bf16"].dtype == torch.bfloat16 and not variants["cache_bf16"].is_contiguous()
    results = {}
    for name, xin in variants.items():
        model = build(cfg, seed=7)
        sd = orc.rerandomize_zero_init({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})
        model.load_state_dict(sd)
        loss, _ = model.vae.get_loss(xin, eps=eps)
        loss.backward()
        torch.cuda.synchronize()
        results[name] = (loss.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()
                                                 if p.grad is not None})
        post = model.vae.encode(xin)
        results[name] += (post.mean.detach().clone(),)
    base = results["nchw_f32"]
    for name, r in results.items():
        assert torch.equal(r[0], base[0]), name
        assert torch.equal(r[2], base[2]), name
        for k in base[1]:
            assert torch.equal(r[1][k], base[1][k]), (name, k)


def test_full_size_step_properties():
    """BASELINE config 2 at its real size (default model, B=256 patches of [1028, 64, 64]) through properties that do not
    need an oracle run: the known initial loss of the reference (2.527e7, SURVEY.md 8a row a9: nll = X * logvar_init
    + sum|x - x_hat| / e^logvar), run-to-run determinism, and per-sample independence (the loss of a batch is the mean
    of the losses of its halves -- the property data parallelism rests on)."""
    import tempo_vae_b200 as t
    sys.path.insert(0, ROOT)
    from bench import DEFAULT_MODEL, synthetic_batch
    dev = torch.device("cuda")
    x = synthetic_batch(torch, 256, (1028, 64, 64), dev, seed=0)
    eps = torch.randn((256, 32, 16, 16), device=dev, generator=torch.Generator(device=dev).manual_seed(1))

    def run(xb, eb):
        t.seed_all(42)
        model = t.get_model(DEFAULT_MODEL, dev)
        loss, m = model.vae.get_loss(xb, eps=eb)
        loss.backward()
        gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters() if p.grad is not None))
        out = (float(loss.detach()), float(m["kl_loss"]), float(model.vae.last_pixel_mse()), float(gn))
        del model, loss
        torch.cuda.empty_cache()
        return out

    a = run(x, eps)
    b = run(x, eps)
    assert a == b                                                   # bit-reproducible (fixed-order reductions everywhere)
    X = 1028 * 64 * 64
    expected = X * 6.0 + float(x.abs().mean()) * X / math.exp(6.0)   # zero-init conv_out => x_hat = 0 at init
    assert abs(a[0] - expected) / expected < 1e-5 and abs(a[0] - 2.527e7) / 2.527e7 < 1e-3
    assert abs(a[2] - float((x ** 2).mean())) < 1e-3                 # pixel_mse of x_hat = 0
    h0, h1 = run(x[:128], eps[:128]), run(x[128:], eps[128:])
    assert abs(0.5 * (h0[0] + h1[0]) - a[0]) / a[0] < 1e-6
    assert abs(0.5 * (h0[1] + h1[1]) - a[1]) / max(a[1], 1e-12) < 1e-4
    assert math.isfinite(a[3]) and a[3] > 0


def test_evaluate_reconstruction_metrics_vs_oracle():
    """src/scripts/evaluate_reconstruction.py:23-42: per-sample MSE / MAE / PSNR of a stochastic reconstruction, against the
    same quantities computed from the oracle's reconstruction with the same injected noise."""
    import tempo_vae_b200 as t
    cfg = orc.TINY_CFG
    fx = gold("tiny_train.pt")
    model = build(cfg, fx["state_dict"])
    x = orc.structured_batch(5, cfg, seed=31)
    eps = torch.randn((5, cfg["embed_dim"], cfg["shape"][1] // 4, cfg["shape"][2] // 4),
                      generator=torch.Generator().manual_seed(9))
    got = t.evaluate_reconstruction(model, x.cuda(), eps=eps.cuda())
    with torch.no_grad():
        mean, logvar, _ = orc.encode(fx["state_dict"], x, cfg)
        z = mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * eps
        recon = orc.decode(fx["state_dict"], z, cfg)
    d = (x - recon).reshape(5, -1)
    ref_mse, ref_mae = (d ** 2).mean(1), d.abs().mean(1)
    ref_psnr = 10.0 * torch.log10(20.0 ** 2 / (ref_mse + 1e-10))
    assert got["mse"].shape == (5,) and got["mse"].is_cuda
    assert rel(got["mse"], ref_mse) < 2e-2 and rel(got["mae"], ref_mae) < 2e-2
    assert (got["psnr"].cpu() - ref_psnr).abs().max() < 0.1          # dB
    mode = t.evaluate_reconstruction(model, x.cuda(), sample_posterior=False)
    assert torch.isfinite(mode["psnr"]).all()


def test_l2_device_tile_cache_batches_match_host_batches(tmp_path):
    """DeviceTileCacheWithL2: dict batches gathered on the device give the same loss, bit for bit, as the same samples
    fed the reference way (fp32 NCHW spectral tensor + [B, H, W] targets with NaN holes)."""
    import tempo_vae_b200 as t
    fx = gold("tiny_l2.pt")
    cfg = fx["cfg"]
    C, H, W = cfg["shape"]
    n = 6
    tiles = orc.structured_batch(n, cfg, seed=41).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).float()
    gen = torch.Generator().manual_seed(43)
    split = tmp_path / "train"
    split.mkdir()
    torch.save(tiles[:4].clone(), split / "g0.pt")
    torch.save(tiles[4:].clone(), split / "g1.pt")
    targets = {}
    for p in ("NO2", "O3TOT", "HCHO", "CLDO4"):
        tg = torch.randn((n, H, W), generator=gen)
        tg[torch.rand((n, H, W), generator=gen) < 0.15] = float("nan")
        targets[p] = tg
        (split / f"l2_{p}").mkdir()
        torch.save(tg[:4].clone(), split / f"l2_{p}" / "g0.pt")
        torch.save(tg[4:].clone(), split / f"l2_{p}" / "g1.pt")
    cache = t.DeviceTileCacheWithL2.from_dir(str(tmp_path), "train", torch.device("cuda"))
    assert len(cache) == n
    batch = next(cache.batches(4, seed=3, epochs=1))
    assert batch["spectral"].dtype == torch.bfloat16 and batch["spectral"].shape == (4, C, H, W)
    # identify which tiles were drawn, rebuild the same batch the reference way
    flat = tiles.reshape(n, -1)
    got = batch["spectral"].float().permute(0, 2, 3, 1).reshape(4, -1).cpu()
    idx = [int(((flat - g[None]).abs().sum(1) == 0).nonzero()[0]) for g in got]
    ref_batch = {"spectral": tiles[idx].permute(0, 3, 1, 2).contiguous().cuda()}
    for p, tg in targets.items():
        ref_batch[p] = tg[idx].cuda()
        assert torch.equal(torch.nan_to_num(batch[p].cpu(), nan=-7.0), torch.nan_to_num(tg[idx], nan=-7.0))
    base = build(cfg)
    model = t.VAEWithL2Supervision(base.vae, latent_channels=cfg["embed_dim"], mlp_hidden=fx["mlp_hidden"]).cuda()
    model.load_state_dict(fx["state_dict"])
    eps = torch.randn((4, cfg["embed_dim"], H // 4, W // 4), generator=gen).cuda()
    eps2 = torch.randn((4, cfg["embed_dim"], H // 4, W // 4), generator=gen).cuda()
    la, ma = model.compute_loss(batch, l2_weights=fx["weights"], eps=eps, eps2=eps2)
    lb, mb = model.compute_loss(ref_batch, l2_weights=fx["weights"], eps=eps, eps2=eps2)
    assert torch.equal(la.detach(), lb.detach()) and ma.keys() == mb.keys()
    for k in ma:
        assert ma[k] == mb[k] or (ma[k] != ma[k] and mb[k] != mb[k]), k
    # the trainer consumes the cached batch without converting it
    opt = t.FusedAdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05)
    tr = t.L2SupervisedTrainer(model, opt, torch.device("cuda"), str(tmp_path / "out"), kl_weight=cfg["kl_weight"],
                               l2_weights=fx["weights"])
    m = tr.train_step(batch)
    assert all(torch.isfinite(torch.tensor(float(v))) or k.endswith("_loss") for k, v in m.items())


def test_data_writes_to_weights_are_seen_by_the_next_forward():
    """`p.data.copy_(...)` / `nn.init.*_(p.data)` do not bump the autograd version counter; the bf16 weight packs
    must still follow (they are refreshed at the start of every top-level forward), like the reference's modules, which
    read the live parameter on every call. `ENGINE.frozen_weights()` is the documented opt-out for inference sweeps."""
    import tempo_vae_b200 as t
    cfg = orc.TINY_CFG
    fx = gold("tiny_train.pt")
    model = build(cfg, fx["state_dict"])
    x = fx["x"][0].cuda()
    with torch.no_grad():
        m0 = model.vae.encode(x).mean.clone()
        w = model.vae.encoder.conv_in.weight
        ver = w._version
        w.data.mul_(1.5)                                        # invisible to autograd's version counter
        assert w._version == ver
        m1 = model.vae.encode(x).mean.clone()
        assert not torch.equal(m0, m1)
        sd = {k: v.clone() for k, v in fx["state_dict"].items()}
        sd["vae.encoder.conv_in.weight"] = sd["vae.encoder.conv_in.weight"] * 1.5
        ref_mean, _, _ = orc.encode(sd, fx["x"][0], cfg)
        assert rel(m1, ref_mean) < 2e-2
        with t.ENGINE.frozen_weights():                         # packs refreshed on entry, then assumed unchanged
            m2 = model.vae.encode(x).mean.clone()
            w.data.mul_(2.0)
            m3 = model.vae.encode(x).mean.clone()
        assert torch.equal(m1, m2) and torch.equal(m2, m3)
        m4 = model.vae.encode(x).mean
        assert not torch.equal(m3, m4)


def test_accumulated_micro_batches_equal_one_step_on_their_concatenation(tmp_path):
    """bench.py's config-4 leg (global batch 2048 as accumulated micro-batches of 256): Trainer.train_step_accumulate over
    two halves gives the gradient / parameters of one train_step on the whole batch (same noise: keyed by sample index)."""
    import tempo_vae_b200 as t
    cfg = orc.TINY_CFG
    fx = gold("tiny_train.pt")
    x = orc.structured_batch(8, cfg, seed=21).cuda()
    res = []
    for halves in (False, True):
        model = build(cfg, fx["state_dict"])
        tr = t.Trainer(model, model.optimizer, torch.device("cuda"), tmp_path / str(halves))
        tr.step = 1
        t.seed_all(3)
        if halves:
            m = tr.train_step_accumulate([x[:4], x[4:]])
        else:
            m = tr.train_step_device(x)
        res.append((model.optimizer.flat_grad.clone(), model.optimizer.flat_param.clone(), float(m["loss"].detach())))
    (g1, p1, l1), (g2, p2, l2) = res
    assert abs(l1 - l2) / abs(l1) < 1e-6
    assert float((g2 / 2 - g1).norm() / g1.norm()) < 2e-3           # accumulated sum = 2 x the mean gradient
    assert float((p2 - p1).abs().max()) < 2.5e-4


def test_default_size_l2_variant_with_shipped_head_vs_oracle_on_gpu(capsys):
    """SURVEY 8a rows a14-a16 at the SHIPPED size: default model + the [512, 512] L2 head of
    configs/training/train_vae_l2_supervised.yaml (the golden fixture uses a tiny model with a [64, 64] head). Checker: the
    fp32 oracle evaluated on the same GPU (TF32 off). compute_loss: total / nll / kl / per-product losses and every
    gradient; forward(): reconstruction and the four l2_predictions against the oracle applied to the z it returned."""
    import tempo_vae_b200 as t
    sys.path.insert(0, ROOT)
    from bench import DEFAULT_MODEL
    dev = torch.device("cuda")
    cfg = orc.DEFAULT_CFG
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        B = 4
        t.seed_all(42)
        base = t.get_model(DEFAULT_MODEL, dev)
        model = t.VAEWithL2Supervision(base.vae, latent_channels=32, mlp_hidden=[512, 512]).to(dev)
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        ref_sd = orc.init_state_dict(cfg, seed=42, l2_hidden=None)
        assert all(torch.equal(sd[k], v) for k, v in ref_sd.items())            # the VAE part is the reference's seed-42 init
        orc.rerandomize_zero_init(sd, seed=1234)
        model.load_state_dict(sd)
        opt = t.FusedAdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05)
        sdg = {k: v.to(dev) for k, v in sd.items()}
        g = torch.Generator().manual_seed(5)
        batch = {"spectral": orc.structured_batch(B, cfg, seed=77).to(dev)}
        for p in ("NO2", "O3TOT", "HCHO", "CLDO4"):
            tg = torch.randn((B, 64, 64), generator=g)
            tg[torch.rand((B, 64, 64), generator=g) < 0.15] = float("nan")
            batch[p] = tg.to(dev)
        batch["CLDO4"][:] = float("nan")                                         # a product without a single valid pixel
        eps = torch.randn((B, 32, 16, 16), generator=g).to(dev)
        eps2 = torch.randn((B, 32, 16, 16), generator=g).to(dev)
        weights = {"NO2": 0.1, "O3TOT": 0.1, "HCHO": 0.1, "CLDO4": 0.1}
        total, metrics = model.compute_loss(batch, l2_weights=weights, eps=eps, eps2=eps2)
        opt.zero_grad()
        total.backward()
        grads, out = orc.grads_of(lambda leaves: orc.l2_supervised_loss(leaves, batch, eps, eps2, cfg, weights), sdg)
        assert abs(total.item() - out["total"].item()) / out["total"].item() < 1e-4
        assert abs(metrics["kl_loss"] - out["kl_loss"].item()) / out["kl_loss"].item() < 2e-2
        assert "CLDO4_loss" not in metrics and set(metrics) == {"loss", "nll_loss", "kl_loss", "NO2_loss", "O3TOT_loss", "HCHO_loss"}
        for p in ("NO2", "O3TOT", "HCHO"):
            assert abs(metrics[f"{p}_loss"] - out["l2_losses"][p].item()) / out["l2_losses"][p].item() < 3e-2, p
        named = dict(model.named_parameters())
        gn = {k: float(v.norm()) for k, v in grads.items() if v is not None}
        floor = 1e-5 * max(v for k, v in gn.items() if not k.endswith("logvar"))
        errs = {k: rel(named[k].grad, v) for k, v in grads.items() if v is not None and gn[k] > floor}
        vals = sorted(errs.values())
        med, top = vals[len(vals) // 2], max(errs.items(), key=lambda kv: kv[1])
        head = {k: v for k, v in errs.items() if k.startswith("l2_head")}
        assert len(head) == 8 and max(head.values()) < 1e-1, head
        assert med < 5e-2 and top[1] < 3e-1, (med, top)
        # forward(): dict outputs against the oracle applied to the z the engine sampled
        with torch.no_grad():
            res = model(batch["spectral"])
            z = res["z"]
            ref_rec = orc.decode(sdg, z, cfg)
            ref_pred = orc.l2_head(sdg, z)
        assert rel(res["reconstruction"], ref_rec) < 1e-2
        for i, p in enumerate(("NO2", "O3TOT", "HCHO", "CLDO4")):
            assert res["l2_predictions"][p].shape == (B, 1, 16, 16)
            assert rel(res["l2_predictions"][p], ref_pred[:, i:i + 1]) < 1e-2, p
        assert rel(res["posterior"].mean, orc.encode(sdg, batch["spectral"], cfg)[0]) < 1e-2
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    with capsys.disabled():
        print(f"\n[default-size L2 variant, [512,512] head, B={B}] total {total.item():.1f} (oracle {out['total'].item():.1f}); "
              f"per-tensor gradient rel-L2 median {med:.3e}, worst {top[1]:.3e} at {top[0]}; L2-head tensors max {max(head.values()):.3e}")


def test_validate_is_the_sample_weighted_mean_of_get_loss(tmp_path):
    """Trainer.validate (src/train_utils.py:185-212): `val_<k>` = sum_k metric * batch_size / total samples over n_batches
    batches of get_loss under no_grad, here with unequal batch sizes and compared value by value."""
    import tempo_vae_b200 as t
    cfg = orc.TINY_CFG
    model = build(cfg, gold("tiny_train.pt")["state_dict"])
    tr = t.Trainer(model, model.optimizer, torch.device("cuda"), tmp_path)
    tr.step = 1
    batches = [orc.structured_batch(b, cfg, seed=50 + b) for b in (2, 5, 3, 4)]
    t.seed_all(7)
    val = tr.validate(batches, n_batches=3)
    t.seed_all(7)
    acc = {"kl_loss": 0.0, "nll_loss": 0.0, "loss": 0.0}
    with torch.no_grad():
        for x in batches[:3]:
            _, m = model.get_loss(x.cuda())
            for k in acc:
                acc[k] += float(m[k]) * x.shape[0]
    assert set(val) == {"val_kl_loss", "val_nll_loss", "val_loss"}
    for k, v in acc.items():
        assert abs(val[f"val_{k}"] - v / 10) / abs(v / 10) < 1e-6, k
    assert not model.training
