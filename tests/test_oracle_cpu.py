"""CPU tests (no GPU): the oracle is pinned against fixtures produced by the real reference, the C-ABI library
loads and exports every symbol include/tvae.h declares, and the host-side logic behaves."""
import math
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tempo_vae_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def gold(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def check_grads(got, ref, tol):
    """Per-tensor relative L2 error; tensors whose true gradient is round-off (e.g. the attention k-bias, whose
    gradient is exactly zero analytically) are compared on the scale of the largest non-logvar gradient."""
    norms = [float(g.norm()) for k, g in ref.items() if g is not None and not k.endswith("logvar")]
    floor = 1e-6 * max(norms)
    for k, g in ref.items():
        if g is None:
            assert got[k] is None or float(got[k].abs().max()) == 0.0, k
        elif float(g.norm()) < floor:
            assert float((got[k] - g).norm()) < floor, k
        else:
            assert rel(got[k], g) < tol, (k, rel(got[k], g))


# ------------------------------------------------------------------------------------------------- oracle pinning
def test_oracle_matches_reference_tiny_forward_and_grads():
    fx = gold("tiny_train.pt")
    cfg, sd = fx["cfg"], fx["state_dict"]
    s0 = fx["steps"][0]
    grads, out = orc.grads_of(lambda leaves: orc.vae_loss(leaves, fx["x"][0], fx["eps"][0], cfg), sd)
    assert rel(out["mean"], s0["mean"]) < 1e-5
    assert rel(out["logvar"], s0["logvar"]) < 1e-5
    assert rel(out["recon"], s0["recon"]) < 1e-5
    assert abs(out["loss"].item() - s0["loss"]) / s0["loss"] < 1e-6
    assert abs(out["kl_loss"].item() - s0["kl_loss"]) / s0["kl_loss"] < 1e-4
    assert abs(out["pixel_mse"].item() - s0["pixel_mse"]) / s0["pixel_mse"] < 1e-5
    check_grads(grads, s0["grads"], 1e-4)


def test_oracle_train_steps_match_reference():
    """3 x (get_loss, backward, clip_grad_norm_(1.0), AdamW.step) with injected noise."""
    fx = gold("tiny_train.pt")
    cfg = fx["cfg"]
    params = {k: v.clone() for k, v in fx["state_dict"].items()}
    state = {}
    for i, s in enumerate(fx["steps"]):
        grads, out = orc.grads_of(lambda leaves: orc.vae_loss(leaves, fx["x"][i], fx["eps"][i], cfg), params)
        assert abs(out["loss"].item() - s["loss"]) / s["loss"] < 1e-6
        total = orc.clip_and_adamw(params, grads, state, step=i + 1)
        assert abs(total.item() - s["grad_norm"]) / s["grad_norm"] < 1e-5
        for k, v in s["params_after"].items():
            assert torch.allclose(params[k], v, rtol=1e-5, atol=1e-7), (i, k)
        assert abs(params["vae.logvar"].item() - s["logvar_after"]) < 1e-6


def test_oracle_l2_variant_matches_reference():
    fx = gold("tiny_l2.pt")
    grads, out = orc.grads_of(
        lambda leaves: orc.l2_supervised_loss(leaves, fx["batch"], fx["eps"], fx["eps2"], fx["cfg"], fx["weights"]),
        fx["state_dict"])
    assert abs(out["total"].item() - fx["total"]) / fx["total"] < 1e-6
    for p in ("NO2", "O3TOT", "HCHO"):
        assert abs(out["l2_losses"][p].item() - fx["metrics"][f"{p}_loss"]) < 1e-5
    assert "CLDO4" not in out["l2_losses"] and "CLDO4_loss" not in fx["metrics"]
    check_grads(grads, fx["grads"], 1e-4)


def _our_default_state_dict():
    from tempo_vae_b200.model import DEFAULT_ENC_DEC, AutoencoderKL, SpectralVAE
    torch.manual_seed(42)
    vae = AutoencoderKL(dict(DEFAULT_ENC_DEC), embed_dim=32, kl_weight=1e-6, nll_loss_type="l1")
    sd = {k: v.clone() for k, v in SpectralVAE(vae).state_dict().items()}
    return orc.rerandomize_zero_init(sd, seed=1234)


def test_oracle_default_config_matches_reference_and_constructor_rng_is_identical():
    """Weights are rebuilt from seed 42 by OUR constructor, so this also pins 'same init as the reference'."""
    fx = gold("default_train_b2.pt")
    cfg = fx["cfg"]
    sd = _our_default_state_dict()
    assert sum(v.numel() for v in sd.values()) == 27_289_893
    x = orc.structured_batch(fx["B"], cfg, seed=fx["x_seeds"][0])
    s0 = fx["steps"][0]
    with torch.no_grad():
        out = orc.vae_loss(sd, x, fx["eps"][0], cfg)
    assert rel(out["mean"], s0["mean"]) < 1e-5
    assert rel(out["logvar"], s0["logvar"]) < 1e-5
    assert rel(out["recon"][:, ::16, ::4, ::4], s0["recon"]) < 1e-5
    assert abs(out["loss"].item() - s0["loss"]) / s0["loss"] < 1e-6


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference is only mounted in the build container")
def test_oracle_matches_live_reference():
    sys.path.insert(0, REF)
    import src.model as ref_model
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from make_golden import EpsInjector, build_ref
    cfg = dict(orc.TINY_CFG, chs=[32, 32, 16], nll_loss_type="l2", shape=(12, 16, 16))
    model = build_ref(cfg, seed=5)
    x = orc.structured_batch(3, cfg, seed=9)
    eps = torch.randn((3, cfg["embed_dim"], 4, 4), generator=torch.Generator().manual_seed(3))
    with EpsInjector([eps]):
        loss, metrics = model.get_loss(x)
    out = orc.vae_loss(model.state_dict(), x, eps, cfg)
    assert abs(out["loss"].item() - loss.item()) / abs(loss.item()) < 1e-6
    assert abs(out["kl_loss"].item() - metrics["kl_loss"].item()) / abs(metrics["kl_loss"].item()) < 1e-4
    assert ref_model.DiagonalGaussianDistribution is not None


# ------------------------------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    import ctypes
    from tempo_vae_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tvae.h")).read()
    declared = set(re.findall(r"\b(tvae_[a-z0-9_]+)\s*\(", header))
    declared -= {"tvae_stream_t"}
    assert len(declared) >= 30
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in include/tvae.h but not exported"
    assert set(_lib.EXPORTED) == declared
    assert _lib.lib.tvae_abi_version() == 1


def test_struct_layouts_match_header_order():
    from tempo_vae_b200._lib import ConvArgs, WgradArgs
    from tempo_vae_b200.ops import PackDesc
    header = open(os.path.join(ROOT, "include", "tvae.h")).read()
    for struct, cls in (("tvae_conv_args", ConvArgs), ("tvae_wgrad_args", WgradArgs), ("tvae_pack_desc", PackDesc)):
        body = header[:header.index("} " + struct)]
        body = body[body.rindex("typedef struct {"):]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.replace("typedef struct {", "").strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(void|float|double|int32_t|int64_t)\s*\*?", "", decl).strip()
            names += [n.strip().lstrip("*") for n in decl.split(",")]
        assert names == [f[0] for f in cls._fields_], struct


# ------------------------------------------------------------------------------------------------- host logic
def test_no_cpu_fallback():
    import tempo_vae_b200 as t
    params = dict(architecture_type="vae", architecture_params=dict(enc_dec_params={}), optimizer_type="AdamW",
                  optimizer_params=dict(lr=1e-4))
    with pytest.raises(t.TvaeError):
        t.get_model(params, torch.device("cpu"))
    from tempo_vae_b200.model import AutoencoderKL, DEFAULT_ENC_DEC
    cfg = dict(DEFAULT_ENC_DEC, shape=(20, 16, 16), chs=[32, 16, 16], z_channels=4)
    vae = AutoencoderKL(cfg, embed_dim=4)
    with pytest.raises(t.TvaeError):
        vae.get_loss(torch.zeros(2, 20, 16, 16))
    with pytest.raises(t.TvaeError):
        vae.encoder.conv_in(torch.zeros(2, 20, 16, 16))
    with pytest.raises(t.TvaeError):
        t.FusedAdamW(vae.parameters(), lr=1e-4)


def test_state_dict_keys_and_init_match_reference_layout():
    from tempo_vae_b200.model import AutoencoderKL, DEFAULT_ENC_DEC, SpectralVAE
    from tempo_vae_b200.model_with_l2 import VAEWithL2Supervision
    cfg = dict(DEFAULT_ENC_DEC, shape=(20, 16, 16), chs=[32, 16, 16], z_channels=4)
    torch.manual_seed(42)
    m = SpectralVAE(AutoencoderKL(cfg, embed_dim=4))
    fx = gold("tiny_train.pt")
    assert list(m.state_dict().keys()) == list(fx["state_dict"].keys())
    for k, v in m.state_dict().items():
        assert v.shape == fx["state_dict"][k].shape, k
    # zero-initialised convs (src/model.py:205,402-408,544-550)
    zk = orc.zero_init_keys(m.state_dict())
    assert len(zk) == 24 and all(float(m.state_dict()[k].abs().max()) == 0.0 for k in zk)
    l2 = VAEWithL2Supervision(m.vae, latent_channels=4, mlp_hidden=[64, 64])
    assert list(l2.state_dict().keys()) == list(gold("tiny_l2.pt")["state_dict"].keys())


def test_posterior_object_matches_reference_semantics():
    from tempo_vae_b200 import DiagonalGaussianDistribution
    mom = torch.randn(2, 8, 4, 4) * 20
    d = DiagonalGaussianDistribution(mom)
    mean, lv = torch.chunk(mom, 2, 1)
    lv = lv.clamp(-30, 20)
    assert torch.equal(d.mean, mean) and torch.equal(d.logvar, lv) and torch.equal(d.mode(), mean)
    assert torch.allclose(d.std, torch.exp(0.5 * lv)) and torch.allclose(d.var, torch.exp(lv))
    assert torch.allclose(d.kl(), orc.kl_per_sample(mean, lv))
    det = DiagonalGaussianDistribution(mom, deterministic=True)
    assert float(det.std.abs().max()) == 0.0 and torch.equal(det.sample(), mean)


def test_pack_geometry_strides():
    from tempo_vae_b200.ops import pack_geometry
    w = torch.arange(6 * 5 * 9, dtype=torch.float32).reshape(6, 5, 3, 3)
    for mode in ("fwd", "dgrad"):
        g = pack_geometry(tuple(w.shape), mode)
        for cr in range(g["Crow"]):
            for c in range(g["C"]):
                for t in range(9):
                    v = w.reshape(-1)[cr * g["s_row"] + c * g["s_col"] + t * g["s_tap"]]
                    exp = w[cr, c].reshape(-1)[t] if mode == "fwd" else w[c, cr].reshape(-1)[t]
                    assert v == exp
    wt = torch.arange(5 * 6 * 4, dtype=torch.float32).reshape(5, 6, 2, 2)   # ConvTranspose2d [Cin][Cout][2][2]
    g = pack_geometry(tuple(wt.shape), "up_fwd")
    assert (g["Crow"], g["TR"], g["C"]) == (6, 4, 5)
    assert wt.reshape(-1)[3 * g["s_row"] + 2 * g["s_col"] + 1 * g["s_tap"]] == wt[2, 3].reshape(-1)[1]
    g = pack_geometry(tuple(wt.shape), "up_dgrad")
    assert wt.reshape(-1)[2 * g["s_row"] + 3 * g["s_col"] + 1 * g["s_tap"]] == wt[2, 3].reshape(-1)[1]


def test_sqrt_schedule_and_random_buffer():
    import numpy as np
    from tempo_vae_b200 import RandomBuffer, get_sqrt_schedule
    s = get_sqrt_schedule(1000, 10)
    assert s[0] == 0 and s[-1] == 1000 and s == sorted(set(s))
    np.random.seed(0)
    buf = RandomBuffer()
    for i in range(50):
        buf.put(i)
    got = [buf.get() for _ in range(50)]
    assert sorted(got) == list(range(50)) and got != list(range(50)) and len(buf) == 0
    with pytest.raises(IndexError):
        buf.get()


def test_tile_loaders(tmp_path):
    from tempo_vae_b200 import TEMPODataLoader, TEMPODataLoaderWithL2
    C = 12
    d = tmp_path / "tiles" / "train"
    d.mkdir(parents=True)
    for f in range(2):
        tiles = torch.arange(64 * 4 * 4 * C, dtype=torch.float32).reshape(64, 4, 4, C) + 1000 * f
        torch.save(tiles, d / f"f{f}.pt")
        for p in ("NO2", "O3TOT", "HCHO", "CLDO4"):
            (d / f"l2_{p}").mkdir(exist_ok=True)
            torch.save(torch.full((64, 4, 4), float(f)), d / f"l2_{p}" / f"f{f}.pt")
    dl = TEMPODataLoader.get_dataloader(str(d), batch_size=8, num_workers=0, min_buffer_size=20, verbose=False)
    b = next(iter(dl))
    assert b.shape == (8, C, 4, 4) and b.dtype == torch.float32
    # sample [c, h, w] is the permuted channels-last tile
    assert torch.equal(b[0, :, 0, 0] - b[0, 0, 0, 0], torch.arange(C, dtype=torch.float32))
    with pytest.raises(ValueError):
        TEMPODataLoader.get_dataloader(str(tmp_path), batch_size=2, num_workers=0, verbose=False)
    dl2 = TEMPODataLoaderWithL2.get_dataloader(str(tmp_path / "tiles"), split="train", batch_size=4, num_workers=0,
                                               min_buffer_size=10, verbose=False)
    b2 = next(iter(dl2))
    assert b2["spectral"].shape == (4, C, 4, 4) and b2["NO2"].shape == (4, 4, 4)
    with pytest.raises(FileNotFoundError):
        TEMPODataLoaderWithL2.get_dataloader(str(tmp_path / "tiles"), split="val", verbose=False)


def test_adamw_live_ranges():
    from tempo_vae_b200.optim import FusedAdamW

    class P:
        def __init__(self, s, e):
            self._tvae_flat_range = (s, e)
    opt = FusedAdamW.__new__(FusedAdamW)
    opt._total = 160
    assert opt._live_ranges([]) == [(0, 160)]
    assert opt._live_ranges([P(16, 30)]) == [(0, 16), (32, 160)]
    assert opt._live_ranges([P(0, 16), P(144, 150)]) == [(16, 144)]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the real reference from oracle/_ref when it was built, else the oracle port, on the
    host cores) prints ONE JSON line with the contract's keys and NEVER imports the product package (its .so must not be
    mapped into the reference arm); a non-zero rank of a torchrun launch exits 0 without work."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    prog = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', "
            "'--ref-batch', '1']; runpy.run_path(%r, run_name='__main__'); "
            "assert not [m for m in sys.modules if m.startswith('tempo_vae_b200')], 'product imported'; "
            "maps = open('/proc/self/maps').read(); assert 'libtvae_b200' not in maps, 'product .so mapped'; "
            "print('NO_PRODUCT_OK', file=sys.stderr)" % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "NO_PRODUCT_OK" in r.stderr
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout                        # exactly one JSON line on stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    assert line["metric"] == "train samples/sec (fwd+bwd+AdamW)" and line["value"] > 0
    built = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "src", "model.bytecode"))
    assert line["cpu_baseline"]["kind"] == ("reference" if built else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["lean_step_samples_per_s"] > 0 and line["cpu_baseline"]["encode_samples_per_s"] > 0
    assert line["config"]["batch_per_step"] == 1            # (the driver's run uses the fixed default: 8)
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_oracle_init_state_dict_is_the_reference_constructor():
    """oracle.init_state_dict (plain torch.nn modules in the reference's construction order) gives the weights of
    `seed_all(42); get_model(...)`: bit-identical to OUR constructor's (which test_..._constructor_rng_is_identical pins to
    the reference's golden outputs), for the default model and for the L2 variant; and, when the reference was built
    into oracle/_ref, to the real reference's."""
    from tempo_vae_b200.model import DEFAULT_ENC_DEC, AutoencoderKL, SpectralVAE
    from tempo_vae_b200.model_with_l2 import VAEWithL2Supervision
    torch.manual_seed(42)
    vae = AutoencoderKL(dict(DEFAULT_ENC_DEC), embed_dim=32, kl_weight=1e-6, nll_loss_type="l1")
    ours = SpectralVAE(vae).state_dict()
    sd = orc.init_state_dict(orc.DEFAULT_CFG, seed=42)
    assert list(sd.keys()) == list(ours.keys())
    assert all(torch.equal(sd[k], ours[k]) for k in sd)
    torch.manual_seed(7)
    vae = AutoencoderKL(dict(DEFAULT_ENC_DEC, shape=(20, 16, 16), chs=[32, 16, 16], z_channels=4), embed_dim=4)
    l2 = VAEWithL2Supervision(vae, latent_channels=4, mlp_hidden=[64, 64]).state_dict()
    sd2 = orc.init_state_dict(orc.TINY_CFG, seed=7, l2_hidden=[64, 64])
    assert sorted(sd2.keys()) == sorted(l2.keys()) and all(torch.equal(sd2[k], l2[k]) for k in sd2)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    ref = build_ref.load()
    if ref is not None:
        import src.model as rm
        torch.manual_seed(42)
        from bench import DEFAULT_MODEL
        live = rm.get_model(DEFAULT_MODEL, torch.device("cpu")).state_dict()
        assert list(live.keys()) == list(sd.keys()) and all(torch.equal(live[k], sd[k]) for k in sd)


def test_oracle_tile_extraction_matches_the_reference_function():
    """oracle.extract_tiles against tests/golden/tile_prep.pt (produced by executing the reference's own extract_tiles,
    oracle/make_golden_data.py); normalisation and spectrum statistics against their closed forms."""
    fx = gold("tile_prep.pt")
    for c in fx["cases"]:
        tiles, specs = orc.extract_tiles(c["z"], (c["tile"], c["tile"]), c["n"], seed=c["seed"])
        assert torch.equal(tiles, c["tiles"]) and len(specs) == c["n"]
    assert orc.extract_tiles(torch.zeros(4, 4, 2), (8, 8), 3, seed=0) == (None, None)
    from tempo_vae_b200.tile_prep import draw_tile_specs
    c = fx["cases"][0]
    _, specs = orc.extract_tiles(c["z"], (c["tile"], c["tile"]), c["n"], seed=c["seed"])
    ours = draw_tile_specs(c["z"].shape[0], c["z"].shape[1], (c["tile"], c["tile"]), c["n"], seed=c["seed"])
    assert [tuple(int(v) for v in row) for row in ours] == [tuple(s) for s in specs]     # same np.random consumption
    g = torch.Generator().manual_seed(2)
    rads = [torch.exp(torch.randn((9, 11, 7), generator=g) * 0.5 + 3.0), torch.exp(torch.randn((5, 4, 7), generator=g))]
    mean, std = orc.spectrum_statistics(rads, min_radiance=1.0)
    allp = torch.cat([torch.log(torch.clamp(r, min=1.0)).reshape(-1, 7) for r in rads]).double()
    assert torch.allclose(mean.double(), allp.mean(0), atol=1e-6)
    assert torch.allclose(std.double(), allp.std(0, unbiased=False), atol=1e-6)
    z = orc.normalize_radiance(rads[0], mean, std)
    assert float(z.abs().max()) <= 10.0 and z.shape == rads[0].shape


def test_oracle_probe_training_matches_the_reference_run():
    """oracle.train_probe against tests/golden/probes.pt (the reference's own train_probe, oracle/make_golden_data.py): same
    seed => the same loss curves to the last bit, dropout included (identical RNG consumption)."""
    fx = gold("probes.pt")
    for name, c in fx["cases"].items():
        cfg = c["config"]
        torch.manual_seed(fx["seed"])
        params = orc.probe_init(32, cfg.get("hidden_dims"))
        trained, tl, vl = orc.train_probe(fx["X_train"], fx["y_train"], fx["X_val"], fx["y_val"], cfg, params)
        assert tl == c["train_losses"] and vl == c["val_losses"], name
        with torch.no_grad():
            pred = orc.probe_forward(trained, fx["X_val"], cfg.get("activation", "relu"), cfg.get("dropout", 0.0)).squeeze(1)
        assert torch.allclose(pred, c["pred_val"], atol=1e-6)
        assert 0.0 < orc.r2_score(fx["y_val"], pred) < 1.0


def test_oracle_probe_targets_match_the_reference_function():
    """oracle.normalize_component / nanmean_pool against tests/golden/probe_targets.pt (the reference's own
    normalize_component, src/scripts/linear_probe_analysis.py:60-110, and its pooling statement :183-190)."""
    import numpy as np
    fx = gold("probe_targets.pt")
    assert {c["norm_type"] for c in fx["cases"]} == {"zscore", "minmax", "asinh", "logit"}
    for c in fx["cases"]:
        field = c["field"].numpy()
        normalized, stats = orc.normalize_component(field, c["norm_type"])
        assert {k: float(v) for k, v in stats.items()} == c["stats"], c["name"]
        assert normalized.dtype == np.float32 == np.dtype(c["normalized_dtype"])
        want = c["normalized"].numpy()
        assert np.array_equal(np.isnan(normalized), np.isnan(want)) and np.array_equal(np.isnan(want), np.isnan(field))
        if c["norm_type"] == "logit":      # scipy's logit against log(p / (1 - p)) in float32
            assert np.allclose(normalized, want, rtol=2e-6, atol=2e-6, equal_nan=True)
        else:
            assert np.array_equal(normalized, want, equal_nan=True), c["name"]
        again, _ = orc.normalize_component(field, c["norm_type"], stats=c["stats"])
        assert np.allclose(again, want, rtol=2e-6, atol=2e-6, equal_nan=True)
        pooled = orc.nanmean_pool(want)
        assert pooled.shape == tuple(c["pooled"].shape) == (field.shape[0] // 4, field.shape[1] // 4)
        assert np.array_equal(np.isnan(pooled), np.isnan(c["pooled"].numpy())) and np.isnan(pooled).any()
        assert np.allclose(pooled, c["pooled"].numpy(), rtol=1e-6, atol=1e-7, equal_nan=True)


def test_radix_select_host_logic_with_an_emulated_histogram_pass(monkeypatch):
    """Host side of the exact median (tempo_vae_b200/probe_targets.py: order_statistics / nan_median) with
    tvae_select_hist emulated in numpy from its documented contract (include/tvae.h): prefix narrowing over four 8-bit
    passes, two ranks sharing a pass while they share a bucket, the even-count average, NaN holes, negative keys."""
    import numpy as np
    from tempo_vae_b200 import ops, probe_targets as pt
    calls = []

    def fake_select_hist(x, center, use_abs, prefix, prefix_mask, shift, hist=None):
        v = x.numpy().astype(np.float32)
        v = v[~np.isnan(v)]
        if use_abs:
            v = np.abs(v - np.float32(center))
        bits = v.view(np.uint32)
        keys = np.where(bits >> 31 == 1, ~bits, bits | np.uint32(0x80000000)).astype(np.uint32)
        keys = keys[(keys & np.uint32(prefix_mask)) == np.uint32(prefix)]
        calls.append(shift)
        return torch.from_numpy(np.bincount((keys >> np.uint32(shift)) & np.uint32(255), minlength=256).astype(np.int64))

    monkeypatch.setattr(ops, "select_hist", fake_select_hist)
    rs = np.random.RandomState(6)
    for a in (rs.standard_normal(501), rs.standard_normal(500), np.array([1.5, -2.5]), np.array([-4.0]),
              np.concatenate([np.zeros(9), rs.standard_normal(8)]), rs.standard_t(2, size=4000) * 1e15,
              rs.randint(-2, 3, size=300).astype(np.float64)):
        a = a.astype(np.float32)
        holes = a.copy()
        if a.size > 4:
            holes[rs.rand(a.size) < 0.25] = np.nan
        for v in (a, holes):
            good = v[~np.isnan(v)]
            del calls[:]
            med = pt.nan_median(torch.from_numpy(v))
            assert med.dtype == np.float32 and med == np.median(good)
            assert 4 <= len(calls) <= 7 and calls[0] == 24            # two ranks: at most one split on the way down
            mad = pt.nan_median(torch.from_numpy(v), center=float(med), use_abs=True)
            assert mad == np.median(np.abs(good - med))
    n, vals = pt.order_statistics(torch.from_numpy(a), lambda n: range(n))
    assert n == a.size and [vals[i] for i in range(n)] == sorted(a.tolist())
    assert np.isnan(pt.nan_median(torch.full((3,), float("nan"))))
    assert pt._transform("zscore", {"mean": 1.0, "std": 2.0})[:2] == (0, 1.0) and pt._transform("asinh", {"scale": 3.0})[0] == 1
    with pytest.raises(ValueError):
        pt._transform("boxcox", {})


def test_epoch_shard_gives_every_rank_the_same_number_of_batches():
    """with n % world != 0, perm[rank::world] alone can hand rank 0 one batch more than the others (n=4089,
    world=8, B=256: 2 vs 1) and the ranks would issue different numbers of all-reduces."""
    from tempo_vae_b200.tempo_data import epoch_shard
    for n, world, B in ((4089, 8, 256), (4096, 8, 256), (1000, 3, 7), (64, 2, 32), (513, 2, 256)):
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(n))
        shards = [epoch_shard(perm, B, r, world) for r in range(world)]
        counts = {s.numel() // B for s in shards}
        assert len(counts) == 1 and all(s.numel() % B == 0 for s in shards), (n, world, B)
        assert counts.pop() == n // (world * B)
        allidx = torch.cat(shards)
        assert allidx.unique().numel() == allidx.numel()                       # disjoint


def test_bf16_floor_evidence_is_reproducible():
    """profiles/bf16_floor_r2.json (tools/bf16_floor.py): on the tiny parity fixture an IDEAL bf16-operand engine is already
    above the north star's 1e-2 -- the committed evidence behind the floor-relative bound of tests/test_model_gpu.py."""
    import json
    fx = gold("tiny_train.pt")
    with torch.no_grad():
        ref = orc.vae_loss(fx["state_dict"], fx["x"][0], fx["eps"][0], fx["cfg"])
        with orc.bf16_operands():
            idl = orc.vae_loss(fx["state_dict"], fx["x"][0], fx["eps"][0], fx["cfg"])
        again = orc.vae_loss(fx["state_dict"], fx["x"][0], fx["eps"][0], fx["cfg"])
    assert torch.equal(again["recon"], ref["recon"])                            # the context manager restores F.conv2d
    e = {k: rel(idl[k], ref[k]) for k in ("mean", "logvar", "recon")}
    rec = json.load(open(os.path.join(ROOT, "profiles", "bf16_floor_r2.json")))
    for k, v in e.items():
        assert abs(v - rec["tiny_ideal"][k]) / rec["tiny_ideal"][k] < 0.05, (k, v, rec["tiny_ideal"][k])
    assert e["mean"] > 1e-2 and rec["default_b2_autocast"]["recon"] > 1e-2 > rec["default_b2_ideal"]["recon"]


def test_header_is_plain_c_and_host_geometry_helpers():
    """include/tvae.h must be consumable by a C compiler (it is the drop-in boundary: plain pointers and sizes), and
    the Python mirrors of the kernels' geometry rules agree with what the C side documents."""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "tvae.h")
    if shutil.which("gcc"):
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)      # comments may mention the reference's PyTorch
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code   # no framework types
    from tempo_vae_b200 import ops
    # vectorised GroupNorm kernels: whole 8-channel octets per group, C/8 dividing 256
    assert ops.gn_fast_ok(512, 8) and ops.gn_fast_ok(256, 8) and ops.gn_fast_ok(128, 8)
    assert not ops.gn_fast_ok(32, 8) and not ops.gn_fast_ok(1028, 4) and not ops.gn_fast_ok(128, 0)
    # fused statistics: group size multiple of 16 dividing the N tile, >= 128 output pixels per image
    assert ops.fused_stats_ok(2, 64, 64, 512, 8, 0, 64, 64) and ops.fused_stats_ok(2, 16, 16, 128, 8, 0, 16, 16)
    assert not ops.fused_stats_ok(2, 8, 8, 128, 8, 0, 8, 8)               # 64 pixels per image
    assert not ops.fused_stats_ok(2, 64, 64, 64, 8, 0, 64, 64)            # groups of 8 channels
    assert ops.fused_stats_ok(2, 32, 32, 256, 8, 2, 16, 16)               # transposed conv: the INPUT grid counts ...
    assert not ops.fused_stats_ok(2, 16, 16, 256, 8, 2, 8, 8)             # ... 64 input pixels per image: no
    assert ops.round_up(1028, 8) == 1032 and ops.round_up(1028, 64) == 1088
