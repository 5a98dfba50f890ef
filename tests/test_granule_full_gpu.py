"""SURVEY.md 8f row 2 at its REAL size: the reference's whole-granule latent extraction
(src/scripts/linear_probe_analysis.py:113-140: `model.get_latent(x).mean` on ONE fully convolutional pass over the
[1, 1028, 128, 2048] crop of a granule; src/model.py:120-152: mid-block attention over all 32 x 512 = 16,384 latent
positions, GroupNorm statistics over the whole granule) on the default model.

Checker: the fp32 oracle evaluated ON THE SAME GPU with stock torch ops, TF32 off (its attention materialises the
[1, 4, 16384, 16384] score tensor, 4.3 GB -- exactly what the reference does), plus the ideal bf16-operand oracle as the
floor. Bound: rel-L2 of the latent means <= max(1e-2, 1.1 x floor). Also reports granules/s and patch-equivalents/s."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tempo_vae_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def model_and_granule(t, dev):
    """Default model (zero-initialised convolutions re-randomised: the latents would be 0 otherwise) and one normalised
    radiance-like granule [131, 2048, 1028]: smooth spatial fields x a smooth spectrum + noise, then the reference's
    normalisation (fused kernel) -> z in [-10, 10]."""
    from bench import DEFAULT_MODEL
    t.seed_all(42)
    model = t.get_model(DEFAULT_MODEL, dev)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    orc.rerandomize_zero_init(sd, seed=1234)
    model.load_state_dict(sd)
    sd = {k: v.to(dev) for k, v in sd.items()}
    g = torch.Generator(device=dev).manual_seed(11)
    rad = torch.exp(0.5 * torch.randn((131, 2048, 1028), device=dev, generator=g) + 3.0)
    coarse = torch.randn((1, 8, 9, 128), device=dev, generator=g)
    field = torch.nn.functional.interpolate(coarse, size=(131, 2048), mode="bilinear")[0]          # [8, 131, 2048]
    basis = torch.cos(torch.linspace(0, 3.14159, 1028, device=dev)[None, :] * torch.arange(1, 9, device=dev)[:, None])
    rad = rad * torch.exp(0.4 * torch.einsum("rhw,rc->hwc", field, basis))
    del coarse, field
    mean_s, std_s = t.granule_statistics(rad)
    z = t.normalize_radiance(rad, mean_s, std_s)
    del rad
    assert z.shape == (131, 2048, 1028) and float(z.abs().max()) <= 10.0
    return model, sd, z


def test_whole_granule_encode_at_full_size_vs_oracle_on_gpu(capsys):
    import tempo_vae_b200 as t
    dev = torch.device("cuda")
    cfg = orc.DEFAULT_CFG
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        model, sd, z = model_and_granule(t, dev)
        lat = t.encode_granule_whole(model, z)
        assert lat.shape == (1, 32, 32, 512)
        for _ in range(12):       # bring the clocks up: three 8 ms calls straight after an idle GPU once measured 38 ms each
            t.encode_granule_whole(model, z)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            lat2 = t.encode_granule_whole(model, z)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        assert torch.equal(lat, lat2)
        x = z[:128, :2048].permute(2, 0, 1).unsqueeze(0).contiguous()
        with torch.no_grad():
            ref_mean, _, _ = orc.encode(sd, x, cfg)
            torch.cuda.synchronize()
            e0.record()
            orc.encode(sd, x, cfg)
            e1.record(); torch.cuda.synchronize()
            ref_ms = e0.elapsed_time(e1)
            with orc.bf16_operands():
                ideal_mean, _, _ = orc.encode(sd, x, cfg)
        err, floor = rel(lat, ref_mean), rel(ideal_mean, ref_mean)
        # the patch-tiled sweep of the same crop is a DIFFERENT function (per-patch statistics and attention): it must
        # differ from the whole-granule result by far more than the numerical error
        patches = t.granule_to_patches(z)
        assert patches.shape == (64, 1028, 64, 64)
        lat_p = t.encode_patches(model, patches, batch_size=64)          # [64, 32, 16, 16]
        tiled = lat_p.reshape(2, 32, 32, 16, 16).permute(2, 0, 3, 1, 4).reshape(1, 32, 32, 512)
        diff = rel(tiled, ref_mean)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    with capsys.disabled():
        print(f"\n[whole granule 128x2048x1028, 16,384-token attention] latent rel-L2 vs fp32 oracle on the GPU {err:.3e} "
              f"(ideal-bf16 floor {floor:.3e}); engine {ms:.1f} ms per granule = {1e3 / ms:.1f} granules/s = "
              f"{64e3 / ms:.0f} patch-equivalents/s; stock torch fp32 on the same GPU {ref_ms:.0f} ms; patch-tiled result "
              f"differs from whole-granule by {diff:.2f} (a different function, by design)")
    assert err < max(1e-2, 1.1 * floor), (err, floor)
    assert diff > 10 * err


def test_whole_granule_reconstruction_at_full_size_vs_oracle_on_gpu(capsys):
    """The other whole-granule caller (src/scripts/analyze_reconstruction.py:111-127): `recon = model(x)` on the
    [1, 1028, 128, 2048] crop -- encoder, one posterior sample, decoder (a second 16,384-token attention, transposed
    convolutions up to 128 x 2048, GroupNorm groups of 16.8 M elements) -- against the fp32 oracle on the same GPU with the
    same noise, and the posterior-mode variant. Bound: rel-L2 of the reconstruction <= max(1e-2, 1.1 x the ideal
    bf16-operand floor of the same variant)."""
    import tempo_vae_b200 as t
    dev = torch.device("cuda")
    cfg = orc.DEFAULT_CFG
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        model, sd, z = model_and_granule(t, dev)
        eps = torch.randn((1, 32, 32, 512), device=dev, generator=torch.Generator(device=dev).manual_seed(5))
        rec = t.reconstruct_granule_whole(model, z, eps=eps)
        assert rec.shape == (128, 2048, 1028) and bool(torch.isfinite(rec).all())
        for _ in range(12):      # clocks up after the idle set-up phase (see the encode test)
            t.reconstruct_granule_whole(model, z, eps=eps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            rec2 = t.reconstruct_granule_whole(model, z, eps=eps)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        assert torch.equal(rec, rec2)
        mode = t.reconstruct_granule_whole(model, z, sample_posterior=False)
        x = z[:128, :2048].permute(2, 0, 1).unsqueeze(0).contiguous()

        def oracle_forward(noise):
            mean, logvar, _ = orc.encode(sd, x, cfg)
            return orc.decode(sd, mean + torch.exp(0.5 * logvar) * noise, cfg)[0].permute(1, 2, 0)

        with torch.no_grad():
            ref = oracle_forward(eps)
            ref_mode = oracle_forward(torch.zeros_like(eps))
            with orc.bf16_operands():
                ideal = oracle_forward(eps)
                ideal_mode = oracle_forward(torch.zeros_like(eps))
        # two floors: with the noise injected exactly on both sides the latent's own error is diluted by std * eps; the
        # posterior mode feeds the decoder the bare mean
        err, err_mode, floor, floor_mode = rel(rec, ref), rel(mode, ref_mode), rel(ideal, ref), rel(ideal_mode, ref_mode)
        gt_err = rel(rec, z[:128, :2048])            # random weights: the reconstruction is nowhere near the input
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    with capsys.disabled():
        print(f"\n[whole-granule reconstruction 128x2048x1028] rel-L2 vs fp32 oracle on the GPU {err:.3e} (ideal-bf16 floor "
              f"{floor:.3e}), posterior mode {err_mode:.3e} (floor {floor_mode:.3e}); engine {ms:.1f} ms per granule = {64e3 / ms:.0f} "
              f"patch-equivalents/s; |recon - input| / |input| = {gt_err:.2f} (random weights)")
    assert err < max(1e-2, 1.1 * floor) and err_mode < max(1e-2, 1.1 * floor_mode), (err, floor, err_mode, floor_mode)


def test_granule_graph_replays_the_eager_pass_bit_for_bit(capsys):
    """`GranuleGraph`: the whole-granule launch chain captured into a CUDA graph. Replays must equal the eager calls bit
    for bit (same kernels, same order), follow new inputs, new noise and NEW WEIGHTS (the pack launch is inside the
    graph), and refuse another shape. Reports the time the host's launch gaps cost."""
    import tempo_vae_b200 as t
    dev = torch.device("cuda")
    model, _, z = model_and_granule(t, dev)
    z2 = torch.roll(z, shifts=(5, 100), dims=(0, 1)).contiguous()

    def timed(fn, n=5):
        for _ in range(12):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    enc = t.GranuleGraph(model, z.shape)
    assert torch.equal(enc(z), t.encode_granule_whole(model, z))
    lat1 = enc(z).clone()                                       # the graph's output buffer is reused by the next call
    assert torch.equal(enc(z2), t.encode_granule_whole(model, z2)) and not torch.equal(lat1, enc(z2))
    rec = t.GranuleGraph(model, z.shape, reconstruct=True)
    eps = torch.randn((1, 32, 32, 512), device=dev, generator=torch.Generator(device=dev).manual_seed(8))
    assert torch.equal(rec(z, eps=eps), t.reconstruct_granule_whole(model, z, eps=eps))
    a = rec(z).clone()
    assert not torch.equal(a, rec(z))                          # fresh noise per call when none is given
    with torch.no_grad():                                      # weights changed behind the graph's back
        model.vae.encoder.conv_in.weight.data.mul_(1.5)
    assert torch.equal(enc(z), t.encode_granule_whole(model, z))
    with pytest.raises(t.TvaeError):
        enc(z[:, :1024])
    ms = {"encode eager": timed(lambda: t.encode_granule_whole(model, z)), "encode graph": timed(lambda: enc(z)),
          "reconstruct eager": timed(lambda: t.reconstruct_granule_whole(model, z, eps=eps)),
          "reconstruct graph": timed(lambda: rec(z, eps=eps))}
    with capsys.disabled():
        print("\n[granule graph] ms per granule: " + ", ".join(f"{k} {v:.2f}" for k, v in ms.items()))
