"""Data-side kernels and loaders against the oracle / golden fixtures, through the C ABI:
tvae_gather_rows, tvae_extract_tiles, tvae_spectrum_stats_*, tvae_batch_stats, HostTileStore, DeviceTileCache.
Byte/index work is held to bit-exactness; the log / z-score arithmetic to 1e-5 absolute (fp32 logf vs torch.log)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tempo_vae_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
DEV = "cuda"


def test_gather_rows_is_bit_exact():
    from tempo_vae_b200 import ops
    g = torch.Generator().manual_seed(0)
    for shape, dtype in (((37, 4, 4, 8), torch.bfloat16), ((50, 16, 16), torch.float32), ((9, 64, 64, 1032), torch.bfloat16)):
        src = torch.randn(shape, generator=g).to(dtype).to(DEV)
        idx = torch.randint(0, shape[0], (23,), generator=g).to(DEV)
        out = torch.full((23,) + shape[1:], 7.0, dtype=dtype, device=DEV)
        ops.gather_rows(src, idx, out)
        assert torch.equal(out, src[idx])
    with pytest.raises(Exception):
        ops.gather_rows(torch.zeros((4, 3), device=DEV), torch.zeros(2, dtype=torch.int64, device=DEV),
                        torch.zeros((2, 3), device=DEV))            # 12-byte rows


def test_extract_tiles_matches_reference_golden_bit_for_bit():
    """Augmentation only (already normalised input): pure index work, so the kernel must reproduce the tiles that the
    reference's own extract_tiles produced (tests/golden/tile_prep.pt), fp32 bit for bit, and their bf16 rounding."""
    import tempo_vae_b200 as t
    fx = torch.load(os.path.join(GOLD, "tile_prep.pt"), weights_only=False)
    for c in fx["cases"]:
        tiles = t.extract_tiles(c["z"].to(DEV), (c["tile"], c["tile"]), c["n"], seed=c["seed"])
        assert tiles.shape == c["tiles"].shape and torch.equal(tiles.cpu(), c["tiles"])
        tiles_h = t.extract_tiles(c["z"], (c["tile"], c["tile"]), c["n"], seed=c["seed"])       # host input is copied over
        assert torch.equal(tiles_h.cpu(), c["tiles"])
    assert t.extract_tiles(torch.zeros(4, 4, 2, device=DEV), (8, 8), 3, seed=0) is None
    with pytest.raises(t.TvaeError):
        t.extract_tiles(torch.zeros(16, 16, 2, device=DEV), (8, 4), 3, seed=0)


def test_process_granule_fused_normalise_and_extract_vs_oracle():
    """Raw radiance -> (log, z-score, clip) -> augmented tiles in ONE kernel, against the oracle's two-step restatement of
    src/scripts/prepare_tempo_tiles.py:61-93, with global statistics and with the per-file fallback; the same call fills
    a DeviceTileCache with the bf16 operand rows."""
    import tempo_vae_b200 as t
    g = torch.Generator().manual_seed(4)
    C, T = 20, 16
    rad = torch.exp(torch.randn((40, 70, C), generator=g) * 0.7 + 2.5)
    rad[3, 5, :4] = 0.2                                                   # below min_radiance: clamped before the log
    params = dict(min_radiance=1.0, clip_min=-2.0, clip_max=2.5, tile_size=[T, T], tiles_per_file=11)
    mean, std = orc.spectrum_statistics([rad])
    z_ref = orc.normalize_radiance(rad, mean, std, 1.0, -2.0, 2.5)
    ref_tiles, _ = orc.extract_tiles(z_ref, (T, T), 11, seed=8)
    cache = t.DeviceTileCache(torch.device(DEV), T, T, C, 32)
    tiles = t.process_granule(rad, params, mean, std, seed=8, cache=cache)
    assert tiles.shape == ref_tiles.shape and float((tiles.cpu() - ref_tiles).abs().max()) < 1e-5
    assert len(cache) == 11
    assert torch.equal(cache.data[:11, :, :, :C].cpu(), tiles.cpu().to(torch.bfloat16))
    assert float(cache.data[:11, :, :, C:].float().abs().max()) == 0.0   # pad lanes zeroed
    n = t.process_granule(rad, params, mean, std, seed=9, cache=cache, want_f32=False)
    assert n == 11 and len(cache) == 22
    # per-file fallback: the granule's own mean / UNBIASED std (torch's .std), src/scripts/prepare_tempo_tiles.py:76-80
    log_rad = torch.log(torch.clamp(rad, 1.0, float("inf")))
    m2, s2 = log_rad.mean(dim=(0, 1)), log_rad.std(dim=(0, 1))
    z2 = torch.clamp((log_rad - m2) / (s2 + 1e-8), -2.0, 2.5)
    ref2, _ = orc.extract_tiles(z2, (T, T), 11, seed=3)
    got2 = t.process_granule(rad, params, seed=3)
    assert float((got2.cpu() - ref2).abs().max()) < 2e-5


def test_spectrum_statistics_vs_numpy_oracle():
    """compute_tempo_stats.py:58-86 over several granules: mean / population std of the log-radiance per channel."""
    import tempo_vae_b200 as t
    g = torch.Generator().manual_seed(6)
    C = 1028
    rads = [torch.exp(torch.randn((13, 300, C), generator=g) * 0.6 + 3.0), torch.exp(torch.randn((7, 129, C), generator=g) + 2.0)]
    st = t.SpectrumStats(C)
    for r in rads:
        st.update(r)
    mean, std = st.finalize()
    ref_mean, ref_std = orc.spectrum_statistics(rads)
    assert st.rows == 13 * 300 + 7 * 129
    # the reference sums ~5,000 float32 rows in float32 (numpy): ITS result is ~1e-5 off the exact mean; the kernel
    # accumulates in fp64, so it is held to 1e-6 of the exact statistics and to the reference's own error of the reference
    allp = torch.cat([torch.log(torch.clamp(r, min=1.0)).reshape(-1, C) for r in rads]).double()
    assert float((mean.cpu().double() - allp.mean(0)).abs().max()) < 1e-6
    assert float((std.cpu().double() - allp.std(0, unbiased=False)).abs().max()) < 1e-6
    assert float((mean.cpu() - ref_mean).abs().max()) < 5e-5 and float((std.cpu() - ref_std).abs().max()) < 5e-5
    st2 = t.SpectrumStats(C)
    for r in rads:
        st2.update(r.to(DEV))
    m2, s2 = st2.finalize()
    assert torch.equal(m2, mean) and torch.equal(s2, std)                  # fixed-order reduction: bit-reproducible


def test_batch_stats_matches_torch():
    from tempo_vae_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn((5, 20, 16, 16), generator=g).clamp_(-3, 3).to(DEV)
    ref = torch.stack([x.min(), x.max(), x.mean(), x.std()]).cpu()
    assert torch.allclose(ops.batch_stats(x).cpu(), ref, atol=1e-5)
    # channels-last bf16 view with pad lanes (what the tile stores yield): the pad lanes must not count
    store = torch.zeros((5, 16, 16, 24), dtype=torch.bfloat16, device=DEV)
    store[..., :20] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    store[..., 20:] = 99.0
    view = store[..., :20].permute(0, 3, 1, 2)
    xb = view.float()
    ref = torch.stack([xb.min(), xb.max(), xb.mean(), xb.std()]).cpu()
    assert torch.allclose(ops.batch_stats(view).cpu(), ref, atol=1e-5)


def test_host_tile_store_batches_feed_the_engine_bit_identically(tmp_path):
    """HostTileStore: every tile exactly once per epoch, rank shards disjoint and equally long, the DMA-gathered bf16
    channels-last batch gives the same loss bit for bit as the fp32 NCHW batch of the same tiles, and the H2D byte
    count is what was declared."""
    import tempo_vae_b200 as t
    from test_model_gpu import build
    cfg = orc.TINY_CFG
    C, H, W = cfg["shape"]
    tiles = orc.structured_batch(10, cfg, seed=11).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).float()
    torch.save(tiles[:6].clone(), tmp_path / "a.pt")
    torch.save(tiles[6:].clone(), tmp_path / "b.pt")
    store = t.TEMPODataLoader.get_host_store(str(tmp_path), verbose=False)
    assert len(store) == 10 and store.data.is_pinned() and store.pitch % 8 == 0
    assert torch.equal(store.data[:10, :, :, :C].float(), tiles)
    key = lambda z: sorted(round(float(v), 3) for v in z.reshape(z.shape[0], -1).sum(1))      # noqa: E731
    seen = [x.float().permute(0, 2, 3, 1).cpu() for x in store.batches(2, DEV, seed=1, epochs=1)]
    assert len(seen) == 5 and key(torch.cat(seen)) == key(tiles)
    assert store.h2d_bytes == 10 * H * W * store.pitch * 2
    halves = [torch.cat([x.float().cpu() for x in store.batches(2, DEV, seed=2, epochs=1, rank=r, world=2)]) for r in (0, 1)]
    assert halves[0].shape[0] == halves[1].shape[0] == 4                   # 10 tiles, world 2, B 2: 2 global batches
    both = torch.cat(halves).permute(0, 2, 3, 1)
    assert len(set(key(both))) == 8
    # a batch stays valid while the next one is being copied (three rotating buffers)
    it = store.batches(4, DEV, seed=5)
    a = next(it); a_copy = a.clone(); b = next(it)
    torch.cuda.synchronize()
    assert torch.equal(a, a_copy) and a.data_ptr() != b.data_ptr()
    # same loss, bit for bit, as the fp32 NCHW batch of the same tiles
    flat = tiles.reshape(10, -1)
    got = a.float().permute(0, 2, 3, 1).reshape(4, -1).cpu()
    idx = [int(((flat - r[None]).abs().sum(1) == 0).nonzero()[0]) for r in got]
    eps = torch.randn((4, cfg["embed_dim"], H // 4, W // 4), generator=torch.Generator().manual_seed(3)).to(DEV)
    model = build(cfg, seed=7)
    la, _ = model.vae.get_loss(a, eps=eps)
    lb, _ = model.vae.get_loss(tiles[idx].permute(0, 3, 1, 2).contiguous().to(DEV), eps=eps)
    assert torch.equal(la.detach(), lb.detach())
    tr = t.Trainer(model, model.optimizer, torch.device(DEV), tmp_path / "out")
    m = tr.train_step(next(it))                                            # step 0: prints batch stats through tvae_batch_stats
    assert set(m) == {"kl_loss", "nll_loss", "loss", "pixel_mse"}
