"""Probe targets (data side) through the C ABI against tests/golden/probe_targets.pt -- produced by the
reference's OWN normalize_component (src/scripts/linear_probe_analysis.py:60-110) and its pooling statement (:183-190),
run by oracle/make_golden_data.py -- and against the numpy oracle on fresh random fields.

Tolerances. Order statistics (median, MAD scale), min and max: EXACT (an exact radix select on the float keys). zscore
mean / std come from fp64 sums where numpy sums pairwise in float32: within 1e-5 relative. With the statistics GIVEN the
affine transforms are the same two float32 operations per pixel: bit-exact; asinh / logit go through CUDA's asinhf / logf
instead of numpy's / scipy's: within 2e-6. Pooled means are float32 sums of <= 16 values in a different order: 2e-6."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tempo_vae_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu
FX = os.path.join(ROOT, "tests", "golden", "probe_targets.pt")
EXACT_STATS = {"median", "scale", "min", "max", "eps"}


def close(a, b, tol=2e-6):
    a, b = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a, b, rtol=tol, atol=tol, equal_nan=True)


def test_probe_targets_match_the_reference_golden():
    from tempo_vae_b200 import probe_targets as pt
    fx = torch.load(FX, weights_only=False)
    for c in fx["cases"]:
        field, nt = c["field"], c["norm_type"]
        want, want_pooled = c["normalized"].numpy(), c["pooled"].numpy()
        # statistics computed on the device
        normalized, stats = pt.normalize_component(field.numpy(), nt)
        assert normalized.is_cuda and normalized.dtype == torch.float32 and set(stats) == set(c["stats"])
        for k, v in c["stats"].items():
            if k in EXACT_STATS:
                assert float(stats[k]) == v, (c["name"], k, float(stats[k]), v)
            else:
                assert abs(float(stats[k]) - v) <= 1e-5 * abs(v), (c["name"], k, float(stats[k]), v)
        assert close(normalized.cpu().numpy(), want, 2e-5), c["name"]
        # statistics given (every file after the first): the affine transforms are bit-exact
        again, same = pt.normalize_component(field.cuda(), nt, stats=c["stats"])
        assert same is c["stats"]
        if nt in ("zscore", "minmax"):
            assert np.array_equal(again.cpu().numpy(), want, equal_nan=True), c["name"]
        else:
            assert close(again.cpu().numpy(), want), c["name"]
        # pooling alone, and normalisation + pooling in one pass
        assert close(pt.pool_component(c["normalized"]).cpu().numpy(), want_pooled), c["name"]
        pooled, _ = pt.component_targets(field, nt, stats=c["stats"])
        assert tuple(pooled.shape) == want_pooled.shape and close(pooled.cpu().numpy(), want_pooled, 4e-6), c["name"]
        assert np.isnan(want_pooled).any() and (~np.isnan(want_pooled)).any()


def test_order_statistics_are_exact():
    """np.median through the radix select: negative values, zeros, duplicates, infinities, NaN holes, odd and even
    counts, one element, and the |x - c| variant (the median absolute deviation)."""
    from tempo_vae_b200 import probe_targets as pt
    rs = np.random.RandomState(4)
    arrays = [
        rs.standard_normal(1001), rs.standard_normal(1000), np.array([3.5]), np.array([2.0, -7.0]),
        np.concatenate([np.zeros(40), -np.zeros(3), rs.standard_normal(30)]),
        np.concatenate([rs.randint(-3, 4, size=500).astype(np.float64), [np.inf, -np.inf, np.inf]]),
        rs.standard_t(2, size=20000) * 1e15, np.full(64, -1.25),
    ]
    for i, a in enumerate(arrays):
        a = a.astype(np.float32)
        holes = a.copy()
        if a.size > 4:
            holes[rs.rand(a.size) < 0.2] = np.nan
        for v in (a, holes):
            good = v[~np.isnan(v)]
            x = torch.from_numpy(v).cuda()
            med = pt.nan_median(x)
            assert med.dtype == np.float32 and med == np.median(good), (i, med, np.median(good))
            c = np.float32(0.37) if not np.isfinite(med) else med
            mad = pt.nan_median(x, center=float(c), use_abs=True)
            want = np.median(np.abs(good - c))
            assert mad == want or (np.isnan(mad) and np.isnan(want)), (i, mad, want)
    n, vals = pt.order_statistics(torch.from_numpy(arrays[0].astype(np.float32)).cuda(), lambda n: (0, n - 1, 17))
    s = np.sort(arrays[0].astype(np.float32))
    assert n == 1001 and vals[0] == s[0] and vals[1000] == s[-1] and vals[17] == s[17]
    assert np.isnan(pt.nan_median(torch.full((5,), float("nan")).cuda()))


def test_probe_targets_vs_oracle_on_random_fields_and_ragged_shapes():
    """Fresh fields, shapes that are not multiples of the pool (the reference crops first; the normalised field keeps its
    edges), a strided view (row pitch > W), pool sizes 1, 2, 4, 8, an all-NaN field."""
    from tempo_vae_b200 import probe_targets as pt
    rs = np.random.RandomState(9)
    for (H, W), nt, pool in [((37, 50), "zscore", 4), ((64, 64), "asinh", 2), ((13, 9), "minmax", 4),
                             ((128, 2048), "asinh", 4), ((40, 24), "logit", 8), ((17, 33), "zscore", 1)]:
        f = (rs.beta(0.7, 1.2, size=(H, W)) if nt == "logit" else rs.standard_t(3, size=(H, W)) * 40.0 + 5.0).astype(np.float32)
        f[rs.rand(H, W) < 0.15] = np.nan
        want, wstats = orc.normalize_component(f, nt)
        got, stats = pt.normalize_component(f, nt)
        for k, v in wstats.items():
            tol = 0.0 if k in EXACT_STATS else 1e-5 * abs(float(v))
            assert abs(float(stats[k]) - float(v)) <= tol, (nt, k, float(stats[k]), float(v))
        assert close(got.cpu().numpy(), want, 2e-5), (H, W, nt)
        pooled, _ = pt.component_targets(f, nt, stats=wstats, pool=pool)
        assert close(pooled.cpu().numpy(), orc.nanmean_pool(want, pool), 4e-6), (H, W, nt, pool)
        # a strided view of a wider buffer, cropped like process_file crops the field to the L1 crop
        wide = torch.full((H + 3, W + 5), 7.0).cuda()
        wide[:H, :W] = torch.from_numpy(f).cuda()
        hc, wc = (H // pool) * pool, (W // pool) * pool
        pooled2, stats2 = pt.component_targets(wide, nt, stats=None, pool=pool, crop=(hc, wc))
        want2, wstats2 = orc.normalize_component(f[:hc, :wc], nt)
        assert all(abs(float(stats2[k]) - float(v)) <= 1e-5 * abs(float(v)) for k, v in wstats2.items())
        assert close(pooled2.cpu().numpy(), orc.nanmean_pool(want2, pool), 2e-5), (H, W, nt, pool)
    nothing = np.full((8, 8), np.nan, dtype=np.float32)
    for nt in ("zscore", "minmax", "asinh"):
        out, stats = pt.normalize_component(nothing, nt)
        assert all(np.isnan(float(v)) for v in stats.values()) and bool(torch.isnan(out).all())


def test_sample_probe_pairs_draws_what_the_reference_draws():
    """main() of the reference (src/scripts/linear_probe_analysis.py:455-486): valid pooled pixels, np.random.choice without
    replacement, latent rows at the same flat positions."""
    from tempo_vae_b200 import probe_targets as pt
    g = torch.Generator().manual_seed(2)
    latent = torch.randn((1, 32, 8, 24), generator=g)
    pooled = torch.randn((8, 24), generator=g)
    pooled[torch.rand((8, 24), generator=g) < 0.3] = float("nan")
    np.random.seed(11)
    X, y = pt.sample_probe_pairs(latent.cuda(), pooled.cuda(), 50)
    np.random.seed(11)
    flat = pooled.flatten().numpy()
    valid = np.where(~np.isnan(flat))[0]
    idx = np.random.choice(valid, min(50, len(valid)), replace=False)
    assert X.is_cuda and X.shape == (50, 32) and torch.equal(X.cpu(), latent[0].reshape(32, -1).T[idx])
    assert torch.equal(y.cpu(), torch.from_numpy(flat[idx])) and not torch.isnan(y).any()
    few, yf = pt.sample_probe_pairs(latent.cuda(), pooled.cuda(), 10 ** 6)
    assert few.shape[0] == len(valid) == yf.shape[0]
    assert pt.sample_probe_pairs(latent.cuda(), torch.full((8, 24), float("nan")).cuda(), 5) == (None, None)


def test_probe_target_entry_points_refuse_bad_arguments():
    import ctypes as C
    from tempo_vae_b200 import TvaeError, ops, probe_targets as pt
    from tempo_vae_b200._lib import lib, last_error
    x = torch.zeros((8, 8)).cuda()
    out = torch.zeros(8, dtype=torch.float64).cuda()
    hist = torch.zeros(256, dtype=torch.int64).cuda()
    assert lib.tvae_nan_moments(None, 4, 0.0, out.data_ptr(), None) != 0
    assert lib.tvae_nan_moments(x.data_ptr(), 0, 0.0, out.data_ptr(), None) != 0
    assert lib.tvae_select_hist(x.data_ptr(), 64, 0.0, 0, 0, 0, 12, hist.data_ptr(), None) != 0 and "shift" in last_error()
    assert lib.tvae_select_hist(x.data_ptr(), 64, 0.0, 0, 0x100, 0xFF000000, 8, hist.data_ptr(), None) != 0
    assert lib.tvae_component_pool(x.data_ptr(), 8, 8, 8, 4, 3, 0.0, 1.0, None, out.data_ptr(), None) != 0 and "mode" in last_error()
    assert lib.tvae_component_pool(x.data_ptr(), 8, 8, 4, 4, 0, 0.0, 1.0, None, out.data_ptr(), None) != 0
    assert lib.tvae_component_pool(x.data_ptr(), 2, 8, 8, 4, 0, 0.0, 1.0, None, out.data_ptr(), None) != 0
    with pytest.raises(ValueError):
        pt.normalize_component(x, "boxcox")
    with pytest.raises(TvaeError):
        pt.normalize_component(torch.zeros((2, 3, 4)), "zscore")
    with pytest.raises(TvaeError):
        ops.nan_moments(torch.zeros(4))
    del C
