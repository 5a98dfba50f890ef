"""Generate tests/golden/tile_prep.pt by RUNNING THE REAL REFERENCE's tile extraction (build container only).

src/scripts/prepare_tempo_tiles.py cannot be imported here (it imports netCDF4 at module level, absent from this image),
so the one function on this path, `extract_tiles` (:21-58), is compiled on its own from the reference file -- its
source is located with `ast`, never written anywhere -- and executed with numpy/torch in its namespace. The fixture
stores input, seed and output, and pins oracle.extract_tiles (tests/test_oracle_cpu.py) and, through it, the CUDA
kernel tvae_extract_tiles (tests/test_data_gpu.py).

  python oracle/make_golden_data.py [tiles] [probes] [probe_targets]
"""
import ast
import os
import sys

import numpy as np
import torch

REF = os.environ.get("TVAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_function(path, name):
    src = open(path).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np, "torch": torch}
    exec(compile(mod, f"<reference>/{os.path.relpath(path, REF)}", "exec"), ns)
    return ns[name]


def main():
    extract_tiles = reference_function(os.path.join(REF, "src/scripts/prepare_tempo_tiles.py"), "extract_tiles")
    g = torch.Generator().manual_seed(17)
    cases = []
    for (M, NT, C, T, n, seed) in [(20, 29, 6, 8, 12, 3), (16, 16, 5, 16, 9, 11), (131, 70, 6, 64, 5, 5)]:
        z = torch.randn((M, NT, C), generator=g)
        tiles = extract_tiles(z, (T, T), n, seed=seed)
        cases.append(dict(z=z, tile=T, n=n, seed=seed, tiles=tiles))
    small = extract_tiles(torch.zeros(4, 4, 2), (8, 8), 3, seed=0)
    assert small is None
    out = os.path.join(ROOT, "tests", "golden", "tile_prep.pt")
    torch.save(dict(cases=cases, source="src/scripts/prepare_tempo_tiles.py:21-58 executed from /root/reference"), out)
    print("wrote", out, os.path.getsize(out), "bytes")



def probe_fixture():
    """tests/golden/probes.pt: the reference's own LinearProbe / MLPProbe / train_probe (compiled on their own from
    src/scripts/linear_probe_analysis.py:212-353, which imports netCDF4 at module level) trained on a small synthetic
    latent -> component regression, CPU, fixed seeds."""
    import ast as _ast
    import torch.nn as nn
    path = os.path.join(REF, "src/scripts/linear_probe_analysis.py")
    tree = _ast.parse(open(path).read())
    want = {"LinearProbe", "MLPProbe", "train_probe"}
    body = [n for n in tree.body if isinstance(n, (_ast.ClassDef, _ast.FunctionDef)) and n.name in want]
    ns = {"np": np, "torch": torch, "nn": nn}
    exec(compile(_ast.Module(body=body, type_ignores=[]), "<reference>/src/scripts/linear_probe_analysis.py", "exec"), ns)
    g = torch.Generator().manual_seed(5)
    n_tr, n_va = 1600, 400
    X = torch.randn((n_tr + n_va, 32), generator=g)
    w = torch.randn((32,), generator=g) / 32 ** 0.5
    y = X @ w + 0.5 * torch.tanh(X[:, 0] * X[:, 1]) + 0.1 * torch.randn((n_tr + n_va,), generator=g)
    Xtr, ytr, Xva, yva = X[:n_tr].numpy(), y[:n_tr].numpy(), X[n_tr:].numpy(), y[n_tr:].numpy()
    cases = {}
    for name, cfg in {
        "linear": dict(architecture="linear", learning_rate=0.01, weight_decay=0.01, batch_size=512, max_epochs=30),
        "mlp_relu_nodrop": dict(architecture="mlp", hidden_dims=[64, 64], dropout=0.0, activation="relu", learning_rate=0.001,
                                weight_decay=0.01, batch_size=512, max_epochs=15),
        "mlp_gelu_drop": dict(architecture="mlp", hidden_dims=[64, 64], dropout=0.1, activation="gelu", learning_rate=0.001,
                              weight_decay=0.01, batch_size=512, max_epochs=15),
    }.items():
        torch.manual_seed(31)
        probe, tl, vl = ns["train_probe"](Xtr, ytr, Xva, yva, cfg)
        probe.eval()
        with torch.no_grad():
            pred = probe(torch.from_numpy(Xva)).squeeze(1)
        cases[name] = dict(config=cfg, train_losses=tl, val_losses=vl, pred_val=pred.clone(),
                           state_dict={k: v.clone() for k, v in probe.state_dict().items()})
        print(name, "final train/val loss", tl[-1], vl[-1])
    out = os.path.join(ROOT, "tests", "golden", "probes.pt")
    torch.save(dict(X_train=torch.from_numpy(Xtr), y_train=torch.from_numpy(ytr), X_val=torch.from_numpy(Xva),
                    y_val=torch.from_numpy(yva), seed=31, cases=cases,
                    source="src/scripts/linear_probe_analysis.py:212-353 executed from /root/reference"), out)
    print("wrote", out, os.path.getsize(out), "bytes")



def probe_target_fixture():
    """tests/golden/probe_targets.pt: the reference's own `normalize_component` (compiled on its own from
    src/scripts/linear_probe_analysis.py:60-110; scipy is in this image) on synthetic component fields with NaN blobs,
    followed by the pooling statement of `process_file` (:183-190: reshape(h//4, 4, w//4, 4) + np.nanmean over axes
    (1, 3)), which cannot run as a function (the rest of process_file reads NetCDF files)."""
    import warnings
    normalize_component = reference_function(os.path.join(REF, "src/scripts/linear_probe_analysis.py"), "normalize_component")
    rs = np.random.RandomState(23)
    cases = []
    for name, norm_type, (H, W), make in [
        ("no2_like", "asinh", (64, 192), lambda s: (rs.standard_t(3, size=s) * 2e15 + 1e15)),
        ("o3_like", "zscore", (64, 192), lambda s: rs.normal(300.0, 25.0, size=s)),
        ("hcho_like", "asinh", (32, 100), lambda s: rs.standard_t(4, size=s) * 8e15),
        ("cloud_like", "logit", (64, 192), lambda s: rs.beta(0.6, 1.5, size=s)),
        ("minmax_odd", "minmax", (36, 52), lambda s: rs.normal(0.0, 1.0, size=s)),
    ]:
        field = make((H, W)).astype(np.float32)
        blob = np.kron(rs.rand(H // 4, W // 4) < 0.12, np.ones((4, 4), dtype=bool))          # whole 4x4 blocks invalid
        field[blob] = np.nan
        field[rs.rand(H, W) < 0.08] = np.nan                                                  # and scattered pixels
        normalized, stats = normalize_component(field.copy(), norm_type)
        h, w = normalized.shape
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pooled = np.nanmean(normalized.reshape(h // 4, 4, w // 4, 4), axis=(1, 3))
        again, _ = normalize_component(field.copy(), norm_type, stats)                        # the `stats` given path
        assert np.array_equal(again, normalized, equal_nan=True)
        cases.append(dict(name=name, norm_type=norm_type, field=torch.from_numpy(field),
                          normalized=torch.from_numpy(np.asarray(normalized, dtype=np.float32)),
                          pooled=torch.from_numpy(np.asarray(pooled, dtype=np.float32)),
                          stats={k: float(v) for k, v in stats.items()},
                          stats_dtype={k: str(np.asarray(v).dtype) for k, v in stats.items()},
                          normalized_dtype=str(normalized.dtype)))
        print(name, norm_type, {k: float(v) for k, v in stats.items()}, normalized.dtype,
              "valid pooled:", int((~np.isnan(pooled)).sum()), "of", pooled.size)
    out = os.path.join(ROOT, "tests", "golden", "probe_targets.pt")
    torch.save(dict(cases=cases, source="src/scripts/linear_probe_analysis.py:60-110 executed from /root/reference; "
                                        ":183-190 restated in oracle/make_golden_data.py"), out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    # python oracle/make_golden_data.py [tiles] [probes] [probe_targets]   (no argument: all three)
    todo = sys.argv[1:] or ["tiles", "probes", "probe_targets"]
    for what in todo:
        {"tiles": main, "probes": probe_fixture, "probe_targets": probe_target_fixture}[what]()
