"""Generate tests/golden/tile_prep.pt by RUNNING THE REAL REFERENCE's tile extraction (build container only).

src/scripts/prepare_tempo_tiles.py cannot be imported here (it imports netCDF4 at module level, absent from this image),
so the one function on this path, `extract_tiles` (:21-58), is compiled on its own from the reference file -- its
source is located with `ast`, never written anywhere -- and executed with numpy/torch in its namespace. The fixture
stores input, seed and output, and pins oracle.extract_tiles (tests/test_oracle_cpu.py) and, through it, the CUDA
kernel tvae_extract_tiles (tests/test_data_gpu.py).

  python oracle/make_golden_data.py
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


if __name__ == "__main__":
    main()
