"""Eager launches against CUDA-graph replay (`tempo_vae_b200.GranuleGraph`) of the whole-granule encode for crops of
different sizes: where is the pass bound by the host's launch cost, where by the GPU?

    python tools/granule_graph_bench.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import tempo_vae_b200 as t  # noqa: E402

dev = torch.device("cuda", 0)
t.seed_all(42)
model = t.get_model(bench.DEFAULT_MODEL, dev)
g = torch.Generator(device=dev).manual_seed(1)


def timed(fn, n=20):
    for _ in range(15):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for M, T in ((64, 64), (64, 256), (128, 512), (128, 2048)):
    try:
        z = torch.randn((M, T, 1028), device=dev, generator=g).clamp_(-10, 10)
        eager = timed(lambda: t.encode_granule_whole(model, z))
        graph = t.GranuleGraph(model, z.shape)
        same = torch.equal(graph(z), t.encode_granule_whole(model, z))
        replay = timed(lambda: graph(z))
        print(f"crop {M}x{T}: eager {eager:.3f} ms, graph replay {replay:.3f} ms ({eager / replay:.2f}x), bit-identical {same}")
        del graph
    except Exception as e:  # noqa: BLE001
        print(f"crop {M}x{T}: {type(e).__name__}: {str(e)[:200]}")
