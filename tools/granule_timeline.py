"""Where does a whole-granule pass go?  Every C-ABI call of `encode_granule_whole` / `reconstruct_granule_whole` on a
[131, 2048, 1028] granule is bracketed with CUDA events (the proxy of tools/step_timeline.py) and listed per entry point
and per GEMM shape.

    python tools/granule_timeline.py [encode|reconstruct] [repeats]
"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import tempo_vae_b200 as t  # noqa: E402
from tempo_vae_b200 import ops  # noqa: E402

WHAT = sys.argv[1] if len(sys.argv) > 1 else "reconstruct"
REP = int(sys.argv[2]) if len(sys.argv) > 2 else 3


class TimedLib:
    def __init__(self, lib):
        self._lib, self.events, self.on = lib, [], False

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("tvae_") or name.endswith("_bytes") or name.endswith("_splits") or "set_" in name:
            return fn

        def call(*a):
            if not self.on:
                return fn(*a)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            key = name
            if name == "tvae_conv_gemm":
                g = a[0]._obj
                key = (f"conv N{g.N} {g.H}x{g.W} {g.C}->{g.Cout} kind{g.kind} R{g.R} flip{int(g.flip)} "
                       + "f32" * bool(g.out_f32) + "+bf16" * bool(g.out_bf16) + "+res" * bool(g.residual)
                       + "+stats" * bool(g.stats_part))
            self.events.append((key, e0, e1))
            return rc
        return call


dev = torch.device("cuda", 0)
t.seed_all(42)
model = t.get_model(bench.DEFAULT_MODEL, dev)
g = torch.Generator(device=dev).manual_seed(1)
z = torch.randn((131, 2048, 1028), device=dev, generator=g).clamp_(-10, 10)
fn = (lambda: t.reconstruct_granule_whole(model, z)) if WHAT == "reconstruct" else (lambda: t.encode_granule_whole(model, z))
proxy = TimedLib(ops.lib)
ops.lib = proxy
for _ in range(6):
    fn()
torch.cuda.synchronize()
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
for _ in range(REP):
    fn()
s1.record()
torch.cuda.synchronize()
plain = s0.elapsed_time(s1) / REP
proxy.on = True
s0.record()
for _ in range(REP):
    fn()
s1.record()
torch.cuda.synchronize()
total = s0.elapsed_time(s1) / REP
agg = collections.defaultdict(lambda: [0, 0.0])
for key, e0, e1 in proxy.events:
    agg[key][0] += 1
    agg[key][1] += e0.elapsed_time(e1)
inside = sum(v[1] for v in agg.values()) / REP
print(f"{WHAT}: {plain:.2f} ms per granule ({total:.2f} ms with {len(proxy.events) // REP} event pairs); inside C-ABI calls "
      f"{inside:.2f} ms, outside {total - inside:.2f} ms")
for key, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"  {ms / REP:8.3f} ms  x{n // REP:<3d} {key}")
