"""Where does the LIVE train step go?  Every C-ABI call of a B=256 step is bracketed with CUDA events on the launching
stream (a proxy around the ctypes library object), then summed per entry point; the time not covered by any call
(launch gaps, torch's own small kernels: zero_grad, bias-gradient copies) is reported as "outside".

    python tools/step_timeline.py [batch] [steps]

Unlike the ncu launch list (cold, serialised, isolated clocks) this is the power-capped steady state of the real step.
"""
import collections
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (DEFAULT_MODEL, synthetic_batch)
import tempo_vae_b200 as t  # noqa: E402
from tempo_vae_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5


class TimedLib:
    def __init__(self, lib):
        self._lib, self.events, self.on, self.shapes = lib, [], False, []

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("tvae_") or name.endswith("_bytes") or name.endswith("_splits") or "set_cta" in name:
            return fn

        def call(*a):
            if not self.on:
                return fn(*a)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            self.events.append((name, e0, e1))
            if name in ("tvae_conv_gemm", "tvae_wgrad_gemm"):       # per-shape table of the GEMM calls
                g = a[0]._obj                                       # ctypes.byref(struct)
                if name == "tvae_conv_gemm":
                    px = g.N * g.H * g.W // (4 if g.kind == 1 else 1)
                    taps = g.R * g.R if g.kind == 0 else 4
                    key = ("conv", px, g.C, g.Cout, g.kind, g.R, int(g.flip), "f32" * bool(g.out_f32) + "+bf16" * bool(g.out_bf16)
                           + "+res" * bool(g.residual) + "+stats" * bool(g.stats_part))
                    flops = 2.0 * px * g.C * g.Cout * (taps if g.kind != 2 else 1) * (4 if g.kind == 2 else 1)
                else:
                    px = g.N * g.H * g.W
                    taps = g.R * g.R if g.kind == 0 else 4
                    key = ("wgrad", px, g.Cm, g.Cn, g.kind, g.R, int(g.flip), "")
                    flops = 2.0 * px * g.Cm * g.Cn * taps
                self.shapes.append((key, flops, e0, e1))
            return rc
        return call


dev = torch.device("cuda", 0)
t.seed_all(42)
model = t.get_model(bench.DEFAULT_MODEL, dev)
trainer = t.Trainer(model, model.optimizer, dev, tempfile.mkdtemp(prefix="tvae_tl_"))
xs = [bench.synthetic_batch(torch, B, (1028, 64, 64), dev, seed=i) for i in range(2)]
proxy = TimedLib(ops.lib)
ops.lib = proxy
for i in range(4):
    trainer.train_step_device(xs[i % 2])
    trainer.step = 1
torch.cuda.synchronize()
proxy.on = True
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
for i in range(STEPS):
    trainer.train_step_device(xs[i % 2])
s1.record()
torch.cuda.synchronize()
total = s0.elapsed_time(s1) / STEPS
agg = collections.defaultdict(lambda: [0, 0.0])
for name, e0, e1 in proxy.events:
    a = agg[name]
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
inside = sum(v[1] for v in agg.values()) / STEPS
print(f"B={B}: {total:.2f} ms/step live (with {len(proxy.events) // STEPS} event pairs per step); "
      f"inside C-ABI calls {inside:.2f} ms, outside {total - inside:.2f} ms")
print("| entry point | calls/step | ms/step | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0] / STEPS:.0f} | {v[1] / STEPS:.2f} | {100 * v[1] / STEPS / total:.1f} % |")

print("\n| GEMM call (pixels, Cin|Cm, Cout|Cn, kind, R, flip, outputs) | calls/step | ms/step | TFLOP/s (algorithmic) |\n|---|---|---|---|")
sh = collections.defaultdict(lambda: [0, 0.0, 0.0])
for key, flops, e0, e1 in proxy.shapes:
    a = sh[key]
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
    a[2] += flops
for k, v in sorted(sh.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {v[0] / STEPS:.0f} | {v[1] / STEPS:.2f} | {v[2] / v[1] / 1e9:.0f} |")
