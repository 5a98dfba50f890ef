"""Per-tile timeline of conv_gemm_kernel on the dominant layer (B=256, 512->512 3x3 @64x64): where do the ~1.5 K idle
tensor cycles per tile switch go?  Uses tvae_conv_set_trace (SM clock cycles recorded by the producer, MMA and epilogue
warps of every leader CTA) and prints per-tile averages over all units, skipping each unit's first and last tile."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tempo_vae_b200 import ops as o  # noqa: E402
from tempo_vae_b200._lib import lib  # noqa: E402

B = int(os.environ.get("TRACE_B", "256"))
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, 64, 64, 512), device="cuda", generator=g).to(torch.bfloat16)
w = torch.randn((512, 512, 3, 3), device="cuda", generator=g) / 68.0
bias = torch.randn((512,), device="cuda", generator=g)
res = torch.randn((B, 64, 64, 512), device="cuda", generator=g)
wp = o.pack_weight(w, "fwd")
CAP = 128
variants = {
    "bf16 out (LEAN dgrad-style epilogue)": dict(want_f32=False, want_bf16=True),
    "fp32 out": dict(want_f32=True),
    "fp32 + bf16 out + residual + stats": dict(want_f32=True, want_bf16=True, residual=res, stats=(8, 1e-6)),
}
for name, kw in variants.items():
    for _ in range(2):
        o.conv_gemm(x, 512, wp, kind=0, R=3, Cout=512, bias=bias, **kw)
    buf = torch.zeros((148 * CAP * 8,), dtype=torch.int64, device="cuda")
    lib.tvae_conv_set_trace(buf.data_ptr(), CAP)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o.conv_gemm(x, 512, wp, kind=0, R=3, Cout=512, bias=bias, **kw)
    e1.record()
    torch.cuda.synchronize()
    lib.tvae_conv_set_trace(None, 0)
    t = buf.view(148, CAP, 8).cpu().double()
    units = [u for u in range(148) if t[u, 0, 3] > 0]
    rows = []
    for u in units:
        n = int((t[u, :, 3] > 0).sum())
        for i in range(1, n - 1):
            r, prev = t[u, i], t[u, i - 1]
            rows.append([
                float(r[1] - r[0]),            # wait for a free accumulator
                float(r[2] - r[1]),            # wait for the tile's first operand stage
                float(r[4]),                   # waits for the other stages
                float(r[3] - r[2]),            # issue window of the tile
                float(r[0] - prev[3]),         # gap between last issue of the previous tile and this tile's start
                float(r[6] - r[5]),            # epilogue duration
                float(r[5] - r[3]),            # accumulator ready after the last K block was issued (MMA drain)
                float(r[2] - r[7]),            # producer lead: first load issued -> first stage consumed
                float(r[3] - prev[3]),         # tile period
            ])
    m = torch.tensor(rows).mean(0).tolist()
    mx = torch.tensor(rows).max(0).values.tolist()
    print(f"== {name}: {e0.elapsed_time(e1):.3f} ms, {len(units)} units x {int((t[units[0], :, 3] > 0).sum())} tiles")
    for lab, a, b in zip(["wait free accumulator", "wait first operand stage", "wait other stages (sum)", "issue window",
                          "gap prev last issue -> start", "epilogue duration", "MMA drain after last issue",
                          "producer lead at first stage", "tile period"], m, mx):
        print(f"   {lab:32s} mean {a:9.0f}  max {b:9.0f} cycles")
