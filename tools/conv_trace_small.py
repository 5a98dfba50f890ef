"""conv_gemm_kernel timeline (tvae_conv_set_trace) on the epilogue-bound shapes of the step: 128->128 3x3 @16x16 and the
transposed conv 256->512 @32x32 (B=256)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tempo_vae_b200 import ops as o  # noqa: E402
from tempo_vae_b200._lib import lib  # noqa: E402

B = 256
g = torch.Generator(device="cuda").manual_seed(0)
CAP = 64


def trace(name, x, cin, wp, **kw):
    for _ in range(2):
        o.conv_gemm(x, cin, wp, **kw)
    buf = torch.zeros((148 * CAP * 8,), dtype=torch.int64, device="cuda")
    lib.tvae_conv_set_trace(buf.data_ptr(), CAP)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o.conv_gemm(x, cin, wp, **kw)
    e1.record()
    torch.cuda.synchronize()
    lib.tvae_conv_set_trace(None, 0)
    t = buf.view(148, CAP, 8).cpu().double()
    units = [u for u in range(148) if t[u, 0, 3] > 0]
    rows = []
    for u in units:
        n = int((t[u, :, 3] > 0).sum())
        for i in range(1, n):
            r, prev = t[u, i], t[u, i - 1]
            rows.append([float(r[1] - r[0]), float(r[2] - r[1]), float(r[3] - r[2]), float(r[6] - r[5]), float(r[3] - prev[3])])
    m = torch.tensor(rows).mean(0).tolist() if rows else [0] * 5
    n0 = int((t[units[0], :, 3] > 0).sum())
    first = t[units[0], 0]
    print(f"== {name}: {e0.elapsed_time(e1) * 1e3:.1f} us, {len(units)} units x {n0} tiles; per tile (cycles): wait free "
          f"accumulator {m[0]:.0f}, wait first stage {m[1]:.0f}, issue window {m[2]:.0f}, epilogue {m[3]:.0f}, tile period "
          f"{m[4]:.0f}; first tile: start->first stage {float(first[2] - first[0]):.0f}, epilogue {float(first[6] - first[5]):.0f}")


x16 = torch.randn((B, 16, 16, 128), device="cuda", generator=g).to(torch.bfloat16)
w16 = torch.randn((128, 128, 3, 3), device="cuda", generator=g) / 34.0
b16 = torch.randn((128,), device="cuda", generator=g)
r16 = torch.randn((B, 16, 16, 128), device="cuda", generator=g)
wp16 = o.pack_weight(w16, "fwd")
trace("128->128 3x3 @16x16, f32 + residual + stats", x16, 128, wp16, kind=0, R=3, Cout=128, bias=b16, want_f32=True,
      residual=r16, stats=(8, 1e-6))
trace("128->128 3x3 @16x16, bf16 + stats", x16, 128, wp16, kind=0, R=3, Cout=128, bias=b16, want_f32=False, want_bf16=True,
      stats=(8, 1e-6))
trace("128->128 3x3 @16x16, bf16 only (LEAN)", x16, 128, wp16, kind=0, R=3, Cout=128, bias=b16, want_f32=False, want_bf16=True)
x32 = torch.randn((B, 32, 32, 256), device="cuda", generator=g).to(torch.bfloat16)
wT = torch.randn((256, 512, 2, 2), device="cuda", generator=g) / 16.0
bT = torch.randn((512,), device="cuda", generator=g)
wpT = o.pack_weight(wT, "up_fwd")
trace("convT 2x2 s2 256->512 @32x32, f32 + stats", x32, 256, wpT, kind=2, R=2, Cout=512, bias=bT, want_f32=True, stats=(8, 1e-6))
trace("convT 2x2 s2 256->512 @32x32, f32", x32, 256, wpT, kind=2, R=2, Cout=512, bias=bT, want_f32=True)
