"""Two launches for ncu: bf16-out-only and f32+bf16+residual+stats on the dominant conv shape."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tempo_vae_b200 import ops as o
B = 256
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, 64, 64, 512), device="cuda", generator=g).to(torch.bfloat16)
w = torch.randn((512, 512, 3, 3), device="cuda", generator=g) / math.sqrt(4608)
bias = torch.randn((512,), device="cuda", generator=g)
res = torch.randn((B, 64, 64, 512), device="cuda", generator=g)
wp = o.pack_weight(w, "fwd")
o.conv_gemm(x, 512, wp, kind=0, R=3, Cout=512, bias=bias, want_f32=False, want_bf16=True)
o.conv_gemm(x, 512, wp, kind=0, R=3, Cout=512, bias=bias, want_f32=True)
o.conv_gemm(x, 512, wp, kind=0, R=3, Cout=512, bias=bias, want_f32=True, want_bf16=True, residual=res, stats=(8, 1e-6))
torch.cuda.synchronize()
print("ok")
