"""Launches for one `ncu --set full` capture (-k regex:gemm_kernel): the dominant conv shape with three epilogues, its
CTA-pair schedule, and the weight-gradient GEMM of the same layer."""
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
# the CTA-pair schedule of the same bf16-out launch, and the weight-gradient GEMM of the same layer
from tempo_vae_b200._lib import lib
lib.tvae_conv_set_cta_pair(1)
o.conv_gemm(x, 512, wp, kind=0, R=3, Cout=512, bias=bias, want_f32=False, want_bf16=True)
lib.tvae_conv_set_cta_pair(0)
dy = torch.randn((B, 64, 64, 512), device="cuda", generator=g).to(torch.bfloat16)
grad = torch.empty((512, 512, 3, 3), device="cuda")
o.wgrad_gemm(dy, 512, x, 512, kind=0, R=3, grad=grad)
torch.cuda.synchronize()
print("ok2")
# GroupNorm + GELU forward / backward on the same tensor size (fp32 residual-stream input), for the HBM-bound roofline
C, G = 512, 8
xf = torch.randn((B, 64, 64, C), device="cuda", generator=g)
gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
stats = o.gn_stats(xf, C, G, 1e-6)
o.gn_act_fwd(xf, stats, gamma, beta, G, 1)
dg, db, cs = (torch.empty(C, device="cuda") for _ in range(3))
o.gn_act_bwd(xf, stats, gamma, beta, dy, res.to(torch.bfloat16), G, 1, dg, db, cs)
torch.cuda.synchronize()
print("ok3")
