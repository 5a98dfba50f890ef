"""Micro-benchmark of the GroupNorm+GELU kernels on the dominant shape (B=256, [64,64,512], 8 groups): ms and GB/s of
algorithmic traffic for forward and backward, fp32 and bf16 input. CUDA events, 10 launches each."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tempo_vae_b200 import ops as o

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
C, G = 512, 8
g = torch.Generator(device="cuda").manual_seed(0)
x32 = torch.randn((B, 64, 64, C), device="cuda", generator=g)
x16 = x32.to(torch.bfloat16)
gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
da = torch.randn((B, 64, 64, C), device="cuda", generator=g).to(torch.bfloat16)
gres = torch.randn((B, 64, 64, C), device="cuda", generator=g).to(torch.bfloat16)
stats = o.gn_stats(x32, C, G, 1e-6)
dg, db, cs = (torch.empty(C, device="cuda") for _ in range(3))
n = x32.numel()


def timed(tag, fn, nbytes):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{tag:44s} {ms:6.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s", flush=True)


for rep in range(2):
    timed("fwd  fp32 in (4+2 B)", lambda: o.gn_act_fwd(x32, stats, gamma, beta, G, 1), 6 * n)
    timed("fwd  bf16 in (2+2 B)", lambda: o.gn_act_fwd(x16, stats, gamma, beta, G, 1), 4 * n)
    timed("bwd  fp32 in, no residual (6 + 8 B)", lambda: o.gn_act_bwd(x32, stats, gamma, beta, da, None, G, 1, dg, db, cs), 14 * n)
    timed("bwd  fp32 in, residual   (6 + 10 B)", lambda: o.gn_act_bwd(x32, stats, gamma, beta, da, gres, G, 1, dg, db, cs), 16 * n)
    timed("bwd  bf16 in, no residual (4 + 6 B)", lambda: o.gn_act_bwd(x16, stats, gamma, beta, da, None, G, 1, dg, db, cs), 10 * n)
    timed("bwd  identity act, fp32 in (6 + 8 B)", lambda: o.gn_act_bwd(x32, stats, gamma, beta, da, None, G, 0, dg, db, cs), 14 * n)
