"""Time the attention core (forward, backward) of the mid block at the train-step size for both TF32 implementations:
tcgen05 kind::tf32 + TMEM (attention_sm100.cu, default) and mma.sync.m16n8k8 (attention_tc.cu).
usage: python tools/attn_bench.py [B] [T] [heads]"""
import sys
import torch
from tempo_vae_b200 import ops as o

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
heads = int(sys.argv[3]) if len(sys.argv) > 3 else 4
C = 32 * heads
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn((B * T, 3 * C), device="cuda", generator=g)
d_out = torch.randn((B * T, C), device="cuda", generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flops_fwd = 4.0 * B * heads * T * T * 32
for impl, name in ((1, "tcgen05"), (0, "mma.sync")):
    o.lib.tvae_attn_set_tcgen05(impl)
    if impl == 0 and B * heads > 65535:
        print(f"{name}: B * heads = {B * heads} exceeds the legacy kernels' grid.y limit, skipped")
        continue
    ob, of, lse = o.attn_fwd(qkv, C, heads, B, T)
    o.attn_bwd(qkv, of, d_out, lse, C, heads, B, T)
    torch.cuda.synchronize()
    tf = tb = 0.0
    n = 10
    for _ in range(n):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        ob, of, lse = o.attn_fwd(qkv, C, heads, B, T)
        e[1].record()
        o.attn_bwd(qkv, of, d_out, lse, C, heads, B, T)
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1])
        tb += e[1].elapsed_time(e[2])
    print(f"{name}: B={B} T={T} heads={heads}  fwd {tf / n:.3f} ms ({flops_fwd / (tf / n) / 1e9:.1f} TFLOP/s)  "
          f"bwd {tb / n:.3f} ms ({2.5 * flops_fwd / (tb / n) / 1e9:.1f} TFLOP/s)")
o.lib.tvae_attn_set_tcgen05(1)
