"""Micro-benchmark of conv_gemm on the dominant shape (B=256, 512->512 3x3 @64x64): what do the epilogue variants and
the operand data cost?  CUDA events, 10 launches each, inputs far larger than L2."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tempo_vae_b200 import ops as o

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, 64, 64, 512), device="cuda", generator=g).to(torch.bfloat16)
w_rand = torch.randn((512, 512, 3, 3), device="cuda", generator=g) / math.sqrt(4608)
w_low = torch.sign(w_rand) * 1e-4
bias = torch.randn((512,), device="cuda", generator=g)
res = torch.randn((B, 64, 64, 512), device="cuda", generator=g)
flops = 2.0 * B * 4096 * 512 * 4608


def run(tag, w, xin, **kw):
    wp = o.pack_weight(w, "fwd")
    for _ in range(3):
        o.conv_gemm(xin, 512, wp, kind=0, R=3, Cout=512, bias=bias, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        o.conv_gemm(xin, 512, wp, kind=0, R=3, Cout=512, bias=bias, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{tag:46s} {ms:6.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)


from tempo_vae_b200._lib import lib

# interleave the two schedules (the board is power-capped and drifts by several % within a sequence)
for rep in range(2):
    for pair in (1, 0):
        lib.tvae_conv_set_cta_pair(pair)
        t = "pair  " if pair else "single"
        run(f"[{t}] bf16 out only", w_rand, x, want_f32=False, want_bf16=True)
        run(f"[{t}] f32 out only", w_rand, x, want_f32=True)
        run(f"[{t}] f32 + bf16 out + residual + stats", w_rand, x, want_f32=True, want_bf16=True, residual=res,
            stats=(8, 1e-6))
lib.tvae_conv_set_cta_pair(1)
run("f32 out + stats, random weights", w_rand, x, want_f32=True, stats=(8, 1e-6))
run("f32 out + residual, random weights", w_rand, x, want_f32=True, residual=res)
run("f32 out only, low-entropy weights (+-1e-4)", w_low, x, want_f32=True)
run("bf16 out only, zero activations", w_rand, torch.zeros_like(x), want_f32=False, want_bf16=True)

# weight-gradient GEMM of the same layer, sustained (same flops as one conv launch)
dy = torch.randn((B, 64, 64, 512), device="cuda", generator=g).to(torch.bfloat16)
grad = torch.empty((512, 512, 3, 3), device="cuda")


def run_wgrad(tag, p, q):
    for _ in range(3):
        o.wgrad_gemm(p, 512, q, 512, kind=0, R=3, grad=grad)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        o.wgrad_gemm(p, 512, q, 512, kind=0, R=3, grad=grad)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{tag:46s} {ms:6.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)


for rep in range(2):
    run_wgrad("wgrad 512x512x3x3, random dY / x", dy, x)
    run("conv bf16 out only (same flops)", w_rand, x, want_f32=False, want_bf16=True)
run_wgrad("wgrad, zero dY", torch.zeros_like(dy), x)
