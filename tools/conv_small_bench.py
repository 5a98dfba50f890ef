"""Epilogue-bound convolutions of the train step, isolated: the 2x2 / stride-2 transposed conv 256 -> 512 (32x32 -> 64x64,
fp32 + statistics), the data gradient of the 512 -> 512 down conv (same GEMM shape, bf16 out) and a 256 -> 256 3x3 layer at
32x32. Prints ms, algorithmic TFLOP/s and output TB/s.  usage: python tools/conv_small_bench.py [B] [which]"""
import math
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tempo_vae_b200 import ops as o

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = sys.argv[2] if len(sys.argv) > 2 else "all"
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(tag, fn, flops, out_bytes, n=5):
    fn(); torch.cuda.synchronize()
    t = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t += e0.elapsed_time(e1)
    ms = t / n
    print(f"{tag:58s} {ms:6.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s  out {out_bytes / ms / 1e9:5.2f} TB/s", flush=True)


if which in ("all", "up"):
    x = torch.randn((B, 32, 32, 256), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn((256, 512, 2, 2), device="cuda", generator=g) / 16          # ConvTranspose2d [Cin][Cout][2][2]
    bias = torch.randn((512,), device="cuda", generator=g)
    wp = o.pack_weight(w, "up_fwd")
    fl = 2.0 * B * 1024 * 256 * 2048
    timed("convT 256->512 2x2/s2, f32 + stats", lambda: o.conv_gemm(x, 256, wp, kind=2, R=2, Cout=512, bias=bias, want_f32=True,
                                                                   stats=(8, 1e-6)), fl, B * 4096 * 512 * 4)
    timed("convT 256->512 2x2/s2, f32", lambda: o.conv_gemm(x, 256, wp, kind=2, R=2, Cout=512, bias=bias, want_f32=True),
          fl, B * 4096 * 512 * 4)
    timed("convT 256->512 2x2/s2, bf16", lambda: o.conv_gemm(x, 256, wp, kind=2, R=2, Cout=512, bias=bias, want_f32=False,
                                                            want_bf16=True), fl, B * 4096 * 512 * 2)
if which in ("all", "3x3"):
    x = torch.randn((B, 32, 32, 256), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn((256, 256, 3, 3), device="cuda", generator=g) / math.sqrt(2304)
    res = torch.randn((B, 32, 32, 256), device="cuda", generator=g)
    bias = torch.randn((256,), device="cuda", generator=g)
    wp = o.pack_weight(w, "fwd")
    fl = 2.0 * B * 1024 * 256 * 2304
    timed("3x3 256->256 @32, f32 + bf16 + residual", lambda: o.conv_gemm(x, 256, wp, kind=0, R=3, Cout=256, bias=bias, want_f32=True,
                                                                        want_bf16=True, residual=res), fl, B * 1024 * 256 * 10)
    timed("3x3 256->256 @32, bf16", lambda: o.conv_gemm(x, 256, wp, kind=0, R=3, Cout=256, bias=bias, want_f32=False,
                                                       want_bf16=True), fl, B * 1024 * 256 * 2)
if which in ("all", "1x1"):
    x = torch.randn((B, 64, 64, 256), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn((512, 256, 1, 1), device="cuda", generator=g) / 16
    bias = torch.randn((512,), device="cuda", generator=g)
    wp = o.pack_weight(w, "fwd")
    fl = 2.0 * B * 4096 * 256 * 512
    timed("1x1 256->512 @64, f32", lambda: o.conv_gemm(x, 256, wp, kind=0, R=1, Cout=512, bias=bias, want_f32=True),
          fl, B * 4096 * 512 * 4)
    timed("1x1 256->512 @64, bf16", lambda: o.conv_gemm(x, 256, wp, kind=0, R=1, Cout=512, bias=bias, want_f32=False,
                                                       want_bf16=True), fl, B * 4096 * 512 * 2)
if which in ("all", "nll"):
    # decoder.conv_out: 512 -> 1028 3x3 @64, fp32 output vs the fused reconstruction-loss epilogue (bf16 gradient out)
    x = torch.randn((B, 64, 64, 512), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn((1028, 512, 3, 3), device="cuda", generator=g) / math.sqrt(4608)
    bias = torch.randn((1028,), device="cuda", generator=g)
    wp = o.pack_weight(w, "fwd")
    target = torch.randn((B, 64, 64, 1032), device="cuda", generator=g).to(torch.bfloat16)
    logvar = torch.zeros((1,), device="cuda")
    out = torch.empty((B, 64, 64, 1032), device="cuda", dtype=torch.bfloat16)
    of32 = torch.empty((B, 64, 64, 1028), device="cuda", dtype=torch.float32)
    fl = 2.0 * B * 4096 * 1028 * 4608
    for rep in range(2):
        timed("conv_out 512->1028, f32", lambda: o.conv_gemm(x, 512, wp, kind=0, R=3, Cout=1028, bias=bias, want_f32=True,
                                                             out_f32=of32), fl, B * 4096 * 1028 * 4)
        timed("conv_out 512->1028, bf16", lambda: o.conv_gemm(x, 512, wp, kind=0, R=3, Cout=1028, bias=bias, want_f32=False,
                                                              want_bf16=True, out_bf16=out), fl, B * 4096 * 1028 * 2)
        timed("conv_out 512->1028, fused loss", lambda: o.conv_gemm(
            x, 512, wp, kind=0, R=3, Cout=1028, bias=bias, want_f32=False, want_bf16=True, out_bf16=out,
            nll={"x": target[..., :1028], "loss_type": 0, "logvar": logvar, "batch": B}), fl, B * 4096 * 1028 * 4)
if which in ("nll1024",):
    # what would decoder.conv_out cost with whole 256-column tiles only (1024 of the 1028 output channels)?
    x = torch.randn((B, 64, 64, 512), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn((1028, 512, 3, 3), device="cuda", generator=g) / math.sqrt(4608)
    bias = torch.randn((1028,), device="cuda", generator=g)
    wp = o.pack_weight(w, "fwd")
    target = torch.randn((B, 64, 64, 1032), device="cuda", generator=g).to(torch.bfloat16)
    logvar = torch.zeros((1,), device="cuda")
    out = torch.empty((B, 64, 64, 1032), device="cuda", dtype=torch.bfloat16)
    for rep in range(3):
        for Cout in (1028, 1024):
            fl = 2.0 * B * 4096 * Cout * 4608
            timed(f"conv_out 512->{Cout}, fused loss", lambda: o.conv_gemm(
                x, 512, wp, kind=0, R=3, Cout=Cout, bias=bias, want_f32=False, want_bf16=True, out_bf16=out,
                nll={"x": target[..., :Cout], "loss_type": 0, "logvar": logvar, "batch": B}), fl, B * 4096 * Cout * 4)
