"""Dump small conv / wgrad GEMM results next to their references for offline inspection (gpurun_out/debug_gemm.pt)."""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tempo_vae_b200 import ops as o  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
out = {}


def nhwc(x, pitch):
    N, C, H, W = x.shape
    t = torch.zeros((N, H, W, pitch), dtype=torch.bfloat16, device=x.device)
    t[..., :C] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return t


def run(tag, fn):
    try:
        fn()
        torch.cuda.synchronize()
        print(tag, "ok", flush=True)
    except Exception as e:  # noqa: BLE001
        print(tag, "EXC", repr(e)[:500], flush=True)
        out[tag + "_exc"] = repr(e)


def conv_case(tag, N, H, W, Cin, Cout, R, onehot=False):
    def f():
        g = torch.Generator(device="cuda").manual_seed(0)
        if onehot:
            x = torch.zeros((N, Cin, H, W), device="cuda")
            pix = torch.arange(N * H * W, device="cuda").reshape(N, H, W)
            x.scatter_(1, (pix % Cin).unsqueeze(1), 1.0)
        else:
            x = torch.randn((N, Cin, H, W), device="cuda", generator=g).to(torch.bfloat16).float()
        w = (torch.randn((Cout, Cin, R, R), device="cuda", generator=g) / math.sqrt(Cin * R * R)).to(torch.bfloat16).float()
        ref = F.conv2d(x, w, None, padding=R // 2)
        of, _ = o.conv_gemm(nhwc(x, o.round_up(Cin, 8)), Cin, o.pack_weight(w, "fwd"), kind=0, R=R, Cout=Cout)
        torch.cuda.synchronize()
        got = of[..., :Cout].permute(0, 3, 1, 2)
        err = ((got - ref).abs().max() / ref.abs().max()).item()
        print(f"{tag}: rel_err={err:.3e}", flush=True)
        out[tag] = dict(got=got.cpu(), ref=ref.cpu(), x=x.cpu(), w=w.cpu())
    run(tag, f)


def wgrad_case(tag, N, H, W, Cin, Cout, R):
    def f():
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn((N, Cin, H, W), device="cuda", generator=g).to(torch.bfloat16).float()
        dy = torch.randn((N, Cout, H, W), device="cuda", generator=g).to(torch.bfloat16).float()
        ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, R, R), dy, padding=R // 2)
        grad = torch.zeros((Cout, Cin, R, R), device="cuda")
        o.wgrad_gemm(nhwc(dy, o.round_up(Cout, 8)), Cout, nhwc(x, o.round_up(Cin, 8)), Cin, kind=0, R=R, grad=grad, splits=1)
        torch.cuda.synchronize()
        err = ((grad - ref).abs().max() / ref.abs().max()).item()
        print(f"{tag}: rel_err={err:.3e}", flush=True)
        out[tag] = dict(got=grad.cpu(), ref=ref.cpu())
    run(tag, f)


conv_case("c1x1_onehot", 1, 8, 16, 64, 64, 1, onehot=True)
conv_case("c1x1_rand", 1, 8, 16, 64, 64, 1)
conv_case("c1x1_k128", 1, 8, 16, 128, 64, 1)
conv_case("c3x3", 1, 16, 16, 64, 64, 3)
conv_case("c1x1_n256", 2, 16, 16, 64, 256, 1)
wgrad_case("w1x1", 1, 8, 8, 64, 64, 1)
wgrad_case("w1x1_m128", 1, 8, 16, 128, 128, 1)
wgrad_case("w3x3", 1, 16, 16, 64, 64, 3)
os.makedirs("gpurun_out", exist_ok=True)
torch.save(out, "gpurun_out/debug_gemm.pt")
