"""Time tvae_wgrad_skinny at the train-step size (B=256, 64x64, 512 wide channels, 4 tail channels) against the padded
single weight-gradient GEMM and the 1024-channel main GEMM.  usage: python tools/wgrad_skinny_bench.py [B]"""
import sys
import torch
from tempo_vae_b200 import ops as o

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, 64, 64, 1032), device="cuda", generator=g, dtype=torch.float32).to(torch.bfloat16)
dy = torch.randn((B, 64, 64, 512), device="cuda", generator=g, dtype=torch.float32).to(torch.bfloat16)
grad = torch.zeros((512, 1028, 3, 3), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    t = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t += e0.elapsed_time(e1)
    return t / n


t_full = timed(lambda: o.wgrad_gemm(x[..., :1028], 1028, dy, 512, kind=0, R=3, grad=grad, flip=True))
t_main = timed(lambda: o.wgrad_gemm(x[..., :1024], 1024, dy, 512, kind=0, R=3, grad=grad, flip=True, grad_ld=1028, grad_off=0))
t_sk = timed(lambda: o.wgrad_skinny(dy, 512, x[..., 1024:1028], 4, sign=+1, grad=grad.view(-1)[1024 * 9:], stride_c=9,
                                    stride_n=9 * 1028))
t_sk2 = timed(lambda: o.wgrad_skinny(dy, 512, x[..., 1024:1028], 4, sign=-1, grad=grad.view(-1)[:4 * 512 * 9], stride_c=9 * 512,
                                     stride_n=9))
gb = dy.numel() * 2 / 1e9
print(f"B={B}: padded 1028 GEMM {t_full:.3f} ms | 1024 main {t_main:.3f} ms + skinny {t_sk:.3f} ms (sign -1: {t_sk2:.3f} ms) "
      f"= {t_main + t_sk:.3f} ms | skinny reads {gb:.2f} GB -> {gb / t_sk:.2f} TB/s")
