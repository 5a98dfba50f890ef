"""GroupNorm-backward probe: the [256,64,64,512] layers of the B=256 train step (fp32 x + residual-branch gradient;
bf16 x without), two-pass kernels vs the single-pass persistent kernel, CUDA-event timed (best of 5 after warm-up).
Under ncu: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gn_bwd ...`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tempo_vae_b200 import _lib, ops  # noqa: E402

N, H, W, C, G = int(os.environ.get("PROBE_N", "256")), 64, 64, 512, 8
reps = int(os.environ.get("PROBE_REPS", "5"))
modes = [int(m) for m in os.environ.get("PROBE_MODES", "0,1").split(",")]
mbs = [int(m) for m in os.environ.get("PROBE_MB", "24").split(",")]
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((N, H, W, C), device=dev, generator=g)
gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
da = torch.randn((N, H, W, C), device=dev, generator=g).to(torch.bfloat16)
gres = torch.randn((N, H, W, C), device=dev, generator=g).to(torch.bfloat16)
stats = ops.gn_stats(x, C, G, 1e-6)
xb = x.to(torch.bfloat16)
dg, db, cs = (torch.empty(C, device=dev) for _ in range(3))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, xin, gr in (("fp32 x + gres", x, gres), ("bf16 x, no gres", xb, None)):
    nbytes = x.numel() * (xin.element_size() + 2 + (2 if gr is not None else 0) + 2)
    for mode in modes:
        for mb in (mbs if mode else [0]):
            _lib.lib.tvae_gn_set_bwd_fused(mode, mb)
            best = 1e9
            for i in range(reps + 2):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.gn_act_bwd(xin, stats, gamma, beta, da, gr, G, 1, dg, db, cs)
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    best = min(best, e0.elapsed_time(e1))
            print(f"{name:18s} mode {'single-pass' if mode else 'two-pass   '} group_mb {mb:3d}: {best:.3f} ms, "
                  f"{nbytes / best / 1e6:.0f} GB/s of algorithmic bytes ({nbytes / 1e9:.2f} GB)", flush=True)
_lib.lib.tvae_gn_set_bwd_fused(0, 24)
