"""Does the conv kernel's tensor-pipe utilisation depend on the number of K blocks per tile (tile switches)?
Three bf16-output launches with 36 / 72 / 144 K blocks per tile and the same FLOP count; run under
`ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum -k regex:conv_gemm`."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tempo_vae_b200 import ops as o
g = torch.Generator(device="cuda").manual_seed(0)
for B, Cin in ((512, 256), (256, 512), (128, 1024)):
    x = torch.randn((B, 64, 64, Cin), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn((512, Cin, 3, 3), device="cuda", generator=g) / math.sqrt(9 * Cin)
    wp = o.pack_weight(w, "fwd")
    o.conv_gemm(x, Cin, wp, kind=0, R=3, Cout=512, want_f32=False, want_bf16=True)
    torch.cuda.synchronize()
    del x
print("ok")
