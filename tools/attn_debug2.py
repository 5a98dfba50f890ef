import ctypes, torch
from tempo_vae_b200 import ops as o
torch.manual_seed(0)
A = torch.randn(128, 64, device="cuda")
B = torch.randn(64, 32, device="cuda")
ref = A @ B
vp = ctypes.c_void_p
for mode, lbo, sbo in [(0, 16, 1024), (1, 16, 1024), (2, 1024, 1024), (2, 16, 1024), (2, 128, 1024), (2, 1024, 128), (2, 4096, 1024),
                       (3, 1024, 1024), (6, 1024, 512), (6, 16, 512), (6, 512, 1024), (6, 128, 512), (7, 1024, 512), (7, 16, 512)]:
    D = torch.full((128, 32), 7.0, device="cuda")
    rc = o.lib.tvae_attn_debug_mma(vp(A.data_ptr()), vp(B.data_ptr()), vp(D.data_ptr()), mode, lbo, sbo)
    err = (D - ref).abs().max().item()
    print(f"mode={mode} (A {'TMEM' if mode & 1 else 'smem'}, B {'MN' if mode & 2 else 'K'}-major) lbo={lbo} sbo={sbo} rc={rc} "
          f"err={err:.4f} refmax={ref.abs().max().item():.2f} D[0,:4]={D[0,:4].tolist()} ref[0,:4]={ref[0,:4].tolist()}")
