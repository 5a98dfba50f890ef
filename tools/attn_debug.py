import ctypes, torch
from tempo_vae_b200 import ops as o
torch.manual_seed(0)
B, T, heads = 1, 128, 1
C = 32 * heads
qkv = torch.randn((B * T, 3 * C), device="cuda")
dbg = torch.zeros(128 * 64 + 128 * 32 + 128 * 32, device="cuda")
print("dbg rc", o.lib.tvae_attn_debug_buffer(ctypes.c_void_p(dbg.data_ptr())))
ob, of, lse = o.attn_fwd(qkv, C, heads, B, T)
torch.cuda.synchronize()
q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
L2E = 1.4426950408889634
S = (q * (32 ** -0.5) * L2E) @ k.t()          # [128 q, 128 k]
Sd = dbg[:128 * 64].view(128, 64)
print("S err", (Sd - S[:, :64]).abs().max().item(), "S ref max", S.abs().max().item())
print("S dump row0[:8]", Sd[0, :8].tolist()); print("S ref  row0[:8]", S[0, :8].tolist())
print("S dump row1[:8]", Sd[1, :8].tolist()); print("S ref  row1[:8]", S[1, :8].tolist())
print("S dump row37[:8]", Sd[37, :8].tolist()); print("S ref  row37[:8]", S[37, :8].tolist())
m = S[:, :64].max(dim=1, keepdim=True).values
P = torch.exp2(S[:, :64] - m)
Ot = P @ v[:64]
Od = dbg[128 * 64:128 * 64 + 128 * 32].view(128, 32)
Pd = dbg[128 * 64 + 128 * 32:].view(128, 32)
print("P(readback) err", (Pd - P[:, :32]).abs().max().item())
print("Otile err", (Od - Ot).abs().max().item(), "ref max", Ot.abs().max().item())
print("O dump row0[:8]", Od[0, :8].tolist()); print("O ref  row0[:8]", Ot[0, :8].tolist())
# try to identify a permutation: which ref (row, col) does dump[0, j] equal?
for j in range(4):
    val = Od[0, j]
    hit = ((Ot - val).abs() < 1e-3 * Ot.abs().max()).nonzero()
    print("O dump[0,%d]=%.4f matches ref at" % (j, val.item()), hit[:4].tolist())
w = torch.softmax((q @ k.t()) * 32 ** -0.5, dim=-1)
print("out err", (of - w @ v).abs().max().item(), "lse err", (lse.view(-1) - torch.logsumexp((q @ k.t()) * 32 ** -0.5, -1)).abs().max().item())
