import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from consistentnerf_b200 import _lib
out = torch.zeros(148, device="cuda")
for mode, n in ((0, 256), (0, 128)):
    for alt in ((0, 2, 32) if n == 256 else (0, 2, 32, 128, 128 | 2, 128 | 32, 128 | 2 | 32 | 8)):
        _lib.call("cnerf_debug_umma_rate", mode, n, 4000, alt, _lib.ptr(out), _lib.stream())
        torch.cuda.synchronize()
        tags = "+".join(t for b, t in ((2, "trywait"), (4, "fence"), (8, "commit"), (16, "commit2"), (32, "alu chain x60"), (64, "8 lds"), (128, "TWO issuers")) if alt & b) or "bare"
        print(f"mode {'SS' if mode == 0 else 'TS'} N={n:3d} per-4-MMA {tags:28s}: {out.mean().item():7.1f} cycles/MMA  ideal {n/2:.0f}")
out2 = torch.zeros(148, device="cuda")
src = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
for pair in (0, 1):
    for traffic in ((4,) if pair else (0,)):
        _lib.call("cnerf_debug_umma_rate_pair", pair, 8000, traffic, _lib.ptr(src), _lib.ptr(out2), _lib.stream())
        torch.cuda.synchronize()
        n = 74 if pair else 148
        print(f"{('pair M=128' if traffic & 4 else 'pair M=256') if pair else 'single M=128'} SS N=256 K=16 traffic(st.shared={traffic & 1}, bulk ring={(traffic >> 1) & 1}, commit per 4={(traffic >> 3) & 1}, random data={(traffic >> 4) & 1}, wait={(traffic >> 5) & 1}, fence={(traffic >> 6) & 1}): {out2[:n].mean().item():7.1f} cycles/MMA  ideal 128")
