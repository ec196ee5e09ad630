import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from consistentnerf_b200 import _lib
out = torch.zeros(148, device="cuda")
for mode, n in ((1, 128), (0, 128), (1, 256)):
    for alt in (0, 2, 4, 8, 2 | 4, 2 | 4 | 8, 2 | 4 | 8 | 16):
        _lib.call("cnerf_debug_umma_rate", mode, n, 4000, alt, _lib.ptr(out), _lib.stream())
        torch.cuda.synchronize()
        tags = "+".join(t for b, t in ((2, "trywait"), (4, "fence"), (8, "commit"), (16, "commit2")) if alt & b) or "bare"
        print(f"mode {'SS' if mode == 0 else 'TS'} N={n:3d} per-4-MMA {tags:28s}: {out.mean().item():7.1f} cycles/MMA  ideal {n/2:.0f}")
