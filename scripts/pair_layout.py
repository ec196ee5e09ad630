"""Locates D(row, col) of an M=128 cta_group::2 accumulator in the two CTAs' tensor memory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from consistentnerf_b200 import _lib
for n in (256, 128, 32):
    k = 16
    a = torch.zeros(128, k); b = torch.zeros(n, k)
    a[:, 0] = torch.arange(128).float(); b[:, 0] = 1.0          # d[r, c] = r + (c + 1) / 512
    a[:, 1] = 1.0; b[:, 1] = (torch.arange(n) + 1.0) / 512.0
    dump = torch.full((2, 128, 256), -1.0, device="cuda")
    ad, bd = a.cuda(), b.cuda()          # keep the device copies alive across the call
    _lib.call("cnerf_debug_pair_layout", _lib.ptr(ad), _lib.ptr(bd), n, k, _lib.ptr(dump), _lib.stream())
    torch.cuda.synchronize()
    d = dump.cpu()
    # hypothesis: CTA c, lane l, column j  <->  row 64 c + (l % 64), col (l // 64) * n/2 + j   for j < n/2
    exp = torch.zeros(2, 128, 256)
    for c in range(2):
        for l in range(128):
            r = 64 * c + l % 64
            j = torch.arange(n // 2)
            exp[c, l, : n // 2] = r + ((l // 64) * (n // 2) + j + 1.0) / 512.0
    bad = (d != exp)
    print(f"N={n}: mismatches {int(bad.sum())} of {bad.numel()}")
    if bad.any():
        idx = bad.nonzero()[:12]
        for c, l, j in idx.tolist():
            print(f"   cta {c} lane {l} col {j}: got {d[c, l, j].item():.6f} expected {exp[c, l, j].item():.6f}")
