"""tcgen05.mma issue-rate microbenchmarks (cycles per instruction, K = 16, fp16 operands in shared memory, all SMs busy).

What they established:
  * M=128 N=256 back to back runs at the pipe rate (128 cycles) and is not slowed by concurrent st.shared / bulk-copy traffic;
  * issue is synchronous with execution: whatever else the issuing warp does per group of four MMAs (an mbarrier try_wait,
    60 dependent integer ops) is ADDED to the time per MMA, it is not hidden behind the running instruction;
  * M=128 N=128 costs the same 128 cycles as N=256 (operand-A fetch bound) and a second issuing warp does not help;
  * a CTA-pair instruction (cta_group::2) runs at 128 cycles for M=256 and 64 cycles for M=128 (64 rows per CTA).
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from consistentnerf_b200 import _lib

out = torch.zeros(148, device="cuda")
TAGS = ((2, "try_wait"), (4, "fence"), (8, "commit"), (32, "60 ALU ops"), (64, "8 LDS"), (128, "two issuing warps"))
print("single CTA per SM, per group of 4 MMAs (cycles per MMA, mean over 148 SMs)")
for n, alts in ((256, (0, 2, 8, 32, 64)), (128, (0, 32, 128, 128 | 32))):
    for alt in alts:
        _lib.call("cnerf_debug_umma_rate", 0, n, 4000, alt, _lib.ptr(out), _lib.stream())
        torch.cuda.synchronize()
        tag = " + ".join(t for b, t in TAGS if alt & b) or "MMAs only"
        print(f"  M=128 N={n:3d}  {tag:32s}: {out.mean().item():7.1f}")
out2 = torch.zeros(148, device="cuda")
src = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
print("N=256, concurrent traffic from the other warps of the CTA (cycles per MMA)")
for pair, m, extra in ((0, 128, 0), (1, 256, 0), (1, 128, 4)):
    for traffic in (0, 1, 2, 3, 3 | 8):
        if (traffic & 8) and not pair:
            continue
        _lib.call("cnerf_debug_umma_rate_pair", pair, 8000, traffic | extra, _lib.ptr(src), _lib.ptr(out2), _lib.stream())
        torch.cuda.synchronize()
        n = 74 if pair else 148
        what = ", ".join(t for b, t in ((1, "st.shared stream"), (2, "bulk-copy ring"), (8, "3 MMAs + multicast commit per group")) if traffic & b) or "none"
        print(f"  {'cta_group::2' if pair else 'cta_group::1'} M={m:3d}  traffic: {what:60s}: {out2[:n].mean().item():7.1f}")
