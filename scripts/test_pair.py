"""CTA-pair fp16 forward (csrc/mlp_fwd6.cu, CNERF_FWD_PAIR=1) against the single-CTA two-tile kernel and the fp64 oracle; timing.
Run as a script (the switch is read once per process):  CNERF_FWD_PAIR=1 python scripts/test_pair.py"""
import os, sys
assert os.environ.get("CNERF_FWD_PAIR") == "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import consistentnerf_b200 as cn
from consistentnerf_b200 import _lib
from oracle import nerf_oracle as O
from util import ARCH, module_from_params, rel_err
import bench

p = O.make_params(7, sigma_bias=0.3, **ARCH)
net = module_from_params(p, ARCH)
packed = net.packed_weights(); packed.refresh({k: v.detach() for k, v in zip(net.spec.param_names(), net.hot_params())})
worst = 0.0
for n, S in ((1, 1), (3, 64), (2, 128), (4, 64), (40, 192), (257, 33), (1200, 64), (700, 192)):
    g = torch.Generator().manual_seed(n * 100 + S)
    pts = torch.randn(n, S, 3, generator=g) * 2.0
    vd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    raw = cn.ops.fused_mlp_forward(packed, pts.cuda(), vd.cuda(), fwd_terms=1)             # pair kernel
    raw3 = cn.ops.fused_mlp_forward(packed, pts.cuda(), vd.cuda(), fwd_terms=3)            # three-term reference on the GPU
    torch.cuda.synchronize()
    e = rel_err(raw, raw3)
    again = cn.ops.fused_mlp_forward(packed, pts.cuda(), vd.cuda(), fwd_terms=1)
    assert torch.equal(raw, again), (n, S)
    if n * S <= 10000:
        ref = O._query({k: v.double() for k, v in p.items()}, ARCH, pts.double(), vd.double(), 10, 4)
        e = max(e, rel_err(raw, ref))
    print(f"{n}x{S}: rel err {e:.2e}", flush=True)
    worst = max(worst, e)
assert worst < 1e-3, worst
dev = torch.device("cuda")
net2 = bench.make_nets(dev)[0]
pk = net2.packed_weights(); pk.refresh(dict(zip(net2.spec.param_names(), [q.detach() for q in net2.hot_params()])))
for n, S in ((4096, 192), (4096, 64)):
    pts = torch.randn(n, S, 3, device=dev); vd = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)
    fn = lambda: cn.ops.fused_mlp_forward(pk, pts, vd, fwd_terms=1)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{n}x{S}  pair fwd6 infer {ms:7.3f} ms   {bench.FLOP_PER_POINT * n * S / (ms * 1e-3) / 1e12:7.1f} algorithmic TFLOP/s")
print("PAIR OK")
