"""Times cnerf_mlp_bwd_heads alone (both libraries when run twice with CNERF_LIB)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, consistentnerf_b200 as cn
from consistentnerf_b200 import _lib
from consistentnerf_b200.ops import _workspace
dev = torch.device("cuda", 0)
net = bench.make_nets(dev)[1]
packed = net.packed_weights(); P = dict(zip(net.spec.param_names(), [p.detach() for p in net.hot_params()])); packed.refresh(P)
n, S = 4096, 192
pts = torch.randn(n, S, 3, device=dev); vd = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)
raw, acts = cn.ops.fused_mlp_forward_train(packed, pts, vd)
d_raw = torch.randn(n * S, 4, device=dev) * 1e-6
g = {k: torch.zeros_like(v) for k, v in P.items()}
ws = _workspace(dev, int(_lib.load().cnerf_mlp_bwd_workspace_bytes()))
def run():
    _lib.call("cnerf_mlp_bwd_heads", _lib.ptr(d_raw), _lib.ptr(acts), n * S, _lib.ptr(g["alpha_linear.weight"]), _lib.ptr(g["alpha_linear.bias"]),
              _lib.ptr(g["rgb_linear.weight"]), _lib.ptr(g["rgb_linear.bias"]), 0, _lib.ptr(ws), _lib.stream())
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print(os.environ.get("CNERF_LIB", "current"), "heads (fine pass, 786432 points): %.3f ms" % (e0.elapsed_time(e1) / 20), "checksum %.6e" % float(g["rgb_linear.weight"].double().abs().sum()))
