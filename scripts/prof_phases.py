"""In-kernel phase profile of the fused MLP kernel (cycles averaged per CTA)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import consistentnerf_b200 as cn
from consistentnerf_b200 import _lib
import bench
dev = "cuda"
net = bench.make_nets(torch.device(dev))[0]
packed = net.packed_weights(); packed.refresh(dict(zip(net.spec.param_names(), [p.detach() for p in net.hot_params()])))
n, S = 4096, 192
pts = torch.randn(n, S, 3, device=dev); vd = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)
names = {0: "mma total", 1: "mma wait A kblocks", 2: "mma wait enc", 3: "mma wait weights", 4: "mma issue+commit", 8: "epilogue total", 9: "epilogue wait D"}
IMPL = os.environ.get("CNERF_MLP_IMPL", "3")
PROF = "cnerf_debug_profile3" if IMPL == "3" else "cnerf_debug_profile4"
for mode, dbg in ([("infer", 0), ("infer", 1)] if IMPL == "4" else [("infer", 0), ("train", 0)]):
    fn = (lambda: cn.ops.fused_mlp_forward(packed, pts, vd, fwd_terms=3)) if mode == "infer" else (lambda: cn.ops.fused_mlp_forward_train(packed, pts, vd, fwd_terms=3))
    for _ in range(2): fn()
    out = (ctypes.c_ulonglong * 16)()
    _lib.call(PROF, 1 | (dbg << 1), out)
    fn()
    _lib.call(PROF, 0, out)
    tiles = n * S / 128 / 148
    print("impl", IMPL, mode, "dbg", dbg, "128-point tiles per SM %.1f" % tiles)
    for k, nm in names.items():
        print(f"   {nm:20s} {out[k] / 148 / 1e3:10.1f} kcycles/CTA   {out[k] / 148 / tiles:10.0f} cycles/tile")

# ---- fp16 two-tile forward (mlp_fwd5.cu)
names5 = {0: "mma total", 1: "mma wait A operand", 3: "mma wait weights", 4: "mma issue+commit", 5: "mma record wait+layer commit",
          8: "epilogue total", 9: "epilogue wait D", 10: "epilogue tcgen05.ld+wait", 11: "epilogue body"}
for mode in ("infer", "train"):
    fn = (lambda: cn.ops.fused_mlp_forward(packed, pts, vd, fwd_terms=1)) if mode == "infer" else (lambda: cn.ops.fused_mlp_forward_train(packed, pts, vd, dw_terms=1, fwd_terms=1))
    for _ in range(2): fn()
    out = (ctypes.c_ulonglong * 16)()
    _lib.call("cnerf_debug_profile5", 1, out)
    fn()
    _lib.call("cnerf_debug_profile5", 0, out)
    tiles = n * S / 128 / 148
    print("fwd5", mode, "128-point tiles per SM %.1f" % tiles)
    for k, nm in names5.items():
        print(f"   {nm:30s} {out[k] / 148 / 1e3:10.1f} kcycles/CTA   {out[k] / 148 / tiles:10.0f} cycles/tile")

# ---- data-gradient chain kernel (training backward)
raw, acts = cn.ops.fused_mlp_forward_train(packed, pts, vd)
P = dict(zip(net.spec.param_names(), [p.detach() for p in net.hot_params()]))
d_raw = torch.randn(n * S, 4, device=dev) * 1e-6
for _ in range(2): cn.ops.fused_mlp_backward(packed, P, acts, d_raw, n * S)
out = (ctypes.c_ulonglong * 16)()
_lib.call("cnerf_debug_profile_chain", 1, out)
cn.ops.fused_mlp_backward(packed, P, acts, d_raw, n * S)
_lib.call("cnerf_debug_profile_chain", 0, out)
tiles = n * S / 128 / 148
print("chain3 128-point tiles per SM %.1f" % tiles)
for k, nm in {0: "mma total", 1: "mma wait A kblocks", 3: "mma wait weights", 4: "mma issue+commit", 8: "epilogue total", 9: "epilogue wait D", 10: "epilogue wait G0 stored"}.items():
    print(f"   {nm:24s} {out[k] / 148 / 1e3:10.1f} kcycles/CTA   {out[k] / 148 / tiles:10.0f} cycles/tile")
