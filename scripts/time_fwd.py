"""Times the two forward kernels (three-term mlp_fwd3 / fp16 mlp_fwd5) alone, inference and training variants, on the fine pass
of workload A (4096 x 192 points) and the coarse pass (4096 x 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import consistentnerf_b200 as cn
import bench
dev = torch.device("cuda")
net = bench.make_nets(dev)[0]
packed = net.packed_weights(); packed.refresh(dict(zip(net.spec.param_names(), [p.detach() for p in net.hot_params()])))
for n, S in ((4096, 192), (4096, 64)):
    pts = torch.randn(n, S, 3, device=dev); vd = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)
    for name, fn in (("fwd3 infer", lambda: cn.ops.fused_mlp_forward(packed, pts, vd, fwd_terms=3)),
                     ("fwd5 infer", lambda: cn.ops.fused_mlp_forward(packed, pts, vd, fwd_terms=1)),
                     ("fwd3 train dw3", lambda: cn.ops.fused_mlp_forward_train(packed, pts, vd, dw_terms=3, fwd_terms=3)),
                     ("fwd3 train dw1", lambda: cn.ops.fused_mlp_forward_train(packed, pts, vd, dw_terms=1, fwd_terms=3)),
                     ("fwd5 train dw1", lambda: cn.ops.fused_mlp_forward_train(packed, pts, vd, dw_terms=1, fwd_terms=1))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tf = bench.FLOP_PER_POINT * n * S / (ms * 1e-3) / 1e12
        print(f"{n}x{S}  {name:16s} {ms:7.3f} ms   {tf:7.1f} algorithmic TFLOP/s")
