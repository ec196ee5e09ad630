"""Two NCCL ranks on two GPUs: the ray batch is sharded, the masked losses use GLOBAL mask counts, the
fine network's gradient segment is reduced under the coarse network's backward -- and the reduced gradients equal the
single-GPU gradients of the whole batch (NP/run_nerf_view.py:1645-1648 semantics).  Skipped with fewer than two devices."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _loss(cn, nets, batch, n_glob, gc, far=6.0, coef=0.2):
    coarse, fine = nets
    e, _ = cn.get_embedder(10, 0)
    ev, _ = cn.get_embedder(4, 0)
    q = lambda i, v, f: cn.run_network(i, v, f, embed_fn=e, embeddirs_fn=ev)
    kw = dict(network_query_fn=q, perturb=0.0, N_importance=128, network_fine=fine, N_samples=64, network_fn=coarse, use_viewdirs=True,
              white_bkgd=True, raw_noise_std=0.0, ndc=False, lindisp=False, near=2.0, far=far)
    o, d, tgt, prior, mask = batch
    rgb, disp, acc, depth, ex = cn.render(1, o.shape[0], None, chunk=32768, rays=(o, d), retraw=True, **kw)
    return (cn.masked_img_loss(rgb, tgt, mask, coef, n_rand=n_glob, global_counts=gc)
            + cn.masked_img_loss(ex["rgb0"], tgt, mask, coef, n_rand=n_glob, global_counts=gc)
            + cn.masked_depth_loss(depth, prior, mask, far, coef, n_rand=n_glob, include_unmasked=True, global_counts=gc)
            + cn.masked_depth_loss(ex["depth0"], prior, mask, far, coef, n_rand=n_glob, include_unmasked=True, global_counts=gc))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import bench
    import consistentnerf_b200 as cn
    from consistentnerf_b200.distributed import FlatGrads, global_mask_counts, shard_bounds
    cn.ops.set_grad_precision("split")
    cn.ops.set_forward_precision("split")
    n = 512
    full = tuple(x.to(dev) for x in bench.make_batch(n, 5))
    full = full[:4] + ((torch.arange(n, device=dev)[:, None] % 5 != 0).float() * (torch.arange(n, device=dev)[:, None] < 300).float(),)   # unequal mask counts per shard
    lo, hi = shard_bounds(n, rank, world)
    mine = tuple(x[lo:hi].contiguous() for x in full)
    nets = bench.make_nets(dev)
    groups = [[p for k, p in net.named_parameters() if k in net.spec.param_names()] for net in nets]
    flat = FlatGrads(groups)
    flat.overlap_with_backward(list(nets))
    flat.zero_()
    gc = global_mask_counts(mine[4], hi - lo)
    assert gc.tolist() == [float((full[4] == 1).sum()), float((full[4] == 0).sum()), float(full[4].sum()), float(n)]
    loss = _loss(cn, nets, mine, n, gc)
    loss.backward()
    assert getattr(flat, "_fired", set()) == {0, 1}          # both segments were launched from inside backward (fine first)
    flat.finish()
    torch.cuda.synchronize()
    total = loss.detach().clone()
    dist.all_reduce(total)
    # single-GPU reference on the whole batch, same weights
    ref_nets = bench.make_nets(dev)
    ref_loss = _loss(cn, ref_nets, full, n, None)
    ref_loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for net in ref_nets for k, p in net.named_parameters() if k in net.spec.param_names()])
    got = flat.flat[:ref.numel()]
    err = float((got - ref).abs().max() / ref.abs().max())
    lerr = abs(float(total) - float(ref_loss)) / abs(float(ref_loss))
    assert err < 2e-5 and lerr < 1e-6, (err, lerr)
    # plain averaging of per-rank masked means (what round 1 did) is NOT the same thing when the mask counts differ
    naive = _loss(cn, ref_nets, mine, hi - lo, None).detach()
    dist.all_reduce(naive)
    assert abs(float(naive) / world - float(ref_loss)) / abs(float(ref_loss)) > 1e-4
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write(f"{err:.3e} {lerr:.3e}")


def test_two_rank_nccl_gradients_equal_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
    print({f: open(tmp_path / f).read() for f in os.listdir(tmp_path)})
