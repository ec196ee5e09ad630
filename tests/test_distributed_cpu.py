"""Host-side multi-GPU logic on CPU: world-size-2 gloo processes exercise ray sharding, the flat-gradient
all-reduce, the mask OR-reduce and the row gather (consistentnerf_b200/distributed.py).  No CUDA needed."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from consistentnerf_b200.distributed import FlatGrads, allreduce_masks, gather_rows, global_mask_counts, shard_bounds, shard_rays
    torch.manual_seed(0)                                   # identical "model" and global batch on every rank
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    rays = torch.randn(101, 6)                             # ragged: 51 + 50
    tgt = torch.randn(101, 3)
    flat = FlatGrads(net.parameters(), extra_slots=2)
    lo, hi = shard_bounds(101, rank, world)
    mine = shard_rays(rays, rank, world)
    assert mine.shape[0] == hi - lo
    flat.zero_()
    loss_sum = ((net(mine) - tgt[lo:hi]) ** 2).sum()       # SUM over the shard; mean is formed after the reduce
    loss_sum.backward()
    flat.extra[0] = float(hi - lo)                         # ray count travels in the same buffer
    flat.allreduce()
    flat.flat[:-2] /= (flat.extra[0] * 3)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    ref.load_state_dict(net.state_dict())
    ((ref(rays) - tgt) ** 2).mean().backward()
    for p, q in zip(net.parameters(), ref.parameters()):
        assert p.grad.data_ptr() >= flat.flat.data_ptr()   # .grad aliases the flat buffer: no packing copies
        torch.testing.assert_close(p.grad, q.grad, rtol=1e-5, atol=1e-6)
    assert float(flat.extra[0]) == 101.0
    # zero_grad(set_to_none=True) detaches .grad from the flat buffer: the next reduce re-attaches and keeps the values
    net.zero_grad(set_to_none=True)
    assert all(p.grad is None for p in net.parameters())
    flat.zero_()
    ((net(mine) - tgt[lo:hi]) ** 2).sum().backward()       # autograd makes fresh .grad tensors
    flat.allreduce()
    for p, q in zip(net.parameters(), ref.parameters()):
        assert flat.flat.data_ptr() <= p.grad.data_ptr() < flat.flat.data_ptr() + flat.flat.numel() * 4
        torch.testing.assert_close(p.grad / (101 * 3), q.grad, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError, match="async_op"):
        flat.allreduce(average=True, async_op=True)
    # two parameter groups (coarse / fine): segment-wise asynchronous reduction == whole-buffer reduction
    a, b = torch.nn.Linear(4, 4), torch.nn.Linear(4, 2)
    fg = FlatGrads([list(a.parameters()), list(b.parameters())], extra_slots=2)
    fg.flat.copy_(torch.arange(fg.flat.numel(), dtype=torch.float32) * (rank + 1))
    fg.allreduce_group(1)                                  # "fine" first, as autograd orders it
    fg.allreduce_group(0)
    fg.wait()
    torch.testing.assert_close(fg.flat, torch.arange(fg.flat.numel(), dtype=torch.float32) * 3)
    assert all(getattr(p, "_cnerf_accumulate_in_place", False) for p in fg.params)
    # global denominators of the masked losses: shard-wise counts sum to the single-process counts
    gmask = (torch.arange(101) % 3 != 0).float()
    c = global_mask_counts(gmask[lo:hi], hi - lo)
    assert c.tolist() == [float((gmask == 1).sum()), float((gmask == 0).sum()), float(gmask.sum()), 101.0]
    c = global_mask_counts(None, hi - lo, device="cpu")
    assert c.tolist() == [101.0, 0.0, 101.0, 101.0]
    # hard-mask partials OR-ed across ranks
    part = {7: torch.zeros(10, dtype=torch.uint8), 3: torch.zeros(10, dtype=torch.uint8)}
    part[7][rank] = 1
    part[3][5 + rank] = 1
    full = allreduce_masks(part)
    assert full[7].tolist() == [True, True] + [False] * 8 and full[3][5:7].all() and int(full[3].sum()) == 2
    # rendered row tiles gathered back in order
    rows = torch.arange(101, dtype=torch.float32)[:, None].repeat(1, 3)
    got = gather_rows(rows[lo:hi], 101)
    assert torch.equal(got, rows)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_world_size_2_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_shard_bounds_cover_and_balance():
    sys.path.insert(0, ROOT)
    from consistentnerf_b200.distributed import shard_bounds
    for n in (0, 1, 7, 4096, 640000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_dropin_patches_reference_globals():
    """dropin.patch swaps exactly the hot-path globals of a script module (module-global lookup)."""
    import types
    sys.path.insert(0, ROOT)
    from consistentnerf_b200 import dropin
    names = ["batchify", "run_network", "batchify_rays", "render", "raw2outputs", "render_rays", "NeRF", "get_embedder",
             "sample_pdf", "get_rays", "ndc_rays", "get_rays_ref", "get_ref_rays", "get_test_label", "train", "config_parser"]
    view = types.ModuleType("run_nerf_view")
    for n in names:
        setattr(view, n, object())
    keep = view.train
    done = dropin.patch(view)
    assert set(done) == set(names) - {"train", "config_parser"} and view.train is keep
    import consistentnerf_b200 as cn
    assert view.NeRF is cn.NeRF and view.get_ref_rays is cn.get_ref_rays
    vanilla = types.ModuleType("run_nerf")
    for n in names[:11]:
        setattr(vanilla, n, object())
    dropin.patch(vanilla)
    assert dropin.wants_depth(view) and not dropin.wants_depth(vanilla)
