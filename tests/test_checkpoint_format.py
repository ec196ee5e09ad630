"""On-disk format the callers expect: the reference's checkpoint is a torch-saved dict
{'global_step', 'network_fn_state_dict', 'network_fine_state_dict', 'optimizer_state_dict'} (NP/run_nerf_view.py:2004-2015),
re-loaded by create_nerf (NP/run_nerf_view.py:340-352).  The product's NeRF module must carry exactly the reference's
state-dict keys, order and shapes (golden list taken from the unmodified reference's NeRF), so that such checkpoints load
in either direction.  CPU only: no kernel runs."""
import io

import torch

from conftest import load_golden
from oracle import nerf_oracle as O
from util import ARCH, SMALL


def _net(cn, arch):
    return cn.NeRF(D=arch["D"], W=arch["W"], input_ch=arch["input_ch"], input_ch_views=arch["input_ch_views"],
                   output_ch=arch["output_ch"], skips=list(arch["skips"]), use_viewdirs=arch["use_viewdirs"])


def test_state_dict_keys_order_and_shapes_match_the_reference():
    import consistentnerf_b200 as cn
    for tag, arch in (("mlp_viewdirs", ARCH), ("mlp_small_noview", SMALL)):
        ref_keys = [str(k) for k in load_golden(tag)["keys"]]
        sd = _net(cn, arch).state_dict()
        assert list(sd.keys()) == ref_keys
        shapes = dict(O.nerf_param_shapes(**arch))
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(shapes[k]), k


def test_reference_format_checkpoint_round_trip():
    import consistentnerf_b200 as cn
    coarse, fine = _net(cn, ARCH), _net(cn, ARCH)
    params = list(coarse.parameters()) + list(fine.parameters())
    opt = torch.optim.Adam(params, lr=5e-4, betas=(0.9, 0.999))
    ckpt = {"global_step": 1234, "network_fn_state_dict": coarse.state_dict(), "network_fine_state_dict": fine.state_dict(),
            "optimizer_state_dict": opt.state_dict()}
    buf = io.BytesIO()
    torch.save(ckpt, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    c2, f2 = _net(cn, ARCH), _net(cn, ARCH)
    c2.load_state_dict(back["network_fn_state_dict"])          # strict: any missing / unexpected key raises
    f2.load_state_dict(back["network_fine_state_dict"])
    torch.optim.Adam(list(c2.parameters()) + list(f2.parameters()), lr=5e-4).load_state_dict(back["optimizer_state_dict"])
    assert back["global_step"] == 1234
    for a, b in zip(coarse.state_dict().values(), c2.state_dict().values()):
        assert torch.equal(a, b)
    # a checkpoint written by the reference's NeRF holds the oracle's parameter dictionary under the same names
    ref_like = {k: v.clone() for k, v in O.make_params(3, **ARCH).items()}
    c2.load_state_dict(ref_like)
    assert torch.equal(c2.state_dict()["pts_linears.5.weight"], ref_like["pts_linears.5.weight"])
