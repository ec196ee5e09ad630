"""The C-ABI library loads and exports every symbol include/cnerf.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cnerf.h")
DEBUG_HEADER = os.path.join(ROOT, "include", "cnerf_debug.h")


def declared_functions(header=HEADER, experiments=False):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    if "#ifdef CNERF_EXPERIMENTS" in src:
        head, _, rest = src.partition("#ifdef CNERF_EXPERIMENTS")
        exp, _, tail = rest.partition("#endif")
        src = exp if experiments else head + tail
    elif experiments:
        src = ""
    return sorted(set(re.findall(r"\b(cnerf_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from consistentnerf_b200.build import build_library
    return build_library()


def test_header_declares_the_hot_path():
    names = declared_functions()
    for need in ("cnerf_stratified_z", "cnerf_posenc", "cnerf_mlp_fwd", "cnerf_composite_fwd", "cnerf_composite_bwd",
                 "cnerf_sample_pdf", "cnerf_sample_fine", "cnerf_project_gather", "cnerf_hard_mask_pair",
                 "cnerf_masked_mse_fwd", "cnerf_masked_mse_bwd", "cnerf_weights_refresh"):
        assert need in names


def test_library_exports_every_declared_symbol(lib_path):
    dll = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(dll, name), f"{name} declared in cnerf.h but not exported"
    dll.cnerf_version.restype = ctypes.c_int
    assert dll.cnerf_version() == 200
    for name in declared_functions(DEBUG_HEADER):
        assert hasattr(dll, name), f"{name} declared in cnerf_debug.h but not exported"


def test_ctypes_table_mirrors_header(lib_path):
    from consistentnerf_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()
    assert sorted(_lib.DEBUG_SIGNATURES) == declared_functions(DEBUG_HEADER)
    assert sorted(_lib.EXPERIMENT_SIGNATURES) == declared_functions(DEBUG_HEADER, experiments=True)
    assert not any(n.startswith(("cnerf_debug", "cnerf_umma")) for n in _lib.SIGNATURES)      # the boundary carries no debug export
    _lib.load()          # resolves every prototype


def test_argument_errors_do_not_touch_the_gpu(lib_path):
    from consistentnerf_b200 import _lib
    dll = _lib.load()
    rc = dll.cnerf_composite_fwd(None, None, None, 3, None, 4, 8, 0, None, None, None, None, None, None)
    assert rc == 1
    assert "null pointer" in _lib.last_error()
    with pytest.raises(RuntimeError, match="null pointer"):
        _lib.call("cnerf_sample_pdf", None, None, None, None, 1, 63, 128, None, None, None, None, None)


def test_product_has_no_oracle_or_cpu_fallback():
    pkg = os.path.join(ROOT, "consistentnerf_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in text and "from oracle" not in text, fn
    import torch
    import consistentnerf_b200 as cn
    with pytest.raises(RuntimeError, match="CUDA"):
        cn.ops.posenc(torch.zeros(4, 3), 10)
