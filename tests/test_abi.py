"""CPU checks of the C-ABI boundary: the library builds/loads without a GPU, exports every symbol that
include/vame_b200.h declares, and its host-only entry points (layouts, sizes, argument validation) behave."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vame_b200 import build, _lib
    build.build()
    return _lib.lib()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "vame_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vame_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in include/vame_b200.h is not exported" % n


def test_signature_table_covers_header(lib):
    from vame_b200 import _lib
    assert sorted(_lib.SIGNATURES.keys()) == declared_symbols()
    assert lib.vame_abi_version() == 1


def test_param_layout_matches_reference_counts(lib):
    from vame_b200.engine import Engine
    e = Engine(24, 30, 30, 256, 256, 256, True, 15)
    assert e.n_flat == 2618476 and len(e.names) == 44
    e2 = Engine(24, 30, 30, 256, 256, 256, False, 0)
    assert e2.n_flat == 2147924 and len(e2.names) == 32
    # tensors do not overlap and stay inside the flat buffer
    spans = sorted(zip(e.offsets, e.sizes))
    for (o1, s1), (o2, _) in zip(spans, spans[1:]):
        assert o1 + s1 <= o2
    assert spans[-1][0] + spans[-1][1] <= e.n_flat
    # the two halves of the fused [mu; logvar] projection are adjacent (one [2Z, 4H] GEMM operand)
    i = e.names.index("lmbda.hidden_to_mean.weight")
    j = e.names.index("lmbda.hidden_to_logvar.weight")
    assert e.offsets[j] == e.offsets[i] + e.sizes[i]


def test_sizes_and_validation(lib):
    from vame_b200.engine import VameDims
    d = VameDims(24, 30, 30, 256, 256, 256, 1, 15, 0)
    assert lib.vame_workspace_bytes(ctypes.byref(d), 256, 1) > lib.vame_workspace_bytes(ctypes.byref(d), 256, 0) > 0
    assert lib.vame_packed_weights_bytes(ctypes.byref(d)) > 0
    assert lib.vame_embed_workspace_bytes(ctypes.byref(d), 100000, 4096) > 0
    assert lib.vame_p16_bytes(128, 64, 128) == 128 * 64 * 4
    bad = VameDims(24, 30, 30, 100, 256, 256, 1, 15, 0)        # hidden size not a multiple of 32
    assert lib.vame_workspace_bytes(ctypes.byref(bad), 256, 1) == 0
    assert b"multiple of 32" in lib.vame_last_error()
    bad2 = VameDims(24, 30, 80, 256, 256, 256, 1, 15, 0)       # zdims > 64
    assert lib.vame_param_layout(ctypes.byref(bad2), None, None) < 0
    # null pointers are rejected before any launch
    assert lib.vame_pack_p16(None, 0, 0, 8, 8, 8, 8, None, None, 128, None, None) != 0
    assert b"null" in lib.vame_last_error()


def test_no_cpu_fallback():
    """The product path must fail loudly without a CUDA device instead of computing on the CPU."""
    import torch
    from vame_b200._lib import VameB200Error
    from vame_b200.rnn_model import RNN_VAE
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = RNN_VAE(60, 30, 24, True, 15, 256, 256, 256, 256, 0, 0, 0, False)
    with pytest.raises(VameB200Error):
        m(torch.zeros(2, 30, 24))
    with pytest.raises(VameB200Error):
        m.encoder(torch.zeros(2, 30, 24))
    with pytest.raises((VameB200Error, RuntimeError, AssertionError)):
        m.cuda()
