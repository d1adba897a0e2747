"""C-ABI checks that need no GPU: the library loads, exports every symbol
include/evcont_b200.h declares, the host-side K0 entry points (strings, addresses,
link tables) are bit-exact against the oracle, and compute entry points fail
loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from evcont_b200 import _lib, cistring
from oracle import cistring as ocs


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "evcont_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(evc_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    handle = C.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert _lib.lib().evc_abi_version() == 2


@pytest.mark.parametrize("norb,nocc", [(1, 0), (1, 1), (2, 1), (4, 2), (5, 3), (6, 3), (8, 4),
                                       (10, 5), (13, 5), (12, 1), (9, 9)])
def test_link_tables_bit_exact(norb, nocc):
    strs = cistring.make_strings(range(norb), nocc)
    assert strs.dtype == np.int64 and np.array_equal(strs, ocs.make_strings(norb, nocc))
    tab = cistring.gen_linkstr_index(range(norb), nocc)
    ref = ocs.gen_linkstr_index(norb, nocc)
    assert tab.dtype == np.int32 and tab.shape == ref.shape
    assert np.array_equal(tab, ref)
    for k in (0, len(strs) // 2, len(strs) - 1):
        assert cistring.str2addr(norb, nocc, int(strs[k])) == k
        assert cistring.addr2str(norb, nocc, k) == int(strs[k])


def test_link_golden_rows():
    g = np.load(os.path.join(GOLDEN, "linkindex_n4_k2.npz"))
    assert np.array_equal(cistring.make_strings(range(4), 2), g["strings"])
    assert np.array_equal(cistring.gen_linkstr_index(range(4), 2)[0], g["row0"])


def test_packed_link_records():
    tab = cistring.gen_linkstr_index(range(6), 3)
    nstr, nlink = tab.shape[:2]
    for link_major in (0, 1):
        out = np.empty(nstr * nlink, dtype=np.uint64)
        _lib.check(_lib.lib().evc_linkindex_pack_host(nstr, nlink, tab.ctypes.data, link_major,
                                                      out.ctypes.data))
        out = out.reshape((nlink, nstr) if link_major else (nstr, nlink))
        if link_major:
            out = out.T
        assert np.array_equal((out & 0xFFFFFFFF).astype(np.int64), tab[:, :, 2])
        assert np.array_equal(((out >> 32) & 0xFF).astype(np.int64), tab[:, :, 0])
        assert np.array_equal(((out >> 40) & 0xFF).astype(np.int64), tab[:, :, 1])
        assert np.array_equal(((out >> 48) & 0xFF).astype(np.uint8).view(np.int8), tab[:, :, 3])


def test_bad_arguments_are_errors():
    assert _lib.lib().evc_num_strings(4, 5) == -1
    with pytest.raises(ValueError):
        cistring.str2addr(4, 2, 0b0111)
    with pytest.raises(ValueError):
        cistring.addr2str(4, 2, 6)
    with pytest.raises(NotImplementedError):
        cistring.make_strings([0, 2, 3], 2)


def test_argument_checks_of_the_training_side_entry_points():
    """NULL handles / bad sizes are rejected before any device work, with a message in evc_last_error; the
    workspace queries are pure host arithmetic."""
    lib = _lib.lib()
    nbytes = C.c_size_t()
    assert lib.evc_transform_ci_workspace_bytes(10, 252, 252, C.byref(nbytes)) == 0
    # Ta^T (252^2) + Tb + two (252 x 252) work matrices, plus GEMM scratch
    assert nbytes.value >= 4 * 252 * 252 * 8
    assert lib.evc_transform_ci_workspace_bytes(13, 1287, 1287, C.byref(nbytes)) == 0
    assert nbytes.value >= (1287 * 1287 + 3 * 1287 * 1288) * 8      # odd row length padded to even
    assert lib.evc_transform_ci_workspace_bytes(0, 1, 1, C.byref(nbytes)) != 0
    assert lib.evc_transform_ci(None, 4, 2, 2, 6, 6, None, None, None, None, None, None, 0) != 0
    assert b"evc_transform_ci" in lib.evc_last_error()
    assert lib.evc_fock_rhf(None, 4, None, None, None, None) != 0
    assert b"evc_fock_rhf" in lib.evc_last_error()
    assert lib.evc_min_sqdist(None, 1, 1, 4, 20, None, None, None) != 0
    assert b"evc_min_sqdist" in lib.evc_last_error()
    assert lib.evc_fci_contract_workspace_bytes(10, 252, 252, C.byref(nbytes)) == 0
    assert nbytes.value >= 2 * 63504 * 100 * 8


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    handle = C.c_void_p()
    rc = _lib.lib().evc_ctx_create(0, None, C.byref(handle))
    assert rc != 0 and b"no CPU fallback" in _lib.lib().evc_last_error()
    from evcont_b200.engine import get_engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        get_engine()
    from evcont_b200.fci import B200FCISolver
    with pytest.raises(RuntimeError):
        B200FCISolver().trans_rdm12(np.eye(2), np.eye(2), 2, (1, 1))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "evcont_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), name


def test_trans_rdm_workspace_plan_is_a_pure_function_of_the_sizes():
    """The alpha-slice plan (host code, no GPU) fixes the summation order of every pair: it must depend on the sizes
    and the SM count only -- not on which kernel form (fused / producer-consumer, ``EVC_TRDM_PIPE``) runs it.
    H10 sizes on 148 SMs: 210 pairs x 9 alpha slices x 28 macro-blocks x 256 doubles."""
    lib = _lib.lib()
    out = C.c_size_t(0)
    assert lib.evc_trans_rdm12_workspace_bytes(10, 252, 252, 210, 148, C.byref(out)) == 0
    assert out.value == 210 * 9 * 28 * 256 * 8
    h10 = out.value
    import subprocess
    import sys
    code = ("import ctypes as C; from evcont_b200 import _lib; o = C.c_size_t(0); "
            "assert _lib.lib().evc_trans_rdm12_workspace_bytes(10, 252, 252, 210, 148, C.byref(o)) == 0; print(o.value)")
    for pipe in ("0", "1", "2"):
        env = dict(os.environ, EVC_TRDM_PIPE=pipe, PYTHONPATH=ROOT)
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=ROOT)
        assert res.returncode == 0, res.stderr
        assert int(res.stdout.strip()) == h10
    # a share of the pairs planned for the whole build is smaller only by the pair count
    assert lib.evc_trans_rdm12_workspace_bytes(13, 1287, 1287, 10, 148, C.byref(out)) == 0
    assert out.value % (10 * 66 * 256 * 8) == 0
    assert lib.evc_trans_rdm12_workspace_bytes(14, 10, 10, 1, 148, C.byref(out)) != 0   # norb > 13: refused
