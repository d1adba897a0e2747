"""Occupation strings and link tables, bit-compatible with ``pyscf.fci.cistring``.

Host-side wrappers over K0 of the C ABI (``evc_make_strings_host``,
``evc_str2addr``, ``evc_linkindex_build_host``).  The reference reaches these
through ``cisolver.trans_rdm12`` (evcont/FCI_EVCont.py:121), where PySCF rebuilds
the table on every call; here it is built once per ``(norb, nocc)`` and cached.
Signatures follow PySCF: ``orb_list`` is ``range(norb)`` (any other orbital
list is rejected -- the FCI path never uses one).
"""
import functools

import numpy as np

from ._lib import check, lib


def _norb_of(orb_list):
    if isinstance(orb_list, (int, np.integer)):
        return int(orb_list)
    orb_list = list(orb_list)
    if orb_list != list(range(len(orb_list))):
        raise NotImplementedError("only orb_list == range(norb) is supported")
    return len(orb_list)


def num_strings(norb, nelec):
    n = lib().evc_num_strings(int(norb), int(nelec))
    if n < 0:
        raise ValueError(f"invalid (norb, nelec) = ({norb}, {nelec})")
    return int(n)


def make_strings(orb_list, nelec):
    norb = _norb_of(orb_list)
    out = np.empty(num_strings(norb, nelec), dtype=np.int64)
    check(lib().evc_make_strings_host(norb, int(nelec), out.ctypes.data))
    return out


def str2addr(norb, nelec, string):
    addr = lib().evc_str2addr(int(norb), int(nelec), int(string))
    if addr < 0:
        raise ValueError(f"{string:#b} is not a {nelec}-electron string over {norb} orbitals")
    return int(addr)


def addr2str(norb, nelec, addr):
    s = lib().evc_addr2str(int(norb), int(nelec), int(addr))
    if s < 0:
        raise ValueError(f"address {addr} out of range for ({norb}, {nelec})")
    return int(s)


@functools.lru_cache(maxsize=32)
def _linkstr_cached(norb, nocc):
    nstr = num_strings(norb, nocc)
    nlink = lib().evc_num_links(norb, nocc)
    tab = np.empty((nstr, nlink, 4), dtype=np.int32)
    check(lib().evc_linkindex_build_host(norb, nocc, tab.ctypes.data))
    tab.setflags(write=False)
    return tab


def gen_linkstr_index(orb_list, nocc, strs=None):
    """``int32 (nstr, nlink, 4)`` rows ``[cre a, des i, addr, sign]`` (read-only, cached)."""
    if strs is not None:
        raise NotImplementedError("custom string lists are not supported")
    return _linkstr_cached(_norb_of(orb_list), int(nocc))
