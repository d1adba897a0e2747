"""Occupation strings and single-excitation link tables (CPU oracle).

Restates the algorithm of ``pyscf.fci.cistring`` (``make_strings``,
``str2addr``, ``gen_linkstr_index``) which the reference reaches through
``cisolver.trans_rdm12`` at evcont/FCI_EVCont.py:121 (link_index=None, so the
table is regenerated on every call).  Layout facts follow SURVEY.md
Appendix A.2:

* strings are integers, bit ``i`` set = orbital ``i`` occupied, listed in
  ascending integer order; the address of a string is its rank in that list;
* the link table is ``int32 (nstr, nlink, 4)`` with rows
  ``[cre a, des i, address of a^+ i |str>, sign]``; the ``nocc`` diagonal rows
  ``(o, o, self, +1)`` come first (``o`` ascending), then for every occupied
  ``i`` (ascending, outer) and every virtual ``a`` (ascending, inner) the
  excitation ``i -> a``.

Pure Python / numpy; test infrastructure only (see oracle/__init__.py).
"""
from math import comb

import numpy as np


def num_strings(norb, nocc):
    return comb(norb, nocc)


def make_strings(norb, nocc):
    """All ``nocc``-electron strings over ``norb`` orbitals, ascending (int64)."""
    if nocc == 0:
        return np.zeros(1, dtype=np.int64)
    out = []
    # Gosper's hack enumerates fixed-popcount integers in ascending order.
    s = (1 << nocc) - 1
    limit = 1 << norb
    while s < limit:
        out.append(s)
        c = s & -s
        r = s + c
        s = (((r ^ s) >> 2) // c) | r
    return np.asarray(out, dtype=np.int64)


def str2addr(norb, nocc, string):
    """Rank of ``string`` in ``make_strings(norb, nocc)``: sum_j C(o_j, j)."""
    addr = 0
    j = 0
    for o in range(norb):
        if (string >> o) & 1:
            j += 1
            addr += comb(o, j)
    return addr


def addr2str(norb, nocc, addr):
    """Inverse of :func:`str2addr`."""
    s = 0
    k = nocc
    for o in range(norb - 1, -1, -1):
        if k == 0:
            break
        c = comb(o, k)
        if addr >= c:
            s |= 1 << o
            addr -= c
            k -= 1
    return s


def gen_linkstr_index(norb, nocc):
    """``int32 (nstr, nlink, 4)`` table of ``[a, i, addr(a^+ i str), sign]``."""
    strs = make_strings(norb, nocc)
    nvir = norb - nocc
    nlink = nocc + nocc * nvir
    tab = np.zeros((len(strs), nlink, 4), dtype=np.int32)
    for k, s0 in enumerate(int(s) for s in strs):
        occ = [o for o in range(norb) if (s0 >> o) & 1]
        vir = [o for o in range(norb) if not (s0 >> o) & 1]
        row = 0
        for o in occ:
            tab[k, row] = (o, o, k, 1)
            row += 1
        for i in occ:
            for a in vir:
                s1 = (s0 ^ (1 << i)) | (1 << a)
                lo, hi = (i, a) if i < a else (a, i)
                between = s0 & ((1 << hi) - (1 << (lo + 1)))
                sign = -1 if bin(between).count("1") & 1 else 1
                tab[k, row] = (a, i, str2addr(norb, nocc, s1), sign)
                row += 1
    return tab
