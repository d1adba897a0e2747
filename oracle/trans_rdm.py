"""Spin-summed transition 1-/2-RDMs between two FCI vectors (CPU oracle).

Restates what ``cisolver.trans_rdm12(cibra, ciket, norb, nelec)`` returns at
the reference call site evcont/FCI_EVCont.py:121-123, following PySCF's chain
``direct_spin1.trans_rdm12 -> rdm.make_rdm12_spin1('FCItdm12kern_sf', ...)
-> rdm.reorder_rdm`` as described in SURVEY.md Appendix A (PySCF itself is not
available in this image: "parity unpinned" w.r.t. the PySCF binary; pinned by
:func:`brute_force_rdm12` and the invariants in tests/test_oracle_trans_rdm.py):

    t1_v[K, p*n+q] = <K| E_pq |v>,   E_pq = sum_sigma p^+_sigma q_sigma
    rdm1_C[p, q]   = sum_K bra[K] t1_ket[K, (p,q)]          = <bra|p^+ q|ket>
    rdm2_C[pq, rs] = sum_K t1_bra[K, (q,p)] t1_ket[K, (r,s)] = <bra|E_pq E_rs|ket>
    dm1 = rdm1_C^T ;  dm2 = rdm2_C ;  dm2[:, k, k, :] -= dm1^T      (reorder)

so that  dm1[p,q] = <bra|q^+ p|ket>,  dm2[p,q,r,s] = <bra|p^+ r^+ s q|ket>.

numpy, vectorised over strings and blocked over alpha strings so that H2O
6-31G (1287 x 1287 determinants) stays within a few hundred MB.
Test infrastructure only (see oracle/__init__.py).
"""
import itertools

import numpy as np

from . import cistring


def _unpack_nelec(nelec):
    if isinstance(nelec, (int, np.integer)):
        nb = int(nelec) // 2
        return int(nelec) - nb, nb
    return int(nelec[0]), int(nelec[1])


def build_t1(civec, norb, nelec, link_a=None, link_b=None, ia_range=None):
    """``t1[Ia, Ib, p*n+q] = <Ia Ib| E_pq |civec>`` for alpha strings in ``ia_range``."""
    neleca, nelecb = _unpack_nelec(nelec)
    if link_a is None:
        link_a = cistring.gen_linkstr_index(norb, neleca)
    if link_b is None:
        link_b = link_a if nelecb == neleca else cistring.gen_linkstr_index(norb, nelecb)
    na, nb = link_a.shape[0], link_b.shape[0]
    c = np.asarray(civec, dtype=np.float64).reshape(na, nb)
    lo, hi = (0, na) if ia_range is None else ia_range
    t1 = np.zeros((hi - lo, nb, norb * norb))
    rows_b = np.arange(nb)
    # beta links: t1[Ia, Ib, i*n+a] += sign * c[Ia, Jb]
    for l in range(link_b.shape[1]):
        a, i, j, s = (link_b[:, l, k] for k in range(4))
        t1[:, rows_b, i * norb + a] += c[lo:hi][:, j] * s
    # alpha links: t1[Ia, Ib, i*n+a] += sign * c[Ja, Ib]
    rows_a = np.arange(hi - lo)
    for l in range(link_a.shape[1]):
        a, i, j, s = (link_a[lo:hi, l, k] for k in range(4))
        t1[rows_a, :, i * norb + a] += c[j, :] * s[:, None]
    return t1


def trans_rdm12(cibra, ciket, norb, nelec, link_index=None, reorder=True, block=64):
    """(dm1, dm2) with PySCF's ``trans_rdm12`` conventions (see module docstring)."""
    neleca, nelecb = _unpack_nelec(nelec)
    if link_index is None:
        link_a = cistring.gen_linkstr_index(norb, neleca)
        link_b = link_a if nelecb == neleca else cistring.gen_linkstr_index(norb, nelecb)
    else:
        link_a, link_b = link_index
    na, nb = link_a.shape[0], link_b.shape[0]
    bra = np.asarray(cibra, dtype=np.float64).reshape(na, nb)
    ket = np.asarray(ciket, dtype=np.float64).reshape(na, nb)
    n2 = norb * norb
    rdm1_c = np.zeros(n2)
    rdm2_c = np.zeros((n2, n2))
    # (p,q) -> (q,p) on the bra side
    swap = np.arange(n2).reshape(norb, norb).T.ravel()
    for lo in range(0, na, block):
        hi = min(na, lo + block)
        t1k = build_t1(ket, norb, (neleca, nelecb), link_a, link_b, (lo, hi)).reshape(-1, n2)
        t1b = build_t1(bra, norb, (neleca, nelecb), link_a, link_b, (lo, hi)).reshape(-1, n2)
        rdm1_c += bra[lo:hi].reshape(-1) @ t1k
        rdm2_c += t1b[:, swap].T @ t1k
    dm1 = rdm1_c.reshape(norb, norb).T.copy()
    dm2 = rdm2_c.reshape(norb, norb, norb, norb)
    if reorder:
        for k in range(norb):
            dm2[:, k, k, :] -= dm1.T
    return dm1, dm2


# ---------------------------------------------------------------------------
# Independent checkers: explicit second quantisation on (alpha, beta) bit
# strings.  Exponentially slow -- small cases only.
# ---------------------------------------------------------------------------

def _apply(op, orb, spin, det):
    """Apply a_{orb,spin} ('d') or a^+_{orb,spin} ('c') to det=(sign, sa, sb)."""
    sign, sa, sb = det
    s = sa if spin == 0 else sb
    bit = 1 << orb
    if op == "d":
        if not s & bit:
            return None
    else:
        if s & bit:
            return None
    # alpha operators sit left of all beta operators in the determinant
    n_before = bin(s & (bit - 1)).count("1")
    if spin == 1:
        n_before += bin(sa).count("1")
    if n_before & 1:
        sign = -sign
    s ^= bit
    return (sign, s, sb) if spin == 0 else (sign, sa, s)


def _expect(bra, ket, ops, norb, nelec):
    """<bra| ops |ket>; ``ops`` = list of (kind, orb, spin), leftmost first."""
    neleca, nelecb = nelec
    sa_list = [int(s) for s in cistring.make_strings(norb, neleca)]
    sb_list = [int(s) for s in cistring.make_strings(norb, nelecb)]
    addr_a = {s: k for k, s in enumerate(sa_list)}
    addr_b = {s: k for k, s in enumerate(sb_list)}
    tot = 0.0
    for ka, sa in enumerate(sa_list):
        for kb, sb in enumerate(sb_list):
            ck = ket[ka, kb]
            if ck == 0.0:
                continue
            det = (1, sa, sb)
            for kind, orb, spin in reversed(ops):
                det = _apply(kind, orb, spin, det)
                if det is None:
                    break
            if det is None:
                continue
            sign, ta, tb = det
            if ta in addr_a and tb in addr_b:
                tot += sign * bra[addr_a[ta], addr_b[tb]] * ck
    return tot


def brute_force_rdm12(cibra, ciket, norb, nelec):
    """dm1[p,q]=<bra|q^+ p|ket>, dm2[p,q,r,s]=<bra|p^+ r^+ s q|ket>, spin-summed."""
    nelec = _unpack_nelec(nelec)
    na = cistring.num_strings(norb, nelec[0])
    nb = cistring.num_strings(norb, nelec[1])
    bra = np.asarray(cibra).reshape(na, nb)
    ket = np.asarray(ciket).reshape(na, nb)
    dm1 = np.zeros((norb, norb))
    dm2 = np.zeros((norb,) * 4)
    for p, q in itertools.product(range(norb), repeat=2):
        dm1[p, q] = sum(
            _expect(bra, ket, [("c", q, s), ("d", p, s)], norb, nelec) for s in (0, 1)
        )
    for p, q, r, s in itertools.product(range(norb), repeat=4):
        dm2[p, q, r, s] = sum(
            _expect(
                bra, ket,
                [("c", p, s1), ("c", r, s2), ("d", s, s2), ("d", q, s1)],
                norb, nelec,
            )
            for s1 in (0, 1) for s2 in (0, 1)
        )
    return dm1, dm2


def hamiltonian_matrix(h1, eri, norb, nelec):
    """Dense FCI Hamiltonian  H = sum h_pq E_pq + 1/2 sum (pq|rs)(E_pq E_rs - d_qr E_ps).

    Built from one-body excitation matrices only (independent of the t1 code
    above); used for the Slater-Condon identity and for exact training vectors
    in the H4/H6 end-to-end checks.
    """
    neleca, nelecb = _unpack_nelec(nelec)
    link_a = cistring.gen_linkstr_index(norb, neleca)
    link_b = cistring.gen_linkstr_index(norb, nelecb)
    na, nb = link_a.shape[0], link_b.shape[0]
    nd = na * nb
    E = np.zeros((norb, norb, nd, nd))
    ida = np.eye(na)
    idb = np.eye(nb)
    for p in range(norb):
        for q in range(norb):
            Ea = np.zeros((na, na))
            for k in range(na):
                for a, i, j, s in link_a[k]:
                    if a == p and i == q:
                        Ea[j, k] += s
            Eb = np.zeros((nb, nb))
            for k in range(nb):
                for a, i, j, s in link_b[k]:
                    if a == p and i == q:
                        Eb[j, k] += s
            E[p, q] = np.kron(Ea, idb) + np.kron(ida, Eb)
    H = np.einsum("pq,pqIJ->IJ", h1, E)
    for p in range(norb):
        for q in range(norb):
            for r in range(norb):
                for s in range(norb):
                    v = eri[p, q, r, s]
                    if v == 0.0:
                        continue
                    H += 0.5 * v * (E[p, q] @ E[r, s])
                    if q == r:
                        H -= 0.5 * v * E[p, s]
    return H
