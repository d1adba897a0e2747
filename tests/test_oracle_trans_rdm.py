"""Pins the CPU oracle of trans_rdm12 / cistring (CPU-only tests).

PySCF cannot run in this image ("parity unpinned" w.r.t. its binary), so the
oracle is pinned against: the definition-based golden vectors under
tests/golden (brute-force second quantisation), RDM sum rules, a Slater-Condon
Hamiltonian identity, the hand value and the hand-derived link-table rows of
SURVEY.md Appendix A.2."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_civec
from oracle import cistring as ocs
from oracle import trans_rdm as otr


def test_strings_and_link_golden():
    g = np.load(os.path.join(GOLDEN, "linkindex_n4_k2.npz"))
    assert np.array_equal(ocs.make_strings(4, 2), g["strings"])
    assert np.array_equal(ocs.gen_linkstr_index(4, 2)[0], g["row0"])


@pytest.mark.parametrize("norb,nocc", [(1, 0), (1, 1), (4, 2), (6, 3), (7, 2), (8, 8), (10, 5)])
def test_addressing_roundtrip(norb, nocc):
    strs = ocs.make_strings(norb, nocc)
    assert len(strs) == ocs.num_strings(norb, nocc)
    assert np.all(np.diff(strs) > 0) or len(strs) == 1
    for k, s in enumerate(int(x) for x in strs):
        assert ocs.str2addr(norb, nocc, s) == k
        assert ocs.addr2str(norb, nocc, k) == s
    tab = ocs.gen_linkstr_index(norb, nocc)
    assert tab.shape == (len(strs), nocc + nocc * (norb - nocc), 4) and tab.dtype == np.int32
    assert set(np.unique(tab[:, :, 3])) <= {-1, 1}


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "trans_rdm_*.npz"))))
def test_definition_goldens(path):
    g = np.load(path)
    norb, nelec = int(g["norb"]), tuple(int(x) for x in g["nelec"])
    d1, d2 = otr.trans_rdm12(g["bra"], g["ket"], norb, nelec)
    assert np.abs(d1 - g["dm1"]).max() < 1e-14
    assert np.abs(d2 - g["dm2"]).max() < 1e-14


def test_brute_force_small():
    bra, ket = random_civec(3, 3, 1), random_civec(3, 3, 2)
    d1, d2 = otr.trans_rdm12(bra, ket, 3, (1, 1))
    b1, b2 = otr.brute_force_rdm12(bra, ket, 3, (1, 1))
    assert np.abs(d1 - b1).max() < 1e-14 and np.abs(d2 - b2).max() < 1e-14


def test_hand_value():
    c = np.zeros((2, 2))
    c[0, 0] = 1.0
    d1, d2 = otr.trans_rdm12(c, c, 2, (1, 1))
    assert np.array_equal(d1, np.diag([2.0, 0.0]))
    ref = np.zeros((2,) * 4)
    ref[0, 0, 0, 0] = 2.0
    assert np.array_equal(d2, ref)


@pytest.mark.parametrize("norb,nelec", [(4, (2, 2)), (5, (3, 2)), (6, (3, 3)), (7, (2, 4))])
def test_sum_rules_and_symmetries(norb, nelec):
    na, nb = ocs.num_strings(norb, nelec[0]), ocs.num_strings(norb, nelec[1])
    bra, ket = random_civec(na, nb, 3), random_civec(na, nb, 4)
    ne = sum(nelec)
    ov = float((bra * ket).sum())
    d1, d2 = otr.trans_rdm12(bra, ket, norb, nelec)
    assert abs(np.trace(d1) - ne * ov) < 1e-13
    assert np.abs(np.einsum("pqrr->pq", d2) - (ne - 1) * d1.T).max() < 1e-13
    assert abs(np.einsum("pprr->", d2) - ne * (ne - 1) * ov) < 1e-12
    assert np.abs(d2 - d2.transpose(2, 3, 0, 1)).max() < 1e-13
    s1, s2 = otr.trans_rdm12(ket, bra, norb, nelec)
    assert np.abs(d1 - s1.T).max() < 1e-13
    assert np.abs(d2 - s2.transpose(1, 0, 3, 2)).max() < 1e-13
    e1, _ = otr.trans_rdm12(bra, bra, norb, nelec)
    w = np.linalg.eigvalsh(0.5 * (e1 + e1.T))
    assert np.abs(e1 - e1.T).max() < 1e-13 and w.min() > -1e-12 and w.max() < 2 + 1e-12


def _sym_integrals(norb, seed):
    rng = np.random.default_rng(seed)
    h = rng.standard_normal((norb, norb))
    h = h + h.T
    e = rng.standard_normal((norb,) * 4)
    e = e + e.transpose(1, 0, 2, 3)
    e = e + e.transpose(0, 1, 3, 2)
    e = e + e.transpose(2, 3, 0, 1)
    return h, e


@pytest.mark.parametrize("norb,nelec", [(4, (2, 2)), (4, (2, 1)), (5, (2, 2))])
def test_hamiltonian_identity(norb, nelec):
    """h1.dm1^T + 1/2 h2.dm2 == <bra|H|ket> with H from an independent construction,
    and a NON-symmetric one-body operator pins the dm1 orientation."""
    na, nb = ocs.num_strings(norb, nelec[0]), ocs.num_strings(norb, nelec[1])
    bra, ket = random_civec(na, nb, 5), random_civec(na, nb, 6)
    h, e = _sym_integrals(norb, 7)
    H = otr.hamiltonian_matrix(h, e, norb, nelec)
    assert np.abs(H - H.T).max() < 1e-12
    d1, d2 = otr.trans_rdm12(bra, ket, norb, nelec)
    lhs = np.einsum("pq,qp->", h, d1) + 0.5 * np.einsum("pqrs,pqrs->", e, d2)
    assert abs(lhs - bra.ravel() @ H @ ket.ravel()) < 1e-11
    t = np.random.default_rng(8).standard_normal((norb, norb))
    T = otr.hamiltonian_matrix(t, np.zeros((norb,) * 4), norb, nelec)
    assert abs(np.einsum("pq,qp->", t, d1) - bra.ravel() @ T @ ket.ravel()) < 1e-12


def test_evcont_exactness_at_training_point():
    """The reference's implicit test (scripts/PES_H_chain/H6_PES/H6_continuation.py:52-82):
    with exact eigenvectors as training states the continuation reproduces the exact
    energy at a training Hamiltonian and is variational elsewhere."""
    from oracle import subspace as osub
    norb, nelec = 4, (2, 2)
    hams = [_sym_integrals(norb, 20 + k) for k in range(3)]
    hams = [(h, 0.1 * e) for h, e in hams]
    vecs, exact = [], []
    for h, e in hams:
        w, v = np.linalg.eigh(otr.hamiltonian_matrix(h, e, norb, nelec))
        vecs.append(v[:, 0].reshape(6, 6))
        exact.append(w[0])
    N = len(vecs)
    S = np.zeros((N, N))
    one = np.zeros((N, N, norb, norb))
    two = np.zeros((N, N) + (norb,) * 4)
    for a in range(N):
        for b in range(N):
            S[a, b] = (vecs[a] * vecs[b]).sum()
            one[a, b], two[a, b] = otr.trans_rdm12(vecs[a], vecs[b], norb, nelec)
    for (h, e), ex in zip(hams, exact):
        en, c = osub.approximate_ground_state(h, e, one, two, S)
        assert abs(en - ex) < 1e-10
    hm = 0.5 * (hams[0][0] + hams[1][0]), 0.5 * (hams[0][1] + hams[1][1])
    en, _ = osub.approximate_ground_state(hm[0], hm[1], one, two, S)
    assert en >= np.linalg.eigvalsh(otr.hamiltonian_matrix(hm[0], hm[1], norb, nelec))[0] - 1e-10
