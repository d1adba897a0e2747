"""Pins oracle/transform_ci.py (restatement of pyscf.fci.addons.transform_ci, called by the reference
at evcont/FCI_EVCont.py:79-85) without PySCF: Leibniz minors, RDM covariance, energy invariance."""
import numpy as np
import pytest

from oracle import cistring as o_cis
from oracle import trans_rdm as o_trdm
from oracle import transform_ci as o_tci


def _rand_orth(n, seed):
    q, r = np.linalg.qr(np.random.default_rng(seed).standard_normal((n, n)))
    return q * np.sign(np.diag(r))


def _rand_ci(norb, nelec, seed):
    na, nb = o_cis.num_strings(norb, nelec[0]), o_cis.num_strings(norb, nelec[1])
    c = np.random.default_rng(seed).standard_normal((na, nb))
    return c / np.linalg.norm(c)


@pytest.mark.parametrize("norb,nocc", [(3, 1), (4, 2), (5, 3), (6, 3)])
def test_minors_against_leibniz(norb, nocc):
    u = np.random.default_rng(norb * 10 + nocc).standard_normal((norb, norb))   # any matrix, not only orthogonal
    assert np.abs(o_tci.minors(u, norb, nocc) - o_tci.minors_leibniz(u, norb, nocc)).max() < 1e-13


def test_hand_values():
    # one electron: the minors are u itself, ci_new = u^T ci (a column of alpha amplitudes)
    u = _rand_orth(3, 0)
    c = np.array([[0.6], [0.0], [-0.8]])
    assert np.allclose(o_tci.transform_ci(c, (1, 0), u), u.T @ c, atol=1e-15)
    # identity rotation, and a permutation with a sign: swapping the two occupied orbitals flips the sign
    c2 = _rand_ci(4, (2, 2), 1)
    assert np.allclose(o_tci.transform_ci(c2, (2, 2), np.eye(4)), c2, atol=1e-15)
    p = np.eye(2)[[1, 0]]
    assert np.allclose(o_tci.transform_ci(np.ones((1, 1)), (2, 0), p), -np.ones((1, 1)))


@pytest.mark.parametrize("norb,nelec", [(4, (2, 2)), (4, (2, 1)), (5, (3, 2)), (6, (3, 3))])
def test_norm_group_property_and_rdm_covariance(norb, nelec):
    u1, u2 = _rand_orth(norb, 3), _rand_orth(norb, 4)
    c = _rand_ci(norb, nelec, 5)
    c1 = o_tci.transform_ci(c, nelec, u1)
    assert abs(np.linalg.norm(c1) - 1.0) < 1e-13
    # old -> mid (u1), mid -> new (u2) equals old -> new (u1 u2)
    c12 = o_tci.transform_ci(c1, nelec, u2)
    assert np.abs(c12 - o_tci.transform_ci(c, nelec, u1 @ u2)).max() < 1e-13
    # back-rotation
    assert np.abs(o_tci.transform_ci(c1, nelec, u1.T) - c).max() < 1e-13
    # covariance of the (transition) density matrices: every index rotates with u
    d = _rand_ci(norb, nelec, 6)
    d1 = o_tci.transform_ci(d, nelec, u1)
    dm1, dm2 = o_trdm.trans_rdm12(c, d, norb, nelec)
    n1, n2 = o_trdm.trans_rdm12(c1, d1, norb, nelec)
    assert np.abs(n1 - u1.T @ dm1 @ u1).max() < 1e-13
    assert np.abs(n2 - np.einsum("pqrs,pa,qb,rc,sd->abcd", dm2, u1, u1, u1, u1)).max() < 1e-13


def test_eigenvector_of_the_rotated_hamiltonian():
    norb, nelec = 5, (2, 2)
    rng = np.random.default_rng(7)
    h1 = rng.standard_normal((norb, norb))
    h1 = h1 + h1.T
    eri = rng.standard_normal((norb,) * 4)
    eri = eri + eri.transpose(1, 0, 2, 3)
    eri = eri + eri.transpose(0, 1, 3, 2)
    eri = eri + eri.transpose(2, 3, 0, 1)
    H = o_trdm.hamiltonian_matrix(h1, eri, norb, nelec)
    w, v = np.linalg.eigh(H)
    u = _rand_orth(norb, 8)
    h1n = u.T @ h1 @ u
    erin = np.einsum("pqrs,pa,qb,rc,sd->abcd", eri, u, u, u, u)
    Hn = o_trdm.hamiltonian_matrix(h1n, erin, norb, nelec)
    na = o_cis.num_strings(norb, 2)
    for k in (0, 3):
        cn = o_tci.transform_ci(v[:, k].reshape(na, na), nelec, u).reshape(-1)
        assert np.abs(Hn @ cn - w[k] * cn).max() < 1e-11
