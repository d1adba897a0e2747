"""Pins of the s+p AO-integral oracle (oracle/integrals_sp.py, McMurchie-Davidson): equality with the
pinned s-only oracle, unit normalisation, the RHF/6-31G energy of water, rotational invariance, and
central finite differences through the reference's assembly formulas
(evcont/ab_initio_gradients_loewdin.py:13-38, 137-152, 234-252)."""
import numpy as np
import pytest

from oracle import integrals as oi
from oracle import integrals_sp as osp

ANG = 1.0 / 0.52917721092
FIELDS = ("ovlp", "hcore", "eri", "ipovlp", "hcore_deriv", "eri_ip1", "e_nuc", "grad_nuc")


def water(rot=None, shift=None):
    r, th = 0.9572 * ANG, np.deg2rad(104.52)
    co = np.array([[0, 0, 0], [r * np.sin(th / 2), 0, r * np.cos(th / 2)], [-r * np.sin(th / 2), 0, r * np.cos(th / 2)]])
    if rot is not None:
        co = co @ rot.T
    if shift is not None:
        co = co + shift
    return [("O", co[0]), ("H", co[1]), ("H", co[2])]


@pytest.fixture(scope="module")
def water_arrays():
    b = osp.SPBasis(water(), "6-31g")
    return b, osp.ao_arrays(b)


def test_reduces_to_the_s_only_oracle():
    h4 = np.array([[0.0, 0.0, 0.0], [0.1, 0.2, 1.7], [0.3, -0.2, 3.5], [1.5, 0.3, 0.5]])
    a = oi.ao_arrays(oi.SBasis([("H", c) for c in h4], "6-31g"))
    b = osp.ao_arrays(osp.SPBasis([("H", c) for c in h4], "6-31g"))
    for k in FIELDS:
        assert np.abs(np.asarray(a[k]) - np.asarray(b[k])).max() < 1e-14, k


def test_water_ao_order_normalisation_and_rhf_energy(water_arrays):
    b, arr = water_arrays
    assert b.nao == 13 and list(b.ao_l) == [0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0]
    assert [tuple(s[2:]) for s in b.aoslices] == [(0, 9), (9, 11), (11, 13)]
    S, h, eri = arr["ovlp"], arr["hcore"], arr["eri"]
    assert np.abs(np.diag(S) - 1).max() < 1e-14
    for perm in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)]:
        assert np.abs(eri - eri.transpose(perm)).max() < 1e-14
    assert np.abs(arr["eri_ip1"] - arr["eri_ip1"].transpose(0, 1, 2, 4, 3)).max() < 1e-14
    assert np.abs(arr["hcore_deriv"].sum(0)).max() < 1e-12
    # restricted Hartree-Fock: -75.9840 Ha at the experimental geometry (HF/6-31G literature value)
    w, v = np.linalg.eigh(S)
    X = v @ np.diag(w ** -0.5) @ v.T
    D = np.zeros_like(S)
    e_old = 0.0
    for it in range(200):
        F = h + 2 * np.einsum("abcd,cd->ab", eri, D) - np.einsum("acbd,cd->ab", eri, D)
        e = np.einsum("ab,ab->", D, h + F) + arr["e_nuc"]
        _, c = np.linalg.eigh(X @ F @ X)
        C = X @ c
        Dn = C[:, :5] @ C[:, :5].T
        if it > 3 and abs(e - e_old) < 1e-10:
            break
        e_old, D = e, (Dn if it == 0 else 0.5 * (D + Dn))
    assert abs(e + 75.9840) < 2e-4


def test_rotational_and_translational_invariance(water_arrays):
    b, arr = water_arrays
    rng = np.random.default_rng(0)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    b2 = osp.SPBasis(water(rot=q, shift=np.array([0.3, -1.0, 2.0])), "6-31g")
    arr2 = osp.ao_arrays(b2, want_ip1=False)
    # s functions are invariant, each p shell rotates with q: U = blockdiag(1.., q, q, 1..)
    U = np.eye(13)
    U[3:6, 3:6] = q
    U[6:9, 6:9] = q
    # phi'_a = sum_b U[a, b] phi_b(rotated frame)  ->  M' = U M U^T
    assert np.abs(arr2["ovlp"] - U @ arr["ovlp"] @ U.T).max() < 1e-13
    assert np.abs(arr2["hcore"] - U @ arr["hcore"] @ U.T).max() < 1e-12
    eri_rot = np.einsum("ai,bj,ck,dl,ijkl->abcd", U, U, U, U, arr["eri"], optimize=True)
    assert np.abs(arr2["eri"] - eri_rot).max() < 1e-13
    assert abs(arr2["e_nuc"] - arr["e_nuc"]) < 1e-13


def test_derivative_integrals_against_finite_differences_oh():
    """O-H fragment (11 AOs, p shells on O): every derivative array against central differences of the
    undifferentiated integrals, assembled as the reference does."""
    co = np.array([[0.1, -0.2, 0.05], [0.4, 0.3, 1.75]])
    b = osp.SPBasis([("O", co[0]), ("H", co[1])], "6-31g")
    arr = osp.ao_arrays(b)
    n = b.nao
    h = 1e-4
    ip, ip1 = arr["ipovlp"], arr["eri_ip1"]
    dS = np.zeros((b.natm, 3, n, n))
    g = np.zeros((b.natm, 3) + (n,) * 4)
    for A, (_, _, p0, p1) in enumerate(b.aoslices):
        dS[A, :, p0:p1, :] -= ip[:, p0:p1, :]
        blk = ip1[:, p0:p1]
        g[A, :, p0:p1] -= blk
        g[A, :, :, p0:p1] -= blk.transpose(0, 2, 1, 3, 4)
        g[A, :, :, :, p0:p1] -= blk.transpose(0, 3, 4, 1, 2)
        g[A, :, :, :, :, p0:p1] -= blk.transpose(0, 3, 4, 2, 1)
    dS = dS + dS.transpose(0, 1, 3, 2)
    for (A, x) in ((0, 0), (0, 2), (1, 1)):
        cp, cm = co.copy(), co.copy()
        cp[A, x] += h
        cm[A, x] -= h
        ap = osp.ao_arrays(b.with_coords(cp), want_ip1=False)
        am = osp.ao_arrays(b.with_coords(cm), want_ip1=False)
        fd = {k: (np.asarray(ap[k]) - np.asarray(am[k])) / (2 * h) for k in ("ovlp", "hcore", "eri", "e_nuc")}
        assert np.abs(dS[A, x] - fd["ovlp"]).max() < 2e-7
        assert np.abs(arr["hcore_deriv"][A, x] - fd["hcore"]).max() < 5e-6   # O core: |h| ~ 30, FD noise
        assert np.abs(g[A, x] - fd["eri"]).max() < 2e-7
        assert abs(arr["grad_nuc"][A, x] - fd["e_nuc"]) < 1e-7
