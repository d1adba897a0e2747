"""Pins of the s-shell AO-integral oracle (oracle/integrals.py).

libcint is not in the image, so the oracle is pinned by published table values
(Szabo & Ostlund, "Modern Quantum Chemistry", sections 3.5.2 / 3.5.3: STO-3G H2 at
R = 1.4 bohr and HeH+ at R = 1.4632 bohr), exact properties, and central finite
differences of the undifferentiated integrals against the derivative integrals
assembled exactly as the reference assembles them
(evcont/ab_initio_gradients_loewdin.py:13-38, 137-152, 234-252).
"""
import numpy as np
import pytest

from oracle import integrals as oi

H4 = np.array([[0.0, 0.0, 0.0], [0.1, 0.2, 1.7], [0.3, -0.2, 3.5], [1.5, 0.3, 0.5]])


def test_szabo_ostlund_h2_sto3g():
    b = oi.SBasis([("H", (0, 0, 0)), ("H", (0, 0, 1.4))], "sto-3g")
    S, T = oi.int1e_ovlp(b), oi.int1e_kin(b)
    V1 = -oi.int1e_rinv(b, b.coords[0])
    eri = oi.int2e(b)
    assert abs(S[0, 0] - 1) < 1e-14 and abs(S[0, 1] - 0.6593) < 5e-5
    assert abs(T[0, 0] - 0.7600) < 5e-5 and abs(T[0, 1] - 0.2365) < 5e-5
    assert abs(V1[0, 0] + 1.2266) < 5e-5 and abs(V1[0, 1] + 0.5974) < 5e-5 and abs(V1[1, 1] + 0.6538) < 5e-5
    assert abs(eri[0, 0, 0, 0] - 0.7746) < 5e-5 and abs(eri[0, 0, 1, 1] - 0.5697) < 5e-5
    assert abs(eri[1, 0, 0, 0] - 0.4441) < 5e-5 and abs(eri[1, 0, 1, 0] - 0.2970) < 5e-5


def test_szabo_ostlund_heh_plus():
    b = oi.SBasis([("He", (0, 0, 0)), ("H", (0, 0, 1.4632))], "sto-3g-so")
    S, T = oi.int1e_ovlp(b), oi.int1e_kin(b)
    eri = oi.int2e(b)
    assert abs(S[0, 1] - 0.4508) < 1e-4
    assert abs(T[0, 0] - 2.1643) < 1e-4 and abs(T[0, 1] - 0.1670) < 1e-4 and abs(T[1, 1] - 0.7600) < 1e-4
    assert abs(eri[0, 0, 0, 0] - 1.3072) < 1e-4 and abs(eri[1, 1, 1, 1] - 0.7746) < 1e-4
    assert abs(eri[0, 0, 1, 1] - 0.6057) < 1e-4 and abs(eri[1, 0, 0, 0] - 0.4373) < 1e-4


def test_hydrogen_atom_sto6g_energy():
    h = oi.SBasis([("H", (0, 0, 0))], "sto-6g")
    assert abs(oi.get_hcore(h)[0, 0] + 0.4710390) < 2e-7
    assert abs(oi.int1e_ovlp(h)[0, 0] - 1) < 1e-14


def test_boys_against_quadrature():
    from scipy import integrate
    for t in (0.0, 1e-9, 0.3, 5.0, 29.9, 30.1, 80.0, 900.0):
        f = oi.boys(2, np.array([t]))
        for m in range(3):
            ref = integrate.quad(lambda u: u ** (2 * m) * np.exp(-t * u * u), 0, 1, epsabs=1e-15, epsrel=1e-14)[0]
            assert abs(f[m, 0] - ref) < 2e-14 * max(1.0, ref)


def _fd(fn, b, h=1e-4):
    co = b.coords
    out = np.zeros((b.natm, 3) + np.shape(fn(b)))
    for A in range(b.natm):
        for x in range(3):
            cp, cm = co.copy(), co.copy()
            cp[A, x] += h
            cm[A, x] -= h
            out[A, x] = (fn(b.with_coords(cp)) - fn(b.with_coords(cm))) / (2 * h)
    return out


@pytest.mark.parametrize("basis", ["sto-6g", "6-31g"])
def test_derivative_integrals_against_finite_differences(basis):
    b = oi.SBasis([("H", c) for c in H4], basis)
    arr = oi.ao_arrays(b)
    n = b.nao
    # dS/dR as the reference builds it from int1e_ipovlp (get_overlap_grad, :13-38)
    ip = arr["ipovlp"]
    d = np.zeros((b.natm, 3, n, n))
    for A, (_, _, p0, p1) in enumerate(b.aoslices):
        d[A, :, p0:p1, :] -= ip[:, p0:p1, :]
    d = d + d.transpose(0, 1, 3, 2)
    assert np.abs(d - _fd(oi.int1e_ovlp, b)).max() < 5e-9
    # hcore_generator()(A) is the full nuclear derivative of the core Hamiltonian
    assert np.abs(arr["hcore_deriv"] - _fd(oi.get_hcore, b)).max() < 5e-9
    # d(ab|cd)/dR_A = - sum over the four index positions sitting on atom A of int2e_ip1
    ip1 = arr["eri_ip1"]
    g = np.zeros((b.natm, 3) + (n,) * 4)
    for A, (_, _, p0, p1) in enumerate(b.aoslices):
        blk = ip1[:, p0:p1]
        g[A, :, p0:p1] -= blk
        g[A, :, :, p0:p1] -= blk.transpose(0, 2, 1, 3, 4)
        g[A, :, :, :, p0:p1] -= blk.transpose(0, 3, 4, 1, 2)
        g[A, :, :, :, :, p0:p1] -= blk.transpose(0, 3, 4, 2, 1)
    assert np.abs(g - _fd(oi.int2e, b)).max() < 5e-9
    assert np.abs(arr["grad_nuc"] - _fd(oi.energy_nuc, b)).max() < 5e-9


def test_symmetries_and_translation_invariance():
    b = oi.SBasis([("H", c) for c in H4], "sto-6g")
    arr = oi.ao_arrays(b)
    eri, ip1 = arr["eri"], arr["eri_ip1"]
    for perm in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)]:
        assert np.abs(eri - eri.transpose(perm)).max() < 1e-15
    assert np.abs(ip1 - ip1.transpose(0, 1, 2, 4, 3)).max() < 1e-15
    # <nabla a|b> + <a|nabla b> = 0
    assert np.abs(arr["ipovlp"] + arr["ipovlp"].transpose(0, 2, 1)).max() < 1e-15
    # sum over atoms of any nuclear derivative vanishes
    assert np.abs(arr["hcore_deriv"].sum(0)).max() < 1e-13
    assert np.abs(arr["grad_nuc"].sum(0)).max() < 1e-13
    # rigid shift leaves every array unchanged
    arr2 = oi.ao_arrays(b.with_coords(b.coords + np.array([0.3, -1.1, 2.0])))
    for k in ("ovlp", "hcore", "eri", "ipovlp", "hcore_deriv", "eri_ip1"):
        assert np.abs(arr[k] - arr2[k]).max() < 2e-13
    # Schwarz: (ab|ab) >= 0 and (ab|cd)^2 <= (ab|ab)(cd|cd)
    n = b.nao
    m = eri.reshape(n * n, n * n)
    dg = np.diag(m)
    assert dg.min() > 0 and (m ** 2 <= np.outer(dg, dg) * (1 + 1e-12)).all()
