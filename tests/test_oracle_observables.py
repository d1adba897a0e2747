"""oracle/observables.py: the dipole integrals and the callback observables of the reference's Zundel MD
script (04_Zundel_continuation_MD.py:71-92,140-159), pinned by properties (libcint is absent)."""
import numpy as np

from oracle import integrals_sp as osp
from oracle import observables as oob


def water(shift=np.zeros(3)):
    return [("O", np.array([0.0, 0.0, 0.2]) + shift), ("H", np.array([0.0, 1.45, -0.9]) + shift),
            ("H", np.array([0.1, -1.40, -0.95]) + shift)]


def _ao_values(b, pts):
    """Contracted AOs on points (npts, nao), straight from the definition."""
    d = pts[:, None, :] - b.centers[None, :, :]
    prim = np.prod(d ** b.p_pow[None, :, :], axis=2) * np.exp(-b.p_exp[None, :] * (d ** 2).sum(-1))
    return prim @ b.cmat


def test_int1e_r_against_quadrature():
    b = osp.SPBasis([("O", [0.0, 0.0, 0.1]), ("H", [0.0, 1.2, -0.7])], "6-31g")
    origin = np.array([0.3, -0.2, 0.5])
    r = oob.int1e_r(b, origin)
    # tensor-product Gauss-Hermite-like grid: plain trapezoid on a box wide enough for the diffuse functions and
    # fine enough for the 6-31G valence shells; the tight O 1s core is excluded from the comparison
    g = np.linspace(-9.0, 9.0, 181)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    w = (g[1] - g[0]) ** 3
    phi = _ao_values(b, pts)
    keep = np.array([k for k in range(b.nao) if k != 0])     # AO 0: O 1s (exponents up to 5e3)
    for x in range(3):
        num = (phi[:, keep] * (pts[:, x] - origin[x])[:, None]).T @ phi[:, keep] * w
        assert np.abs(num - r[x][np.ix_(keep, keep)]).max() < 5e-6


def test_translation_rule_and_hermiticity():
    b = osp.SPBasis(water(), "6-31g")
    S = osp.ao_arrays(b, want_ip1=False)["ovlp"]
    O1, O2 = np.array([0.1, 0.2, -0.3]), np.array([-1.0, 0.5, 2.0])
    r1, r2 = oob.int1e_r(b, O1), oob.int1e_r(b, O2)
    for x in range(3):
        assert np.abs(r1[x] - r1[x].T).max() < 1e-13
        assert np.abs(r2[x] - (r1[x] - (O2[x] - O1[x]) * S)).max() < 1e-12
    # moving the molecule and the origin together changes nothing
    sh = np.array([0.7, -1.1, 0.4])
    r3 = oob.int1e_r(osp.SPBasis(water(sh), "6-31g"), O1 + sh)
    assert np.abs(r3 - r1).max() < 1e-12


def test_dipole_of_symmetric_and_charged_densities():
    # H2, one doubly occupied bonding orbital: no dipole; charges zero
    b = osp.SPBasis([("H", [0, 0, -0.7]), ("H", [0, 0, 0.7])], "sto-6g")
    S = osp.ao_arrays(b, want_ip1=False)["ovlp"]
    c = np.array([1.0, 1.0]) / np.sqrt(2.0 + 2.0 * S[0, 1])
    dm = 2.0 * np.outer(c, c)
    assert np.abs(oob.dip_moment(b, dm)).max() < 1e-13
    for m in ("mulliken", "loewdin"):
        q = oob.atomic_charges(b, dm, S, m)
        assert np.abs(q).max() < 1e-13
    # both electrons on atom 0: charges (-1, +1) by Loewdin in the orthogonalised basis, dipole along -z ... +z
    w, V = np.linalg.eigh(S)
    X = (V / np.sqrt(w)) @ V.T
    dm = X @ np.diag([2.0, 0.0]) @ X.T
    q = oob.atomic_charges(b, dm, S, "loewdin")
    assert np.abs(q - np.array([-1.0, 1.0])).max() < 1e-12
    assert abs(oob.atomic_charges(b, dm, S, "mulliken").sum()) < 1e-12        # neutral molecule
    d = oob.dip_moment(b, dm, unit="au")
    assert d[2] > 0.5 and abs(d[0]) < 1e-13 and abs(d[1]) < 1e-13             # electrons sit at z < 0
