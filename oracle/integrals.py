"""AO integrals over contracted s-type Gaussians -- CPU oracle for the device
integral engine.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

What the reference asks PySCF / libcint for on the prediction path, restated for
s shells in closed form (Boys function F_m, Gaussian product theorem):

============================  =================================================  =======================
array                         definition                                         reference call site
============================  =================================================  =======================
``int1e_ovlp``  (n,n)         <a|b>                                              ab_initio_gradients_loewdin.py:338 (via get_loewdin_trafo)
``int1e_ipovlp`` (3,n,n)      <nabla a|b>                                        :25
``int1e_kin`` / ``int1e_nuc``  <a|-1/2 nabla^2|b>, <a|sum_C -Z_C/|r-C||b>         scf.hf.get_hcore, :177,:338
``int1e_ipkin``/``ipnuc``      <nabla a|T|b>, <nabla a|V_nuc|b>                   grad.RHF.hcore_generator, :147
``int1e_iprinv`` @ R_C        <nabla a|1/|r-C||b>                                same
``int2e`` (n,n,n,n)           (ab|cd), chemists' notation                        :339 (ao2mo.kernel)
``int2e_ip1`` (3,n,n,n,n)     (nabla a b|cd)                                     :283-284
``hcore_generator()(A)``      v + v^T(0,2,1),  v = -Z_A iprinv@R_A,              :147
                              v[:, p0:p1] -= (ipkin + ipnuc)[:, p0:p1]
``energy_nuc``, ``grad_nuc``  sum_{A<B} Z_A Z_B/R_AB and its gradient            :370,:378
============================  =================================================  =======================

libcint (a PySCF dependency) is absent from this image, so these are pinned by
(tests/test_oracle_integrals.py): the Szabo-Ostlund STO-3G H2 / HeH+ table values,
unit self-overlap, central finite differences of every undifferentiated integral
against the derivative integrals *through the reference's own assembly formulas*
(``get_overlap_grad``, ``hcore_generator``, the four-position sum of ``int2e_ip1``),
translational invariance, and the H-atom STO-6G energy.  **Parity unpinned with
respect to the libcint binary.**

Basis data (Appendix B of SURVEY.md; EMSL/BSE STO-6G, STO-3G and 6-31G for H):
contraction coefficients refer to normalised primitives and the contracted
function is renormalised to unit self-overlap, as ``pyscf.gto`` does.
"""
import math

import numpy as np
from scipy import special

BASIS_S = {
    ("H", "sto-6g"): [
        ([35.52322122, 6.513143725, 1.822142904, 0.625955266, 0.243076747, 0.100112428],
         [0.00916359628, 0.04936149294, 0.1685383049, 0.3705627997, 0.4164915298, 0.1303340841]),
    ],
    ("H", "sto-3g"): [
        ([3.42525091, 0.62391373, 0.16885540], [0.15432897, 0.53532814, 0.44463454]),
    ],
    # Szabo & Ostlund's HeH+ example: Slater exponent 2.0925 for He (scaled STO-3G)
    ("He", "sto-3g-so"): [
        ([0.109818 * 2.0925 ** 2, 0.405771 * 2.0925 ** 2, 2.22766 * 2.0925 ** 2],
         [0.444635, 0.535328, 0.154329]),
    ],
    ("H", "sto-3g-so"): [
        ([0.109818 * 1.24 ** 2, 0.405771 * 1.24 ** 2, 2.22766 * 1.24 ** 2],
         [0.444635, 0.535328, 0.154329]),
    ],
    ("H", "6-31g"): [
        ([18.7311370, 2.8253937, 0.6401217], [0.03349460, 0.23472695, 0.81375733]),
        ([0.1612778], [1.0]),
    ],
}
CHARGE = {"H": 1, "He": 2}


def boys(mmax, t):
    """F_m(t) = int_0^1 u^(2m) exp(-t u^2) du for m = 0..mmax; returns (mmax+1, *t.shape)."""
    t = np.asarray(t, dtype=np.float64)
    out = np.empty((mmax + 1,) + t.shape)
    small = t < 30.0
    ts = np.where(small, t, 0.0)
    tl = np.where(small, 1.0, t)
    for m in range(mmax + 1):
        # all-positive Kummer series: F_m = exp(-t) sum_k (2t)^k / ((2m+1)(2m+3)...(2m+2k+1))
        term = np.full(t.shape, 1.0 / (2 * m + 1))
        acc = term.copy()
        for k in range(1, 200):
            term = term * (2.0 * ts) / (2 * m + 2 * k + 1)
            acc += term
        series = np.exp(-ts) * acc
        big = special.gamma(m + 0.5) * special.gammainc(m + 0.5, tl) / (2.0 * tl ** (m + 0.5))
        out[m] = np.where(small, series, big)
    return out


class SBasis:
    """Primitive-level description of a molecule of s shells.

    ``atoms``: list of (symbol, (x, y, z)) in bohr; ``basis``: key into BASIS_S.
    """

    def __init__(self, atoms, basis="sto-6g"):
        self.symbols = [a[0] for a in atoms]
        self.coords = np.array([a[1] for a in atoms], dtype=np.float64).reshape(-1, 3)
        self.charges = np.array([CHARGE[s] for s in self.symbols], dtype=np.float64)
        self.natm = len(atoms)
        ao_atom, exps, wts, prim_ao = [], [], [], []
        slices = []
        for ia, sym in enumerate(self.symbols):
            p0 = len(ao_atom)
            for e, c in BASIS_S[(sym, basis.lower())]:
                e = np.asarray(e, dtype=np.float64)
                c = np.asarray(c, dtype=np.float64) * (2.0 * e / np.pi) ** 0.75
                ss = (c[:, None] * c[None, :] * (np.pi / (e[:, None] + e[None, :])) ** 1.5).sum()
                c = c / math.sqrt(ss)
                for ek, ck in zip(e, c):
                    exps.append(ek)
                    wts.append(ck)
                    prim_ao.append(len(ao_atom))
                ao_atom.append(ia)
            slices.append((0, 0, p0, len(ao_atom)))
        self.ao_atom = np.array(ao_atom)
        self.nao = len(ao_atom)
        self.exps = np.array(exps)
        self.wts = np.array(wts)
        self.prim_ao = np.array(prim_ao)
        self.aoslices = np.array(slices, dtype=np.int64)
        self.nprim = len(exps)
        #: contraction matrix (nprim, nao)
        self.cmat = np.zeros((self.nprim, self.nao))
        self.cmat[np.arange(self.nprim), self.prim_ao] = self.wts

    def with_coords(self, coords):
        other = object.__new__(SBasis)
        other.__dict__.update(self.__dict__)
        other.coords = np.array(coords, dtype=np.float64).reshape(-1, 3)
        return other

    @property
    def prim_centers(self):
        return self.coords[self.ao_atom[self.prim_ao]]

    def contract2(self, prim):
        """(..., nprim, nprim) -> (..., nao, nao)."""
        return np.einsum("...ij,ia,jb->...ab", prim, self.cmat, self.cmat, optimize=True)


def _pairs(b):
    a = b.exps
    A = b.prim_centers
    p = a[:, None] + a[None, :]
    mu = a[:, None] * a[None, :] / p
    AB = A[:, None, :] - A[None, :, :]
    r2 = (AB ** 2).sum(-1)
    P = (a[:, None, None] * A[:, None, :] + a[None, :, None] * A[None, :, :]) / p[..., None]
    K = np.exp(-mu * r2)
    return p, mu, AB, r2, P, K


def int1e_ovlp(b):
    p, mu, AB, r2, P, K = _pairs(b)
    return b.contract2((np.pi / p) ** 1.5 * K)


def int1e_ipovlp(b):
    """<nabla a|b> = -d/dA <a|b>."""
    p, mu, AB, r2, P, K = _pairs(b)
    s = (np.pi / p) ** 1.5 * K
    prim = 2.0 * mu[None] * AB.transpose(2, 0, 1) * s[None]
    return b.contract2(prim)


def int1e_kin(b):
    p, mu, AB, r2, P, K = _pairs(b)
    s = (np.pi / p) ** 1.5 * K
    return b.contract2(mu * (3.0 - 2.0 * mu * r2) * s)


def int1e_ipkin(b):
    p, mu, AB, r2, P, K = _pairs(b)
    s = (np.pi / p) ** 1.5 * K
    prim = 2.0 * (mu ** 2 * (5.0 - 2.0 * mu * r2) * s)[None] * AB.transpose(2, 0, 1)
    return b.contract2(prim)


def _rinv_prim(b, origin):
    """Primitive <a|1/|r-C||b> and <nabla a|1/|r-C||b> for the point C = origin."""
    p, mu, AB, r2, P, K = _pairs(b)
    PC = P - np.asarray(origin, dtype=np.float64)[None, None, :]
    t = p * (PC ** 2).sum(-1)
    f = boys(1, t)
    pref = 2.0 * np.pi / p * K
    val = pref * f[0]
    a = b.exps
    # d/dA_x of val, then the minus sign of nabla_r = -nabla_A
    dA = pref[None] * (-2.0 * mu[None] * AB.transpose(2, 0, 1) * f[0][None]
                       - 2.0 * a[None, :, None] * PC.transpose(2, 0, 1) * f[1][None])
    return val, -dA


def int1e_rinv(b, origin):
    return b.contract2(_rinv_prim(b, origin)[0])


def int1e_iprinv(b, origin):
    return b.contract2(_rinv_prim(b, origin)[1])


def int1e_nuc(b):
    out = np.zeros((b.nao, b.nao))
    for z, c in zip(b.charges, b.coords):
        out -= z * int1e_rinv(b, c)
    return out


def int1e_ipnuc(b):
    out = np.zeros((3, b.nao, b.nao))
    for z, c in zip(b.charges, b.coords):
        out -= z * int1e_iprinv(b, c)
    return out


def get_hcore(b):
    return int1e_kin(b) + int1e_nuc(b)


def hcore_generator(b):
    """What ``pyscf.grad.rhf.Gradients.hcore_generator`` returns: atom -> (3, n, n)."""
    h1 = -(int1e_ipkin(b) + int1e_ipnuc(b))

    def hcore_deriv(atm):
        p0, p1 = b.aoslices[atm][2:4]
        v = -b.charges[atm] * int1e_iprinv(b, b.coords[atm])
        v[:, p0:p1] += h1[:, p0:p1]
        return v + v.transpose(0, 2, 1)

    return hcore_deriv


def energy_nuc(b):
    e = 0.0
    for i in range(b.natm):
        for j in range(i):
            e += b.charges[i] * b.charges[j] / np.linalg.norm(b.coords[i] - b.coords[j])
    return e


def grad_nuc(b):
    g = np.zeros((b.natm, 3))
    for i in range(b.natm):
        for j in range(b.natm):
            if i != j:
                d = b.coords[i] - b.coords[j]
                g[i] -= b.charges[i] * b.charges[j] * d / np.linalg.norm(d) ** 3
    return g


def int2e_and_ip1(b, want_ip1=True):
    """(ab|cd) as (n,n,n,n) and (nabla a b|cd) as (3,n,n,n,n), one bra AO at a time."""
    p, mu, AB, r2, P, K = _pairs(b)
    n, npr = b.nao, b.nprim
    a = b.exps
    A = b.prim_centers
    eri = np.zeros((n, n, n, n))
    ip1 = np.zeros((3, n, n, n, n)) if want_ip1 else None
    # ket side flattened over primitive pairs (k, l)
    q = p.reshape(-1)
    Q = P.reshape(-1, 3)
    Kq = K.reshape(-1)
    ckl = (b.cmat[:, None, :, None] * b.cmat[None, :, None, :]).reshape(npr * npr, n * n)
    for i in range(npr):  # bra primitive i (on AO prim_ao[i])
        pi = p[i][:, None]  # (npr_j, 1)
        rho = pi * q[None, :] / (pi + q[None, :])
        PQ = P[i][:, None, :] - Q[None, :, :]  # (j, kl, 3)
        t = rho * (PQ ** 2).sum(-1)
        f = boys(1, t)
        pref = 2.0 * np.pi ** 2.5 / (pi * q[None, :] * np.sqrt(pi + q[None, :])) * K[i][:, None] * Kq[None, :]
        i0 = pref * f[0]
        # contract j -> b, kl -> cd
        tmp = np.einsum("jk,jb->bk", i0, b.cmat, optimize=True) @ ckl
        eri[b.prim_ao[i]] += b.wts[i] * tmp.reshape(n, n, n)
        if want_ip1:
            i1 = pref * f[1]
            # d/dA_x [ab|cd] = 2 alpha [(a+1_x) b|cd],  [p_x b|cd] = (P-A)_x I0 + (W-P)_x I1,
            # W - P = -(rho/p) (P - Q);  (nabla a b|cd) = -d/dA
            PA = P[i] - A[i][None, :]  # (j, 3)
            for x in range(3):
                d = 2.0 * a[i] * (PA[:, x][:, None] * i0 - (rho / pi) * PQ[:, :, x] * i1)
                tmp = np.einsum("jk,jb->bk", d, b.cmat, optimize=True) @ ckl
                ip1[x, b.prim_ao[i]] -= b.wts[i] * tmp.reshape(n, n, n)
    return eri, ip1


def int2e(b):
    return int2e_and_ip1(b, want_ip1=False)[0]


def ao_arrays(b):
    """Everything the prediction path reads, as the ``evcont_b200.mol.ArrayMol`` arguments."""
    eri, ip1 = int2e_and_ip1(b)
    gen = hcore_generator(b)
    return dict(ovlp=int1e_ovlp(b), hcore=get_hcore(b), eri=eri, ipovlp=int1e_ipovlp(b),
                hcore_deriv=np.array([gen(A) for A in range(b.natm)]), eri_ip1=ip1,
                aoslices=b.aoslices, e_nuc=energy_nuc(b), grad_nuc=grad_nuc(b))
