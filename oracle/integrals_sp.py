"""AO integrals over contracted Cartesian s and p Gaussians (McMurchie-Davidson) -- CPU oracle for
the s+p device integral engine (second half: 6-31G H / O, i.e. the H2O
and Zundel configurations).  TEST INFRASTRUCTURE ONLY.

Same arrays and conventions as oracle/integrals.py (which it reproduces for s-only molecules): what
``mol.intor('int1e_ovlp' | 'int1e_ipovlp' | 'int2e' | 'int2e_ip1')``, ``scf.hf.get_hcore``,
``grad.RHF.hcore_generator`` / ``grad_nuc`` and ``mol.energy_nuc`` return in PySCF
(evcont/ab_initio_gradients_loewdin.py:25,130,147,177,283-284,338-339,370,378).  AO order as
pyscf.gto: atoms in input order, per atom all s shells, then the p shells, p components x, y, z;
every contracted function normalised to unit self-overlap.

Hermite expansion:  G_i(x; a, A) G_j(x; b, B) = sum_t E^{ij}_t Lambda_t(x; p, P),
  E^{00}_0 = exp(-mu X_AB^2),  E^{i+1,j}_t = E^{ij}_{t-1}/(2p) + X_PA E^{ij}_t + (t+1) E^{ij}_{t+1},
  (ab|cd) = 2 pi^{5/2} / (p q sqrt(p+q)) sum_{tuv} E^{ab}_{tuv} sum_{t'u'v'} (-1)^{t'+u'+v'} E^{cd}_{t'u'v'}
            R_{t+t',u+u',v+v'}(rho, P - Q),   R^n_{000} = (-2 rho)^n F_n(rho |PQ|^2).
Centre derivative of a Cartesian Gaussian: d/dA_x G_l = 2a G_{l+1} - l G_{l-1};  nabla_r = -d/dA.

libcint is absent: pinned by tests/test_oracle_integrals_sp.py (equality with the s-only oracle,
unit normalisation, rotational invariance, finite differences through the reference's assembly
formulas, the RHF/6-31G energy of water).  **Parity unpinned w.r.t. the libcint binary.**
"""
import itertools
import math

import numpy as np

from .integrals import BASIS_S, boys

CHARGE = {"H": 1, "He": 2, "O": 8}

#: (element, basis) -> list of (l, exponents, coefficients); "sp" shells already split
BASIS = {(k[0], k[1]): [(0, e, c) for e, c in v] for k, v in BASIS_S.items()}
BASIS[("O", "6-31g")] = [
    (0, [5484.6717, 825.23495, 188.04696, 52.9645, 16.89757, 5.7996353],
     [0.0018311, 0.0139501, 0.0684451, 0.2327143, 0.470193, 0.3585209]),
    (0, [15.539616, 3.5999336, 1.0137618], [-0.1107775, -0.1480263, 1.130767]),
    (0, [0.2700058], [1.0]),
    (1, [15.539616, 3.5999336, 1.0137618], [0.0708743, 0.3397528, 0.7271586]),
    (1, [0.2700058], [1.0]),
]

_P = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]


class SPBasis:
    """Primitive Cartesian functions of a molecule of s and p shells."""

    def __init__(self, atoms, basis="6-31g"):
        self.symbols = [a[0] for a in atoms]
        self.coords = np.array([a[1] for a in atoms], dtype=np.float64).reshape(-1, 3)
        self.charges = np.array([CHARGE[s] for s in self.symbols], dtype=np.float64)
        self.natm = len(atoms)
        self.basis = basis.lower()
        p_atom, p_exp, p_wt, p_pow, p_ao = [], [], [], [], []
        ao_atom, ao_l = [], []
        slices = []
        for ia, sym in enumerate(self.symbols):
            p0 = len(ao_atom)
            shells = sorted(BASIS[(sym, self.basis)], key=lambda s: s[0])  # stable: s shells, then p
            for l, e, c in shells:
                e = np.asarray(e, dtype=np.float64)
                c = np.asarray(c, dtype=np.float64)
                if l == 0:
                    c = c * (2.0 * e / np.pi) ** 0.75
                    ss = (c[:, None] * c[None, :] * (np.pi / (e[:, None] + e[None, :])) ** 1.5).sum()
                    comps = [(0, 0, 0)]
                else:
                    c = c * (2.0 * e / np.pi) ** 0.75 * 2.0 * np.sqrt(e)
                    pp = e[:, None] + e[None, :]
                    ss = (c[:, None] * c[None, :] * (np.pi / pp) ** 1.5 / (2.0 * pp)).sum()
                    comps = _P
                c = c / math.sqrt(ss)
                for pw in comps:
                    for ek, ck in zip(e, c):
                        p_atom.append(ia); p_exp.append(ek); p_wt.append(ck); p_pow.append(pw)
                        p_ao.append(len(ao_atom))
                    ao_atom.append(ia)
                    ao_l.append(l)
            slices.append((0, 0, p0, len(ao_atom)))
        self.p_atom = np.array(p_atom); self.p_exp = np.array(p_exp); self.p_wt = np.array(p_wt)
        self.p_pow = np.array(p_pow, dtype=int).reshape(-1, 3); self.p_ao = np.array(p_ao)
        self.ao_atom = np.array(ao_atom); self.ao_l = np.array(ao_l)
        self.nao, self.nprim = len(ao_atom), len(p_exp)
        self.aoslices = np.array(slices, dtype=np.int64)
        self.cmat = np.zeros((self.nprim, self.nao))
        self.cmat[np.arange(self.nprim), self.p_ao] = self.p_wt

    def with_coords(self, coords):
        other = object.__new__(SPBasis)
        other.__dict__.update(self.__dict__)
        other.coords = np.array(coords, dtype=np.float64).reshape(-1, 3)
        return other

    @property
    def centers(self):
        return self.coords[self.p_atom]


# ---- Hermite machinery, vectorised over arrays of primitive pairs / quartets ------------------------
def _E(i, j, t, Qx, a, b, cache):
    """E^{ij}_t for arrays (Qx = A_x - B_x, exponents a, b); without the exp(-mu Qx^2) factor."""
    key = (i, j, t)
    if key in cache:
        return cache[key]
    p = a + b
    if t < 0 or t > i + j:
        r = np.zeros_like(Qx)
    elif i == 0 and j == 0:
        r = np.ones_like(Qx)
    elif i > 0:
        r = (_E(i - 1, j, t - 1, Qx, a, b, cache) / (2 * p) - (b / p) * Qx * _E(i - 1, j, t, Qx, a, b, cache)
             + (t + 1) * _E(i - 1, j, t + 1, Qx, a, b, cache))
    else:
        r = (_E(i, j - 1, t - 1, Qx, a, b, cache) / (2 * p) + (a / p) * Qx * _E(i, j - 1, t, Qx, a, b, cache)
             + (t + 1) * _E(i, j - 1, t + 1, Qx, a, b, cache))
    cache[key] = r
    return r


def _R(tmax, alpha, PC):
    """dict (t,u,v) -> R^0_{tuv}(alpha, PC) for t+u+v <= tmax; arrays over the leading dims of PC."""
    T = alpha * (PC ** 2).sum(-1)
    F = boys(tmax, T)
    Rn = {}
    for n in range(tmax + 1):
        Rn[(n, 0, 0, 0)] = (-2.0 * alpha) ** n * F[n]

    def get(n, t, u, v):
        key = (n, t, u, v)
        if key in Rn:
            return Rn[key]
        if t < 0 or u < 0 or v < 0:
            return 0.0
        if t > 0:
            r = (t - 1) * get(n + 1, t - 2, u, v) + PC[..., 0] * get(n + 1, t - 1, u, v)
        elif u > 0:
            r = (u - 1) * get(n + 1, t, u - 2, v) + PC[..., 1] * get(n + 1, t, u - 1, v)
        else:
            r = (v - 1) * get(n + 1, t, u, v - 2) + PC[..., 2] * get(n + 1, t, u, v - 1)
        Rn[key] = r
        return r

    out = {}
    for t in range(tmax + 1):
        for u in range(tmax + 1 - t):
            for v in range(tmax + 1 - t - u):
                out[(t, u, v)] = get(0, t, u, v)
    return out


class _Pairs:
    """All ordered primitive pairs (i, j) with lazily cached E tables per (power_i, power_j)."""

    def __init__(self, b):
        self.b = b
        A = b.centers
        self.a = b.p_exp[:, None] * np.ones((1, b.nprim))
        self.bb = np.ones((b.nprim, 1)) * b.p_exp[None, :]
        self.p = self.a + self.bb
        self.Q = A[:, None, :] - A[None, :, :]
        self.P = (self.a[..., None] * A[:, None, :] + self.bb[..., None] * A[None, :, :]) / self.p[..., None]
        self.K = np.exp(-(self.a * self.bb / self.p) * (self.Q ** 2).sum(-1))
        self.caches = [dict(), dict(), dict()]

    def E(self, d, i, j, t):
        return _E(i, j, t, self.Q[..., d], self.a, self.bb, self.caches[d])


def _shift_terms(pw, alpha, x):
    """d/dA_x of a Cartesian Gaussian with powers pw: list of (factor array, new powers)."""
    up = list(pw); up[x] += 1
    terms = [(2.0 * alpha, tuple(up))]
    if pw[x] > 0:
        dn = list(pw); dn[x] -= 1
        terms.append((-float(pw[x]) * np.ones_like(alpha), tuple(dn)))
    return terms


def _one_electron(b):
    """Primitive-level overlap-type quantities through explicit power bookkeeping (loops over the
    distinct power pairs, vectorised over primitives)."""
    n = b.nprim
    pr = _Pairs(b)
    pows = [tuple(x) for x in b.p_pow]

    def ovl(pa, pb):  # <G_pa | G_pb> for all primitive pairs, as if every primitive had those powers
        r = (np.pi / pr.p) ** 1.5 * pr.K
        for d in range(3):
            r = r * pr.E(d, pa[d], pb[d], 0)
        return r

    def kin(pa, pb):
        bexp = pr.bb
        r = bexp * (2 * sum(pb) + 3) * ovl(pa, pb)
        for d in range(3):
            up = list(pb); up[d] += 2
            r = r - 2.0 * bexp ** 2 * ovl(pa, tuple(up))
            if pb[d] >= 2:
                dn = list(pb); dn[d] -= 2
                r = r - 0.5 * pb[d] * (pb[d] - 1) * ovl(pa, tuple(dn))
        return r

    def rinv(pa, pb, C):
        L = sum(pa) + sum(pb)
        R = _R(L, pr.p, pr.P - np.asarray(C)[None, None, :])
        r = 0.0
        for t in range(pa[0] + pb[0] + 1):
            for u in range(pa[1] + pb[1] + 1):
                for v in range(pa[2] + pb[2] + 1):
                    r = r + pr.E(0, pa[0], pb[0], t) * pr.E(1, pa[1], pb[1], u) * pr.E(2, pa[2], pb[2], v) * R[(t, u, v)]
        return 2.0 * np.pi / pr.p * pr.K * r

    def assemble(fn, deriv):
        """Matrix over primitives of fn (deriv=False) or of <nabla a|fn|b> (3, n, n)."""
        out = np.zeros((3, n, n)) if deriv else np.zeros((n, n))
        upow = sorted(set(pows))
        for pa in upow:
            ia = np.array([k for k in range(n) if pows[k] == pa])
            for pb in upow:
                ib = np.array([k for k in range(n) if pows[k] == pb])
                sel = np.ix_(ia, ib)
                if not deriv:
                    out[sel] = fn(pa, pb)[sel]
                else:
                    for x in range(3):
                        acc = 0.0
                        for fac, pw in _shift_terms(pa, pr.a, x):
                            acc = acc + fac * fn(pw, pb)
                        out[x][sel] = -acc[sel]     # nabla_r = -d/dA
        return out

    return pr, assemble, ovl, kin, rinv


def ao_arrays(b, want_ip1=True):
    """Everything the prediction path reads, as the ``evcont_b200.mol.ArrayMol`` arguments."""
    cm = b.cmat
    pr, assemble, ovl, kin, rinv = _one_electron(b)
    c2 = lambda m: np.einsum("...ij,ia,jb->...ab", m, cm, cm, optimize=True)
    S = c2(assemble(ovl, False))
    T = c2(assemble(kin, False))
    ipovlp = c2(assemble(ovl, True))
    ipkin = c2(assemble(kin, True))
    V = np.zeros_like(S)
    ipnuc = np.zeros_like(ipovlp)
    iprinv = []
    for z, C in zip(b.charges, b.coords):
        f = lambda pa, pb, C=C: rinv(pa, pb, C)
        V -= z * c2(assemble(f, False))
        d = c2(assemble(f, True))
        iprinv.append(d)
        ipnuc -= z * d
    hcore = T + V
    h1 = -(ipkin + ipnuc)
    hd = []
    for A in range(b.natm):
        p0, p1 = b.aoslices[A][2:4]
        v = -b.charges[A] * iprinv[A].copy()
        v[:, p0:p1] += h1[:, p0:p1]
        hd.append(v + v.transpose(0, 2, 1))
    eri, ip1 = int2e_and_ip1(b, pr, want_ip1)
    e_nuc, g_nuc = 0.0, np.zeros((b.natm, 3))
    for i in range(b.natm):
        for j in range(b.natm):
            if i == j:
                continue
            d = b.coords[i] - b.coords[j]
            r = np.linalg.norm(d)
            if j < i:
                e_nuc += b.charges[i] * b.charges[j] / r
            g_nuc[i] -= b.charges[i] * b.charges[j] * d / r ** 3
    return dict(ovlp=S, hcore=hcore, eri=eri, ipovlp=ipovlp, hcore_deriv=np.array(hd), eri_ip1=ip1,
                aoslices=b.aoslices, e_nuc=e_nuc, grad_nuc=g_nuc)


def int2e_and_ip1(b, pr=None, want_ip1=True):
    """(ab|cd) and (nabla a b|cd) over contracted AOs.  Loops over the distinct power 4-tuples; each is
    evaluated for the primitives carrying those powers, vectorised over the primitive quartets."""
    pr = pr or _Pairs(b)
    n, npr = b.nao, b.nprim
    pows = [tuple(x) for x in b.p_pow]
    upow = sorted(set(pows))
    idx = {pw: np.array([k for k in range(npr) if pows[k] == pw]) for pw in upow}
    eri = np.zeros((n, n, n, n))
    ip1 = np.zeros((3, n, n, n, n)) if want_ip1 else None

    def prim_block(pa, pb, pc, pd, ia, ib, ic, id_):
        """[pa pb|pc pd] for the primitive index sets (as if they carried those powers)."""
        sel_ab = np.ix_(ia, ib)
        sel_cd = np.ix_(ic, id_)
        p = pr.p[sel_ab][:, :, None, None]
        q = pr.p[sel_cd][None, None, :, :]
        PQ = pr.P[sel_ab][:, :, None, None, :] - pr.P[sel_cd][None, None, :, :, :]
        rho = p * q / (p + q)
        L = sum(pa) + sum(pb) + sum(pc) + sum(pd)
        R = _R(L, rho, PQ)
        Eab = {}
        for t in range(pa[0] + pb[0] + 1):
            for u in range(pa[1] + pb[1] + 1):
                for v in range(pa[2] + pb[2] + 1):
                    Eab[(t, u, v)] = (pr.E(0, pa[0], pb[0], t) * pr.E(1, pa[1], pb[1], u)
                                      * pr.E(2, pa[2], pb[2], v))[sel_ab][:, :, None, None]
        acc = 0.0
        for t2 in range(pc[0] + pd[0] + 1):
            for u2 in range(pc[1] + pd[1] + 1):
                for v2 in range(pc[2] + pd[2] + 1):
                    ecd = ((-1.0) ** (t2 + u2 + v2) * pr.E(0, pc[0], pd[0], t2) * pr.E(1, pc[1], pd[1], u2)
                           * pr.E(2, pc[2], pd[2], v2))[sel_cd][None, None, :, :]
                    for (t, u, v), eab in Eab.items():
                        acc = acc + eab * ecd * R[(t + t2, u + u2, v + v2)]
        pref = 2.0 * np.pi ** 2.5 / (p * q * np.sqrt(p + q))
        return pref * pr.K[sel_ab][:, :, None, None] * pr.K[sel_cd][None, None, :, :] * acc

    for pa, pb, pc, pd in itertools.product(upow, repeat=4):
        ia, ib, ic, id_ = idx[pa], idx[pb], idx[pc], idx[pd]
        ca, cb, cc, cd = b.cmat[ia], b.cmat[ib], b.cmat[ic], b.cmat[id_]
        blk = prim_block(pa, pb, pc, pd, ia, ib, ic, id_)
        eri += np.einsum("ijkl,ia,jb,kc,ld->abcd", blk, ca, cb, cc, cd, optimize=True)
        if want_ip1:
            alpha = b.p_exp[ia]
            for x in range(3):
                acc = 0.0
                for fac, pw in _shift_terms(pa, alpha, x):
                    acc = acc + fac[:, None, None, None] * prim_block(pw, pb, pc, pd, ia, ib, ic, id_)
                ip1[x] -= np.einsum("ijkl,ia,jb,kc,ld->abcd", acc, ca, cb, cc, cd, optimize=True)
    return eri, ip1
