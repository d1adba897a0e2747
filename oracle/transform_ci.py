"""CI-vector change of the one-particle basis (CPU oracle).

Restates ``pyscf.fci.addons.transform_ci(ci, nelec, u)`` which the reference calls at
evcont/FCI_EVCont.py:79-85 to bring FCI vectors solved in the canonical / split basis
into the OAO basis (``u = basis^T S basis_oao``, rows = old orbitals, columns = new
orbitals).  PySCF (third-party, unpinned in pyproject.toml:10) is absent from this
image, so this is **parity unpinned w.r.t. the PySCF binary**; its published algorithm:

    trans_a[I, J] = det( u[occ(I), :][:, occ(J)] )      (nalpha x nalpha minors; same for beta)
    ci_new        = trans_a^T . ci . trans_b

which is the statement  |J>_new = sum_I det(u[I, J]) |I>_old  for Slater determinants
built from  a^+_{new, j} = sum_i u[i, j] a^+_{old, i}.  Pinned by
tests/test_oracle_transform_ci.py: the Leibniz-formula minors below (no LAPACK),
covariance of the RDMs (``dm1_new = u^T dm1_old u``), invariance of the energy under
the simultaneous rotation of the Hamiltonian, the group property
``T(u1 u2) = T(u2) o T(u1)`` and the identity rotation.

Test infrastructure only (see oracle/__init__.py).
"""
import itertools

import numpy as np

from . import cistring


def _occ_lists(norb, nocc):
    strs = cistring.make_strings(norb, nocc)
    return [[o for o in range(norb) if (int(s) >> o) & 1] for s in strs]


def minors(u, norb, nocc):
    """``M[I, J] = det(u[occ(I), occ(J)])`` over all ``nocc``-electron strings (LAPACK det)."""
    u = np.asarray(u, dtype=np.float64)
    if nocc == 0:
        return np.ones((1, 1))
    occ = np.asarray(_occ_lists(norb, nocc))          # (ns, nocc)
    sub = u[occ[:, None, :, None], occ[None, :, None, :]]   # (ns, ns, nocc, nocc)
    return np.linalg.det(sub)


def minors_leibniz(u, norb, nocc):
    """The same minors from the permutation expansion (independent of LAPACK; small cases)."""
    u = np.asarray(u, dtype=np.float64)
    if nocc == 0:
        return np.ones((1, 1))
    occ = _occ_lists(norb, nocc)
    perms = []
    for p in itertools.permutations(range(nocc)):
        inv = sum(1 for a in range(nocc) for b in range(a + 1, nocc) if p[a] > p[b])
        perms.append((p, -1.0 if inv & 1 else 1.0))
    out = np.zeros((len(occ), len(occ)))
    for i, oi in enumerate(occ):
        for j, oj in enumerate(occ):
            acc = 0.0
            for p, sgn in perms:
                term = sgn
                for r in range(nocc):
                    term *= u[oi[r], oj[p[r]]]
                acc += term
            out[i, j] = acc
    return out


def transform_ci(ci, nelec, u, minor_fn=minors):
    """``ci`` (na, nb) in the old basis -> the same state in the new basis (see module docstring)."""
    u = np.asarray(u, dtype=np.float64)
    norb = u.shape[0]
    assert u.shape == (norb, norb)
    if isinstance(nelec, (int, np.integer)):
        nb_ = int(nelec) // 2
        nelec = (int(nelec) - nb_, nb_)
    na, nb = cistring.num_strings(norb, nelec[0]), cistring.num_strings(norb, nelec[1])
    ci = np.asarray(ci, dtype=np.float64).reshape(na, nb)
    ta = minor_fn(u, norb, nelec[0])
    tb = ta if nelec[1] == nelec[0] else minor_fn(u, norb, nelec[1])
    return ta.T @ ci @ tb
