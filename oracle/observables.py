"""Dipole moment and atomic charges of an AO density matrix -- CPU oracle for evcont_b200/observables.py and
csrc/observables.cu.  TEST INFRASTRUCTURE ONLY.

Follows the reference's MD callback (scripts/MD/Zundel_thermodynamics/continuation/
04_Zundel_continuation_MD.py:71-92 ``dip_moment``: ``mol.intor_symmetric("int1e_r", comp=3)`` inside
``mol.with_common_orig(centre of mass)``, ``el_dip = einsum("xij,ji->x", ao_dip, dm)``,
``mol_dip = sum_A Z_A (R_A - R_com) - el_dip``, Debye via ``nist.AU2DEBYE``; :140-159 ``callback``:
``dm_ao = X gamma X^T`` with ``X = get_basis(mol)``).

``int1e_r`` is built from the Hermite-expansion overlaps of oracle/integrals_sp.py:
``<a| x - O_x |b> = <a + 1_x | b> + (A_x - O_x) <a|b>``.  libcint is absent: the integrals are pinned by
tests/test_oracle_observables.py (numerical quadrature, the translation rule r(O') = r(O) - (O' - O) S,
hermiticity, the dipole of a point-symmetric density).  **Parity unpinned w.r.t. the libcint binary.**

Atomic charges: the reference's ``hf.mulliken_meta`` needs PySCF's ANO tables (absent); ``mulliken`` is
``hf.mulliken_pop`` (``n_mu = (dm S)_mu,mu``), ``loewdin`` the symmetric-orthogonalisation populations.
"""
import numpy as np

from . import integrals_sp as osp

AU2DEBYE = 2.541746473
MASSES = {"H": 1.00782503223, "He": 4.00260325413, "O": 15.99491461957}


def center_of_mass(b):
    m = np.array([MASSES[s] for s in b.symbols])
    return (m[:, None] * b.coords).sum(0) / m.sum()


def int1e_r(b, origin):
    """(3, nao, nao): <i| r - origin |j> for an ``oracle.integrals_sp.SPBasis``."""
    origin = np.asarray(origin, dtype=np.float64)
    pr, assemble, ovl, _, _ = osp._one_electron(b)
    A = b.centers
    out = []
    for x in range(3):
        def fn(pa, pb, x=x):
            up = list(pa); up[x] += 1
            return ovl(tuple(up), pb) + (A[:, x] - origin[x])[:, None] * ovl(pa, pb)
        out.append(np.einsum("ij,ia,jb->ab", assemble(fn, False), b.cmat, b.cmat, optimize=True))
    return np.array(out)


def dip_moment(b, dm, unit="Debye"):
    com = center_of_mass(b)
    r = int1e_r(b, com)
    el = np.einsum("xij,ji->x", r, dm)
    nuc = np.einsum("i,ix->x", b.charges, b.coords - com)
    d = nuc - el
    return d * AU2DEBYE if unit.upper() == "DEBYE" else d


def atomic_charges(b, dm, S, method="mulliken"):
    if method == "mulliken":
        pop = np.einsum("ij,ji->i", dm, S)
    else:
        w, V = np.linalg.eigh(S)
        Sh = (V * np.sqrt(w)) @ V.T
        pop = np.diag(Sh @ dm @ Sh)
    return np.array([b.charges[A] - pop[b.aoslices[A][2]:b.aoslices[A][3]].sum() for A in range(b.natm)])
