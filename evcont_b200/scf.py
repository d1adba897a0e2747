"""Closed-shell RHF for ``get_basis(mol, "canonical")`` (evcont/electron_integral_utils.py:103-106:
``scf.RHF(mol).scf(); basis = myhf.mo_coeff``), the reference's default ``cibasis``.

Roothaan iteration with Pulay DIIS.  The AO integrals come from the device engine (K9 through
``MolLite``), the Loewdin matrix from K3 and every Fock build ``F = hcore + J - K/2`` from
``evc_fock_rhf``; the n x n eigenproblem of each cycle and the DIIS mixing are solved on the host, as
the projected problem of the FCI Davidson solver is.  The canonical orbitals only serve as the
basis the FCI problem is solved in: the state that reaches the t-RDM stack is rotated to the OAO
basis by ``transform_ci`` (evcont/FCI_EVCont.py:79-85) and does not depend on which SCF solution
was found.
"""
import numpy as np

from ._lib import check
from .engine import _ptr, get_engine


class RHFResult:
    def __init__(self, e_tot, mo_energy, mo_coeff, converged, cycles):
        self.e_tot, self.mo_energy, self.mo_coeff = e_tot, mo_energy, mo_coeff
        self.converged, self.cycles = converged, cycles


def _hcore(mol):
    if hasattr(mol, "get_hcore"):
        return mol.get_hcore()
    from pyscf import scf
    return scf.hf.get_hcore(mol)


def rhf(mol, conv_tol=1e-10, max_cycle=200, diis_space=8):
    """RHF of a closed-shell ``mol`` (``MolLite`` / ``ArrayMol`` / PySCF ``Mole``): :class:`RHFResult`."""
    nocc = int(mol.nelec[0])
    if int(mol.nelec[1]) != nocc:
        raise NotImplementedError("rhf: closed-shell molecules only")
    eng = get_engine()
    n = int(mol.nao)
    S = np.ascontiguousarray(mol.intor("int1e_ovlp"), dtype=np.float64)
    h = np.ascontiguousarray(_hcore(mol), dtype=np.float64)
    eri = eng.to_device(np.ascontiguousarray(mol.intor("int2e"), dtype=np.float64).reshape(n, n, n, n))
    hd = eng.to_device(h)
    X = eng.loewdin(eng.to_device(S)[None])[0][0].cpu().numpy()
    fock_d, dm_d = eng.empty(n, n), eng.empty(n, n)
    enuc = float(mol.energy_nuc())

    def fock(dm):
        dm_d.copy_(eng.to_device(dm))
        eng._bind_stream()
        check(eng.lib.evc_fock_rhf(eng._ctx, n, _ptr(hd), _ptr(eri), _ptr(dm_d), _ptr(fock_d)))
        return fock_d.cpu().numpy()

    def solve(F):
        w, c = np.linalg.eigh(X.T @ F @ X)
        c = X @ c
        return w, c, 2.0 * c[:, :nocc] @ c[:, :nocc].T

    w, c, dm = solve(h)           # core-Hamiltonian guess
    e_old, fs, errs = None, [], []
    converged = False
    for cycle in range(1, max_cycle + 1):
        F = fock(dm)
        e = 0.5 * float(np.sum(dm * (h + F))) + enuc
        err = X.T @ (F @ dm @ S - S @ dm @ F) @ X
        fs.append(F)
        errs.append(err)
        if len(fs) > diis_space:
            fs.pop(0)
            errs.pop(0)
        if e_old is not None and abs(e - e_old) < conv_tol and np.abs(err).max() < np.sqrt(conv_tol):
            converged = True
            break
        e_old = e
        if len(fs) > 1:
            m = len(fs)
            B = -np.ones((m + 1, m + 1))
            B[m, m] = 0.0
            for i in range(m):
                for j in range(m):
                    B[i, j] = np.sum(errs[i] * errs[j])
            rhs = np.zeros(m + 1)
            rhs[m] = -1.0
            try:
                coef = np.linalg.solve(B, rhs)[:m]
                F = sum(ci * Fi for ci, Fi in zip(coef, fs))
            except np.linalg.LinAlgError:
                pass
        w, c, dm = solve(F)
    w, c, _ = solve(fock(dm))
    return RHFResult(e, w, c, converged, cycle)
