"""Subspace generalized eigenproblem (evcont/ab_initio_eigenvector_continuation.py).

``approximate_ground_state`` / ``approximate_multistate`` and their ``_OAO``
wrappers with the reference's signatures.  H assembly is K5 (one streaming pass
over the stack), the eigenproblem is K6 (Cholesky of S once per stack + batched
Jacobi).  ``hermitian=False`` (scipy ``eig``) is only reachable from the DMRG
builder in the reference (DMRG_EVCont.py:99) and is not implemented.
"""
import numpy as np

from .electron_integral_utils import get_basis, get_integrals  # noqa: F401 (re-export, see
#   scripts/MD/Zundel_thermodynamics/continuation/05_Zundel_test_potential_energy.py:7)
from .engine import get_engine
from .stackcache import as_device_stack


def _solve(h1, h2, one_RDM, two_RDM, S, nroots, hermitian):
    if hermitian is not True:
        raise NotImplementedError(
            "hermitian=False (scipy.linalg.eig) is not on the FCI continuation path and is "
            "not implemented on the device")
    stack = as_device_stack(one_RDM, two_RDM, S)
    eng = stack.engine
    n = stack.norb
    h1d = eng.to_device(np.asarray(h1, dtype=np.float64)).reshape(1, n, n)
    h2d = eng.to_device(np.asarray(h2, dtype=np.float64)).reshape(1, n, n, n, n)
    H = eng.subspace_H(stack, h1d, h2d)
    E, C = eng.geneig(H, stack.linv, nroots)
    return E[0].cpu().numpy(), C[0].cpu().numpy()


def approximate_ground_state(h1, h2, one_RDM, two_RDM, S, hermitian=True):
    """``(E0, c0)``: lowest root of ``H c = E S c`` with
    ``H = one_RDM.h1 + 1/2 two_RDM.h2`` (evcont/...continuation.py:12-90)."""
    E, C = _solve(h1, h2, one_RDM, two_RDM, S, 1, hermitian)
    return float(E[0]), C[0]


def approximate_multistate(h1, h2, one_RDM, two_RDM, S, nroots=1, hermitian=True):
    """The ``nroots`` lowest roots: ``(E[nroots], C[nroots, N])`` (:93-175)."""
    N = np.asarray(S).shape[0]
    assert N >= nroots
    return _solve(h1, h2, one_RDM, two_RDM, S, nroots, hermitian)


def approximate_ground_state_OAO(mol, one_RDM, two_RDM, S, hermitian=True):
    """Total energy (with nuclear repulsion) and subspace vector at ``mol`` (:178-211)."""
    h1, h2 = get_integrals(mol, get_basis(mol))
    en, vec = approximate_ground_state(h1, h2, one_RDM, two_RDM, S, hermitian=hermitian)
    return en + mol.energy_nuc(), vec


def approximate_multistate_OAO(mol, one_RDM, two_RDM, S, nroots=1, hermitian=True):
    """Multi-root version of :func:`approximate_ground_state_OAO` (:214-250)."""
    h1, h2 = get_integrals(mol, get_basis(mol))
    en, vec = approximate_multistate(h1, h2, one_RDM, two_RDM, S, nroots=nroots,
                                     hermitian=hermitian)
    return en + mol.energy_nuc(), vec
