"""Integral / basis utilities (evcont/electron_integral_utils.py), on the B200.

Same names and argument meaning as the reference; numpy in, numpy out.
"""
import numpy as np

from .engine import get_engine
from .mol import ao_bundle


def get_loewdin_trafo(overlap_mat):
    """``S^-1/2`` through an eigendecomposition, eigenvalues <= 1e-15 dropped
    (evcont/electron_integral_utils.py:6-18).  GPU: batched Jacobi (K3)."""
    eng = get_engine()
    s = eng.to_device(np.asarray(overlap_mat, dtype=np.float64))[None]
    x, _, _ = eng.loewdin(s)
    return x[0].cpu().numpy()


def transform_integrals(h1, h2, trafo):
    """Rotate ``h1, h2`` with ``trafo[a, i]`` (new index first), batched over leading
    axes (evcont/electron_integral_utils.py:21-35).

    Note: the reference's h1 line passes one ``trafo`` operand to a
    three-subscript einsum and therefore raises for every input; this
    implements the evident intent (``trafo`` on both indices).
    """
    eng = get_engine()
    h1 = np.asarray(h1, dtype=np.float64)
    h2 = np.asarray(h2, dtype=np.float64)
    trafo = np.asarray(trafo, dtype=np.float64)
    n = trafo.shape[0]
    if trafo.shape != (n, n):
        raise NotImplementedError("only square transformations are supported")
    lead = h2.shape[:-4]
    G = int(np.prod(lead)) if lead else 1
    c = eng.to_device(np.broadcast_to(trafo, (G, n, n)).copy())
    o1, o2, _ = eng.ao2oao(eng.to_device(h1.reshape(G, n, n)), eng.to_device(h2.reshape(G, n, n, n, n)),
                           c, transpose_c=True)
    return o1.cpu().numpy().reshape(lead + (n, n)), o2.cpu().numpy().reshape(lead + (n,) * 4)


def compress_electron_exchange_symmetry(h2, diag_multiplier=1.0):
    """Lower triangle (``np.tril_indices`` order) of ``h2`` viewed as ``(n^2, n^2)``,
    diagonal scaled (evcont/electron_integral_utils.py:38-66).  Host-side layout
    utility; the prediction step packs ``h2`` on the device itself."""
    h2 = np.asarray(h2)
    assert np.all(np.array(h2.shape) == h2.shape[0])
    n2 = h2.shape[0] ** 2
    m = h2.reshape(n2, n2)
    out = m[np.tril_indices(n2)].copy()
    if diag_multiplier != 1.0:
        k = np.arange(n2)
        out[k * (k + 1) // 2 + k] *= diag_multiplier
    return out


def restore_electron_exchange_symmetry(h2, norb):
    """Inverse of the compression, mirrored (evcont/electron_integral_utils.py:69-88)."""
    n2 = norb * norb
    m = np.zeros((n2, n2))
    il = np.tril_indices(n2)
    m[il] = h2
    m[(il[1], il[0])] = h2
    return m.reshape((norb,) * 4)


def get_basis(mol, basis_type="OAO"):
    """Orthogonal basis as AO coefficients (evcont/electron_integral_utils.py:91-119).
    ``"OAO"``: Loewdin on the GPU (K3).  ``"canonical"``: RHF orbitals -- PySCF's RHF for a PySCF
    ``Mole``, otherwise :func:`evcont_b200.scf.rhf` (Fock builds on the device).  ``"split"``
    (Boys-localised occupied / virtual blocks) needs PySCF's localiser."""
    if basis_type == "OAO":
        return get_loewdin_trafo(mol.intor("int1e_ovlp"))
    if basis_type == "canonical" and not type(mol).__module__.startswith("pyscf."):
        from .scf import rhf
        res = rhf(mol)
        if not res.converged:  # PySCF only warns here; an unconverged basis silently changes the FCI solve basis
            import warnings
            warnings.warn(f"RHF for the canonical basis did not converge in {res.cycles} cycles", RuntimeWarning)
        return res.mo_coeff
    try:
        from pyscf import lo, scf
    except ImportError as exc:
        raise NotImplementedError(
            f"basis_type={basis_type!r} needs PySCF (Boys localisation); 'OAO' and 'canonical' "
            "are built in") from exc
    myhf = scf.RHF(mol)
    myhf.scf()
    basis = myhf.mo_coeff
    if basis_type == "split":
        parts = []
        for block in (basis[:, : mol.nelec[0]], basis[:, mol.nelec[0]:]):
            localizer = lo.Boys(mol, block)
            localizer.init_guess = None
            parts.append(localizer.kernel())
        basis = np.concatenate(parts, axis=1)
    else:
        assert basis_type == "canonical"
    return basis


def get_integrals(mol, basis):
    """``h1 = C^T hcore C``, ``h2 = (ij|kl)`` transformed with ``C`` on all four
    indices, unpacked (evcont/electron_integral_utils.py:122-138).  GPU: K4."""
    eng = get_engine()
    b = ao_bundle_light(mol)
    basis = np.asarray(basis, dtype=np.float64)
    n = b["eri"].shape[0]
    if basis.shape != (n, n):
        raise NotImplementedError("only square (nao x nao) bases are supported")
    h1, h2, _ = eng.ao2oao(eng.to_device(b["hcore"])[None], eng.to_device(b["eri"])[None],
                           eng.to_device(basis)[None])
    return h1[0].cpu().numpy(), h2[0].cpu().numpy()


def ao_bundle_light(mol):
    """hcore and eri only (what ``get_integrals`` needs)."""
    if hasattr(mol, "get_hcore") and not type(mol).__module__.startswith("pyscf."):
        hcore = mol.get_hcore()
    else:
        from pyscf import scf
        hcore = scf.hf.get_hcore(mol)
    n = int(mol.nao)
    eri = np.ascontiguousarray(mol.intor("int2e"), dtype=np.float64).reshape(n, n, n, n)
    return dict(hcore=np.ascontiguousarray(hcore, dtype=np.float64), eri=eri)


__all__ = ["get_loewdin_trafo", "transform_integrals", "compress_electron_exchange_symmetry",
           "restore_electron_exchange_symmetry", "get_basis", "get_integrals", "ao_bundle"]
