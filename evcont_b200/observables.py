"""Observables of the predicted one-body density matrix: what the reference's
MD callback evaluates every step from ``scanner.base.predicted_one_rdm``
(scripts/MD/Zundel_thermodynamics/continuation/04_Zundel_continuation_MD.py:71-92 ``dip_moment``,
:140-159 ``callback``), on the device and batched over geometries.

``dip_moment(mol, dm, unit)`` and ``atomic_charges(mol, dm)`` are the per-molecule calls of the reference's
callback (same arguments: the density matrix in the AO basis).  ``predicted_observables`` is the batched form:
the continuation's density matrices at G geometries (one batched prediction, no n^4 arrays), their dipole
moments and atomic charges in one pass -- a whole trajectory's frames at once instead of one callback per step.

Atomic charges: the reference calls ``pyscf.scf.hf.mulliken_meta`` (populations in meta-Loewdin orthogonalised
AOs whose pre-orthogonalisation projects PySCF's ANO tables).  Those tables are not available here; the two
populations that need no external data are provided: ``"mulliken"`` (``(dm S)_mu,mu``, ``hf.mulliken_pop``)
and ``"loewdin"`` (``(S^1/2 dm S^1/2)_mu,mu`` -- what meta-Loewdin reduces to when every AO is a valence
function, e.g. hydrogen chains in a minimal basis).
"""
import numpy as np

AU2DEBYE = 2.541746473  # pyscf.data.nist.AU2DEBYE


def _table(mol, eng):
    return eng.aotable([mol.atom_symbol(i) for i in range(mol.natm)], mol.basis)


def dip_moment(mol, dm, unit="Debye"):
    """Dipole moment (3,) of the AO density matrix ``dm`` about the centre of mass
    (04_Zundel_continuation_MD.py:71-92)."""
    from .engine import get_engine
    eng = get_engine()
    dm = np.asarray(dm)
    if dm.ndim != 2:          # UHF density matrices
        dm = dm[0] + dm[1]
    t = _table(mol, eng)
    coords = mol.atom_coords()[None]
    origin = eng.center_of_mass(t, coords)
    r = eng.int1e_r(t, coords, origin)[0].cpu().numpy()
    el = np.einsum("xij,ji->x", r, dm).real
    nuc = np.einsum("i,ix->x", t.tables["charges"], mol.atom_coords() - origin[0].cpu().numpy())
    d = nuc - el
    return d * AU2DEBYE if unit.upper() == "DEBYE" else d


def atomic_charges(mol, dm, method="mulliken"):
    """Atomic charges (natm,) of the AO density matrix ``dm``: ``Z_A - sum_{mu on A} n_mu``."""
    dm = np.asarray(dm)
    if dm.ndim != 2:
        dm = dm[0] + dm[1]
    S = mol.intor("int1e_ovlp")
    if method.lower() == "mulliken":
        pop = np.einsum("ij,ji->i", dm, S).real
    elif method.lower() in ("loewdin", "lowdin"):
        w, V = np.linalg.eigh(S)
        Sh = (V * np.sqrt(w)) @ V.T
        pop = np.einsum("ij,jk,ki->i", Sh, dm, Sh).real
    else:
        raise ValueError("method: 'mulliken' or 'loewdin'")
    sl = mol.aoslice_by_atom()
    return np.array([mol.atom_charges()[A] - pop[sl[A, 2]:sl[A, 3]].sum() for A in range(mol.natm)])


def predicted_observables(mol, geometries, one_rdm, two_rdm, overlap, method="mulliken", unit="Debye",
                          return_dm_ao=False):
    """Dipole moments (G, 3) and atomic charges (G, natm) of the continuation's predicted density matrices at
    ``geometries`` (G, natm, 3), bohr -- the reference's per-step callback for every frame at once:
    integrals (K9), Loewdin (K3), subspace problem (K4-K6), gamma = sum_ab c_a c_b one_rdm[a, b],
    dm_ao = X gamma X^T, then ``evc_rdm1_observables``."""
    import torch
    from .stackcache import as_device_stack
    stack = as_device_stack(one_rdm, two_rdm, overlap)
    eng = stack.engine
    geometries = np.ascontiguousarray(geometries, dtype=np.float64).reshape(-1, mol.natm, 3)
    coords = eng.to_device(geometries)
    ao = eng.ao_integrals(mol.sbasis(eng), coords)
    _, cvec = eng.energies(stack, ao)
    N, n = stack.ntrain, stack.norb
    cc = (cvec[:, :, None] * cvec[:, None, :]).reshape(-1, N * N)
    gamma = (cc @ stack.one_rdm).reshape(-1, n, n)
    x, _, _ = eng.loewdin(ao.ovlp)
    out = eng.rdm1_observables(_table(mol, eng), coords, x, gamma, ao.ovlp, method=method, want_dm_ao=return_dm_ao)
    dip = out[0].cpu().numpy() * (AU2DEBYE if unit.upper() == "DEBYE" else 1.0)
    chg = out[1].cpu().numpy()
    return (dip, chg, out[2].cpu().numpy()) if return_dm_ao else (dip, chg)


def write_observables(dipoles, charges, dipole_file="dipole_moment_continuation.txt",
                      charges_file="atom_charges_continuation.txt", mode="a"):
    """The callback's two text files (04_Zundel_continuation_MD.py:150-159): one line per step, values
    separated by two blanks."""
    with open(dipole_file, mode) as fl:
        for d in np.atleast_2d(dipoles):
            fl.write("".join("{}  ".format(el) for el in d) + "\n")
    with open(charges_file, mode) as fl:
        for q in np.atleast_2d(charges):
            fl.write("".join("{}  ".format(el) for el in q) + "\n")
