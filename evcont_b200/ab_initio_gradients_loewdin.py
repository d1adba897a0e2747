"""Energy + nuclear gradient in the Loewdin basis
(evcont/ab_initio_gradients_loewdin.py), on the B200.

``get_energy_with_grad`` is the hot entry point (one call per MD step in the
reference, evcont/MD_utils.py:43): here it is ONE fused device step (K3..K8);
:func:`get_energy_with_grad_batch` exposes the same step for many geometries at
once, which is how the GPU is actually kept busy.
"""
import numpy as np
import torch

from .engine import DeviceAO, HostAO, get_engine
from .mol import MolLite, ao_bundle
from .stackcache import as_device_stack


def get_overlap_grad(mol):
    """dS/dR as ``(n, n, natm, 3)`` from ``int1e_ipovlp`` (:13-38).  Pure index
    bookkeeping on a 3 n^2 array; done on the host."""
    ip = np.asarray(mol.intor("int1e_ipovlp", comp=3))
    n, natm = int(mol.nao), int(mol.natm)
    d = np.zeros((3, natm, n, n))
    for A, s in enumerate(mol.aoslice_by_atom()):
        d[:, A, s[2]:s[3], :] -= ip[:, s[2]:s[3], :]
    d = d + d.transpose(0, 1, 3, 2)
    return d.transpose(2, 3, 1, 0)


def loewdin_trafo_grad(overlap_mat):
    """``dX_kl / dS_ab`` as ``(n, n, n, n)`` [a, b, k, l] (:41-112).

    GPU: ``dX = V (G o (V^T E_ab V)) V^T`` for the n^2 symmetrised unit
    perturbations ``E_ab`` with exact divided differences ``G`` (K3).  The
    reference instead treats eigenvalues sharing a ``round(.,5)`` bucket as
    degenerate; the two agree unless two eigenvalues of S fall within 1e-5 of
    each other without being equal, where the reference is the approximate one.
    """
    eng = get_engine()
    s = eng.to_device(np.asarray(overlap_mat, dtype=np.float64))[None]
    n = s.shape[-1]
    _, evals, evecs = eng.loewdin(s)
    eye = np.eye(n)
    unit = 0.5 * (np.einsum("ak,bl->abkl", eye, eye) + np.einsum("bk,al->abkl", eye, eye))
    dX = eng.loewdin_grad(evals, evecs, eng.to_device(unit.reshape(1, n * n, n, n)))
    return dX[0].cpu().numpy().reshape(n, n, n, n)


def get_derivative_ao_mo_trafo(mol):
    """dX/dR as ``(n, n, natm, 3)`` (:115-134), without the n^4 intermediate."""
    eng = get_engine()
    n, natm = int(mol.nao), int(mol.natm)
    s = eng.to_device(np.asarray(mol.intor("int1e_ovlp"), dtype=np.float64))[None]
    _, evals, evecs = eng.loewdin(s)
    dS = get_overlap_grad(mol).transpose(2, 3, 0, 1).reshape(1, natm * 3, n, n)
    dX = eng.loewdin_grad(evals, evecs, eng.to_device(np.ascontiguousarray(dS)))
    return np.ascontiguousarray(dX[0].cpu().numpy().reshape(natm, 3, n, n).transpose(2, 3, 0, 1))


def get_one_el_grad_ao(mol):
    """Core-Hamiltonian derivative in the AO basis, ``(n, n, natm, 3)`` (:137-152)."""
    return np.ascontiguousarray(ao_bundle(mol)["hcore_deriv"].transpose(2, 3, 0, 1))


def _device_inputs(mol, one_RDM, two_RDM, S):
    stack = as_device_stack(one_RDM, two_RDM, S)
    if isinstance(mol, MolLite):
        # integrals on the device straight from the coordinates (K9): nothing but
        # natm x 3 doubles crosses PCIe
        eng = stack.engine
        ao = eng.ao_integrals(mol.sbasis(eng), mol.atom_coords()[None])
    else:
        ao = DeviceAO.from_bundles(stack.engine, [ao_bundle(mol)])
    return stack, ao


def get_grad_elec_OAO(mol, one_rdm, two_rdm, ao_mo_trafo=None, ao_mo_trafo_grad=None):
    """Electronic gradient ``(natm, 3)`` from 1-/2-RDMs given in the OAO basis
    (:255-305).  GPU: K4 (three-quarter transform) + K8 in adjoint form.  The
    transformation is always the Loewdin one; ``ao_mo_trafo`` is checked against
    it and ``ao_mo_trafo_grad`` is not needed."""
    eng = get_engine()
    b = ao_bundle(mol)
    n = b["nao"]
    dev = {k: eng.to_device(b[k])[None] for k in ("ovlp", "hcore", "eri", "ipovlp", "hcore_deriv",
                                                  "eri_ip1")}
    x, evals, evecs = eng.loewdin(dev["ovlp"])
    if ao_mo_trafo is not None and not np.allclose(ao_mo_trafo, x[0].cpu().numpy(), atol=1e-9):
        raise NotImplementedError("get_grad_elec_OAO supports the Loewdin transformation only")
    _, _, t3 = eng.ao2oao(None, dev["eri"], x, want_t3=True)
    gamma = eng.to_device(np.asarray(one_rdm, dtype=np.float64)).reshape(1, n, n)
    Gamma = eng.to_device(np.asarray(two_rdm, dtype=np.float64)).reshape(1, n, n, n, n)
    aosl = torch.from_numpy(np.ascontiguousarray(b["aoslices"], dtype=np.int32)).to(eng.device)
    grad = eng.grad_elec(aosl, evals, evecs, x, dev["hcore"], t3, gamma, Gamma, dev["ipovlp"],
                         dev["hcore_deriv"], dev["eri_ip1"])
    return grad[0].cpu().numpy()


def get_energy_with_grad(mol, one_RDM, two_RDM, S, hermitian=True, return_density_matrices=False):
    """Potential energy and nuclear gradient from the continuation (:308-379).

    Returns ``(E + E_nuc, grad + grad_nuc)`` and, if asked, the predicted
    ``(gamma, Gamma)``.  All four ``two_RDM`` layouts are accepted."""
    if hermitian is not True:
        raise NotImplementedError("hermitian=False is not implemented on the device")
    stack, ao = _device_inputs(mol, one_RDM, two_RDM, S)
    E, grad, gamma, Gamma, _ = stack.engine.energy_with_grad(stack, ao, want_rdms=return_density_matrices)
    out = (float(E[0].item()), grad[0].cpu().numpy())
    if return_density_matrices:
        out = out + (gamma[0].cpu().numpy(), Gamma[0].cpu().numpy())
    return out


def get_energy_with_grad_batch(mols, one_RDM, two_RDM, S, return_density_matrices=False):
    """The same step for a list of geometries in one launch sequence:
    ``(E[G], grad[G, natm, 3][, gamma[G], Gamma[G]])``."""
    stack = as_device_stack(one_RDM, two_RDM, S)
    bundles = [ao_bundle(m) for m in mols]
    if not return_density_matrices:
        # host arrays in, host arrays out: the chunked copy/compute pipeline
        host = HostAO.from_bundles(bundles)
        E, grad = stack.engine.energy_with_grad_host(stack, host)
        return E.numpy().copy(), grad.numpy().copy()
    ao = DeviceAO.from_bundles(stack.engine, bundles)
    E, grad, gamma, Gamma, _ = stack.engine.energy_with_grad(stack, ao, want_rdms=True)
    return E.cpu().numpy(), grad.cpu().numpy(), gamma.cpu().numpy(), Gamma.cpu().numpy()


def get_energy_with_grad_coords(mol, coords, one_RDM, two_RDM, S):
    """The step for many geometries of one molecule given as coordinates only:
    ``mol`` (:class:`evcont_b200.mol.MolLite`) fixes atoms and basis, ``coords`` is
    ``(G, natm, 3)`` in bohr.  Integrals (K9) and prediction (K3..K8) run on the device;
    returns ``(E[G], grad[G, natm, 3])`` as numpy arrays."""
    stack = as_device_stack(one_RDM, two_RDM, S)
    eng = stack.engine
    coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, mol.natm, 3)
    E, grad, _, _, _ = eng.energy_with_grad_coords(stack, mol.sbasis(eng), coords)
    return E.cpu().numpy(), grad.cpu().numpy()
