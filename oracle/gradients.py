"""Energy + Loewdin-basis nuclear gradient from a t-RDM stack -- CPU oracle.

numpy restatement of evcont/ab_initio_gradients_loewdin.py, function by function
(the algebra is kept literal -- including the n^4 derivative tensor of the
Loewdin transform and its rounding-bucket degenerate perturbation theory -- so
that it checks the GPU path's rearranged algebra rather than sharing it):

* :func:`get_overlap_grad`           <- :13-38
* :func:`loewdin_trafo_grad`         <- :41-112
* :func:`get_derivative_ao_mo_trafo` <- :115-134
* :func:`get_one_el_grad_ao`         <- :137-152
* :func:`get_one_el_grad`            <- :155-187
* :func:`two_el_grad`                <- :190-252
* :func:`get_grad_elec_OAO`          <- :255-305
* :func:`get_energy_with_grad`       <- :308-379

``mol`` is a duck type carrying AO arrays (``evcont_b200.mol.ArrayMol`` or any
object with the same methods): ``nao, natm, intor(name), get_hcore(),
hcore_generator(), grad_nuc(), energy_nuc(), aoslice_by_atom()`` -- the calls
the reference makes into PySCF (``scf.hf.get_hcore``, ``ao2mo.kernel``,
``grad.RHF(...).hcore_generator/grad_nuc``) are routed to those methods.

Pinned against the reference itself (imported under a pyscf stub) by
tests/golden/make_golden.py.  Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

from .subspace import (
    approximate_ground_state,
    ao_to_oao,
    get_loewdin_trafo,
    restore_exchange,
)


def get_overlap_grad(mol):
    """dS/dR as (n, n, natm, 3) from <nabla mu|nu> (``int1e_ipovlp``)."""
    ip = mol.intor("int1e_ipovlp", comp=3)
    n, natm = mol.nao, mol.natm
    d = np.zeros((3, natm, n, n))
    for A, (_, _, p0, p1) in enumerate(mol.aoslice_by_atom()):
        d[:, A, p0:p1, :] -= ip[:, p0:p1, :]
    d = d + d.transpose(0, 1, 3, 2)
    return d.transpose(2, 3, 1, 0)


def loewdin_trafo_grad(s_ao):
    """dX_kl / dS_ab as (n, n, n, n) [a, b, k, l]; degenerate PT on 1e-5 buckets."""
    w, v = np.linalg.eigh(s_ao)
    n = len(w)
    bucket = np.round(w, decimals=5)
    same = bucket[:, None] == bucket[None, :]
    # per perturbation (a,b): symmetrised projector  V_ab[i,j] = 1/2 (v_ai v_bj + v_bi v_aj)
    rot = np.zeros((n, n, n, n))
    for val in np.unique(bucket):
        ids = np.flatnonzero(bucket == val)
        sub = v[:, ids]
        pert = 0.5 * (np.einsum("ai,bj->abij", sub, sub) + np.einsum("bi,aj->abij", sub, sub))
        _, u = np.linalg.eigh(pert)
        rot[:, :, ids[:, None], ids[None, :]] = u
    vr = np.einsum("ij,abjk->abik", v, rot)                      # rotated eigenvectors per (a,b)
    a_idx = np.arange(n)
    va = vr[a_idx, :, a_idx, :]                                  # [a, b, i] = vr[a,b,a,i]
    vb = vr[:, a_idx, a_idx, :]                                  # [a, b, j] = vr[a,b,b,j]
    pert_r = 0.5 * (np.einsum("abi,abj->abij", va, vb) + np.einsum("abi,abj->abij", vb, va))
    gap = w[None, :] - w[:, None]                                # gap[i,j] = w_j - w_i
    z = np.zeros((n, n, n, n))
    z[:, :, ~same] = pert_r[:, :, ~same] / gap[~same]
    dvec = np.einsum("abij,abjk->abik", vr, z)
    dval = np.einsum("abii->abi", pert_r)
    keep = w > 1.0e-15
    f = np.where(keep, 1.0 / np.sqrt(np.where(keep, w, 1.0)), 0.0)
    df = np.where(keep, -0.5 / np.sqrt(np.where(keep, w, 1.0)) ** 3, 0.0)
    dX = (
        np.einsum("abij,abkj->abik", dvec * f, vr)
        + np.einsum("abij,abkj->abik", vr * (df * dval)[:, :, None, :], vr)
        + np.einsum("abij,abkj->abik", vr * f, dvec)
    )
    return dX.transpose(2, 3, 0, 1)


def get_derivative_ao_mo_trafo(mol):
    """dX/dR as (n, n, natm, 3)."""
    return np.einsum(
        "ijkl,ijmn->klmn", loewdin_trafo_grad(mol.intor("int1e_ovlp")), get_overlap_grad(mol)
    )


def get_one_el_grad_ao(mol):
    gen = mol.hcore_generator()
    return np.array([gen(A) for A in range(mol.natm)]).transpose(2, 3, 0, 1)


def get_one_el_grad(mol, ao_mo_trafo=None, ao_mo_trafo_grad=None):
    """d h1_OAO / dR as (n, n, natm, 3)."""
    x = get_loewdin_trafo(mol.intor("int1e_ovlp")) if ao_mo_trafo is None else ao_mo_trafo
    dx = get_derivative_ao_mo_trafo(mol) if ao_mo_trafo_grad is None else ao_mo_trafo_grad
    h = mol.get_hcore()
    part = np.einsum("ijkl,im,mn->jnkl", dx, h, x)
    part = part + part.swapaxes(0, 1)
    return part + np.einsum("ij,iklm,kn->jnlm", x, get_one_el_grad_ao(mol), x)


def two_el_grad(h2_ao, two_rdm, ao_mo_trafo, ao_mo_trafo_grad, h2_ao_deriv, atm_slices):
    """(natm, 3) two-electron part (before the factor 1/2)."""
    x, dx = ao_mo_trafo, ao_mo_trafo_grad
    gsym = (
        two_rdm
        + two_rdm.transpose(1, 0, 2, 3)
        + two_rdm.transpose(3, 2, 1, 0)
        + two_rdm.transpose(2, 3, 0, 1)
    )
    # sum_{ijkl,abcd} gsym[ijkl] (ab|cd) dX[a,i,A,x] X[b,j] X[c,k] X[d,l]
    t = np.einsum("abcd,dl->abcl", h2_ao, x)
    t = np.einsum("abcl,ck->abkl", t, x)
    t = np.einsum("abkl,bj->ajkl", t, x)
    y = np.einsum("ajkl,ijkl->ai", t, gsym)
    from_trafo = np.einsum("ai,aimn->mn", y, dx)

    g_ao = np.einsum("ijkl,ai->ajkl", two_rdm, x)
    g_ao = np.einsum("ajkl,bj->abkl", g_ao, x)
    g_ao = np.einsum("abkl,ck->abcl", g_ao, x)
    g_ao = np.einsum("abcl,dl->abcd", g_ao, x)
    g_ao_sym = (
        g_ao
        + g_ao.transpose(1, 0, 3, 2)
        + g_ao.transpose(2, 3, 0, 1)
        + g_ao.transpose(3, 2, 1, 0)
    )
    per_ao = np.einsum("nmbcd,abcd->nma", h2_ao_deriv, g_ao_sym)
    n = two_rdm.shape[0]
    acc = np.zeros((3, len(atm_slices), n, n))
    for A, (p0, p1) in enumerate(atm_slices):
        acc[:, A, p0:p1, :] -= per_ao[:, p0:p1, :]
    return from_trafo + np.einsum("nmbb->mn", acc)


def get_grad_elec_OAO(mol, one_rdm, two_rdm, ao_mo_trafo=None, ao_mo_trafo_grad=None):
    x = get_loewdin_trafo(mol.intor("int1e_ovlp")) if ao_mo_trafo is None else ao_mo_trafo
    dx = get_derivative_ao_mo_trafo(mol) if ao_mo_trafo_grad is None else ao_mo_trafo_grad
    h1_jac = get_one_el_grad(mol, ao_mo_trafo=x, ao_mo_trafo_grad=dx)
    slices = tuple((s[2], s[3]) for s in mol.aoslice_by_atom())
    two = two_el_grad(
        mol.intor("int2e"), two_rdm, x, dx, mol.intor("int2e_ip1", comp=3), slices
    )
    return np.einsum("ij,ijkl->kl", one_rdm, h1_jac) + 0.5 * two


def predict_rdms(vec, one_rdm_stack, two_rdm_stack, norb):
    """gamma, Gamma = (c (x) c) . stack for the four layouts (:343-361)."""
    cc = np.outer(vec, vec)
    gamma = np.tensordot(cc, one_rdm_stack, axes=2)
    if two_rdm_stack.ndim in (2, 5):
        wmat = 2.0 * cc
        wmat[np.diag_indices(len(vec))] *= 0.5
        Gamma = np.tensordot(wmat[np.tril_indices(len(vec))], two_rdm_stack, axes=1)
    else:
        Gamma = np.tensordot(cc, two_rdm_stack, axes=2)
    if Gamma.ndim != 4:
        Gamma = restore_exchange(Gamma, norb)
    return gamma, Gamma


def get_energy_with_grad(mol, one_RDM, two_RDM, S, hermitian=True, return_density_matrices=False):
    x = get_loewdin_trafo(mol.intor("int1e_ovlp"))
    h1, h2 = ao_to_oao(mol.get_hcore(), mol.intor("int2e"), x)
    en, vec = approximate_ground_state(h1, h2, one_RDM, two_RDM, S, hermitian=hermitian)
    gamma, Gamma = predict_rdms(vec, one_RDM, two_RDM, mol.nao)
    g_el = get_grad_elec_OAO(mol, gamma, Gamma, ao_mo_trafo=x)
    out = (en.real + mol.energy_nuc(), g_el + mol.grad_nuc())
    if return_density_matrices:
        out = out + (gamma, Gamma)
    return out
