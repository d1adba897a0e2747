"""Device FCI (sigma vector, diagonal, Davidson) against the dense Slater-Condon Hamiltonian of the
oracle, and the reference's H6 workflow end to end without PySCF: train at three bond lengths
(scripts/PES_H_chain/H6_PES/H6_continuation.py: d = 1.0, 1.8, 2.6 bohr), predict the PES and forces."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sym_integrals(norb, seed):
    rng = np.random.default_rng(seed)
    h = rng.standard_normal((norb, norb))
    h = h + h.T
    e = rng.standard_normal((norb,) * 4)
    e = e + e.transpose(1, 0, 2, 3)
    e = e + e.transpose(0, 1, 3, 2)
    e = e + e.transpose(2, 3, 0, 1)
    return h, 0.2 * e


@pytest.mark.parametrize("norb,nelec", [(4, (2, 2)), (5, (3, 2)), (6, (3, 3)), (5, (1, 4)), (3, (3, 0))])
def test_sigma_and_diagonal_against_dense_hamiltonian(norb, nelec):
    import torch
    from evcont_b200.engine import get_engine
    from oracle import trans_rdm as otr
    h1, eri = _sym_integrals(norb, 3 + norb)
    H = otr.hamiltonian_matrix(h1, eri, norb, nelec)
    eng = get_engine()
    ham = eng.fci_hamiltonian(h1, eri, norb, nelec)
    assert ham.ndet == H.shape[0]
    c = np.random.default_rng(1).standard_normal(ham.ndet)
    sig = ham.contract(eng.to_device(c)).cpu().numpy()
    assert np.abs(sig - H @ c).max() < 1e-12 * max(1.0, np.abs(H @ c).max())
    assert np.abs(ham.hdiag().cpu().numpy() - np.diag(H)).max() < 1e-12
    sig2 = ham.contract(eng.to_device(c)).cpu().numpy()
    assert np.array_equal(sig, sig2)  # deterministic


def _h_chain(n, d):
    from evcont_b200.mol import MolLite
    xs = (np.arange(n) - np.median(np.arange(n))) * d
    return MolLite([("H", (x, 0.0, 0.0)) for x in xs], basis="sto-6g", unit="Bohr")


def _oao_integrals(mol):
    from evcont_b200.electron_integral_utils import get_basis, get_integrals
    return get_integrals(mol, get_basis(mol, "OAO"))


def test_davidson_h6_against_dense_diagonalisation():
    from evcont_b200.fci import B200FCISolver
    from oracle import trans_rdm as otr
    mol = _h_chain(6, 1.8)
    h1, h2 = _oao_integrals(mol)
    H = otr.hamiltonian_matrix(h1, h2, 6, (3, 3))
    w, v = np.linalg.eigh(H)
    solver = B200FCISolver()
    e0, c0 = solver.kernel(h1, h2, 6, (3, 3))
    assert c0.shape == (20, 20) and abs(np.linalg.norm(c0) - 1) < 1e-12
    assert abs(e0 - w[0]) < 1e-10
    assert abs(abs(c0.ravel() @ v[:, 0]) - 1) < 1e-9
    # excited states: the lowest eigenvalues whose vectors are symmetric under alpha <-> beta exchange
    symm = [k for k in range(len(w)) if np.abs(v[:, k].reshape(20, 20) - v[:, k].reshape(20, 20).T).max() < 1e-8]
    es, cs = solver.kernel(h1, h2, 6, (3, 3), nroots=3)
    assert np.abs(np.array(es) - w[symm[:3]]).max() < 1e-9
    assert all(np.abs(c - c.T).max() < 1e-10 for c in cs)


def test_h6_workflow_without_pyscf():
    """append_to_rdms(MolLite) x 3 -> exactness at the training points, PES error inside the
    training range, forces = finite differences of the exact FCI energy at a training point."""
    from evcont_b200.FCI_EVCont import FCI_EVCont_obj
    from evcont_b200.ab_initio_eigenvector_continuation import approximate_ground_state_OAO
    from evcont_b200.ab_initio_gradients_loewdin import get_energy_with_grad
    from oracle import trans_rdm as otr

    def exact(mol):
        h1, h2 = _oao_integrals(mol)
        return np.linalg.eigvalsh(otr.hamiltonian_matrix(h1, h2, 6, (3, 3)))[0] + mol.energy_nuc()

    cont = FCI_EVCont_obj(cibasis="OAO")
    train = [1.0, 1.8, 2.6]
    for d in train:
        cont.append_to_rdms(_h_chain(6, d))
    assert cont.overlap.shape == (3, 3) and cont.two_rdm.shape == (3, 3, 6, 6, 6, 6)
    for k, d in enumerate(train):
        mol = _h_chain(6, d)
        e_fci = exact(mol)
        assert abs(cont.ens[k] - e_fci) < 1e-9
        e_pred, _ = approximate_ground_state_OAO(mol, cont.one_rdm, cont.two_rdm, cont.overlap)
        assert abs(e_pred - e_fci) < 1e-8
    # in between: variational upper bound, close to the exact surface
    for d in (1.4, 2.2):
        mol = _h_chain(6, d)
        e_pred, grad = get_energy_with_grad(mol, cont.one_rdm, cont.two_rdm, cont.overlap)
        e_fci = exact(mol)
        assert -1e-10 < e_pred - e_fci < 2e-2
    # forces at a training point = derivative of the exact energy (Hellmann-Feynman in the subspace)
    mol = _h_chain(6, 1.8)
    _, grad = get_energy_with_grad(mol, cont.one_rdm, cont.two_rdm, cont.overlap)
    co, h = mol.atom_coords(), 1e-3
    for (A, x) in ((0, 0), (2, 0), (3, 1)):
        cp, cm = co.copy(), co.copy()
        cp[A, x] += h
        cm[A, x] -= h
        fd = (exact(mol.copy().set_geom_(cp)) - exact(mol.copy().set_geom_(cm))) / (2 * h)
        assert abs(grad[A, x] - fd) < 2e-5
