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


def test_davidson_does_not_depend_on_rounding_noise_to_find_a_root():
    """H6 at 1.4 bohr: the second root of the alpha <-> beta symmetric sector has the other inversion parity than the
    ground state.  A search space without that parity can only acquire it through rounding noise, so whether the
    root was found used to follow the last bits of the integrals.  Perturbing them at the 1e-14 level (different
    rounding paths, same spectrum to 1e-12) must not change which roots come back."""
    from evcont_b200.fci import B200FCISolver
    from oracle import trans_rdm as otr
    mol = _h_chain(6, 1.4)
    h1, h2 = _oao_integrals(mol)
    w, v = np.linalg.eigh(otr.hamiltonian_matrix(h1, h2, 6, (3, 3)))
    symm = [k for k in range(len(w)) if np.abs(v[:, k].reshape(20, 20) - v[:, k].reshape(20, 20).T).max() < 1e-7]
    rng = np.random.default_rng(0)
    for trial in range(6):
        e1 = rng.standard_normal(h1.shape) * 1e-14
        h1p = h1 + e1 + e1.T
        for nroots in (2, 3, 4):
            es, _ = B200FCISolver().kernel(h1p, h2, 6, (3, 3), nroots=nroots)
            assert np.abs(np.array(es) - w[symm[:nroots]]).max() < 1e-8, (trial, nroots, es, w[symm[:nroots + 2]])


@pytest.mark.parametrize("d", [1.0, 1.4, 1.8, 2.6, 3.0])
def test_davidson_excited_roots_at_every_training_distance(d):
    """nroots = 2 and 3 along the H6 scan: no root of the alpha <-> beta symmetric sector may be skipped."""
    from evcont_b200.fci import B200FCISolver
    from oracle import trans_rdm as otr
    mol = _h_chain(6, d)
    h1, h2 = _oao_integrals(mol)
    w, v = np.linalg.eigh(otr.hamiltonian_matrix(h1, h2, 6, (3, 3)))
    symm = [k for k in range(len(w)) if np.abs(v[:, k].reshape(20, 20) - v[:, k].reshape(20, 20).T).max() < 1e-7]
    for nroots in (2, 3):
        es, cs = B200FCISolver().kernel(h1, h2, 6, (3, 3), nroots=nroots)
        assert np.abs(np.array(es) - w[symm[:nroots]]).max() < 1e-8, (d, nroots, es, w[symm[:nroots + 2]])


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


@pytest.mark.parametrize("roots_train,nroots", [([1], 3), ([0, 1], 6)])
def test_h6_excited_state_workflow(roots_train, nroots):
    """scripts/PES_H_chain/H6_PES_excited/H6_continuation_excited.py: roots_train = [1] / [0, 1], three training
    distances, approximate_multistate_OAO with nroots = 3 / 6 -- against dense diagonalisation of the
    Slater-Condon Hamiltonian (the alpha <-> beta symmetric sector direct_spin0 solves in).
    (evcont/FCI_EVCont.py:43-47, 95-131: every trained root becomes its own training state.)"""
    from evcont_b200.FCI_EVCont import FCI_EVCont_obj
    from evcont_b200.ab_initio_eigenvector_continuation import approximate_multistate_OAO
    from oracle import trans_rdm as otr

    def exact_sym(mol, k):
        h1, h2 = _oao_integrals(mol)
        w, v = np.linalg.eigh(otr.hamiltonian_matrix(h1, h2, 6, (3, 3)))
        symm = [q for q in range(len(w)) if np.abs(v[:, q].reshape(20, 20) - v[:, q].reshape(20, 20).T).max() < 1e-7]
        return w[symm[:k]] + mol.energy_nuc()

    cont = FCI_EVCont_obj(roots_train=roots_train, cibasis="canonical")
    train = [1.0, 1.8, 2.6]
    for d in train:
        cont.append_to_rdms(_h_chain(6, d))
    nt = len(roots_train)
    N = nt * len(train)
    assert cont.overlap.shape == (N, N) and cont.two_rdm.shape == (N, N, 6, 6, 6, 6)
    assert len(cont.fcivecs) == N and len(cont.ens) == N
    assert cont.mol_index == [k for k in range(len(train)) for _ in range(nt)]
    # the stored energies are the FCI energies of the trained roots at their geometry
    for k, d in enumerate(train):
        ex = exact_sym(_h_chain(6, d), max(roots_train) + 1)
        for j, r in enumerate(roots_train):
            assert abs(cont.ens[k * nt + j] - ex[r]) < 1e-8
    # training vectors of one geometry are orthonormal; the stack reproduces its own training energies
    assert np.abs(np.diag(cont.overlap) - 1).max() < 1e-10
    if nt == 2:
        assert abs(cont.overlap[0, 1]) < 1e-9
    nr = min(nroots, N)
    for k, d in enumerate(train):
        mol = _h_chain(6, d)
        en, c = approximate_multistate_OAO(mol, cont.one_rdm, cont.two_rdm, cont.overlap, nroots=nr)
        assert en.shape == (nr,) and c.shape == (nr, N)
        ex = exact_sym(mol, max(roots_train) + 1)
        for r in roots_train:   # every trained root is an exact eigenvalue of the subspace problem there
            assert np.abs(en - ex[r]).min() < 1e-8
        assert np.all(np.diff(en) > -1e-12)
    # between the training points: variational (each Ritz value bounds the matching exact one from above)
    mol = _h_chain(6, 2.2)
    en, _ = approximate_multistate_OAO(mol, cont.one_rdm, cont.two_rdm, cont.overlap, nroots=nr)
    ex = exact_sym(mol, 1)
    assert en[0] > ex[0] - 1e-10
