"""BASELINE configs[2] at full size without PySCF: water / 6-31G, 13 orbitals, (5, 5) electrons, 1 656 369
determinants.  RHF -> device FCI (canonical basis) -> transform_ci -> transition RDMs -> prediction.  The dense
Hamiltonian does not exist at this size; the checks are the size-independent identities the path offers:
the t-RDM contraction <c|H|c> reproduces the Davidson energy (sigma kernel vs RDM kernel), the state does not
depend on the basis it was solved in, sum rules of the RDMs, and the forces are the derivative of the energy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ANG = 1.0 / 0.52917721092


def _water(scale=1.0):
    from evcont_b200.mol import MolLite
    r, th = 0.9572 * ANG * scale, np.deg2rad(104.52)
    return MolLite([("O", (0.0, 0.0, 0.0)), ("H", (r * np.sin(th / 2), 0.0, r * np.cos(th / 2))),
                    ("H", (-r * np.sin(th / 2), 0.0, r * np.cos(th / 2)))], basis="6-31g", unit="Bohr")


@pytest.fixture(scope="module")
def trained():
    from evcont_b200.FCI_EVCont import FCI_EVCont_obj
    cont = FCI_EVCont_obj()                 # the reference's default: cibasis="canonical"
    for s in (1.0, 1.25):
        cont.append_to_rdms(_water(s))
    return cont


def test_fci_energy_below_rhf_and_reproduced_by_the_rdms(trained):
    from evcont_b200.ab_initio_eigenvector_continuation import approximate_ground_state_OAO
    from evcont_b200.scf import rhf
    cont = trained
    assert cont.two_rdm.shape == (2, 2, 13, 13, 13, 13) and cont.fcivecs[0].shape == (1287, 1287)
    e_hf = rhf(_water(1.0)).e_tot
    # FCI/6-31G correlation energy of water at equilibrium is about -0.13 Ha
    assert -0.16 < cont.ens[0] - e_hf < -0.10
    for k, s in enumerate((1.0, 1.25)):
        mol = _water(s)
        # diagonal block alone: <c_k|H(R_k)|c_k> through h1.dm1 + h2.dm2/2 equals the Davidson eigenvalue
        e_k, _ = approximate_ground_state_OAO(mol, cont.one_rdm[k:k + 1, k:k + 1], cont.two_rdm[k:k + 1, k:k + 1],
                                              cont.overlap[k:k + 1, k:k + 1])
        assert abs(e_k - cont.ens[k]) < 1e-8
        # and the two-state continuation is exact at its training points
        e2, _ = approximate_ground_state_OAO(mol, cont.one_rdm, cont.two_rdm, cont.overlap)
        assert -1e-9 < cont.ens[k] - e2 < 1e-8


def test_rdm_sum_rules_and_overlaps(trained):
    cont = trained
    ne = 10
    for a in range(2):
        for b in range(2):
            ov = cont.overlap[a, b]
            assert abs(ov - np.vdot(cont.fcivecs[a], cont.fcivecs[b])) < 1e-12
            assert abs(np.trace(cont.one_rdm[a, b]) - ne * ov) < 1e-10
            assert abs(np.einsum("pprr->", cont.two_rdm[a, b]) - ne * (ne - 1) * ov) < 1e-9
            assert np.abs(np.einsum("pqrr->pq", cont.two_rdm[a, b]) - (ne - 1) * cont.one_rdm[a, b].T).max() < 1e-10
    assert abs(cont.overlap[0, 0] - 1) < 1e-12 and 0.5 < abs(cont.overlap[1, 0]) < 1.0


def test_forces_are_the_derivative_of_the_energy(trained):
    from evcont_b200.ab_initio_gradients_loewdin import get_energy_with_grad
    cont = trained
    mol = _water(1.1)
    e0, grad = get_energy_with_grad(mol, cont.one_rdm, cont.two_rdm, cont.overlap)
    co, h = mol.atom_coords(), 1e-3
    for (A, x) in ((0, 2), (1, 0), (2, 2)):
        cp, cm = co.copy(), co.copy()
        cp[A, x] += h
        cm[A, x] -= h
        ep, _ = get_energy_with_grad(mol.copy().set_geom_(cp), cont.one_rdm, cont.two_rdm, cont.overlap)
        em, _ = get_energy_with_grad(mol.copy().set_geom_(cm), cont.one_rdm, cont.two_rdm, cont.overlap)
        assert abs(grad[A, x] - (ep - em) / (2 * h)) < 1e-6
    assert np.abs(grad.sum(axis=0)).max() < 1e-8     # translational invariance
