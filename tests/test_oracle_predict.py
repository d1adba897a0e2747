"""Pins the CPU oracle of the prediction path (oracle/subspace.py, oracle/gradients.py)
against golden vectors produced by the REFERENCE ITSELF (tests/golden/make_golden.py,
run where /root/reference exists).  CPU-only."""
import numpy as np
import pytest

from conftest import PREDICT_CASES, load_predict_golden, synthetic_stack
from evcont_b200.mol import synthetic_mol
from oracle import gradients as og
from oracle import subspace as osub


@pytest.mark.parametrize("norb,natm,ntrain", PREDICT_CASES)
def test_integral_utils(norb, natm, ntrain):
    g = load_predict_golden(norb, natm, ntrain)
    mol = synthetic_mol(norb, natm, seed=int(g["seed"]))
    x = osub.get_loewdin_trafo(mol.intor("int1e_ovlp"))
    assert np.abs(x - g["loewdin_X"]).max() < 1e-13
    assert np.abs(x @ mol.intor("int1e_ovlp") @ x - np.eye(norb)).max() < 1e-12
    h1, h2 = osub.ao_to_oao(mol.get_hcore(), mol.intor("int2e"), x)
    assert np.abs(h1 - g["h1"]).max() < 1e-12 and np.abs(h2 - g["h2"]).max() < 1e-12
    t1, t2 = osub.transform_integrals(g["h1"], g["h2"], g["loewdin_X"])
    assert np.abs(t1 - g["trafo_h1"]).max() < 1e-11 and np.abs(t2 - g["trafo_h2"]).max() < 1e-11
    c = osub.compress_exchange(g["h2"], 0.5)
    assert np.array_equal(c, g["h2_compressed_half"])
    assert np.array_equal(osub.restore_exchange(osub.compress_exchange(g["h2"]), norb)
                          .reshape(norb * norb, -1)[np.tril_indices(norb * norb)],
                          g["h2"].reshape(norb * norb, -1)[np.tril_indices(norb * norb)])


@pytest.mark.parametrize("norb,natm,ntrain", PREDICT_CASES[:3])
def test_gradient_pieces(norb, natm, ntrain):
    g = load_predict_golden(norb, natm, ntrain)
    mol = synthetic_mol(norb, natm, seed=int(g["seed"]))
    dx = og.get_derivative_ao_mo_trafo(mol)
    assert np.abs(dx - g["dX_dR"]).max() < 1e-11
    h1j = og.get_one_el_grad(mol, g["loewdin_X"], dx)
    assert np.abs(h1j - g["h1_jac"]).max() < 1e-11


@pytest.mark.parametrize("layout", [6, 5, 3, 2])
@pytest.mark.parametrize("norb,natm,ntrain", PREDICT_CASES[:3])
def test_energy_with_grad(norb, natm, ntrain, layout):
    g = load_predict_golden(norb, natm, ntrain)
    mol = synthetic_mol(norb, natm, seed=int(g["seed"]))
    ovlp, one, two = synthetic_stack(norb, ntrain, int(g["seed"]) + 100, layout)
    e, c = osub.approximate_ground_state(g["h1"], g["h2"], one, two, ovlp)
    assert abs(e - float(g[f"L{layout}_E0"])) < 1e-11
    em, _ = osub.approximate_multistate(g["h1"], g["h2"], one, two, ovlp, nroots=min(3, ntrain))
    assert np.abs(em - g[f"L{layout}_Ems"]).max() < 1e-11
    en, grad, gam, Gam = og.get_energy_with_grad(mol, one, two, ovlp, return_density_matrices=True)
    assert abs(en - float(g[f"L{layout}_Etot"])) < 1e-11
    assert np.abs(grad - g[f"L{layout}_grad"]).max() < 1e-10
    assert np.abs(gam - g[f"L{layout}_gamma"]).max() < 1e-10
    if f"L{layout}_Gamma" in g:
        assert np.abs(Gam - g[f"L{layout}_Gamma"]).max() < 1e-10


def test_gradient_is_derivative_of_energy():
    """Finite-difference check of the oracle's S-dependence: dX from the analytic
    formula vs central differences of get_loewdin_trafo."""
    rng = np.random.default_rng(3)
    n = 5
    a = rng.standard_normal((n, n))
    s = np.eye(n) + 0.1 * (a + a.T)
    d = rng.standard_normal((n, n))
    d = d + d.T
    full = og.loewdin_trafo_grad(s)
    ana = np.einsum("abkl,ab->kl", full, d)
    h = 1e-6
    num = (osub.get_loewdin_trafo(s + h * d) - osub.get_loewdin_trafo(s - h * d)) / (2 * h)
    assert np.abs(ana - num).max() < 1e-8
