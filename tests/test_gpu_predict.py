"""K3..K8 parity through the C ABI / the reference-named Python API, against the
golden vectors produced by the reference itself (tests/golden/make_golden.py) and
the CPU oracle.  Tolerances from BASELINE.json north_star: 1e-10 Ha on subspace
energies, 1e-8 Ha/bohr on forces."""
import numpy as np
import pytest

from conftest import PREDICT_CASES, load_predict_golden, synthetic_stack

pytestmark = pytest.mark.gpu
E_TOL, F_TOL = 1e-10, 1e-8


def _mol(norb, natm, seed):
    from evcont_b200.mol import synthetic_mol
    return synthetic_mol(norb, natm, seed=seed)


@pytest.mark.parametrize("norb,natm,ntrain", PREDICT_CASES)
def test_loewdin_and_integrals(norb, natm, ntrain):
    from evcont_b200 import electron_integral_utils as eiu
    g = load_predict_golden(norb, natm, ntrain)
    mol = _mol(norb, natm, int(g["seed"]))
    x = eiu.get_loewdin_trafo(mol.intor("int1e_ovlp"))
    assert np.abs(x - g["loewdin_X"]).max() < 1e-12
    h1, h2 = eiu.get_integrals(mol, g["loewdin_X"])
    assert np.abs(h1 - g["h1"]).max() < 1e-11
    assert np.abs(h2 - g["h2"]).max() < 1e-11
    t1, t2 = eiu.transform_integrals(g["h1"], g["h2"], g["loewdin_X"])
    assert np.abs(t1 - g["trafo_h1"]).max() < 1e-11
    assert np.abs(t2 - g["trafo_h2"]).max() < 1e-11
    # batched leading axes
    b1, b2 = eiu.transform_integrals(np.stack([g["h1"], 2 * g["h1"]]), np.stack([g["h2"], 2 * g["h2"]]),
                                     g["loewdin_X"])
    assert np.abs(b2[1] - 2 * g["trafo_h2"]).max() < 1e-10 and np.abs(b1[0] - g["trafo_h1"]).max() < 1e-11


@pytest.mark.parametrize("norb,natm,ntrain", PREDICT_CASES)
def test_loewdin_derivative(norb, natm, ntrain):
    from evcont_b200 import ab_initio_gradients_loewdin as agl
    from oracle import gradients as og
    g = load_predict_golden(norb, natm, ntrain)
    mol = _mol(norb, natm, int(g["seed"]))
    dx = agl.get_derivative_ao_mo_trafo(mol)
    assert dx.shape == (norb, norb, natm, 3)
    assert np.abs(dx - g["dX_dR"]).max() < 1e-10
    full = agl.loewdin_trafo_grad(mol.intor("int1e_ovlp"))
    assert np.abs(full - og.loewdin_trafo_grad(mol.intor("int1e_ovlp"))).max() < 1e-10
    assert np.array_equal(agl.get_overlap_grad(mol), og.get_overlap_grad(mol))


@pytest.mark.parametrize("layout", [6, 5, 3, 2])
@pytest.mark.parametrize("norb,natm,ntrain", PREDICT_CASES)
def test_subspace_solvers(norb, natm, ntrain, layout):
    from evcont_b200 import ab_initio_eigenvector_continuation as evc
    g = load_predict_golden(norb, natm, ntrain)
    ovlp, one, two = synthetic_stack(norb, ntrain, int(g["seed"]) + 100, layout)
    e, c = evc.approximate_ground_state(g["h1"], g["h2"], one, two, ovlp)
    assert abs(e - float(g[f"L{layout}_E0"])) < E_TOL
    ref = g[f"L{layout}_c0"]
    assert np.abs(c * np.sign(c[0]) - ref * np.sign(ref[0])).max() < 1e-8
    assert abs(c @ ovlp @ c - 1.0) < 1e-12
    nr = min(3, ntrain)
    em, cm = evc.approximate_multistate(g["h1"], g["h2"], one, two, ovlp, nroots=nr)
    assert em.shape == (nr,) and cm.shape == (nr, ntrain)
    assert np.abs(em - g[f"L{layout}_Ems"]).max() < E_TOL
    for r in range(nr):
        ref = g[f"L{layout}_Cms"][r]
        k = np.argmax(np.abs(ref))
        assert np.abs(cm[r] * np.sign(cm[r][k]) - ref * np.sign(ref[k])).max() < 1e-7


@pytest.mark.parametrize("layout", [6, 5, 3, 2])
@pytest.mark.parametrize("norb,natm,ntrain", PREDICT_CASES)
def test_energy_with_grad_golden(norb, natm, ntrain, layout):
    from evcont_b200.ab_initio_gradients_loewdin import get_energy_with_grad
    g = load_predict_golden(norb, natm, ntrain)
    mol = _mol(norb, natm, int(g["seed"]))
    ovlp, one, two = synthetic_stack(norb, ntrain, int(g["seed"]) + 100, layout)
    e, grad, gam, Gam = get_energy_with_grad(mol, one, two, ovlp, return_density_matrices=True)
    assert abs(e - float(g[f"L{layout}_Etot"])) < E_TOL
    assert grad.shape == (natm, 3)
    assert np.abs(grad - g[f"L{layout}_grad"]).max() < F_TOL
    assert np.abs(gam - g[f"L{layout}_gamma"]).max() < 1e-9
    if f"L{layout}_Gamma" in g:
        assert np.abs(Gam - g[f"L{layout}_Gamma"]).max() < 1e-9
    # without the RDMs the step runs in the packed (8-fold symmetric) form
    e2, grad2 = get_energy_with_grad(mol, one, two, ovlp)
    assert abs(e2 - float(g[f"L{layout}_Etot"])) < E_TOL
    assert np.abs(grad2 - g[f"L{layout}_grad"]).max() < F_TOL
    e3, grad3 = get_energy_with_grad(mol, one, two, ovlp)
    assert e3 == e2 and np.array_equal(grad3, grad2)  # deterministic


@pytest.mark.parametrize("layout", [6, 5, 3, 2])
@pytest.mark.parametrize("norb,natm,ntrain", PREDICT_CASES)
def test_packed_and_full_steps_agree(norb, natm, ntrain, layout):
    """evc_energy_with_grad_packed vs evc_energy_with_grad on the same device inputs."""
    from evcont_b200.engine import DeviceAO
    from evcont_b200.mol import ao_bundle
    from evcont_b200.stackcache import as_device_stack
    ovlp, one, two = synthetic_stack(norb, ntrain, 40 + layout, layout)
    stack = as_device_stack(one, two, ovlp)
    mols = [_mol(norb, natm, 500 + k) for k in range(5)]
    ao = DeviceAO.from_bundles(stack.engine, [ao_bundle(m) for m in mols])
    Ep, gp, _, _, cp = stack.engine.energy_with_grad(stack, ao, packed=True)
    Ef, gf, _, _, cf = stack.engine.energy_with_grad(stack, ao, packed=False)
    assert (Ep - Ef).abs().max().item() < 1e-11
    assert (gp - gf).abs().max().item() < 1e-10
    sgn = torch_sign(cp, cf)
    assert (cp * sgn - cf).abs().max().item() < 1e-9


def torch_sign(a, b):
    import torch
    k = b.abs().argmax(dim=1, keepdim=True)
    return torch.sign(a.gather(1, k)) * torch.sign(b.gather(1, k))


@pytest.mark.parametrize("layout", [6, 3])
def test_packed_step_asymmetric_full_stack(layout):
    """Full layouts whose [a,b] and [b,a] blocks differ (a user-assigned stack): H reads the
    lower-triangle blocks only, the predicted RDMs read both (evcont FCI_EVCont.py:117-127,
    ab_initio_gradients_loewdin.py:343-353)."""
    from evcont_b200.ab_initio_gradients_loewdin import get_energy_with_grad
    from oracle import gradients as og
    norb, natm, ntrain = 6, 3, 4
    ovlp, one, two = synthetic_stack(norb, ntrain, 8, layout)
    rng = np.random.default_rng(3)
    one = one + 0.3 * rng.standard_normal(one.shape)
    pert = rng.standard_normal((ntrain, ntrain, norb * norb, norb * norb)) / norb
    pert = pert + pert.transpose(0, 1, 3, 2)
    if layout == 6:
        two = two + 0.3 * pert.reshape(two.shape)
    else:
        ic = np.tril_indices(norb * norb)
        two = two + 0.3 * pert[:, :, ic[0], ic[1]]
    mol = _mol(norb, natm, 77)
    e, grad = get_energy_with_grad(mol, one, two, ovlp)
    oe, ogr = og.get_energy_with_grad(mol, one, two, ovlp)
    assert abs(e - oe) < E_TOL and np.abs(grad - ogr).max() < F_TOL


@pytest.mark.parametrize("norb,natm,ntrain,layout", [(13, 3, 4, 5), (11, 4, 3, 2), (12, 5, 3, 6),
                                                     (14, 4, 3, 5), (16, 5, 3, 2)])
def test_packed_step_larger_norb_vs_oracle(norb, natm, ntrain, layout):
    """n = 13 is the H2O/6-31G size (largest shared-memory case); n > 13 takes the
    full-tensor transform/gradient kernels around the packed stack contractions."""
    from evcont_b200.ab_initio_gradients_loewdin import get_energy_with_grad
    from oracle import gradients as og
    ovlp, one, two = synthetic_stack(norb, ntrain, 19, layout)
    mol = _mol(norb, natm, 31)
    e, grad = get_energy_with_grad(mol, one, two, ovlp)
    oe, ogr = og.get_energy_with_grad(mol, one, two, ovlp)
    assert abs(e - oe) < E_TOL
    assert np.abs(grad - ogr).max() < F_TOL


@pytest.mark.parametrize("norb,natm,ntrain,layout", [(28, 7, 6, 2), (30, 30, 5, 6), (30, 30, 4, 5),
                                                     (28, 7, 4, 3), (20, 6, 5, 2)])
def test_large_norb_configs_vs_oracle(norb, natm, ntrain, layout):
    """BASELINE configs[3] (H30/STO-6G: 30 orbitals, 30 atoms) and configs[4] (Zundel/6-31G:
    28 orbitals, 7 atoms, exchange-compressed tril layout) at their orbital counts: the
    HBM-resident transform / gradient kernels, single geometry and a small batch."""
    from evcont_b200.ab_initio_gradients_loewdin import get_energy_with_grad, get_energy_with_grad_batch
    from oracle import gradients as og
    ovlp, one, two = synthetic_stack(norb, ntrain, 41, layout)
    mols = [_mol(norb, natm, 700 + k) for k in range(3)]
    E, G = get_energy_with_grad_batch(mols, one, two, ovlp)
    for k, m in enumerate(mols[:2]):
        oe, ogr = og.get_energy_with_grad(m, one, two, ovlp)
        assert abs(E[k] - oe) < E_TOL * max(1.0, abs(oe))
        assert np.abs(G[k] - ogr).max() < F_TOL * max(1.0, np.abs(ogr).max())
    e1, g1 = get_energy_with_grad(mols[2], one, two, ovlp)
    assert abs(E[2] - e1) < 1e-11 * max(1.0, abs(e1)) and np.abs(G[2] - g1).max() < 1e-10 * max(1.0, np.abs(g1).max())


def test_grad_elec_OAO_against_oracle():
    from evcont_b200.ab_initio_gradients_loewdin import get_grad_elec_OAO
    from oracle import gradients as og
    rng = np.random.default_rng(5)
    mol = _mol(7, 3, 21)
    gamma = rng.standard_normal((7, 7))
    Gamma = rng.standard_normal((7,) * 4)
    ref = og.get_grad_elec_OAO(mol, gamma, Gamma)
    got = get_grad_elec_OAO(mol, gamma, Gamma)
    assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_batch_matches_single_and_oracle():
    """A batch of geometries through the fused step == one-by-one == CPU oracle."""
    from evcont_b200.ab_initio_gradients_loewdin import get_energy_with_grad, get_energy_with_grad_batch
    from oracle import gradients as og
    norb, natm, ntrain = 6, 4, 5
    ovlp, one, two = synthetic_stack(norb, ntrain, 77, 5)
    mols = [_mol(norb, natm, 300 + k) for k in range(7)]
    E, G = get_energy_with_grad_batch(mols, one, two, ovlp)
    E2, G2 = get_energy_with_grad_batch(mols, one, two, ovlp)
    assert np.array_equal(E, E2) and np.array_equal(G, G2)  # run-to-run bit identity
    for k, m in enumerate(mols):
        e1, g1 = get_energy_with_grad(m, one, two, ovlp)
        # the batch and the single call use different register tiles of the streaming
        # kernels: same numbers up to summation order
        assert abs(E[k] - e1) < 1e-12 and np.abs(G[k] - g1).max() < 1e-11
        oe, ogr = og.get_energy_with_grad(m, one, two, ovlp)
        assert abs(E[k] - oe) < E_TOL and np.abs(G[k] - ogr).max() < F_TOL


def test_h10_size_vs_oracle():
    """configs[1] sizes: norb 10, 10 atoms, N = 20, tril layout."""
    from evcont_b200.ab_initio_gradients_loewdin import get_energy_with_grad
    from oracle import gradients as og
    ovlp, one, two = synthetic_stack(10, 20, 9, 5)
    mol = _mol(10, 10, 123)
    e, grad = get_energy_with_grad(mol, one, two, ovlp)
    oe, ogr = og.get_energy_with_grad(mol, one, two, ovlp)
    assert abs(e - oe) < E_TOL
    assert np.abs(grad - ogr).max() < F_TOL


def test_not_positive_definite_overlap_raises():
    from evcont_b200.ab_initio_eigenvector_continuation import approximate_ground_state
    g = load_predict_golden(4, 2, 3)
    ovlp, one, two = synthetic_stack(4, 3, 1, 6)
    ovlp = ovlp.copy()
    ovlp[2, 2] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        approximate_ground_state(g["h1"], g["h2"], one, two, ovlp)


def test_bad_layout_asserts():
    from evcont_b200.ab_initio_eigenvector_continuation import approximate_ground_state
    g = load_predict_golden(4, 2, 3)
    ovlp, one, two = synthetic_stack(4, 3, 1, 6)
    with pytest.raises(AssertionError):
        approximate_ground_state(g["h1"], g["h2"], one, two.reshape(3, 3, 4, 64), ovlp)


def test_scanner_surface():
    from evcont_b200.MD_utils import get_scanner
    from oracle import gradients as og
    ovlp, one, two = synthetic_stack(6, 3, 3, 6)
    mol = _mol(6, 6, 12)
    sc = get_scanner(mol, one, two, ovlp)
    e, grad = sc(mol)
    oe, ogr, ogam, oGam = og.get_energy_with_grad(mol, one, two, ovlp, return_density_matrices=True)
    assert abs(e - oe) < E_TOL and np.abs(grad - ogr).max() < F_TOL
    assert sc.base.converged and sc.mol is mol
    assert np.abs(sc.base.predicted_one_rdm - ogam).max() < 1e-9
    assert np.abs(sc.base.predicted_two_rdm - oGam).max() < 1e-9
    e0, g0 = get_scanner(mol, None, None, None)(mol)
    assert e0 == mol.energy_nuc() and np.array_equal(g0, mol.grad_nuc())


@pytest.mark.parametrize("layout", [6, 5, 3, 2])
@pytest.mark.parametrize("norb,ntrain,G", [(5, 3, 9), (7, 5, 70), (10, 6, 33), (13, 4, 20),
                                           (6, 4, 1), (6, 4, 2), (6, 5, 3), (7, 5, 8), (7, 3, 13),
                                           (13, 4, 16), (10, 6, 17)])
def test_batched_stack_contractions(norb, ntrain, G, layout):
    """K5/K7 against numpy on both paths -- the HBM-streaming kernels (batch <= 16, every
    register-tile variant 1/2/4/8 with ragged last tiles) and the DMMA GEMM kernels
    (batch > 16) -- for all four layouts, odd leading dimensions (norb 5, 7, 13) and
    ragged tile edges."""
    from evcont_b200.engine import DeviceStack, get_engine
    from oracle import gradients as og
    from oracle import subspace as osub
    rng = np.random.default_rng(norb * 100 + ntrain)
    ovlp, one, two = synthetic_stack(norb, ntrain, 40 + norb, layout)
    eng = get_engine()
    stack = DeviceStack(ovlp, one, two, engine=eng, norb=norb)
    h1 = rng.standard_normal((G, norb, norb))
    h1 = h1 + h1.transpose(0, 2, 1)
    h2 = rng.standard_normal((G,) + (norb,) * 4)
    h2 = h2 + h2.transpose(0, 3, 4, 1, 2)
    H = eng.subspace_H(stack, eng.to_device(h1), eng.to_device(h2)).cpu().numpy()
    cv = rng.standard_normal((G, ntrain))
    gam, Gam = eng.predict_rdm(stack, eng.to_device(cv))
    gam, Gam = gam.cpu().numpy(), Gam.cpu().numpy()
    il = np.tril_indices(ntrain)
    for g in range(G):
        ref = osub.subspace_hamiltonian(h1[g], h2[g], one, two)
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(H[g][il] - ref[il]).max() < 1e-12 * scale * norb ** 2
        rg, rG = og.predict_rdms(cv[g], one, two, norb)
        assert np.abs(gam[g] - rg).max() < 1e-11 * max(1.0, np.abs(rg).max())
        assert np.abs(Gam[g] - rG).max() < 1e-11 * max(1.0, np.abs(rG).max())


@pytest.mark.parametrize("G,chunk", [(1, 256), (23, 8), (16, 16), (9, 2)])
def test_host_buffer_pipeline(G, chunk):
    """evc_energy_with_grad_host (chunked H2D / compute / D2H pipeline) == device-resident
    call == CPU oracle, including a ragged last chunk."""
    from evcont_b200.engine import DeviceAO, DeviceStack, HostAO, get_engine
    from evcont_b200.mol import ao_bundle
    from oracle import gradients as og
    norb, natm, ntrain = 6, 3, 4
    ovlp, one, two = synthetic_stack(norb, ntrain, 55, 2)
    eng = get_engine()
    stack = DeviceStack(ovlp, one, two, engine=eng, norb=norb)
    mols = [_mol(norb, natm, 900 + k) for k in range(G)]
    bundles = [ao_bundle(m) for m in mols]
    host = HostAO.from_bundles(bundles)
    host.E.fill_(float("nan"))
    E, grad = eng.energy_with_grad_host(stack, host, chunk=chunk)
    E, grad = E.numpy().copy(), grad.numpy().copy()
    Ed, gd, _, _, _ = eng.energy_with_grad(stack, DeviceAO.from_bundles(eng, bundles))
    assert np.abs(E - Ed.cpu().numpy()).max() < 1e-12
    assert np.abs(grad - gd.cpu().numpy()).max() < 1e-11
    for k in (0, G // 2, G - 1):
        oe, ogr = og.get_energy_with_grad(mols[k], one, two, ovlp)
        assert abs(E[k] - oe) < E_TOL and np.abs(grad[k] - ogr).max() < F_TOL
    E2, grad2 = eng.energy_with_grad_host(stack, host, chunk=chunk)
    assert np.array_equal(E, E2.numpy()) and np.array_equal(grad, grad2.numpy())


def test_energy_only_mode_matches_the_full_step():
    """``evc_energy_with_grad_packed`` with grad = NULL (approximate_ground_state_OAO for a batch)."""
    from evcont_b200.engine import DeviceAO, DeviceStack, get_engine
    from evcont_b200.mol import ao_bundle, synthetic_mol
    eng = get_engine()
    for norb, natm, ntrain in ((6, 6, 3), (10, 10, 7), (15, 4, 4)):
        rng = np.random.default_rng(norb)
        b = rng.standard_normal((ntrain, ntrain))
        S = np.eye(ntrain) + 0.01 * (b + b.T)
        one = rng.standard_normal((ntrain, ntrain, norb, norb))
        two = rng.standard_normal((ntrain, ntrain) + (norb,) * 4)
        stack = DeviceStack(S, one, two, engine=eng, norb=norb)
        ao = DeviceAO.from_bundles(eng, [ao_bundle(synthetic_mol(norb, natm, seed=70 + k)) for k in range(5)])
        E, _, _, _, cvec = eng.energy_with_grad(stack, ao)
        E2, c2 = eng.energies(stack, ao)
        assert np.array_equal(E.cpu().numpy(), E2.cpu().numpy())
        assert np.array_equal(cvec.cpu().numpy(), c2.cpu().numpy())


def test_near_degenerate_overlap_eigenvalues():
    """An AO overlap with two eigenvalues 2e-6 apart (inside one round(., 5) bucket of
    evcont/ab_initio_gradients_loewdin.py:55-56).  The reference then treats the pair as exactly degenerate;
    the device path uses exact divided differences.  Both must agree with the central
    finite difference of the Loewdin transform; the size of the reference's own deviation is recorded."""
    from evcont_b200 import ab_initio_gradients_loewdin as agl
    from evcont_b200.electron_integral_utils import get_loewdin_trafo
    from evcont_b200.mol import synthetic_mol
    from oracle import gradients as og
    n, natm = 6, 3
    mol = synthetic_mol(n, natm, seed=31)
    rng = np.random.default_rng(8)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    w = np.array([0.61, 0.8, 0.800002, 1.05, 1.3, 1.45])   # the 2nd and 3rd share the 1e-5 bucket
    S = (q * w) @ q.T
    S = 0.5 * (S + S.T)
    mol._ovlp = np.ascontiguousarray(S)
    dS = og.get_overlap_grad(mol)                      # (n, n, natm, 3) from the molecule's int1e_ipovlp
    dev = agl.get_derivative_ao_mo_trafo(mol)          # device: exact divided differences
    ref = np.einsum("ijkl,ijmn->klmn", og.loewdin_trafo_grad(S), dS)   # the reference's bucketed formula
    # central finite differences of X(S + h dS_xi)
    h = 1e-6
    worst_dev = worst_ref = 0.0
    for A in range(natm):
        for x in range(3):
            d = dS[:, :, A, x]
            fd = (get_loewdin_trafo(S + h * d) - get_loewdin_trafo(S - h * d)) / (2 * h)
            worst_dev = max(worst_dev, np.abs(dev[:, :, A, x] - fd).max())
            worst_ref = max(worst_ref, np.abs(ref[:, :, A, x] - fd).max())
    assert worst_dev < 5e-8                 # the device derivative is the true derivative
    # the reference deviates by O(coupling between the two near-degenerate vectors); it is small but not zero
    assert 1e-8 < worst_ref < 1e-3
    print(f"near-degenerate S: |dX - FD| device {worst_dev:.2e}, reference formula {worst_ref:.2e}")
