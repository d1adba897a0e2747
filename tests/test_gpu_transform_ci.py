"""``transform_ci`` on the device against the oracle restatement of
pyscf.fci.addons.transform_ci (evcont/FCI_EVCont.py:79-85), the device Fock build / RHF behind
``get_basis(mol, "canonical")`` (evcont/electron_integral_utils.py:103-106), and the reference's
default workflow ``FCI_EVCont_obj()`` (cibasis="canonical") end to end without PySCF."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_orth(n, seed):
    q, r = np.linalg.qr(np.random.default_rng(seed).standard_normal((n, n)))
    return q * np.sign(np.diag(r))


def _rand_ci(na, nb, seed):
    c = np.random.default_rng(seed).standard_normal((na, nb))
    return c / np.linalg.norm(c)


@pytest.mark.parametrize("norb,nelec", [(2, (1, 1)), (4, (2, 1)), (5, (2, 1)), (5, (3, 2)), (6, (3, 3)), (7, (4, 0)),
                                        (10, (5, 5))])
def test_against_oracle(norb, nelec):
    from evcont_b200.fci import transform_ci
    from oracle import cistring as ocs
    from oracle import transform_ci as otc
    na, nb = ocs.num_strings(norb, nelec[0]), ocs.num_strings(norb, nelec[1])
    c = _rand_ci(na, nb, 11 + norb)
    for u in (_rand_orth(norb, norb), np.random.default_rng(5).standard_normal((norb, norb)) / np.sqrt(norb)):
        ref = otc.transform_ci(c, nelec, u)
        out = transform_ci(c, nelec, u)
        assert out.shape == (na, nb)
        assert np.abs(out - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
    assert np.array_equal(transform_ci(c, nelec, u), out)   # deterministic
    if nelec[0] - nelec[1] in (0, 1):   # integer electron count, identity rotation
        assert np.abs(transform_ci(c, sum(nelec), np.eye(norb)) - c).max() < 1e-14


def test_h2o_size_properties():
    """13 orbitals, (5, 5) electrons: 1287 x 1287 determinants (odd row length), by norm conservation,
    the back-rotation and the group property instead of the oracle."""
    from evcont_b200.fci import transform_ci
    norb, nelec = 13, (5, 5)
    c = _rand_ci(1287, 1287, 2)
    u1, u2 = _rand_orth(norb, 3), _rand_orth(norb, 4)
    c1 = transform_ci(c, nelec, u1)
    assert abs(np.linalg.norm(c1) - 1.0) < 1e-12
    assert np.abs(transform_ci(c1, nelec, u1.T) - c).max() < 1e-12
    assert np.abs(transform_ci(c1, nelec, u2) - transform_ci(c, nelec, u1 @ u2)).max() < 1e-12


def test_rdm_covariance_on_device():
    from evcont_b200.fci import B200FCISolver, transform_ci
    norb, nelec = 6, (3, 3)
    u = _rand_orth(norb, 9)
    a, b = _rand_ci(20, 20, 1), _rand_ci(20, 20, 2)
    s = B200FCISolver()
    d1, d2 = s.trans_rdm12(a, b, norb, nelec)
    n1, n2 = s.trans_rdm12(transform_ci(a, nelec, u), transform_ci(b, nelec, u), norb, nelec)
    assert np.abs(n1 - u.T @ d1 @ u).max() < 1e-12
    assert np.abs(n2 - np.einsum("pqrs,pa,qb,rc,sd->abcd", d2, u, u, u, u)).max() < 1e-12


def _h_chain(n, d, basis="sto-6g"):
    from evcont_b200.mol import MolLite
    xs = (np.arange(n) - np.median(np.arange(n))) * d
    return MolLite([("H", (x, 0.0, 0.0)) for x in xs], basis=basis, unit="Bohr")


@pytest.mark.parametrize("n", [3, 10, 13, 28])
def test_fock_build_against_einsum(n):
    from evcont_b200._lib import check
    from evcont_b200.engine import _ptr, get_engine
    rng = np.random.default_rng(n)
    h = rng.standard_normal((n, n))
    eri = rng.standard_normal((n,) * 4)
    dm = rng.standard_normal((n, n))
    eng = get_engine()
    out = eng.empty(n, n)
    hd, ed, dd = eng.to_device(h), eng.to_device(eri), eng.to_device(dm)   # keep the buffers alive
    eng._bind_stream()
    check(eng.lib.evc_fock_rhf(eng._ctx, n, _ptr(hd), _ptr(ed), _ptr(dd), _ptr(out)))
    ref = h + np.einsum("pqrs,rs->pq", eri, dm) - 0.5 * np.einsum("pqrs,qr->ps", eri, dm)
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-12 * n * n


def test_rhf_known_energies():
    """H2 / STO-3G at 1.4 bohr: E_RHF = -1.1167 Ha (Szabo-Ostlund, table 3.11); water / 6-31G at the
    experimental geometry: -75.9840 Ha (the value tests/test_oracle_integrals_sp.py pins the oracle with)."""
    from evcont_b200.mol import MolLite
    from evcont_b200.scf import rhf
    r = rhf(MolLite([("H", (0, 0, 0)), ("H", (1.4, 0, 0))], basis="sto-3g"))
    assert r.converged and abs(r.e_tot + 1.1167) < 1e-4
    ang = 1.0 / 0.52917721092
    rr, th = 0.9572 * ang, np.deg2rad(104.52)
    w = MolLite([("O", (0, 0, 0)), ("H", (rr * np.sin(th / 2), 0, rr * np.cos(th / 2))),
                 ("H", (-rr * np.sin(th / 2), 0, rr * np.cos(th / 2)))], basis="6-31g")
    r = rhf(w)
    assert r.converged and abs(r.e_tot + 75.9840) < 2e-4
    S = w.intor("int1e_ovlp")
    assert np.abs(r.mo_coeff.T @ S @ r.mo_coeff - np.eye(13)).max() < 1e-10
    # stationarity: the occupied-virtual block of the Fock matrix vanishes; energies are its diagonal
    from evcont_b200.electron_integral_utils import get_basis
    assert np.abs(get_basis(w, "canonical").T @ S @ get_basis(w, "canonical") - np.eye(13)).max() < 1e-10


def test_canonical_basis_workflow_matches_oao_workflow():
    """The reference's default ``cibasis='canonical'`` (RHF orbitals -> FCI -> transform_ci to the OAO
    basis) and ``cibasis='OAO'`` describe the same states: identical overlaps / t-RDMs up to the sign of
    each training vector, identical predictions."""
    from evcont_b200.FCI_EVCont import FCI_EVCont_obj
    from evcont_b200.ab_initio_gradients_loewdin import get_energy_with_grad
    a, b = FCI_EVCont_obj(), FCI_EVCont_obj(cibasis="OAO")
    assert a.cibasis == "canonical"
    for d in (1.0, 1.8, 2.6):
        a.append_to_rdms(_h_chain(6, d))
        b.append_to_rdms(_h_chain(6, d))
    assert np.abs(np.array(a.ens) - np.array(b.ens)).max() < 1e-9
    sg = np.array([np.sign(np.vdot(x, y)) for x, y in zip(a.fcivecs, b.fcivecs)])
    for x, y, s in zip(a.fcivecs, b.fcivecs, sg):
        assert np.abs(x - s * y).max() < 1e-7
    ss = sg[:, None] * sg[None, :]
    assert np.abs(a.overlap - ss * b.overlap).max() < 1e-7
    assert np.abs(a.one_rdm - ss[:, :, None, None] * b.one_rdm).max() < 1e-6
    assert np.abs(a.two_rdm - ss[:, :, None, None, None, None] * b.two_rdm).max() < 1e-6
    mol = _h_chain(6, 1.5)
    ea, ga = get_energy_with_grad(mol, a.one_rdm, a.two_rdm, a.overlap)
    eb, gb = get_energy_with_grad(mol, b.one_rdm, b.two_rdm, b.overlap)
    assert abs(ea - eb) < 1e-8 and np.abs(ga - gb).max() < 1e-6
