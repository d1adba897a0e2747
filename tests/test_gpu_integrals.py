"""K9 (device AO integrals over s shells) against the CPU oracle (oracle/integrals.py, pinned
by tests/test_oracle_integrals.py), and the whole MD step from coordinates."""
import numpy as np
import pytest

from conftest import synthetic_stack

pytestmark = pytest.mark.gpu

H4 = np.array([[0.0, 0.0, 0.0], [0.1, 0.2, 1.7], [0.3, -0.2, 3.5], [1.5, 0.3, 0.5]])
FIELDS = ("ovlp", "hcore", "eri", "ipovlp", "hcore_deriv", "eri_ip1", "e_nuc", "grad_nuc")
TOL = 2e-13


def _chain(n, d=1.78596, radius=0.3, seed=1):
    rng = np.random.default_rng(seed)
    co = np.zeros((n, 3))
    co[:, 0] = d * np.arange(n)
    v = rng.standard_normal((n, 3))
    return co + radius * v / np.linalg.norm(v, axis=1)[:, None]


def _device_arrays(symbols, basis, coords):
    from evcont_b200.engine import get_engine
    eng = get_engine()
    sb = eng.sbasis(symbols, basis)
    ao = eng.ao_integrals(sb, np.asarray(coords))
    return {k: getattr(ao, k).cpu().numpy() for k in FIELDS}, sb


@pytest.mark.parametrize("basis", ["sto-6g", "sto-3g", "6-31g"])
def test_h4_every_array_against_oracle(basis):
    from oracle import integrals as oi
    got, sb = _device_arrays(["H"] * 4, basis, H4[None])
    ref = oi.ao_arrays(oi.SBasis([("H", c) for c in H4], basis))
    assert np.array_equal(sb.aoslices_host, ref["aoslices"][:, 2:])
    for k in FIELDS:
        assert np.abs(got[k][0] - ref[k]).max() < TOL, k


def test_h10_batch_against_oracle_and_split_invariance():
    """configs[1] geometry sampler (H10 chain, atoms displaced by 0.3 bohr); the same geometry
    inside a large batch (one CTA per geometry) and alone (several CTAs per geometry)."""
    from oracle import integrals as oi
    geoms = np.stack([_chain(10, seed=s) for s in range(1, 4)])
    big = np.concatenate([geoms] * 100)  # 300 geometries: no quartet split
    got_big, _ = _device_arrays(["H"] * 10, "sto-6g", big)
    got_one, _ = _device_arrays(["H"] * 10, "sto-6g", geoms[:1])
    for k in FIELDS:
        assert np.array_equal(got_big[k][:3], got_big[k][297:]), k   # run-to-run / position independent
        assert np.array_equal(got_big[k][0], got_one[k][0]), k      # split-independent, bit for bit
    for g in range(2):
        ref = oi.ao_arrays(oi.SBasis([("H", c) for c in geoms[g]], "sto-6g"))
        for k in FIELDS:
            assert np.abs(got_big[k][g] - ref[k]).max() < TOL, k


def test_large_system_global_pair_tables():
    """14 hydrogens: 3780 primitive pairs do not fit the shared-memory budget, so the pair tables are
    built in the caller's workspace (spairs_kernel) -- the path H30 (BASELINE configs[3]) takes."""
    from oracle import integrals as oi
    co = _chain(14, d=1.9, radius=0.2, seed=11)
    got, _ = _device_arrays(["H"] * 14, "sto-6g", np.stack([co] * 3))
    ref = oi.ao_arrays(oi.SBasis([("H", c) for c in co], "sto-6g"))
    for k in FIELDS:
        assert np.abs(got[k][0] - ref[k]).max() < TOL, k
        assert np.array_equal(got[k][0], got[k][2]), k
    big, _ = _device_arrays(["H"] * 14, "sto-6g", np.stack([co] * 600))   # no quartet split
    for k in FIELDS:
        assert np.array_equal(got[k][0], big[k][599]), k


def test_symmetries_on_device():
    got, _ = _device_arrays(["H"] * 6, "sto-6g", _chain(6, seed=7)[None])
    eri, ip1 = got["eri"][0], got["eri_ip1"][0]
    for perm in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)]:
        assert np.array_equal(eri, eri.transpose(perm))
    assert np.array_equal(ip1, ip1.transpose(0, 1, 2, 4, 3))
    assert np.abs(got["hcore_deriv"][0].sum(0)).max() < 1e-12
    # d(ab|cd)/dR summed over the four positions and all atoms vanishes
    tot = ip1 + ip1.transpose(0, 2, 1, 3, 4) + ip1.transpose(0, 3, 4, 1, 2) + ip1.transpose(0, 3, 4, 2, 1)
    assert np.abs(tot).max() < 1e-12


def test_mollite_step_against_oracle_and_finite_differences():
    """The reference-named entry point on a MolLite: energy and forces from coordinates only,
    against the numpy port fed the oracle's integrals, and the forces against central finite
    differences of the energy (the reference's implicit exactness check)."""
    from evcont_b200.ab_initio_gradients_loewdin import get_energy_with_grad, get_energy_with_grad_coords
    from evcont_b200.mol import ArrayMol, MolLite
    from oracle import gradients as og
    from oracle import integrals as oi
    n, N = 6, 4
    ovlp, one, two = synthetic_stack(n, N, 5, 5)
    co = _chain(n, seed=3)
    mol = MolLite([("H", tuple(c)) for c in co], basis="sto-6g", unit="Bohr")
    assert mol.nao == n and mol.natm == n and mol.nelec == (3, 3)
    e, g = get_energy_with_grad(mol, one, two, ovlp)
    ref = oi.ao_arrays(oi.SBasis([("H", c) for c in co], "sto-6g"))
    amol = ArrayMol(**ref)
    oe, ogr = og.get_energy_with_grad(amol, one, two, ovlp)
    assert abs(e - oe) < 1e-10 and np.abs(g - ogr).max() < 1e-8
    # batch from coordinates == single calls
    h = 1e-4
    disp = []
    for A in range(2):
        for x in range(3):
            for sgn in (1, -1):
                c = co.copy()
                c[A, x] += sgn * h
                disp.append(c)
    E, G = get_energy_with_grad_coords(mol, np.stack([co] + disp), one, two, ovlp)
    assert abs(E[0] - e) < 1e-12 and np.abs(G[0] - g).max() < 1e-11
    fd = (E[1::2] - E[2::2]) / (2 * h)
    assert np.abs(fd - g[:2].reshape(-1)).max() < 5e-7
    # mol.intor / set_geom_ surface
    m2 = mol.copy().set_geom_(disp[0])
    assert np.abs(m2.intor("int1e_ovlp") - oi.int1e_ovlp(oi.SBasis([("H", c) for c in disp[0]], "sto-6g"))).max() < TOL
    assert abs(mol.energy_nuc() - ref["e_nuc"]) < 1e-12


def test_angstrom_units_and_unknown_element():
    from evcont_b200.mol import BOHR, MolLite
    m = MolLite("H 0 0 0; H 0 0 0.74", basis="sto-3g", unit="Angstrom")
    assert abs(m.atom_coords()[1, 2] - 0.74 / BOHR) < 1e-14
    with pytest.raises(NotImplementedError):
        MolLite("N 0 0 0; H 0 0 1", basis="6-31g")
