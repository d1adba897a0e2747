"""K9g (device AO integrals over s and p shells, McMurchie-Davidson) against the CPU oracle
(oracle/integrals_sp.py, pinned by tests/test_oracle_integrals_sp.py): water and O-H in 6-31G, a
Zundel-sized cation through invariants, and the whole step from coordinates on water."""
import numpy as np
import pytest

from conftest import synthetic_stack

pytestmark = pytest.mark.gpu
ANG = 1.0 / 0.52917721092
FIELDS = ("ovlp", "hcore", "eri", "ipovlp", "hcore_deriv", "eri_ip1", "e_nuc", "grad_nuc")


def water_coords(noise=0.0, seed=0):
    r, th = 0.9572 * ANG, np.deg2rad(104.52)
    co = np.array([[0, 0, 0.1], [r * np.sin(th / 2), 0.05, r * np.cos(th / 2)], [-r * np.sin(th / 2), -0.1, r * np.cos(th / 2)]])
    return co + noise * np.random.default_rng(seed).standard_normal(co.shape)


def _device(symbols, coords):
    from evcont_b200.engine import get_engine
    eng = get_engine()
    sb = eng.sbasis(symbols, "6-31g")
    assert sb.general
    ao = eng.ao_integrals(sb, np.asarray(coords))
    return {k: getattr(ao, k).cpu().numpy() for k in FIELDS}, sb


def _tol(ref):
    return 2e-13 * max(1.0, float(np.abs(np.asarray(ref)).max()))


def test_water_every_array_against_oracle():
    from oracle import integrals_sp as osp
    co = water_coords()
    got, sb = _device(["O", "H", "H"], np.stack([co, co + 0.01]))
    ref = osp.ao_arrays(osp.SPBasis([("O", co[0]), ("H", co[1]), ("H", co[2])], "6-31g"))
    assert sb.nao == 13 and np.array_equal(sb.aoslices_host, ref["aoslices"][:, 2:])
    for k in FIELDS:
        assert np.abs(got[k][0] - ref[k]).max() < _tol(ref[k]), k
    # a rigid shift leaves everything unchanged
    for k in ("ovlp", "hcore", "eri", "eri_ip1", "hcore_deriv"):
        assert np.abs(got[k][0] - got[k][1]).max() < 50 * _tol(ref[k]), k


def test_oh_against_oracle_and_batch_consistency():
    from oracle import integrals_sp as osp
    co = np.array([[0.1, -0.2, 0.05], [0.4, 0.3, 1.75]])
    big, _ = _device(["O", "H"], np.stack([co] * 700))   # no quartet split
    one, _ = _device(["O", "H"], co[None])                # split over many CTAs
    ref = osp.ao_arrays(osp.SPBasis([("O", co[0]), ("H", co[1])], "6-31g"))
    for k in FIELDS:
        assert np.abs(one[k][0] - ref[k]).max() < _tol(ref[k]), k
        assert np.array_equal(one[k][0], big[k][0]) and np.array_equal(big[k][0], big[k][699]), k


def test_two_oxygens_every_array_against_oracle():
    """O2 in 6-31G at a general orientation: every quartet class with p shells on DIFFERENT centres (ppps, pppp,
    psps, ppss across the two atoms) is compared element by element with the oracle."""
    from oracle import integrals_sp as osp
    co = np.array([[0.05, 0.10, -0.20], [0.85, -0.60, 2.15]])
    got, sb = _device(["O", "O"], co[None])
    ref = osp.ao_arrays(osp.SPBasis([("O", co[0]), ("O", co[1])], "6-31g"))
    assert sb.nao == 18
    for k in FIELDS:
        assert np.abs(got[k][0] - ref[k]).max() < _tol(ref[k]), k


def test_ooh_against_golden_fixture():
    """O, O, H (20 AOs: two p centres and a hydrogen between them, the Zundel motif): the small arrays whole and
    6000 seeded random elements of int2e / int2e_ip1 against tests/golden/integrals_sp_OOH.npz (written by
    tests/golden/make_integrals_sp_golden.py from the oracle)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "integrals_sp_OOH.npz"))
    got, sb = _device([str(x) for x in g["symbols"]], g["coords"][None])
    assert sb.nao == 20 and np.array_equal(sb.aoslices_host, g["aoslices"][:, 2:])
    for k in ("ovlp", "hcore", "ipovlp", "hcore_deriv", "e_nuc", "grad_nuc"):
        assert np.abs(got[k][0] - g[k]).max() < _tol(g[k]), k
    eri, ip1 = got["eri"][0], got["eri_ip1"][0]
    assert np.abs(eri[tuple(g["idx_eri"].T)] - g["eri_vals"]).max() < _tol(g["eri_vals"])
    assert np.abs(ip1[tuple(g["idx_ip1"].T)] - g["ip1_vals"]).max() < _tol(g["ip1_vals"])
    assert abs(np.abs(eri).sum() - float(g["eri_abs_sum"])) < 1e-10 * float(g["eri_abs_sum"])
    assert abs(np.abs(ip1).sum() - float(g["ip1_abs_sum"])) < 1e-10 * float(g["ip1_abs_sum"])


def test_zundel_sized_invariants():
    """H5O2+ (28 AOs, 7 atoms): permutational symmetry, translational invariance of the derivative
    arrays and rotational invariance of scalar contractions, device only."""
    rng = np.random.default_rng(4)
    co = np.array([[-2.25, 0.0, 0.0], [2.25, 0.0, 0.0], [0.0, 0.1, 0.0], [-2.9, 1.45, 0.3], [-2.9, -1.45, -0.3],
                   [2.9, 0.3, 1.45], [2.9, -0.3, -1.45]]) + 0.05 * rng.standard_normal((7, 3))
    sym = ["O", "O", "H", "H", "H", "H", "H"]
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    got, sb = _device(sym, np.stack([co, co @ q.T + 1.0]))
    assert sb.nao == 28
    eri, ip1 = got["eri"][0], got["eri_ip1"][0]
    for perm in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)]:
        assert np.array_equal(eri, eri.transpose(perm))
    assert np.array_equal(ip1, ip1.transpose(0, 1, 2, 4, 3))
    tot = ip1 + ip1.transpose(0, 2, 1, 3, 4) + ip1.transpose(0, 3, 4, 1, 2) + ip1.transpose(0, 3, 4, 2, 1)
    assert np.abs(tot).max() < 1e-10
    assert np.abs(got["hcore_deriv"][0].sum(0)).max() < 1e-10
    assert np.abs(np.diag(got["ovlp"][0]) - 1).max() < 1e-13
    # rotation: eigenvalues of S, of S^-1/2 h S^-1/2, and the trace-like ERI contractions are invariant
    for g in (0, 1):
        w, v = np.linalg.eigh(got["ovlp"][g])
        x = v @ np.diag(w ** -0.5) @ v.T
        got[("ev", g)] = np.linalg.eigvalsh(x @ got["hcore"][g] @ x)
        got[("sw", g)] = w
        e = np.einsum("ai,bj,ck,dl,abcd->ijkl", x, x, x, x, got["eri"][g], optimize=True)
        got[("j", g)] = (np.einsum("iijj->", e), np.einsum("ijji->", e))
    assert np.abs(got[("sw", 0)] - got[("sw", 1)]).max() < 1e-12
    assert np.abs(got[("ev", 0)] - got[("ev", 1)]).max() < 1e-10
    assert np.abs(np.array(got[("j", 0)]) - np.array(got[("j", 1)])).max() < 1e-9


def test_water_step_from_coordinates():
    from evcont_b200.ab_initio_gradients_loewdin import get_energy_with_grad, get_energy_with_grad_coords
    from evcont_b200.mol import ArrayMol, MolLite
    from oracle import gradients as og
    from oracle import integrals_sp as osp
    n, N = 13, 4
    ovlp, one, two = synthetic_stack(n, N, 17, 2)
    co = water_coords()
    mol = MolLite([("O", tuple(co[0])), ("H", tuple(co[1])), ("H", tuple(co[2]))], basis="6-31g", unit="Bohr")
    assert mol.nao == 13 and mol.nelec == (5, 5)
    e, g = get_energy_with_grad(mol, one, two, ovlp)
    ref = osp.ao_arrays(osp.SPBasis([("O", co[0]), ("H", co[1]), ("H", co[2])], "6-31g"))
    oe, ogr = og.get_energy_with_grad(ArrayMol(**ref), one, two, ovlp)
    assert abs(e - oe) < 1e-10 * max(1.0, abs(oe)) and np.abs(g - ogr).max() < 1e-8 * max(1.0, np.abs(ogr).max())
    h = 1e-4
    disp = []
    for (A, x) in ((0, 0), (1, 2), (2, 1)):
        for sgn in (1, -1):
            c = co.copy()
            c[A, x] += sgn * h
            disp.append(c)
    E, _ = get_energy_with_grad_coords(mol, np.stack(disp), one, two, ovlp)
    fd = (E[0::2] - E[1::2]) / (2 * h)
    assert np.abs(fd - np.array([g[0, 0], g[1, 2], g[2, 1]])).max() < 2e-6 * max(1.0, np.abs(g).max())


def test_water_md_graph_replay_equals_eager_small_and_large_batches():
    """The s+p integrals inside the device MD step: with few replicas the class kernels of one call run
    concurrently on side streams (event fork / join), which must be capturable in the CUDA graph of the step and
    give bit-identical trajectories; with many replicas the single-stream path."""
    from evcont_b200.md import DeviceNVE
    from evcont_b200.mol import MolLite
    ang = 1.0 / 0.52917721092
    r, th = 0.9572 * ang, np.deg2rad(104.52)
    mol = MolLite([("O", (0, 0, 0)), ("H", (r * np.sin(th / 2), 0, r * np.cos(th / 2))),
                   ("H", (-r * np.sin(th / 2), 0, r * np.cos(th / 2)))], basis="6-31g")
    n, N = mol.nao, 3
    rng = np.random.default_rng(5)
    b = rng.standard_normal((N, N))
    ovlp = np.eye(N) + 0.01 * (b + b.T)
    one = rng.standard_normal((N, N, n, n))
    one = one + one.transpose(1, 0, 3, 2)
    two = 0.1 * rng.standard_normal((N, N) + (n,) * 4)
    two = two + two.transpose(1, 0, 3, 2, 5, 4)
    for B in (2, 24):      # <= 16: concurrent class kernels; > 16: one stream
        x0 = mol.atom_coords()[None] + 0.02 * rng.standard_normal((B, 3, 3))
        runs = {}
        for graph in (False, True):
            runs[graph] = DeviceNVE(mol, one, two, ovlp, x0, None, dt=2.0, max_frames=6, use_graph=graph).run(5).frames()
        for a, c in zip(runs[False], runs[True]):
            assert a.shape[0] == 6 and np.array_equal(a, c)
        assert np.isfinite(runs[True][1]).all()
