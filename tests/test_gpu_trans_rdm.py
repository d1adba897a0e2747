"""K0/K1/K2 parity: the CUDA trans_rdm12 (through the C ABI) against the CPU oracle,
the definition-based goldens and size-independent invariants.  Tolerance from
BASELINE.json north_star: 1e-12 absolute on transition RDMs."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, random_civec

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def solver():
    from evcont_b200.fci import B200FCISolver
    return B200FCISolver()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "trans_rdm_*.npz"))))
def test_golden_definition_vectors(solver, path):
    g = np.load(path)
    norb, nelec = int(g["norb"]), tuple(int(x) for x in g["nelec"])
    dm1, dm2 = solver.trans_rdm12(g["bra"], g["ket"], norb, nelec)
    assert np.abs(dm1 - g["dm1"]).max() < TOL
    assert np.abs(dm2 - g["dm2"]).max() < TOL


def test_hand_value(solver):
    c = np.zeros((2, 2))
    c[0, 0] = 1.0
    dm1, dm2 = solver.trans_rdm12(c, c, 2, (1, 1))
    assert np.array_equal(dm1, np.diag([2.0, 0.0]))
    ref = np.zeros((2,) * 4)
    ref[0, 0, 0, 0] = 2.0
    assert np.array_equal(dm2, ref)


@pytest.mark.parametrize("norb,nelec", [
    (1, (1, 1)), (2, (1, 1)), (3, (1, 1)), (3, (2, 1)), (4, (2, 2)), (4, (1, 3)), (5, (2, 2)),
    (5, (3, 2)), (6, (3, 3)), (6, (1, 1)), (7, (3, 3)), (7, (4, 2)), (8, (4, 4)), (8, (2, 2)),
    (9, (3, 3)), (10, (2, 2)), (11, (2, 2)), (12, (2, 2)), (13, (2, 2)), (13, (1, 2)),
    (6, (0, 0)), (6, (6, 6)), (5, (5, 0)),
    # producer / consumer kernel (norb = 10, 11): even and odd beta counts, na != nb, one- and two-step CTAs
    (10, (2, 3)), (10, (3, 2)), (10, (0, 1)), (10, (1, 0)), (10, (10, 10)), (10, (1, 4)), (11, (1, 3)), (11, (3, 3)),
])
def test_against_oracle(solver, norb, nelec):
    from oracle import cistring as ocs, trans_rdm as otr
    na, nb = ocs.num_strings(norb, nelec[0]), ocs.num_strings(norb, nelec[1])
    bra, ket = random_civec(na, nb, 1), random_civec(na, nb, 2)
    dm1, dm2 = solver.trans_rdm12(bra, ket, norb, nelec)
    o1, o2 = otr.trans_rdm12(bra, ket, norb, nelec)
    assert dm1.shape == (norb, norb) and dm2.shape == (norb,) * 4
    assert np.abs(dm1 - o1).max() < TOL
    assert np.abs(dm2 - o2).max() < TOL


def test_int_nelec_and_reorder_false(solver):
    from oracle import trans_rdm as otr
    bra, ket = random_civec(6, 6, 3), random_civec(6, 6, 4)
    dm1, dm2 = solver.trans_rdm12(bra, ket, 4, 4, reorder=False)
    o1, o2 = otr.trans_rdm12(bra, ket, 4, 4, reorder=False)
    assert np.abs(dm1 - o1).max() < TOL and np.abs(dm2 - o2).max() < TOL


def test_h6_batch_all_pairs(solver):
    """H6 STO-6G size (configs[0]): 3 vectors, all 9 ordered pairs in one launch."""
    from oracle import trans_rdm as otr
    vecs = np.stack([random_civec(20, 20, 1000 + k, symmetric=True) for k in range(3)])
    pairs = [(a, b) for a in range(3) for b in range(3)]
    ovlp, dm1, dm2 = solver.trans_rdm12_batch(vecs, pairs, 6, (3, 3))
    for k, (a, b) in enumerate(pairs):
        o1, o2 = otr.trans_rdm12(vecs[a], vecs[b], 6, (3, 3))
        assert abs(ovlp[k] - (vecs[a] * vecs[b]).sum()) < TOL
        assert np.abs(dm1[k] - o1).max() < TOL
        assert np.abs(dm2[k] - o2).max() < TOL


def test_h10_pair_vs_oracle(solver):
    """H10 STO-6G size (configs[1]): 63 504 determinants per vector."""
    from oracle import trans_rdm as otr
    bra, ket = random_civec(252, 252, 1000), random_civec(252, 252, 1001)
    dm1, dm2 = solver.trans_rdm12(bra, ket, 10, (5, 5))
    o1, o2 = otr.trans_rdm12(bra, ket, 10, (5, 5))
    assert np.abs(dm1 - o1).max() < TOL
    assert np.abs(dm2 - o2).max() < TOL


def _check_invariants(dm1, dm2, ovlp, nelec_tot, tol):
    n = dm1.shape[0]
    assert abs(np.trace(dm1) - nelec_tot * ovlp) < tol
    assert np.abs(np.einsum("pqrr->pq", dm2) - (nelec_tot - 1) * dm1.T).max() < tol
    assert abs(np.einsum("pprr->", dm2) - nelec_tot * (nelec_tot - 1) * ovlp) < tol
    assert np.abs(dm2 - dm2.transpose(2, 3, 0, 1)).max() < tol
    return n


def test_h10_full_stack_invariants_and_determinism(solver):
    """BASELINE size: 20 training vectors, all 210 pairs a >= b; sum rules, bra<->ket
    swap symmetry, and run-to-run bit identity."""
    N = 20
    vecs = np.stack([random_civec(252, 252, 1000 + k) for k in range(N)])
    pairs = [(a, b) for a in range(N) for b in range(a + 1)]
    ovlp, dm1, dm2 = solver.trans_rdm12_batch(vecs, pairs, 10, (5, 5))
    for k, (a, b) in enumerate(pairs):
        assert abs(ovlp[k] - (vecs[a] * vecs[b]).sum()) < TOL
        _check_invariants(dm1[k], dm2[k], ovlp[k], 10, 2e-11)
    # swap symmetry on a few pairs: dm(bra,ket)[p,q] = dm(ket,bra)[q,p]
    sw = [(3, 7), (0, 19)]
    o2, d1s, d2s = solver.trans_rdm12_batch(vecs, [(b, a) for a, b in sw], 10, (5, 5))
    for j, (a, b) in enumerate(sw):
        k = pairs.index((b, a)) if (b, a) in pairs else pairs.index((a, b))
        assert np.abs(dm1[k] - d1s[j].T).max() < TOL or np.abs(dm1[k] - d1s[j]).max() < TOL
    ovlp2, dm1b, dm2b = solver.trans_rdm12_batch(vecs, pairs, 10, (5, 5))
    assert np.array_equal(dm2, dm2b) and np.array_equal(dm1, dm1b) and np.array_equal(ovlp, ovlp2)


def test_swap_symmetry_small(solver):
    bra, ket = random_civec(10, 10, 5), random_civec(10, 10, 6)
    a1, a2 = solver.trans_rdm12(bra, ket, 5, (2, 2))
    b1, b2 = solver.trans_rdm12(ket, bra, 5, (2, 2))
    assert np.abs(a1 - b1.T).max() < TOL
    assert np.abs(a2 - b2.transpose(1, 0, 3, 2)).max() < TOL


def test_h2o_size_one_pair_invariants(solver):
    """H2O 6-31G size (configs[2]): norb 13, (5,5), 1287^2 determinants, odd row length."""
    from oracle import trans_rdm as otr
    bra, ket = random_civec(1287, 1287, 1000), random_civec(1287, 1287, 1001)
    dm1, dm2 = solver.trans_rdm12(bra, ket, 13, (5, 5))
    _check_invariants(dm1, dm2, float((bra * ket).sum()), 10, 5e-11)
    o1, o2 = otr.trans_rdm12(bra, ket, 13, (5, 5), block=32)
    assert np.abs(dm1 - o1).max() < TOL
    assert np.abs(dm2 - o2).max() < TOL
