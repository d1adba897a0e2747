"""f4 observables on the device (csrc/observables.cu, evcont_b200/observables.py) against oracle/observables.py:
the dipole integrals, the dipole moment and atomic charges of a density matrix, and the batched form over the
continuation's predicted density matrices (the reference's per-step MD callback,
04_Zundel_continuation_MD.py:71-92,140-159)."""
import numpy as np
import pytest

from conftest import synthetic_stack

pytestmark = pytest.mark.gpu

H2O = np.array([[0.0, 0.0, 0.2], [0.0, 1.45, -0.9], [0.1, -1.40, -0.95]])
ZUNDEL = np.array([[0.0, 0.05, -2.27], [0.0, -0.05, 2.27], [0.0, 0.0, 0.02], [1.5, 0.3, -3.0], [-1.5, 0.4, -3.05],
                   [1.45, -0.35, 3.0], [-1.52, -0.3, 3.02]])


def _cases():
    from evcont_b200.mol import MolLite
    from oracle import integrals_sp as osp
    h4 = [[0.0, 0.0, 1.7 * k + 0.05 * k * k] for k in range(4)]
    yield (MolLite([("H", c) for c in h4], "sto-6g"), osp.SPBasis([("H", c) for c in h4], "sto-6g"))
    syms = ["O", "H", "H"]
    yield (MolLite(list(zip(syms, H2O)), "6-31g"), osp.SPBasis(list(zip(syms, H2O)), "6-31g"))
    syms = ["O", "O", "H", "H", "H", "H", "H"]
    yield (MolLite(list(zip(syms, ZUNDEL)), "6-31g"), osp.SPBasis(list(zip(syms, ZUNDEL)), "6-31g"))


def test_int1e_r_and_com():
    from evcont_b200.engine import get_engine
    from oracle import observables as oob
    eng = get_engine()
    for mol, b in _cases():
        t = eng.aotable([mol.atom_symbol(i) for i in range(mol.natm)], mol.basis)
        com = eng.center_of_mass(t, mol.atom_coords()[None])[0].cpu().numpy()
        assert np.abs(com - oob.center_of_mass(b)).max() < 1e-13
        assert np.abs(com - (mol.atom_mass_list()[:, None] * mol.atom_coords()).sum(0) / mol.atom_mass_list().sum()).max() < 1e-13
        for origin in (np.zeros(3), np.array([0.3, -1.2, 0.8])):
            with mol.with_common_orig(origin):
                r = mol.intor_symmetric("int1e_r", comp=3)
            assert np.abs(r - oob.int1e_r(b, origin)).max() < 1e-12


def test_dipole_and_charges_of_a_density_matrix():
    from evcont_b200 import observables as ob
    from oracle import observables as oob
    rng = np.random.default_rng(8)
    for mol, b in _cases():
        n = mol.nao
        S = mol.intor("int1e_ovlp")
        w, V = np.linalg.eigh(S)
        X = (V / np.sqrt(w)) @ V.T
        g = rng.standard_normal((n, n))
        g = g + g.T
        g *= mol.nelectron / np.trace(g)          # a symmetric "density matrix" with the right electron count
        dm = X @ g @ X.T
        for unit in ("Debye", "au"):
            assert np.abs(ob.dip_moment(mol, dm, unit) - oob.dip_moment(b, dm, unit)).max() < 1e-10
        for m in ("mulliken", "loewdin"):
            q = ob.atomic_charges(mol, dm, m)
            assert np.abs(q - oob.atomic_charges(b, dm, S, m)).max() < 1e-11
            assert abs(q.sum()) < 1e-10


def test_predicted_observables_along_a_trajectory(tmp_path):
    """The batched device form vs the reference's callback run frame by frame through the oracle
    (get_energy_with_grad -> predicted_one_rdm -> X gamma X^T -> dip_moment / charges)."""
    from evcont_b200 import observables as ob
    from oracle import integrals_sp as osp
    from oracle import observables as oob
    from oracle import subspace as osub
    rng = np.random.default_rng(4)
    for (mol, b), ntrain in zip(_cases(), (4, 3, 2)):
        layout = 2 if mol.nao > 13 else 6
        ovlp, one, two = synthetic_stack(mol.nao, ntrain, 77, layout)
        frames = mol.atom_coords()[None] + 0.05 * rng.standard_normal((5, mol.natm, 3))
        refs = []
        for k in range(len(frames) if mol.nao <= 13 else 1):      # one Zundel frame: its oracle integrals take a minute
            bk = b.with_coords(frames[k])
            arr = osp.ao_arrays(bk, want_ip1=False)
            x = osub.get_loewdin_trafo(arr["ovlp"])
            h1, h2 = osub.ao_to_oao(arr["hcore"], arr["eri"], x)
            _, vec = osub.approximate_ground_state(h1, h2, one, two, ovlp)
            gamma = np.tensordot(np.outer(vec, vec), one, axes=2)
            refs.append((bk, arr["ovlp"], x @ gamma @ x.T))
        for method in ("mulliken", "loewdin"):
            dip, chg, dm = ob.predicted_observables(mol, frames, one, two, ovlp, method=method, return_dm_ao=True)
            for k, (bk, S, dm_ref) in enumerate(refs):
                assert np.abs(dm[k] - dm_ref).max() < 1e-9
                assert np.abs(dip[k] - oob.dip_moment(bk, dm_ref)).max() < 1e-8
                assert np.abs(chg[k] - oob.atomic_charges(bk, dm_ref, S, method)).max() < 1e-9
    # the callback's two text files: one line per frame
    ob.write_observables(dip, chg, tmp_path / "dip.txt", tmp_path / "chg.txt", mode="w")
    assert np.abs(np.loadtxt(tmp_path / "dip.txt") - dip).max() < 1e-12
    assert np.loadtxt(tmp_path / "chg.txt").shape == chg.shape
