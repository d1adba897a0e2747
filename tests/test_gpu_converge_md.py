"""``converge_EVCont_MD`` (evcont/MD_utils.py:128-502) with its heavy steps batched
on the device -- the batched frame energies and the farthest-point selection against the reference's
per-frame formulas, and the whole active-learning loop on an H4 chain: files, stopping rule, resume."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _h_chain(n, d):
    from evcont_b200.mol import MolLite
    xs = (np.arange(n) - np.median(np.arange(n))) * d
    return MolLite([("H", (x, 0.0, 0.0)) for x in xs], basis="sto-6g", unit="Bohr")


def _mk(path):
    os.makedirs(str(path), exist_ok=True)
    return str(path)


def _frames(mol, nframes, seed):
    rng = np.random.default_rng(seed)
    return mol.atom_coords()[None] + 0.15 * rng.standard_normal((nframes, mol.natm, 3))


def test_batched_frame_energies_match_the_per_frame_loop():
    from evcont_b200.FCI_EVCont import FCI_EVCont_obj
    from evcont_b200.MD_utils import predict_energies
    from evcont_b200.ab_initio_eigenvector_continuation import approximate_ground_state_OAO
    cont = FCI_EVCont_obj(cibasis="OAO")
    for d in (1.4, 2.0):
        cont.append_to_rdms(_h_chain(4, d))
    mol = _h_chain(4, 1.7)
    frames = _frames(mol, 7, 1)
    for sl in (slice(None), slice(0, 1)):   # full training set, and the "[:-1, :-1]" one of the loop
        one, two, ov = cont.one_rdm[sl, sl], cont.two_rdm[sl, sl], cont.overlap[sl, sl]
        e = predict_energies(mol, frames, one, two, ov)
        ref = [approximate_ground_state_OAO(mol.copy().set_geom_(g), one, two, ov)[0] for g in frames]
        assert np.abs(e - np.array(ref)).max() < 1e-10


def test_farthest_point_in_hamiltonian_space():
    from evcont_b200.MD_utils import farthest_point_ham
    from evcont_b200.electron_integral_utils import get_basis, get_integrals
    mol = _h_chain(4, 1.7)
    frames, trn = _frames(mol, 9, 2), _frames(mol, 3, 3)
    idx, d = farthest_point_ham(mol, trn, frames)
    # the reference's loop (evcont/MD_utils.py:383-405)
    ht = [get_integrals(m, get_basis(m)) for m in (mol.copy().set_geom_(g) for g in trn)]
    h1t, h2t = np.array([h[0] for h in ht]), np.array([h[1] for h in ht])
    best, ref_idx, ref_d = None, 0, []
    for j, g in enumerate(frames):
        m = mol.copy().set_geom_(g)
        h1, h2 = get_integrals(m, get_basis(m))
        dist = np.sum(abs(h1 - h1t) ** 2, axis=(-1, -2)) + 0.5 * np.sum(abs(h2 - h2t) ** 2, axis=(-1, -2, -3, -4))
        ref_d.append(np.min(dist))
        if best is None or ref_d[-1] > best:
            best, ref_idx = ref_d[-1], j
    assert idx == ref_idx
    assert np.abs(d - np.array(ref_d)).max() < 1e-12


@pytest.mark.parametrize("data_addition", ["farthest_point_ham", "energy", "farthest_point"])
def test_active_learning_loop_h4(tmp_path, data_addition):
    from evcont_b200.FCI_EVCont import FCI_EVCont_obj
    from evcont_b200.MD_utils import converge_EVCont_MD, predict_energies
    wd = str(tmp_path)
    mol = _h_chain(4, 1.5)         # compressed chain: the atoms fly apart, the surface changes along the way
    cont = FCI_EVCont_obj(cibasis="OAO")
    traj = converge_EVCont_MD(cont, mol, steps=25, dt=8.0, convergence_thresh=1e-4, data_addition=data_addition,
                              workdir=wd)
    N = len(cont.fcivecs)
    assert 2 <= N <= 12 and cont.overlap.shape == (N, N) and traj.shape == (25, 4, 3)
    assert np.abs(traj[0] - mol.atom_coords()).max() < 1e-14
    # files of the reference protocol
    for name in ("overlap.npy", "one_rdm.npy", "two_rdm.npy", "trn_times.txt", "traj_EVCont_0.npy", "ens_EVCont_0.xyz",
                 "en_diff_0.txt", f"traj_EVCont_{N - 1}.npy", f"en_diff_{N - 1}.txt"):
        assert os.path.exists(os.path.join(wd, name)), name
    assert np.load(os.path.join(wd, "two_rdm.npy")).shape == cont.two_rdm.shape
    trn_times = np.loadtxt(os.path.join(wd, "trn_times.txt")).astype(int)
    assert len(trn_times) == N and trn_times[0] == 0
    # stopping rule: the last two energy-difference files are below the threshold, earlier ones are not all
    last = [np.max(np.loadtxt(os.path.join(wd, f"en_diff_{k}.txt"))) for k in (N - 2, N - 1)]
    assert max(last) <= 1e-4
    assert np.max(np.loadtxt(os.path.join(wd, "en_diff_0.txt"))) > 1e-4
    # the recorded difference is |E(previous set) - E(current set)| along the last trajectory
    e_prev = predict_energies(mol, traj, cont.one_rdm[:-1, :-1], cont.two_rdm[:-1, :-1], cont.overlap[:-1, :-1])
    e_cur = predict_energies(mol, traj, cont.one_rdm, cont.two_rdm, cont.overlap)
    assert np.abs(np.abs(e_prev - e_cur) - np.loadtxt(os.path.join(wd, f"en_diff_{N - 1}.txt"))).max() < 1e-8
    # every added training geometry is the chosen frame of the previous trajectory
    for k in range(1, N):
        prev = np.load(os.path.join(wd, f"traj_EVCont_{k - 1}.npy"))
        assert prev.shape == (25, 4, 3) and 0 <= trn_times[k] < 25


def test_resume_from_files(tmp_path):
    from evcont_b200.FCI_EVCont import FCI_EVCont_obj
    from evcont_b200.MD_utils import converge_EVCont_MD
    wd = str(tmp_path)
    mol = _h_chain(4, 1.5)
    cont = FCI_EVCont_obj(cibasis="OAO")
    converge_EVCont_MD(cont, mol, steps=25, dt=8.0, convergence_thresh=1e-4, workdir=wd, max_iterations=1)
    assert len(cont.fcivecs) == 2
    times = [int(t) for t in np.loadtxt(os.path.join(wd, "trn_times.txt"))]
    full = FCI_EVCont_obj(cibasis="OAO")
    ref_traj = converge_EVCont_MD(full, mol, steps=25, dt=8.0, convergence_thresh=1e-4, workdir=_mk(tmp_path / "ref"))
    # resume the truncated run: same training times and final trajectory as the uninterrupted one
    traj = converge_EVCont_MD(cont, mol, steps=25, dt=8.0, convergence_thresh=1e-4, workdir=wd, trn_times=times)
    assert len(cont.fcivecs) == len(full.fcivecs)
    assert np.abs(traj - ref_traj).max() < 1e-8
