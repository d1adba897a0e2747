"""Host logic of ``converge_EVCont_MD`` (evcont/MD_utils.py:128-502) without a GPU: the device pieces
(trajectory, batched frame energies, farthest-point selection, FCI append) are replaced by cheap stand-ins
and the control flow is checked -- files, the "twice below the threshold" stopping rule, the three
data-addition criteria, pruning, resume."""
import os

import numpy as np
import pytest


class FakeMol:
    natm, nao = 2, 2

    def __init__(self, coords=None):
        self._c = np.zeros((2, 3)) if coords is None else np.array(coords, dtype=float)

    def copy(self):
        return FakeMol(self._c)

    def set_geom_(self, coords):
        self._c = np.array(coords, dtype=float).reshape(2, 3)
        return self

    def atom_coords(self):
        return self._c.copy()


class FakeCont:
    """Training set = list of scalars x_k (first coordinate of the training geometry)."""

    def __init__(self):
        self.points, self.overlap, self.one_rdm, self.two_rdm = [], None, None, None

    def append_to_rdms(self, mol):
        self.points.append(float(mol.atom_coords()[0, 0]))
        self._refresh()

    def _refresh(self):
        n = len(self.points)
        p = np.array(self.points)
        self.overlap = np.eye(n)
        self.one_rdm = np.tile(p[:, None, None, None], (1, n, 1, 1)) * np.ones((n, n, 2, 2))
        self.two_rdm = np.zeros((n, n, 2, 2, 2, 2))

    def prune_datapoints(self, keep_ids):
        self.points = [self.points[i] for i in keep_ids]
        self._refresh()


def _model_energy(points, x):
    """'Continuation' energy: exact value x^2 plus the squared distance to the nearest training point."""
    d = np.min(np.abs(np.asarray(points)[:, None] - x[None, :]), axis=0)
    return x ** 2 + d ** 2


@pytest.fixture
def patched(monkeypatch, tmp_path):
    pytest.importorskip("torch")
    try:
        from evcont_b200 import MD_utils as mu
    except ImportError as exc:   # the shared library is not built
        pytest.skip(str(exc))

    def points_of(one_rdm):
        return [float(one_rdm[k, 0, 0, 0]) for k in range(one_rdm.shape[0])]

    def fake_trajectory(mol, overlap, one_rdm, two_rdm, steps, dt, trajectory_output, energy_output):
        x = mol.atom_coords()[0, 0] + 0.1 * np.arange(steps)      # the atom drifts along x
        traj = np.zeros((steps, 2, 3))
        traj[:, 0, 0] = x
        e = _model_energy(points_of(one_rdm), x)
        np.savetxt(trajectory_output, np.column_stack([np.arange(steps), traj.reshape(steps, -1)]))
        np.savetxt(energy_output, np.column_stack([np.arange(steps), e, np.zeros(steps), e]))
        return traj

    monkeypatch.setattr(mu, "get_trajectory", fake_trajectory)
    monkeypatch.setattr(mu, "predict_energies",
                        lambda mol, geoms, one, two, ov: _model_energy(points_of(one), np.asarray(geoms)[:, 0, 0]))
    monkeypatch.setattr(mu, "farthest_point_ham", lambda mol, trn, traj: (
        int(np.argmax(np.min(np.abs(np.asarray(trn)[:, None, 0, 0] - np.asarray(traj)[None, :, 0, 0]), axis=0))), None))
    return mu, str(tmp_path)


@pytest.mark.parametrize("criterion", ["farthest_point_ham", "farthest_point", "energy"])
def test_loop_files_and_stopping_rule(patched, criterion):
    mu, wd = patched
    cont = FakeCont()
    traj = mu.converge_EVCont_MD(cont, FakeMol(), steps=11, dt=1.0, convergence_thresh=2e-3, data_addition=criterion,
                                 workdir=wd)
    n = len(cont.points)
    assert traj.shape == (11, 2, 3) and 3 <= n <= 13
    times = np.loadtxt(os.path.join(wd, "trn_times.txt")).astype(int)
    assert len(times) == n and times[0] == 0 and cont.points[0] == 0.0
    # every training point is the chosen frame of the previous trajectory (x = 0.1 * frame)
    assert np.allclose(cont.points[1:], 0.1 * times[1:])
    # stop only after two consecutive iterations below the threshold; the first file is the start-point reference
    diffs = [np.max(np.atleast_1d(np.loadtxt(os.path.join(wd, f"en_diff_{k}.txt")))) for k in range(n)]
    assert diffs[-1] <= 2e-3 and diffs[-2] <= 2e-3 and max(diffs[1:-2] + [1.0]) > 2e-3
    for k in range(n):
        assert os.path.exists(os.path.join(wd, f"traj_EVCont_{k}.npy")) and os.path.exists(os.path.join(wd, f"ens_EVCont_{k}.xyz"))
    assert np.load(os.path.join(wd, "overlap.npy")).shape == (n, n)
    # en_diff_k = |E(previous set) - E(current set)| along trajectory k
    x = 0.1 * np.arange(11)
    ref = np.abs(_model_energy(cont.points[:-1], x) - _model_energy(cont.points, x))
    assert np.allclose(np.loadtxt(os.path.join(wd, f"en_diff_{n - 1}.txt")), ref)


def test_first_addition_is_the_farthest_frame(patched):
    mu, wd = patched
    cont = FakeCont()
    mu.converge_EVCont_MD(cont, FakeMol(), steps=11, dt=1.0, convergence_thresh=1e-6, workdir=wd, max_iterations=2)
    times = np.loadtxt(os.path.join(wd, "trn_times.txt")).astype(int)
    assert list(times) == [0, 10, 5]        # farthest from {0}: frame 10; then from {0, 1.0}: frame 5


def test_resume_continues_the_interrupted_run(patched):
    mu, wd = patched
    a = FakeCont()
    mu.converge_EVCont_MD(a, FakeMol(), steps=11, dt=1.0, convergence_thresh=2e-3, workdir=wd, max_iterations=2)
    times = [int(t) for t in np.loadtxt(os.path.join(wd, "trn_times.txt"))]
    assert len(a.points) == 3
    mu.converge_EVCont_MD(a, FakeMol(), steps=11, dt=1.0, convergence_thresh=2e-3, workdir=wd, trn_times=times)
    ref_dir = os.path.join(wd, "ref")
    os.makedirs(ref_dir)
    b = FakeCont()
    mu.converge_EVCont_MD(b, FakeMol(), steps=11, dt=1.0, convergence_thresh=2e-3, workdir=ref_dir)
    assert a.points == b.points


def test_pruning_removes_redundant_points(patched):
    mu, wd = patched
    cont = FakeCont()
    mu.converge_EVCont_MD(cont, FakeMol(), steps=11, dt=1.0, convergence_thresh=0.3, prune_irrelevant_data=True,
                          workdir=wd)
    # with a loose threshold most points are redundant: the pruned set is small and the files carry the suffix
    assert 1 <= len(cont.points) <= 3
    assert os.path.exists(os.path.join(wd, "overlap_0.npy")) and os.path.exists(os.path.join(wd, "trn_times_1.txt"))


def test_unknown_criterion_asserts(patched):
    mu, wd = patched
    with pytest.raises(AssertionError):
        mu.converge_EVCont_MD(FakeCont(), FakeMol(), steps=5, dt=1.0, data_addition="nonsense", workdir=wd)
