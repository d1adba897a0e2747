"""Device-resident velocity Verlet (evcont_b200.md) against the numpy oracle (oracle/md.py) fed by
the oracle integrals + the numpy port of get_energy_with_grad; energy conservation; CUDA-graph
replay == eager; batch == single."""
import numpy as np
import pytest

from conftest import synthetic_stack

pytestmark = pytest.mark.gpu

H4 = np.array([[0.0, 0.0, 0.0], [0.1, 0.2, 1.7], [0.3, -0.2, 3.5], [1.5, 0.3, 0.5]])


def _setup(n=4, N=3):
    from evcont_b200.mol import MolLite
    ovlp, one, two = synthetic_stack(n, N, 11, 5)
    # a physically shaped stack is not needed for integrator parity, but keep forces moderate
    mol = MolLite([("H", tuple(c)) for c in H4[:n]], basis="sto-6g", unit="Bohr")
    return mol, ovlp, 0.05 * one, 0.05 * two


def _oracle_force(mol, one, two, ovlp):
    from evcont_b200.mol import ArrayMol
    from oracle import gradients as og
    from oracle import integrals as oi

    def f(x):
        arr = oi.ao_arrays(oi.SBasis([("H", c) for c in x], mol.basis))
        return og.get_energy_with_grad(ArrayMol(**arr), one, two, ovlp)
    return f


def test_trajectory_against_oracle():
    from evcont_b200.MD_utils import get_trajectory
    from evcont_b200.md import atomic_masses
    from oracle import md as omd
    mol, ovlp, one, two = _setup()
    rng = np.random.default_rng(2)
    v0 = 1e-4 * rng.standard_normal((mol.natm, 3))
    steps, dt = 12, 5.0
    traj = get_trajectory(mol, ovlp, one, two, dt=dt, steps=steps, init_veloc=v0)
    m = atomic_masses(mol)
    assert abs(m[0] - omd.COMMON_ISOTOPE_MASSES["H"] * omd.AMU2AU) < 1e-9
    ref, epot, ekin = omd.velocity_verlet(mol.atom_coords(), v0, m, dt, steps, _oracle_force(mol, one, two, ovlp))
    assert traj.shape == (steps, mol.natm, 3)
    assert np.array_equal(traj[0], mol.atom_coords())
    assert np.abs(traj - ref).max() < 1e-9


def test_graph_replay_equals_eager_and_batch_equals_single():
    from evcont_b200.md import DeviceNVE
    mol, ovlp, one, two = _setup()
    rng = np.random.default_rng(3)
    x0 = mol.atom_coords()[None] + 0.05 * rng.standard_normal((5, mol.natm, 3))
    v0 = 1e-4 * rng.standard_normal((5, mol.natm, 3))
    runs = {}
    for graph in (False, True):
        nve = DeviceNVE(mol, one, two, ovlp, x0, v0, dt=5.0, max_frames=9, use_graph=graph).run(8)
        runs[graph] = nve.frames()
    for a, b in zip(runs[False], runs[True]):
        assert a.shape[0] == 9 and np.array_equal(a, b)
    single = DeviceNVE(mol, one, two, ovlp, x0[2:3], v0[2:3], dt=5.0, max_frames=9, use_graph=False).run(8).frames()
    assert np.abs(single[0][:, 0] - runs[False][0][:, 2]).max() < 1e-10


def test_energy_conservation_second_order():
    """NVE on a continuation surface built from exact training vectors is beyond a unit test's
    budget; a random symmetric stack still defines a smooth surface E(R), on which velocity Verlet
    must conserve E_pot + E_kin to O(dt^2)."""
    from evcont_b200.md import DeviceNVE
    mol, ovlp, one, two = _setup()
    drift = []
    for dt in (4.0, 2.0):
        nve = DeviceNVE(mol, one, two, ovlp, mol.atom_coords()[None], None, dt=dt, max_frames=int(160 / dt) + 1)
        nve.run(int(160 / dt))
        _, epot, ekin = nve.frames()
        etot = epot[:, 0] + ekin[:, 0]
        drift.append(np.abs(etot - etot[0]).max())
    assert drift[0] < 1e-2 and drift[1] < 0.4 * drift[0] + 1e-12, drift


def test_berendsen_thermostat_against_oracle():
    from evcont_b200.md import DeviceNVT, atomic_masses
    from oracle import md as omd
    mol, ovlp, one, two = _setup()
    rng = np.random.default_rng(5)
    v0 = 3e-4 * rng.standard_normal((mol.natm, 3))
    steps, dt = 10, 5.0
    nvt = DeviceNVT(mol, one, two, ovlp, mol.atom_coords()[None], v0[None], dt=dt, T=298.15, taut=250.0,
                    max_frames=steps).run(steps - 1)
    traj, _, ekin = nvt.frames()
    ref, _, ekin_ref = omd.velocity_verlet(mol.atom_coords(), v0, atomic_masses(mol), dt, steps,
                                          _oracle_force(mol, one, two, ovlp), berendsen=(298.15, 250.0))
    assert np.abs(traj[:, 0] - ref).max() < 1e-9
    assert np.abs(ekin[:, 0] - ekin_ref).max() < 1e-11
