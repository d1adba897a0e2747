"""The files get_trajectory writes (evcont/MD_utils.py:109-120 hands ``trajectory_output`` / ``energy_output`` to
pyscf.md.NVE): xyz frames headed "MD Time", and the time / Epot / Ekin / Etot table the reference reads back with
np.genfromtxt(...)[:, 1] (evcont/MD_utils.py:208)."""
import io

import numpy as np

from evcont_b200.MD_utils import read_md_energies, write_md_energies, write_xyz_trajectory
from evcont_b200.mol import BOHR


def test_xyz_frames(tmp_path):
    traj = np.arange(2 * 3 * 3, dtype=float).reshape(2, 3, 3) * 0.25
    path = tmp_path / "traj.xyz"
    write_xyz_trajectory(path, ["O", "H", "H"], traj, [0.0, 10.0])
    lines = path.read_text().splitlines()
    assert len(lines) == 2 * (2 + 3)
    assert lines[0] == "3" and lines[1] == "MD Time 0.00" and lines[6] == "MD Time 10.00"
    sym, x, y, z = lines[7].split()
    assert sym == "O" and np.allclose([float(x), float(y), float(z)], traj[1, 0] * BOHR, atol=1e-8)
    buf = io.StringIO()                      # file objects, as the reference passes them (:185-186)
    write_xyz_trajectory(buf, ["O", "H", "H"], traj, [0.0, 10.0])
    assert buf.getvalue() == path.read_text()


def test_energy_table_round_trip(tmp_path):
    t = 5.0 * np.arange(4)
    ep, ek = -1.0 - 0.1 * np.arange(4), 0.01 * np.arange(4)
    path = tmp_path / "ens.xyz"
    write_md_energies(path, t, ep, ek)
    first = path.read_text().splitlines()
    assert first[0].split() == ["time", "Epot", "Ekin", "Etot"]
    assert first[1] == "%8.2f  %.12E  %.12E  %.12E" % (0.0, -1.0, 0.0, -1.0)
    tab = read_md_energies(path)
    assert tab.shape == (4, 4)
    assert np.allclose(tab[:, 1], ep) and np.allclose(tab[:, 3], ep + ek) and np.allclose(tab[:, 0], t)
    np.savetxt(path, np.column_stack([t, ep, ek, ep + ek]))      # a table without the header reads the same
    assert np.allclose(read_md_energies(path)[:, 1], ep)
