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


class _PyscfLikeMole:
    """The slice of ``pyscf.gto.Mole`` that the reference's scripts hand to ``get_trajectory``
    (evcont/MD_utils.py:60-125): geometry in bohr, a named basis, charge and spin."""

    def __init__(self, symbols, coords, basis, charge=0, spin=0):
        self._s, self._c = list(symbols), np.asarray(coords, dtype=float)
        self.natm, self.basis, self.charge, self.spin = len(self._s), basis, charge, spin

    def atom_symbol(self, i):
        return self._s[i]

    def atom_coords(self):
        return self._c.copy()


def test_mollite_from_pyscf_like_mole():
    from evcont_b200.mol import MolLite
    co = np.array([[0.0, 0.0, 0.0], [1.8, 0.1, 0.0], [3.5, 0.0, -0.2], [5.4, 0.0, 0.0]])
    ref = MolLite([("H", tuple(c)) for c in co], basis="sto-6g", unit="Bohr")
    got = MolLite.from_mol(_PyscfLikeMole(["H"] * 4, co, "sto-6g"))
    assert isinstance(got, MolLite) and got.natm == 4 and got.nelec == ref.nelec and got.basis == "sto-6g"
    assert np.array_equal(got.atom_coords(), ref.atom_coords())
    assert [got.atom_symbol(i) for i in range(4)] == ["H"] * 4
    assert MolLite.from_mol(ref) is ref                                  # a MolLite passes through untouched
    ion = MolLite.from_mol(_PyscfLikeMole(["H"] * 3, co[:3], "sto-6g", charge=1))
    assert ion.charge == 1 and sum(ion.nelec) == 2
    import pytest
    with pytest.raises(TypeError):
        MolLite.from_mol(_PyscfLikeMole(["H"] * 4, co, {"H": [[0, (1.0, 1.0)]]}))   # explicit shells: no device table
    with pytest.raises(TypeError):
        MolLite.from_mol(object())
