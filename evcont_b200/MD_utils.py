"""``get_scanner``: the drop-in boundary for MD drivers (evcont/MD_utils.py:20-57).

Returns an object PySCF's ``md.NVE`` / ``md.NVTBerendson`` (or any velocity-Verlet
loop) can call as ``scanner(mol) -> (E_total, grad)``.  ``get_trajectory`` runs the NVE
trajectory itself on the device (``evcont_b200.md.DeviceNVE``) when given a
:class:`evcont_b200.mol.MolLite`; ``converge_EVCont_MD`` (the active-learning outer loop over
FCI solves, evcont/MD_utils.py:128-502) is host orchestration and not part of this package.
"""
import numpy as np

from .ab_initio_gradients_loewdin import get_energy_with_grad
from .mol import ao_bundle

try:  # subclass PySCF's GradScanner when it exists so isinstance checks in pyscf.md pass
    from pyscf.lib import GradScanner as _ScannerBase
except ImportError:
    _ScannerBase = object


def get_scanner(mol, one_rdm, two_rdm, overlap, hermitian=True):
    """Fake gradient scanner over an eigenvector continuation (same attributes as the
    reference: ``.mol``, ``.base.converged/ovlp/one_trdm/two_trdm/predicted_*``)."""

    class Base:
        converged = True
        ovlp = overlap
        one_trdm = one_rdm
        two_trdm = two_rdm
        predicted_one_rdm = None
        predicted_two_rdm = None

    class Scanner(_ScannerBase):
        def __init__(self):
            self.mol = mol
            self.base = Base()

        def __call__(self, mol):
            self.mol = mol
            if one_rdm is not None and two_rdm is not None and overlap is not None:
                en, grad, rdm_o, rdm_t = get_energy_with_grad(
                    mol, one_rdm, two_rdm, overlap, hermitian=hermitian,
                    return_density_matrices=True)
                self.base.predicted_one_rdm = rdm_o
                self.base.predicted_two_rdm = rdm_t
                return en, grad
            b = ao_bundle(mol)
            return b["e_nuc"], b["grad_nuc"]

    return Scanner()


def get_trajectory(init_mol, overlap, one_rdm, two_rdm, dt=10.0, steps=10, init_veloc=None, hermitian=True,
                   trajectory_output=None, energy_output=None):
    """MD trajectory from the continuation, ``(steps, natm, 3)`` in bohr (evcont/MD_utils.py:60-125:
    ``pyscf.md.NVE`` with ``frames=[]``; frame 0 is the initial geometry).  The integrator runs on
    the device; ``trajectory_output`` / ``energy_output`` (file names or file objects) receive plain
    text tables (step, coordinates) / (step, E_pot, E_kin, E_tot) instead of PySCF's formats."""
    if hermitian is not True:
        raise NotImplementedError("hermitian=False is not implemented on the device")
    from .md import DeviceNVE
    from .mol import MolLite
    if not isinstance(init_mol, MolLite):
        raise TypeError("get_trajectory integrates on the device and needs an evcont_b200.mol.MolLite "
                        "(use get_scanner with pyscf.md for a pyscf Mole)")
    veloc = None if init_veloc is None else np.asarray(init_veloc, dtype=np.float64)[None]
    nve = DeviceNVE(init_mol, one_rdm, two_rdm, overlap, init_mol.atom_coords()[None], veloc, dt=dt,
                    max_frames=steps)
    nve.run(steps - 1)
    traj, epot, ekin = nve.frames()
    if trajectory_output is not None:
        np.savetxt(trajectory_output, np.column_stack([np.arange(len(traj)), traj[:, 0].reshape(len(traj), -1)]))
    if energy_output is not None:
        np.savetxt(energy_output, np.column_stack([np.arange(len(traj)), epot[:, 0], ekin[:, 0],
                                                   epot[:, 0] + ekin[:, 0]]))
    return traj[:, 0]
