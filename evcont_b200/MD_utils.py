"""``get_scanner``: the drop-in boundary for MD drivers (evcont/MD_utils.py:20-57).

Returns an object PySCF's ``md.NVE`` / ``md.NVTBerendson`` (or any velocity-Verlet
loop) can call as ``scanner(mol) -> (E_total, grad)``.  The trajectory drivers
``get_trajectory`` / ``converge_EVCont_MD`` are host orchestration over PySCF's
integrators and are not part of this package.
"""
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
