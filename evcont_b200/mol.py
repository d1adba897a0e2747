"""``mol`` duck types for the EVCont prediction path.

The reference drives its prediction path off a ``pyscf.gto.Mole``; the methods
it actually touches are listed in SURVEY.md section 8(b) ("mol duck type"):
``nao, natm, nelec, intor(name, comp=), aoslice_by_atom(), energy_nuc()`` plus
the PySCF free functions ``scf.hf.get_hcore(mol)``,
``grad.RHF(scf.RHF(mol)).hcore_generator() / .grad_nuc()`` and
``ao2mo.kernel(mol, C)`` (evcont/ab_initio_gradients_loewdin.py:25,130,147,
177,283,284,338,339,370).

:class:`ArrayMol` carries those AO arrays directly (synthetic inputs for tests
and the benchmark, or arrays exported once from a real ``Mole``), exposing the
PySCF free functions as methods.  :func:`ao_bundle` collects the arrays of any
supported ``mol`` into one dict, which is what the device path consumes.
"""
import numpy as np

_INTOR_NAMES = ("int1e_ovlp", "int1e_ipovlp", "int2e", "int2e_ip1")


class ArrayMol:
    """A molecule reduced to the AO arrays the EVCont prediction path reads."""

    def __init__(self, ovlp, hcore, eri, ipovlp, hcore_deriv, eri_ip1, aoslices,
                 e_nuc=0.0, grad_nuc=None, nelec=None, atom_coords=None):
        self._ovlp = np.ascontiguousarray(ovlp, dtype=np.float64)
        self._hcore = np.ascontiguousarray(hcore, dtype=np.float64)
        self._eri = np.ascontiguousarray(eri, dtype=np.float64)
        self._ipovlp = np.ascontiguousarray(ipovlp, dtype=np.float64)
        #: (natm, 3, n, n): what ``hcore_generator()(atom)`` returns per atom
        self._hcore_deriv = np.ascontiguousarray(hcore_deriv, dtype=np.float64)
        self._eri_ip1 = np.ascontiguousarray(eri_ip1, dtype=np.float64)
        self._aoslices = np.asarray(aoslices, dtype=np.int64).reshape(-1, 4)
        self.nao = self._ovlp.shape[0]
        self.natm = self._aoslices.shape[0]
        self._e_nuc = float(e_nuc)
        self._grad_nuc = (np.zeros((self.natm, 3)) if grad_nuc is None
                          else np.ascontiguousarray(grad_nuc, dtype=np.float64))
        self.nelec = nelec
        self._coords = atom_coords
        n, natm = self.nao, self.natm
        assert self._hcore.shape == (n, n) and self._eri.shape == (n, n, n, n)
        assert self._ipovlp.shape == (3, n, n) and self._eri_ip1.shape == (3, n, n, n, n)
        assert self._hcore_deriv.shape == (natm, 3, n, n)
        assert self._grad_nuc.shape == (natm, 3)

    # --- the gto.Mole surface -------------------------------------------------
    def intor(self, name, comp=None):
        if name == "int1e_ovlp":
            return self._ovlp
        if name == "int1e_ipovlp":
            return self._ipovlp
        if name == "int2e":
            return self._eri
        if name == "int2e_ip1":
            return self._eri_ip1
        raise KeyError(f"ArrayMol carries only {_INTOR_NAMES}, not {name!r}")

    def aoslice_by_atom(self):
        return self._aoslices

    def energy_nuc(self):
        return self._e_nuc

    def atom_coords(self):
        return self._coords

    # --- PySCF free functions the reference calls, as methods ----------------
    def get_hcore(self):
        return self._hcore

    def hcore_generator(self):
        return lambda atm_id: self._hcore_deriv[atm_id]

    def grad_nuc(self):
        return self._grad_nuc


def synthetic_mol(norb, natm, seed=0, nelec=None):
    """Seeded synthetic AO arrays with the symmetries of real integrals.

    The recipe of SURVEY.md section 8(d): ``S = I + (0.3/n)(A + A^T)``, random
    symmetric ``hcore``, 8-fold symmetric ``eri``, ``int2e_ip1`` symmetric in
    its last index pair, random ``<nabla mu|nu>`` and core-Hamiltonian
    derivatives (symmetric per (atom, xyz) as ``hcore_generator`` returns).
    AOs are dealt to atoms as evenly as possible, in order.
    """
    rng = np.random.default_rng(seed)
    n = norb
    a = rng.standard_normal((n, n))
    ovlp = np.eye(n) + (0.3 / n) * (a + a.T)
    h = rng.standard_normal((n, n))
    hcore = 0.5 * (h + h.T)
    e = rng.standard_normal((n, n, n, n)) / n
    e = e + e.transpose(1, 0, 2, 3)
    e = e + e.transpose(0, 1, 3, 2)
    eri = e + e.transpose(2, 3, 0, 1)
    ipovlp = 0.1 * rng.standard_normal((3, n, n))
    hd = 0.1 * rng.standard_normal((natm, 3, n, n))
    hcore_deriv = hd + hd.transpose(0, 1, 3, 2)
    ip1 = 0.1 * rng.standard_normal((3, n, n, n, n)) / n
    eri_ip1 = ip1 + ip1.transpose(0, 1, 2, 4, 3)
    bounds = np.linspace(0, n, natm + 1).round().astype(int)
    aoslices = [(0, 0, int(bounds[k]), int(bounds[k + 1])) for k in range(natm)]
    grad_nuc = 0.1 * rng.standard_normal((natm, 3))
    return ArrayMol(ovlp, hcore, eri, ipovlp, hcore_deriv, eri_ip1, aoslices,
                    e_nuc=float(rng.standard_normal()), grad_nuc=grad_nuc,
                    nelec=nelec)


def ao_bundle(mol):
    """Collect the AO arrays the device prediction step consumes from ``mol``.

    Accepts an :class:`ArrayMol` (or any object exposing ``get_hcore`` /
    ``hcore_generator`` / ``grad_nuc`` as methods); a real ``pyscf.gto.Mole`` is
    routed through PySCF's own free functions when PySCF is importable.
    """
    if hasattr(mol, "get_hcore") and hasattr(mol, "hcore_generator") and not _is_pyscf_mole(mol):
        hcore = mol.get_hcore()
        gen = mol.hcore_generator()
        grad_nuc = mol.grad_nuc()
    else:
        try:
            from pyscf import scf, grad  # noqa: WPS433 (optional dependency)
        except ImportError as exc:  # pragma: no cover - PySCF absent in this image
            raise TypeError(
                "mol must be an evcont_b200.mol.ArrayMol/MolLite, or a pyscf Mole "
                "with PySCF importable") from exc
        hcore = scf.hf.get_hcore(mol)
        g = grad.RHF(scf.RHF(mol))
        gen = g.hcore_generator()
        grad_nuc = g.grad_nuc()
    natm = mol.natm
    return dict(
        nao=int(mol.nao),
        natm=int(natm),
        ovlp=np.ascontiguousarray(mol.intor("int1e_ovlp"), dtype=np.float64),
        hcore=np.ascontiguousarray(hcore, dtype=np.float64),
        eri=np.ascontiguousarray(mol.intor("int2e"), dtype=np.float64).reshape(
            (int(mol.nao),) * 4),
        ipovlp=np.ascontiguousarray(mol.intor("int1e_ipovlp", comp=3), dtype=np.float64),
        hcore_deriv=np.ascontiguousarray([gen(a) for a in range(natm)], dtype=np.float64),
        eri_ip1=np.ascontiguousarray(mol.intor("int2e_ip1", comp=3), dtype=np.float64),
        aoslices=np.asarray([(s[2], s[3]) for s in mol.aoslice_by_atom()], dtype=np.int32),
        e_nuc=float(mol.energy_nuc()),
        grad_nuc=np.ascontiguousarray(grad_nuc, dtype=np.float64),
    )


def _is_pyscf_mole(mol):
    return type(mol).__module__.startswith("pyscf.")
