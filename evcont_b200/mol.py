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


BOHR = 0.52917721092  # Angstrom per bohr (the value pyscf.data.nist.BOHR carries)


def _parse_atoms(atom):
    if isinstance(atom, str):
        out = []
        for line in atom.replace(";", "\n").splitlines():
            f = line.replace(",", " ").split()
            if f:
                out.append((f[0], tuple(float(v) for v in f[1:4])))
        return out
    return [(a[0], tuple(float(v) for v in (a[1] if len(a) == 2 else a[1:4]))) for a in atom]


class MolLite:
    """A ``pyscf.gto.Mole`` stand-in whose integrals come from the device engine (K9).

    Covers what the reference touches on the prediction path ("mol duck
    type"): ``nao, natm, nelec, atom_coords(), set_geom_(), copy(), energy_nuc(),
    intor(name, comp=), aoslice_by_atom()`` and, as methods, the PySCF free functions
    ``get_hcore / hcore_generator / grad_nuc``.  s and p shells (``evcont_b200.basis``: H, He, O).
    The reference scripts build their molecules as
    ``gto.Mole().build(atom=[("H", (x, 0, 0)), ...], basis="sto-6g", unit="Bohr")``
    (examples/H10_continuation_3D_replacements.py:84-102); ``MolLite(atom, basis, unit)``
    takes the same arguments.
    """

    def __init__(self, atom, basis="sto-6g", unit="Bohr", charge=0, spin=0):
        atoms = _parse_atoms(atom)
        self._symbols = [a[0].capitalize() for a in atoms]
        scale = 1.0 if unit.lower().startswith(("b", "au")) else 1.0 / BOHR
        self._coords = np.array([a[1] for a in atoms], dtype=np.float64).reshape(-1, 3) * scale
        self.basis, self.charge, self.spin, self.unit = basis, charge, spin, unit
        from .basis import CHARGES, has_p_shells, s_basis_tables, sp_basis_tables
        self._tables = (sp_basis_tables if has_p_shells(self._symbols, basis) else s_basis_tables)(
            self._symbols, basis)
        self.natm = len(atoms)
        self.nao = int(len(self._tables["ao_atom"]))
        nel = int(sum(CHARGES[s] for s in self._symbols)) - int(charge)
        self.nelectron = nel
        self.nelec = ((nel + spin) // 2, (nel - spin) // 2)
        self._cache = None

    # --- geometry --------------------------------------------------------------
    @classmethod
    def from_mol(cls, mol):
        """``mol`` itself if it is a :class:`MolLite`, else a :class:`MolLite` with the geometry, basis NAME, charge and
        spin of a PySCF-like molecule (``natm``, ``atom_symbol(i)``, ``atom_coords()`` in bohr, ``basis``): what
        ``get_trajectory`` (evcont/MD_utils.py:60-125) is handed by the reference's scripts.  A basis given as a dict
        of explicit shells cannot be mapped onto the device basis tables and raises ``TypeError``."""
        if isinstance(mol, cls):
            return mol
        try:
            natm = int(mol.natm)
            symbols = [str(mol.atom_symbol(i)) for i in range(natm)]
            coords = np.asarray(mol.atom_coords(), dtype=np.float64).reshape(natm, 3)
            basis = mol.basis
        except AttributeError as exc:
            raise TypeError("expected an evcont_b200.mol.MolLite or a PySCF-like molecule with natm, atom_symbol(), "
                            f"atom_coords() and basis ({exc})") from None
        if not isinstance(basis, str):
            raise TypeError("the device integral kernels need a NAMED basis (sto-6g, 6-31g); got " +
                            type(basis).__name__)
        return cls([(sym, tuple(map(float, c))) for sym, c in zip(symbols, coords)], basis=basis, unit="Bohr",
                   charge=int(getattr(mol, "charge", 0)), spin=int(getattr(mol, "spin", 0)))

    def atom_coords(self, unit="Bohr"):
        c = self._coords.copy()
        return c if unit.lower().startswith(("b", "au")) else c * BOHR

    def atom_symbol(self, i):
        return self._symbols[i]

    def atom_charges(self):
        return self._tables["charges"].astype(int)

    def atom_mass_list(self, isotope_avg=False):
        """Masses in atomic mass units (most common isotope, PySCF's default)."""
        if isotope_avg:
            raise NotImplementedError("isotope-averaged masses are not tabulated")
        from .md import COMMON_ISOTOPE_MASSES
        return np.array([COMMON_ISOTOPE_MASSES[s] for s in self._symbols])

    def with_common_orig(self, origin):
        """Context manager: origin of the ``int1e_r`` operator (``pyscf.gto.Mole.with_common_orig``)."""
        mol = self

        class _Ctx:
            def __enter__(self):
                self.old = getattr(mol, "_common_orig", None)
                mol._common_orig = np.asarray(origin, dtype=np.float64).reshape(3)
                return mol

            def __exit__(self, *exc):
                mol._common_orig = self.old
                return False

        return _Ctx()

    def intor_symmetric(self, name, comp=None):
        return self.intor(name, comp=comp)

    def set_geom_(self, coords, unit="Bohr", inplace=True):
        mol = self if inplace else self.copy()
        scale = 1.0 if unit.lower().startswith(("b", "au")) else 1.0 / BOHR
        mol._coords = np.array(coords, dtype=np.float64).reshape(self.natm, 3) * scale
        mol._cache = None
        return mol

    def copy(self):
        other = object.__new__(MolLite)
        other.__dict__.update(self.__dict__)
        other._coords = self._coords.copy()
        other._cache = None
        return other

    def aoslice_by_atom(self):
        t = self._tables["ao_atom"]
        out = np.zeros((self.natm, 4), dtype=np.int64)
        for A in range(self.natm):
            idx = np.flatnonzero(t == A)
            if len(idx):
                out[A, 2:] = (idx[0], idx[-1] + 1)
        return out

    # --- integrals (device, cached per geometry) -----------------------------------
    def sbasis(self, engine=None):
        from .engine import get_engine
        return (engine or get_engine()).sbasis(self._symbols, self.basis)

    def _arrays(self):
        if self._cache is None:
            from .engine import DeviceAO, get_engine
            eng = get_engine()
            ao = eng.ao_integrals(self.sbasis(eng), self._coords[None])
            self._cache = {k: getattr(ao, k)[0].cpu().numpy() for k in DeviceAO.FIELDS}
        return self._cache

    def intor(self, name, comp=None):
        if name == "int1e_r":
            from .engine import get_engine
            eng = get_engine()
            origin = getattr(self, "_common_orig", None)
            origin = np.zeros(3) if origin is None else origin
            return eng.int1e_r(eng.aotable(self._symbols, self.basis), self._coords[None], origin)[0].cpu().numpy()
        key = {"int1e_ovlp": "ovlp", "int1e_ipovlp": "ipovlp", "int2e": "eri", "int2e_ip1": "eri_ip1"}.get(name)
        if key is None:
            raise KeyError(f"MolLite serves {_INTOR_NAMES}, not {name!r}")
        return self._arrays()[key]

    def energy_nuc(self):
        return float(self._arrays()["e_nuc"])

    def get_hcore(self):
        return self._arrays()["hcore"]

    def hcore_generator(self):
        hd = self._arrays()["hcore_deriv"]
        return lambda atm_id: hd[atm_id]

    def grad_nuc(self):
        return self._arrays()["grad_nuc"]


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
