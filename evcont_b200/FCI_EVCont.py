"""``FCI_EVCont_obj``: the FCI training-set builder (evcont/FCI_EVCont.py:10-151).

Same constructor, attributes and methods as the reference.  Differences, all
behind the same surface:

* the N+1 ``trans_rdm12`` calls of one append (reference loop at :117-127) go to
  the GPU as ONE batched launch over the pair list;
* link tables are cached instead of being rebuilt per call;
* :meth:`append_civec` feeds a CI vector directly (synthetic vectors, vectors
  from another solver);
* ``kernel`` (device Davidson), ``transform_ci`` (device) and the canonical basis
  (``evcont_b200.scf.rhf``) are built in, so ``append_to_rdms(MolLite)`` needs no PySCF.

Reference quirks kept on purpose: mirror blocks
``[i, -1]`` hold the *untransposed* RDMs (:125,127); ``mol_index`` is not pruned.
"""
import numpy as np

from .electron_integral_utils import get_basis, get_integrals
from .fci import B200FCISolver


class FCI_EVCont_obj:
    """Holds the data structure for the continuation from FCI states."""

    def __init__(self, cisolver=None, cibasis="canonical", nroots=1, roots_train=None):
        self.cisolver = B200FCISolver() if cisolver is None else cisolver
        self.cibasis = cibasis
        self.nroots = nroots
        if roots_train is None:
            self.roots_train = list(range(nroots))
        else:
            assert isinstance(roots_train, list)
            self.roots_train = roots_train
        self.fcivecs = []
        self.ens = []
        self.mol_index = []
        self.overlap = None
        self.one_rdm = None
        self.two_rdm = None
        self._dev_civecs = None  # the training vectors resident in HBM (only the newest one is uploaded per append)

    # -- reference API -----------------------------------------------------------
    def append_to_rdms(self, mol):
        """Solve FCI at ``mol`` and grow the t-RDM stack (evcont/FCI_EVCont.py:58-131)."""
        basis = get_basis(mol, basis_type=self.cibasis)
        h1, h2 = get_integrals(mol, basis)
        nroots_train = max(self.roots_train) + 1
        e_all, fcivec_all = self.cisolver.kernel(h1, h2, mol.nao, mol.nelec, nroots=nroots_train)
        if nroots_train == 1:
            e_all, fcivec_all = [e_all], [fcivec_all]
        if self.cibasis != "OAO":
            from .fci import transform_ci  # device minors + DMMA products
            S = mol.intor("int1e_ovlp")
            u = np.einsum("ji,jk,kl->il", basis, S, get_basis(mol))
            fcivec_all = [transform_ci(c, mol.nelec, u) for c in fcivec_all]
        mindex = 0 if len(self.mol_index) == 0 else max(self.mol_index) + 1
        for ind in range(len(e_all)):
            if ind in self.roots_train:
                self.append_civec(fcivec_all[ind], e_all[ind] + mol.energy_nuc(), mol.nao,
                                  mol.nelec, mol_index=mindex)

    def prune_datapoints(self, keep_ids):
        """Keep only ``keep_ids`` (evcont/FCI_EVCont.py:133-151)."""
        if self.overlap is not None:
            self.overlap = self.overlap[np.ix_(keep_ids, keep_ids)]
        if self.one_rdm is not None:
            self.one_rdm = self.one_rdm[np.ix_(keep_ids, keep_ids)]
        if self.two_rdm is not None:
            self.two_rdm = self.two_rdm[np.ix_(keep_ids, keep_ids)]
        self.fcivecs = [self.fcivecs[i] for i in keep_ids]
        self.ens = [self.ens[i] for i in keep_ids]
        self._dev_civecs = None

    # -- extension ---------------------------------------------------------------
    def append_civec(self, fcivec, energy=None, norb=None, nelec=None, mol_index=None):
        """Append one training vector given directly in the OAO basis.

        Grows ``overlap / one_rdm / two_rdm`` exactly like the inner loop of
        ``append_to_rdms`` (:95-131): new row ``[-1, i]`` from
        ``trans_rdm12(new, i)`` and the same (untransposed) block at ``[i, -1]``.
        """
        fcivec = np.asarray(fcivec, dtype=np.float64)
        if norb is None or nelec is None:
            if not self.fcivecs:
                raise ValueError("norb and nelec are required for the first vector")
            norb, nelec = self._norb, self._nelec
        self._norb, self._nelec = int(norb), nelec
        self.fcivecs.append(fcivec)
        self.ens.append(energy)
        if mol_index is None:
            mol_index = 0 if len(self.mol_index) == 0 else max(self.mol_index) + 1
        self.mol_index.append(mol_index)
        N = len(self.fcivecs)
        n = int(norb)
        pairs = [(N - 1, i) for i in range(N)]
        solver = self.cisolver
        if hasattr(solver, "device_civecs"):
            self._dev_civecs = solver.device_civecs(self._dev_civecs, self.fcivecs)
            ovlp, dm1, dm2 = solver.trans_rdm12_batch(self._dev_civecs[:N], pairs, n, nelec)
        elif hasattr(solver, "trans_rdm12_batch"):
            ovlp, dm1, dm2 = solver.trans_rdm12_batch(np.stack(self.fcivecs), pairs, n, nelec)
        else:  # a foreign cisolver: the reference's pair-by-pair loop
            ovlp = np.array([fcivec.ravel().dot(self.fcivecs[i].ravel()) for i in range(N)])
            res = [solver.trans_rdm12(fcivec, self.fcivecs[i], n, nelec) for i in range(N)]
            dm1 = np.array([r[0] for r in res])
            dm2 = np.array([r[1] for r in res])
        overlap_new = np.ones((N, N))
        one_new = np.ones((N, N, n, n))
        two_new = np.ones((N, N, n, n, n, n))
        if self.overlap is not None:
            overlap_new[:-1, :-1] = self.overlap
            one_new[:-1, :-1] = self.one_rdm
            two_new[:-1, :-1] = self.two_rdm
        overlap_new[-1, :] = ovlp
        overlap_new[:, -1] = ovlp
        one_new[-1, :] = dm1
        one_new[:, -1] = dm1
        two_new[-1, :] = dm2
        two_new[:, -1] = dm2
        self.overlap, self.one_rdm, self.two_rdm = overlap_new, one_new, two_new

    @classmethod
    def from_civecs(cls, civecs, norb, nelec, energies=None, **kwargs):
        """Build the whole stack from ``civecs`` with ONE launch over all pairs a >= b."""
        self = cls(**kwargs)
        civecs = [np.asarray(c, dtype=np.float64) for c in civecs]
        N, n = len(civecs), int(norb)
        pairs = [(a, b) for a in range(N) for b in range(a + 1)]
        ovlp, dm1, dm2 = self.cisolver.trans_rdm12_batch(np.stack(civecs), pairs, n, nelec)
        self.overlap = np.empty((N, N))
        self.one_rdm = np.empty((N, N, n, n))
        self.two_rdm = np.empty((N, N, n, n, n, n))
        for k, (a, b) in enumerate(pairs):
            self.overlap[a, b] = self.overlap[b, a] = ovlp[k]
            self.one_rdm[a, b] = self.one_rdm[b, a] = dm1[k]
            self.two_rdm[a, b] = self.two_rdm[b, a] = dm2[k]
        self.fcivecs = civecs
        self.ens = list(energies) if energies is not None else [None] * N
        self.mol_index = list(range(N))
        self._norb, self._nelec = n, nelec
        return self
