"""``cisolver`` drop-in: the object injected as ``FCI_EVCont_obj(cisolver=...)``.

The reference's default is ``pyscf.fci.direct_spin0.FCI()`` (evcont/FCI_EVCont.py:17);
the two methods ``append_to_rdms`` calls on it are ``kernel`` (:70) and
``trans_rdm12`` (:121).  :class:`B200FCISolver` runs ``trans_rdm12`` on the GPU
(K1+K2 of the C ABI) and delegates ``kernel`` to a wrapped PySCF solver when one
is available.
"""
import numpy as np

from .engine import get_engine


def _unpack_nelec(nelec):
    if isinstance(nelec, (int, np.integer)):
        nb = int(nelec) // 2
        return int(nelec) - nb, nb
    return int(nelec[0]), int(nelec[1])


class B200FCISolver:
    """FCI solver facade whose transition RDMs run on the B200."""

    def __init__(self, base_solver=None, device=None):
        self._base = base_solver
        self._device = device

    # -- the part the reference delegates to PySCF's Davidson --------------------
    def kernel(self, h1e, eri, norb, nelec, nroots=1, **kwargs):
        base = self._base
        if base is None:
            try:
                from pyscf import fci
            except ImportError as exc:
                raise NotImplementedError(
                    "B200FCISolver.kernel needs a wrapped FCI eigensolver (pass "
                    "base_solver=pyscf.fci.direct_spin0.FCI()) -- PySCF is not importable. "
                    "Training vectors obtained elsewhere can be added with "
                    "FCI_EVCont_obj.append_civec().") from exc
            base = self._base = fci.direct_spin0.FCI()
        return base.kernel(h1e, eri, norb, nelec, nroots=nroots, **kwargs)

    # -- K1 + K2 -------------------------------------------------------------------
    def trans_rdm12(self, cibra, ciket, norb, nelec, link_index=None, reorder=True):
        """``(dm1, dm2)`` with PySCF's conventions: ``dm1[p,q] = <bra|q^+ p|ket>``,
        ``dm2[p,q,r,s] = <bra|p^+ r^+ s q|ket>`` (spin-summed)."""
        nelec = _unpack_nelec(nelec)
        eng = get_engine(self._device)
        na, _, _, _ = eng.link_tables(norb, nelec[0])
        nb, _, _, _ = eng.link_tables(norb, nelec[1])
        vecs = np.stack([np.asarray(cibra, dtype=np.float64).reshape(na, nb),
                         np.asarray(ciket, dtype=np.float64).reshape(na, nb)])
        _, dm1, dm2 = eng.trans_rdm12_batch(vecs, [(0, 1)], norb, nelec)
        dm1, dm2 = dm1[0].cpu().numpy(), dm2[0].cpu().numpy()
        if not reorder:
            for k in range(norb):
                dm2[:, k, k, :] += dm1.T
        return dm1, dm2

    def trans_rdm12_batch(self, civecs, pairs, norb, nelec):
        """All ``pairs`` [(bra, ket)] among ``civecs`` in one launch; numpy results
        ``(ovlp[np], dm1[np,n,n], dm2[np,n,n,n,n])``."""
        eng = get_engine(self._device)
        ovlp, dm1, dm2 = eng.trans_rdm12_batch(np.asarray(civecs, dtype=np.float64), pairs, norb,
                                               _unpack_nelec(nelec))
        return ovlp.cpu().numpy(), dm1.cpu().numpy(), dm2.cpu().numpy()

    def make_rdm12(self, fcivec, norb, nelec, link_index=None, reorder=True):
        return self.trans_rdm12(fcivec, fcivec, norb, nelec, link_index, reorder)

    def __getattr__(self, name):  # anything else: the wrapped solver, if any
        base = self.__dict__.get("_base")
        if base is None:
            raise AttributeError(name)
        return getattr(base, name)
