"""``cisolver`` drop-in: the object injected as ``FCI_EVCont_obj(cisolver=...)``.

The reference's default is ``pyscf.fci.direct_spin0.FCI()`` (evcont/FCI_EVCont.py:17);
the two methods ``append_to_rdms`` calls on it are ``kernel`` (:70) and
``trans_rdm12`` (:121).  :class:`B200FCISolver` runs ``trans_rdm12`` on the GPU
(K1+K2 of the C ABI); ``kernel`` is a Davidson solver over the device ``H c``
(``csrc/fci.cu``) unless a PySCF solver is wrapped.
"""
import numpy as np

from .engine import get_engine


def davidson(ham, nroots=1, tol=1e-10, max_cycle=200, max_space=12, spin0=False):
    """Davidson-Liu for the lowest ``nroots`` eigenpairs of ``ham`` (:class:`FCIHamiltonian`).

    Vectors live on the device as torch tensors (dot products and axpys are torch calls -- plumbing);
    the small projected eigenproblem is solved on the host.  Diagonal preconditioner, two passes of
    Gram-Schmidt, restart with the current Ritz vectors when the space exceeds
    ``max_space * nblock``.  Converged when the residual norms of the lowest ``nroots`` are below ``tol``.

    For ``nroots > 1`` the iteration carries a block of ``nroots + 2`` Ritz vectors (corrections and restarts keep
    all of them) and starts from ``4 nroots + 4`` P-space states: a symmetry sector that is absent from the search
    space can never enter it in exact arithmetic, and the lowest few P-space states alone do not always cover the
    sectors of the lowest roots (H6 at 1.4 bohr: the second root of the alpha <-> beta symmetric sector has the
    other inversion parity and was skipped or found depending on rounding noise).
    """
    import torch
    na, nb, nd = ham.na, ham.nb, ham.ndet
    hdiag = ham.hdiag()

    def sym(v):
        if spin0 and na == nb:
            m = v.reshape(na, nb)
            return (0.5 * (m + m.T)).reshape(-1)
        return v

    def orth(v, basis):
        for _ in range(2):
            for b in basis:
                v = v - torch.dot(b, v) * b
        nrm = torch.linalg.vector_norm(v)
        return v, float(nrm)

    # Initial space from a small "P space", as pyscf.fci.direct_spin0 starts (pspace_size = 400 there): the P
    # determinants of lowest diagonal energy (closed under alpha <-> beta exchange), H restricted to them
    # built with P applications of the device H c, diagonalised on the host; its lowest eigenvectors seed the
    # iteration.  Unit vectors on the lowest diagonal elements alone let the iteration skip roots that are
    # weakly coupled to them (H6 at 1.4 bohr, nroots = 2: the second root of the symmetric sector was missed).
    ninit = nroots + 2 if nroots == 1 else 4 * nroots + 4
    P = min(nd, 48 + 16 * nroots if nd > 500000 else 96 + 32 * nroots)
    idx = torch.argsort(hdiag)[:P].tolist()
    if spin0 and na == nb:   # add the exchange partners (Ib, Ia) of the chosen (Ia, Ib)
        have = set(idx)
        for k in list(idx):
            t = (k % nb) * nb + (k // nb)
            if t not in have:
                have.add(t)
                idx.append(t)
    P = len(idx)
    idx_t = torch.tensor(idx, device=hdiag.device)
    hpp = np.empty((P, P))
    unit = torch.zeros(nd, dtype=torch.float64, device=hdiag.device)
    for j, k in enumerate(idx):
        unit[k] = 1.0
        hpp[:, j] = ham.contract(unit)[idx_t].cpu().numpy()
        unit[k] = 0.0
    hpp = 0.5 * (hpp + hpp.T)
    _, pv = np.linalg.eigh(hpp)
    V = []
    for k in range(P):
        g = torch.zeros(nd, dtype=torch.float64, device=hdiag.device)
        g[idx_t] = torch.from_numpy(np.ascontiguousarray(pv[:, k])).to(hdiag.device)
        gs = sym(g)
        if spin0 and na == nb and float(torch.linalg.vector_norm(gs)) < 0.5:
            continue   # an exchange-antisymmetric P-space state: not in the sector direct_spin0 solves
        g, nrm = orth(gs, V)
        if nrm > 1e-8:
            V.append(g / nrm)
        if len(V) == ninit:
            break
    if len(V) < nroots:
        raise RuntimeError(f"FCI Davidson: only {len(V)} independent initial guesses for nroots={nroots}")
    W = [ham.contract(v) for v in V]
    nblock = nroots if nroots == 1 else min(nroots + 2, len(V))
    theta, X = None, None
    for _cycle in range(max_cycle):
        m = len(V)
        Vm, Wm = torch.stack(V), torch.stack(W)
        hsub = (Vm @ Wm.T).cpu().numpy()
        hsub = 0.5 * (hsub + hsub.T)
        w, s = np.linalg.eigh(hsub)
        theta = w[:nblock]
        S = torch.from_numpy(np.ascontiguousarray(s[:, :nblock].T)).to(Vm.device)
        X = S @ Vm            # Ritz vectors (nblock, nd)
        HX = S @ Wm
        R = HX - torch.from_numpy(theta).to(Vm.device)[:, None] * X
        rn = torch.linalg.vector_norm(R, dim=1).cpu().numpy()
        if rn[:nroots].max() < tol:
            break
        if m + nblock > max_space * nblock:   # restart from the Ritz vectors
            V, W = [], []
            for k in range(nblock):
                v, nrm = orth(X[k].clone(), V)
                if nrm > 1e-10:
                    V.append(v / nrm)
            W = [ham.contract(v) for v in V]
        added = 0
        for k in range(nblock):
            if rn[k] < tol:
                continue
            den = theta[k] - hdiag
            den = torch.where(den.abs() < 1e-8, torch.full_like(den, -1e-8), den)
            t, nrm = orth(sym(R[k] / den), V)
            if nrm < 1e-10:
                continue
            V.append(t / nrm)
            W.append(ham.contract(V[-1]))
            added += 1
        if added == 0:
            # every correction vector was linearly dependent on the space although a residual is still above
            # tol: collapse onto the Ritz vectors plus the (unpreconditioned) residuals instead of giving up
            V, W = [], []
            for k in range(nblock):
                v, nrm = orth(X[k].clone(), V)
                if nrm > 1e-10:
                    V.append(v / nrm)
            for k in range(nblock):
                if rn[k] >= tol:
                    t, nrm = orth(sym(R[k].clone()), V)
                    if nrm > 1e-12:
                        V.append(t / nrm)
                        added += 1
            if added == 0:
                raise RuntimeError(f"FCI Davidson stalled with residuals {rn} (tol {tol})")
            W = [ham.contract(v) for v in V]
    else:
        raise RuntimeError(f"FCI Davidson did not converge in {max_cycle} cycles (residuals {rn})")
    out = []
    for k in range(nroots):
        v = sym(X[k])
        out.append(v / torch.linalg.vector_norm(v))
    return np.asarray(theta[:nroots], dtype=np.float64), out


def _unpack_nelec(nelec):
    if isinstance(nelec, (int, np.integer)):
        nb = int(nelec) // 2
        return int(nelec) - nb, nb
    return int(nelec[0]), int(nelec[1])


def transform_ci(ci, nelec, u):
    """Drop-in for ``pyscf.fci.addons.transform_ci`` as the reference uses it
    (evcont/FCI_EVCont.py:79-85): the CI vector ``ci`` (na, nb), solved in an orthonormal basis,
    expressed in the basis reached by the rotation ``u[old, new]``.  Runs on the device
    (``evc_transform_ci``: minors of ``u`` per string pair, two DMMA products)."""
    u = np.asarray(u, dtype=np.float64)
    if u.ndim != 2:
        raise NotImplementedError("transform_ci: one rotation for both spins only")
    return get_engine().transform_ci(ci, _unpack_nelec(nelec), u).cpu().numpy()


class B200FCISolver:
    """FCI solver facade whose transition RDMs run on the B200."""

    def __init__(self, base_solver=None, device=None):
        self._base = base_solver
        self._device = device

    # -- the FCI eigensolver (PySCF's Davidson in the reference, evcont/FCI_EVCont.py:70) ----
    def kernel(self, h1e, eri, norb, nelec, nroots=1, tol=1e-10, max_cycle=200, max_space=12,
               **kwargs):
        """Lowest ``nroots`` FCI states: ``(e, c)`` for one root, ``(list e, list c)`` otherwise, ``c``
        of shape ``(na, nb)`` -- the return convention of ``pyscf.fci.direct_spin0.FCI().kernel``.
        A wrapped ``base_solver`` is used when one was given; otherwise a Davidson iteration whose
        ``H c`` and ``diag H`` run on the device (``evc_fci_contract_2e`` / ``evc_fci_hdiag``).  For
        ``n_alpha == n_beta`` the vectors are kept symmetric under alpha <-> beta exchange, as
        ``direct_spin0`` does."""
        if self._base is not None:
            return self._base.kernel(h1e, eri, norb, nelec, nroots=nroots, **kwargs)
        nelec = _unpack_nelec(nelec)
        eng = get_engine(self._device)
        ham = eng.fci_hamiltonian(h1e, eri, norb, nelec)
        e, vecs = davidson(ham, nroots=nroots, tol=tol, max_cycle=max_cycle, max_space=max_space,
                           spin0=(nelec[0] == nelec[1]))
        cs = [v.reshape(ham.na, ham.nb).cpu().numpy() for v in vecs]
        if nroots == 1:
            return float(e[0]), cs[0]
        return [float(x) for x in e], cs

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
        if not hasattr(civecs, "data_ptr"):  # numpy in; a device tensor (FCI_EVCont_obj's resident copy) passes through
            civecs = np.asarray(civecs, dtype=np.float64)
        ovlp, dm1, dm2 = eng.trans_rdm12_batch(civecs, pairs, norb, _unpack_nelec(nelec))
        return ovlp.cpu().numpy(), dm1.cpu().numpy(), dm2.cpu().numpy()

    def device_civecs(self, cache, fcivecs):
        """Keep the training vectors resident in HBM across appends: ``cache`` is the (capacity, ndet) device
        tensor holding ``fcivecs[:-1]`` (or ``None``); only the newest vector is uploaded.  Returns the new cache."""
        import torch
        eng = get_engine(self._device)
        N, ndet = len(fcivecs), fcivecs[-1].size
        ids = [id(v) for v in fcivecs]
        valid = cache is not None and cache.shape[1] == ndet and getattr(cache, "_evc_ids", None) == ids[:-1]
        if not valid or cache.shape[0] < N:
            new = torch.empty(max(8, 2 * N), ndet, dtype=torch.float64, device=eng.device)
            if valid:
                new[:N - 1] = cache[:N - 1]
                first = N - 1
            else:
                first = 0
            for k in range(first, N):
                new[k] = torch.from_numpy(np.ascontiguousarray(fcivecs[k], dtype=np.float64).reshape(-1))
            cache = new
        else:
            cache[N - 1] = torch.from_numpy(np.ascontiguousarray(fcivecs[-1], dtype=np.float64).reshape(-1))
        cache._evc_ids = ids  # the host arrays this copy mirrors (a re-assigned fcivecs list invalidates it)
        return cache

    def make_rdm12(self, fcivec, norb, nelec, link_index=None, reorder=True):
        return self.trans_rdm12(fcivec, fcivec, norb, nelec, link_index, reorder)

    def __getattr__(self, name):  # anything else: the wrapped solver, if any
        base = self.__dict__.get("_base")
        if base is None:
            raise AttributeError(name)
        return getattr(base, name)
