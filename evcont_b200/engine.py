"""Device engine: thin object layer over the C ABI.

PyTorch is used only for what the C ABI leaves to the caller: device buffers
(``torch.empty(..., device=...)`` + ``data_ptr()``), the current CUDA stream and
(in ``evcont_b200.distributed``) the NCCL process group.  All arithmetic runs in
``libevcont_b200.so``.
"""
import contextlib
import ctypes as C
import threading

import numpy as np
import torch

from . import _lib
from ._lib import AoBundle, check

LAYOUT_FULL, LAYOUT_TRIL, LAYOUT_FULL_EXCH, LAYOUT_TRIL_EXCH = 6, 5, 3, 2

#: use the 8-fold-symmetric packed prediction step unless the full RDMs are wanted
USE_PACKED = True

_engines = {}
_engines_lock = threading.Lock()


def get_engine(device=None):
    """The per-device :class:`Engine` singleton (``device``: int, str or torch.device)."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            "evcont_b200 needs a CUDA device (B200, sm_100a): torch.cuda.is_available() is "
            "False and there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise ValueError(f"evcont_b200 runs on CUDA devices only, not {dev}")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    with _engines_lock:
        if idx not in _engines:
            _engines[idx] = Engine(idx)
        return _engines[idx]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Engine:
    """One ``evc_ctx`` bound to one GPU, plus caches (link tables, workspace)."""

    def __init__(self, device_index):
        self.lib = _lib.lib()
        self.device = torch.device("cuda", device_index)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.evc_ctx_create(device_index, None, C.byref(handle)))
        self._ctx = handle
        self.sm_count = self.lib.evc_ctx_sm_count(handle)
        self._links = {}
        self._sbases = {}
        self._aotables = {}
        self._ws = {}            # shared scratch, one Workspace per CUDA stream
        self._ws_private = None  # a caller-owned Workspace (using_workspace)

    # -- plumbing ------------------------------------------------------------
    def _bind_stream(self):
        """Make this engine's device current (the ``device=`` arguments of the public API work
        without a prior ``torch.cuda.set_device``) and launch on torch's current stream of it."""
        if torch.cuda.current_device() != self.device.index:
            torch.cuda.set_device(self.device)
        stream = torch.cuda.current_stream(self.device)
        check(self.lib.evc_ctx_set_stream(self._ctx, C.c_void_p(stream.cuda_stream)))

    def workspace(self, nbytes):
        """Scratch buffer of at least ``nbytes``: the caller-owned :class:`Workspace` inside a
        :meth:`using_workspace` block, else this engine's shared one for the current stream
        (grown geometrically; work queued on one stream is ordered, so sharing per stream is safe)."""
        if self._ws_private is not None:
            return self._ws_private.get(nbytes)
        key = torch.cuda.current_stream(self.device).cuda_stream
        if key not in self._ws:
            self._ws[key] = Workspace(self.device)
        return self._ws[key].get(nbytes)

    @contextlib.contextmanager
    def using_workspace(self, ws):
        """Route every scratch request of the enclosed engine calls to ``ws`` (a :class:`Workspace`
        the caller keeps alive -- e.g. the owner of a captured CUDA graph, whose replays use the
        pointer that was current at capture time)."""
        prev, self._ws_private = self._ws_private, ws
        try:
            yield ws
        finally:
            self._ws_private = prev

    STAGES = ("loewdin", "ao2oao", "subspace_H", "geneig", "predict_rdm", "grad", "grad_stream")

    def launch_count(self):
        """Kernel launches issued by the library so far (process-wide)."""
        return int(self.lib.evc_launch_count())

    def stage_timing(self, enable=True):
        """Switch per-stage CUDA-event timing of :meth:`energy_with_grad` on/off."""
        check(self.lib.evc_ctx_stage_timing(self._ctx, 1 if enable else 0))

    def stage_times(self):
        """``({stage: accumulated ms}, calls)`` since :meth:`stage_timing` was enabled."""
        ms = (C.c_double * len(self.STAGES))()
        calls = C.c_int64()
        check(self.lib.evc_ctx_stage_times(self._ctx, ms, C.byref(calls)))
        return dict(zip(self.STAGES, list(ms))), int(calls.value)

    def empty(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def to_device(self, a, dtype=torch.float64):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a)).to(device=self.device, dtype=dtype)

    # -- K0 --------------------------------------------------------------------
    def link_tables(self, norb, nocc):
        """(nstr, nlink, string-major packed table, link-major packed table) on device."""
        key = (norb, nocc)
        if key not in self._links:
            from . import cistring
            tab = cistring.gen_linkstr_index(range(norb), nocc)
            nstr, nlink = tab.shape[0], tab.shape[1]
            # at least one record so that empty tables (nocc == 0) still have an address
            sm = np.zeros(max(1, nstr * nlink), dtype=np.uint64)
            lm = np.zeros(max(1, nstr * nlink), dtype=np.uint64)
            if nlink:
                check(self.lib.evc_linkindex_pack_host(nstr, nlink, tab.ctypes.data, 0, sm.ctypes.data))
                check(self.lib.evc_linkindex_pack_host(nstr, nlink, tab.ctypes.data, 1, lm.ctypes.data))
            self._links[key] = (nstr, nlink,
                                torch.from_numpy(sm.view(np.int64)).to(self.device),
                                torch.from_numpy(lm.view(np.int64)).to(self.device))
        return self._links[key]

    # -- K1 + K2 -----------------------------------------------------------------
    def trans_rdm12_batch(self, civecs, pairs, norb, nelec):
        """Transition RDMs for ``pairs`` [(bra, ket), ...] among ``civecs`` (nvec, na, nb).

        Returns device tensors ``(ovlp[np], dm1[np,n,n], dm2[np,n,n,n,n])``.
        """
        neleca, nelecb = nelec
        na, nlink_a, la_sm, _ = self.link_tables(norb, neleca)
        nb, nlink_b, _, lb_lm = self.link_tables(norb, nelecb)
        civecs = self.to_device(civecs).reshape(-1, na * nb)
        nvec = civecs.shape[0]
        stride = na * nb + ((na * nb) & 1)
        if stride != na * nb:  # keep every vector 16-byte aligned
            padded = self.empty(nvec, stride)
            padded[:, : na * nb] = civecs
            padded[:, na * nb:] = 0
            civecs = padded
        pairs_h = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        if pairs_h.min() < 0 or pairs_h.max() >= nvec:
            raise IndexError("pair index out of range")
        npairs = pairs_h.shape[0]
        pairs_d = torch.from_numpy(pairs_h).to(self.device)
        n = norb
        ovlp = self.empty(npairs)
        dm1 = self.empty(npairs, n, n)
        dm2 = self.empty(npairs, n, n, n, n)
        nbytes = C.c_size_t()
        check(self.lib.evc_trans_rdm12_workspace_bytes(n, na, nb, npairs, self.sm_count,
                                                       C.byref(nbytes)))
        ws = self.workspace(nbytes.value)
        self._bind_stream()
        check(self.lib.evc_trans_rdm12_batch(
            self._ctx, n, na, nb, _ptr(civecs), stride, nvec, _ptr(pairs_d), npairs,
            _ptr(la_sm), nlink_a, _ptr(lb_lm), nlink_b, _ptr(ovlp), _ptr(dm1), _ptr(dm2),
            _ptr(ws), ws.numel()))
        return ovlp, dm1, dm2

    def trans_rdm12_rows(self, civecs, pairs, norb, nelec, plan_pairs=0):
        """Like :meth:`trans_rdm12_batch`, but every pair's results form one contiguous row
        ``[dm2 (n^4) | dm1 (n^2) | ovlp]`` of a ``(npairs, evc_stack_row_len(n))`` tensor -- the send buffer of the
        single all_gather of the multi-GPU stack build (``evc_trans_rdm12_batch_strided``).  ``plan_pairs``: pair
        count of the whole build when ``pairs`` is one rank's share of it (``evc_trans_rdm12_plan_pairs``)."""
        neleca, nelecb = nelec
        na, nlink_a, la_sm, _ = self.link_tables(norb, neleca)
        nb, nlink_b, _, lb_lm = self.link_tables(norb, nelecb)
        civecs = self.to_device(civecs).reshape(-1, na * nb)
        nvec = civecs.shape[0]
        stride = na * nb + ((na * nb) & 1)
        if stride != na * nb:
            padded = self.empty(nvec, stride)
            padded[:, : na * nb] = civecs
            padded[:, na * nb:] = 0
            civecs = padded
        pairs_h = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        if pairs_h.min() < 0 or pairs_h.max() >= nvec:
            raise IndexError("pair index out of range")
        npairs, n = pairs_h.shape[0], norb
        pairs_d = torch.from_numpy(pairs_h).to(self.device)
        width = int(self.lib.evc_stack_row_len(n))
        rows = self.empty(npairs, width)
        n2 = n * n
        nbytes = C.c_size_t()
        check(self.lib.evc_trans_rdm12_workspace_bytes(n, na, nb, npairs, self.sm_count, C.byref(nbytes)))
        ws = self.workspace(nbytes.value)
        self._bind_stream()
        base = rows.data_ptr()
        check(self.lib.evc_trans_rdm12_plan_pairs(self._ctx, int(plan_pairs)))
        check(self.lib.evc_trans_rdm12_batch_strided(
            self._ctx, n, na, nb, _ptr(civecs), stride, nvec, _ptr(pairs_d), npairs,
            _ptr(la_sm), nlink_a, _ptr(lb_lm), nlink_b,
            C.c_void_p(base + 8 * (n2 * n2 + n2)), width, C.c_void_p(base + 8 * n2 * n2), width,
            C.c_void_p(base), width, _ptr(ws), ws.numel()))
        check(self.lib.evc_trans_rdm12_plan_pairs(self._ctx, 0))
        return rows

    def stack_scatter_rows(self, rows, row_pairs, ntrain, norb):
        """Gathered rows -> ``(overlap, one_rdm, two_rdm)`` in the reference's (N, N, ...) layout
        (``evc_stack_scatter_rows``); ``row_pairs``: (nrows, 2) int32 device tensor of (a, b)."""
        N, n = int(ntrain), int(norb)
        overlap, one, two = self.empty(N, N), self.empty(N, N, n, n), self.empty(N, N, n, n, n, n)
        self._bind_stream()
        check(self.lib.evc_stack_scatter_rows(self._ctx, N, n, _ptr(rows), rows.stride(0), rows.shape[0],
                                              _ptr(row_pairs), _ptr(overlap), _ptr(one), _ptr(two)))
        return overlap, one, two

    def trans_rdm12_issued_flops(self):
        return self.lib.evc_trans_rdm12_last_issued_flops(self._ctx)

    # -- FCI Hamiltonian action (training side) --------------------------------------------
    def fci_hamiltonian(self, h1e, eri, norb, nelec):
        """Device operator for ``H c`` and ``diag H`` of one geometry: :class:`FCIHamiltonian`."""
        return FCIHamiltonian(self, h1e, eri, norb, nelec)

    def transform_ci(self, ci, nelec, u):
        """``pyscf.fci.addons.transform_ci``: ``ci`` (na, nb) in the old orbitals -> device tensor
        (na, nb) in the new ones, ``u[old, new]`` (``evc_transform_ci``)."""
        from . import cistring
        u = self.to_device(np.asarray(u, dtype=np.float64))
        n = u.shape[0]
        if u.shape != (n, n):
            raise NotImplementedError("transform_ci: only square rotations (dimension conserved)")
        nea, neb = int(nelec[0]), int(nelec[1])
        key = ("strs", n, nea, neb)
        if key not in self._links:
            self._links[key] = tuple(torch.from_numpy(cistring.make_strings(range(n), k)).to(self.device)
                                     for k in (nea, neb))
        sa, sb = self._links[key]
        na, nb = sa.numel(), sb.numel()
        ci = self.to_device(ci).reshape(na, nb)
        out = self.empty(na, nb)
        nbytes = C.c_size_t()
        check(self.lib.evc_transform_ci_workspace_bytes(n, na, nb, C.byref(nbytes)))
        ws = self.workspace(nbytes.value)
        self._bind_stream()
        check(self.lib.evc_transform_ci(self._ctx, n, nea, neb, na, nb, _ptr(sa), _ptr(sb), _ptr(u), _ptr(ci),
                                        _ptr(out), _ptr(ws), ws.numel()))
        return out

    # -- K3 ------------------------------------------------------------------------
    def loewdin(self, s_ao):
        """Batched Loewdin: ``s_ao`` (G, n, n) device -> (X, evals, evecs)."""
        G, n = s_ao.shape[0], s_ao.shape[-1]
        x, evals, evecs = self.empty(G, n, n), self.empty(G, n), self.empty(G, n, n)
        self._bind_stream()
        check(self.lib.evc_loewdin(self._ctx, G, n, _ptr(s_ao), _ptr(x), _ptr(evals), _ptr(evecs)))
        return x, evals, evecs

    def loewdin_grad(self, evals, evecs, dS):
        """``dS`` (G, nder, n, n) -> dX (G, nder, n, n)."""
        G, nder, n = dS.shape[0], dS.shape[1], dS.shape[-1]
        dX = self.empty(G, nder, n, n)
        self._bind_stream()
        check(self.lib.evc_loewdin_grad(self._ctx, G, n, nder, _ptr(evals), _ptr(evecs), _ptr(dS),
                                        _ptr(dX)))
        return dX

    # -- K4 --------------------------------------------------------------------------
    def ao2oao(self, hcore, eri, c, transpose_c=False, want_t3=False):
        """``h1 = C^T h C`` and the four-index transform, batched (G, ...)."""
        G, n = c.shape[0], c.shape[-1]
        h1 = self.empty(G, n, n) if hcore is not None else None
        h2 = self.empty(G, n, n, n, n) if eri is not None else None
        t3 = self.empty(G, n, n, n, n) if (want_t3 and eri is not None) else None
        ws = self.workspace(2 * (G * n ** 4 * 8 + 256))
        self._bind_stream()
        check(self.lib.evc_ao2oao(self._ctx, G, n, _ptr(hcore), _ptr(eri), _ptr(c),
                                  1 if transpose_c else 0, _ptr(h1), _ptr(h2), _ptr(t3),
                                  _ptr(ws), ws.numel()))
        return h1, h2, t3

    # -- K5 ---------------------------------------------------------------------------
    def subspace_H(self, stack, h1, h2):
        G = h1.shape[0]
        H = self.empty(G, stack.ntrain, stack.ntrain)
        nbytes = C.c_size_t()
        check(self.lib.evc_subspace_workspace_bytes(stack.layout, stack.ntrain, stack.norb, G,
                                                    C.byref(nbytes)))
        ws = self.workspace(nbytes.value)
        self._bind_stream()
        check(self.lib.evc_subspace_H(self._ctx, stack.layout, stack.ntrain, stack.norb,
                                      _ptr(stack.one_rdm), _ptr(stack.two_rdm), G, _ptr(h1),
                                      _ptr(h2), _ptr(H), _ptr(ws), ws.numel()))
        return H

    # -- K6 ----------------------------------------------------------------------------
    def geneig_prepare(self, S):
        N = S.shape[0]
        linv = self.empty(N, N)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._bind_stream()
        check(self.lib.evc_geneig_prepare(self._ctx, N, _ptr(S), _ptr(linv), _ptr(info)))
        k = int(info.item())
        if k != 0:
            raise np.linalg.LinAlgError(
                f"The leading minor of order {k} of the subspace overlap S is not positive "
                "definite; the generalized eigenproblem cannot be solved (as scipy.linalg.eigh)")
        return linv

    def geneig(self, H, linv, nroots=1):
        G, N = H.shape[0], H.shape[-1]
        E, Cv = self.empty(G, nroots), self.empty(G, nroots, N)
        self._bind_stream()
        check(self.lib.evc_geneig(self._ctx, G, N, _ptr(H), _ptr(linv), nroots, _ptr(E), _ptr(Cv)))
        return E, Cv

    # -- K7 ------------------------------------------------------------------------------
    def predict_rdm(self, stack, cvec):
        """``cvec`` (G, N) -> gamma (G, n, n), Gamma (G, n, n, n, n)."""
        G, n = cvec.shape[0], stack.norb
        gamma, Gamma = self.empty(G, n, n), self.empty(G, n, n, n, n)
        nbytes = C.c_size_t()
        check(self.lib.evc_predict_workspace_bytes(stack.layout, stack.ntrain, n, G, C.byref(nbytes)))
        ws = self.workspace(nbytes.value)
        self._bind_stream()
        check(self.lib.evc_predict_rdm(self._ctx, stack.layout, stack.ntrain, n, _ptr(stack.one_rdm),
                                       _ptr(stack.two_rdm), G, _ptr(cvec), cvec.stride(0),
                                       _ptr(gamma), _ptr(Gamma), _ptr(ws), ws.numel()))
        return gamma, Gamma

    # -- K8 --------------------------------------------------------------------------------
    def grad_elec(self, aoslices, evals, evecs, x, hcore, t3, gamma, Gamma, ipovlp, hcore_deriv,
                  eri_ip1):
        G, n, natm = x.shape[0], x.shape[-1], aoslices.shape[0]
        grad = self.empty(G, natm, 3)
        nbytes = C.c_size_t()
        check(self.lib.evc_grad_workspace_bytes(n, natm, G, C.byref(nbytes)))
        ws = self.workspace(nbytes.value)
        self._bind_stream()
        check(self.lib.evc_grad_elec(self._ctx, G, n, natm, _ptr(aoslices), _ptr(evals), _ptr(evecs),
                                     _ptr(x), _ptr(hcore), _ptr(t3), _ptr(gamma), _ptr(Gamma),
                                     _ptr(ipovlp), _ptr(hcore_deriv), _ptr(eri_ip1), _ptr(grad),
                                     _ptr(ws), ws.numel()))
        return grad

    # -- K9 --------------------------------------------------------------------------------
    def sbasis(self, symbols, basis):
        """Device handle (:class:`SBasis`) of an s-shell basis for the atoms ``symbols``."""
        key = (tuple(s.capitalize() for s in symbols), basis.lower())
        if key not in self._sbases:
            self._sbases[key] = SBasis(self, symbols, basis)
        return self._sbases[key]

    # -- f4: observables of the predicted one-body density matrix ---------------------------
    def aotable(self, symbols, basis, masses=None):
        """Device AO table (:class:`AOTable`) of the atoms ``symbols``: basis functions + nuclear charges + masses
        (atomic mass units; default: most common isotope, ``mol.atom_mass_list()``)."""
        from .md import COMMON_ISOTOPE_MASSES
        syms = tuple(s.capitalize() for s in symbols)
        m = tuple(COMMON_ISOTOPE_MASSES[s] for s in syms) if masses is None else tuple(float(v) for v in masses)
        key = (syms, basis.lower(), m)
        if key not in self._aotables:
            self._aotables[key] = AOTable(self, syms, basis, np.asarray(m, dtype=np.float64))
        return self._aotables[key]

    def center_of_mass(self, table, coords):
        coords = self.to_device(coords).reshape(-1, table.natm, 3)
        out = self.empty(coords.shape[0], 3)
        self._bind_stream()
        check(self.lib.evc_center_of_mass(self._ctx, table.handle, coords.shape[0], _ptr(coords), _ptr(out)))
        return out

    def int1e_r(self, table, coords, origin=None):
        """``<i| r - origin |j>`` (G, 3, n, n); ``origin`` (G, 3) or (3,), default: the centre of mass."""
        coords = self.to_device(coords).reshape(-1, table.natm, 3)
        G = coords.shape[0]
        if origin is None:
            origin = self.center_of_mass(table, coords)
        else:
            origin = self.to_device(np.broadcast_to(np.asarray(origin, dtype=np.float64).reshape(-1, 3), (G, 3)).copy()
                                    if not torch.is_tensor(origin) else origin).reshape(G, 3).contiguous()
        out = self.empty(G, 3, table.nao, table.nao)
        self._bind_stream()
        check(self.lib.evc_int1e_r(self._ctx, table.handle, G, _ptr(coords), _ptr(origin), _ptr(out)))
        return out

    def rdm1_observables(self, table, coords, x, gamma, ovlp=None, method="mulliken", want_dm_ao=False):
        """Dipole moment about the centre of mass (G, 3), atomic units, and atomic charges (G, natm) of the
        one-body density matrices ``gamma`` (G, n, n) given in the orthogonalised basis ``x`` (G, n, n)."""
        meth = {"mulliken": 0, "loewdin": 1, "lowdin": 1}[method.lower()]
        coords = self.to_device(coords).reshape(-1, table.natm, 3)
        G, n = coords.shape[0], table.nao
        x, gamma = self.to_device(x).reshape(G, n, n), self.to_device(gamma).reshape(G, n, n)
        ovlp = None if ovlp is None else self.to_device(ovlp).reshape(G, n, n)
        origin = self.center_of_mass(table, coords)
        rint = self.int1e_r(table, coords, origin)
        dm = self.empty(G, n, n) if want_dm_ao else None
        dip, chg = self.empty(G, 3), self.empty(G, table.natm)
        self._bind_stream()
        check(self.lib.evc_rdm1_observables(self._ctx, table.handle, G, meth, _ptr(coords), _ptr(origin), _ptr(x),
                                            _ptr(gamma), _ptr(ovlp) if ovlp is not None else None, _ptr(rint),
                                            _ptr(dm) if dm is not None else None, _ptr(dip), _ptr(chg)))
        return (dip, chg, dm) if want_dm_ao else (dip, chg)

    def ao_integrals(self, sbasis, coords, out=None, packed=False):
        """AO integrals of a batch of geometries on the device: ``coords`` (G, natm, 3) in
        bohr (device tensor or numpy) -> :class:`DeviceAO` (filled in place if given).
        ``packed`` (s-shell bases): emit the two-electron arrays in the packed layouts
        (``evc_ao_integrals_s_packed``); with ``out`` given, its own ``packed`` flag decides."""
        coords = self.to_device(coords).reshape(-1, sbasis.natm, 3)
        G = coords.shape[0]
        if out is None:
            packed = bool(packed) and not sbasis.general
            ao = DeviceAO(self, G, sbasis.nao, sbasis.natm, sbasis.aoslices_host, packed=packed)
        else:
            ao = out
        if ao.nbatch != G or ao.nao != sbasis.nao or ao.natm != sbasis.natm:
            raise ValueError("DeviceAO does not match the basis / batch size")
        if ao.packed and sbasis.general:
            raise ValueError("packed integral output exists for s-shell bases only (use DeviceAO.to_packed())")
        nbytes = C.c_size_t()
        if sbasis.general:
            ws_fn, fn = self.lib.evc_ao_integrals_sp_workspace_bytes, self.lib.evc_ao_integrals_sp
        else:
            ws_fn = self.lib.evc_ao_integrals_s_workspace_bytes
            fn = self.lib.evc_ao_integrals_s_packed if ao.packed else self.lib.evc_ao_integrals_s
        check(ws_fn(sbasis.handle, G, C.byref(nbytes)))
        ws = self.workspace(nbytes.value)
        self._bind_stream()
        e2, d2 = (ao.erip, ao.eri_ip1p) if ao.packed else (ao.eri, ao.eri_ip1)
        check(fn(self._ctx, sbasis.handle, G, _ptr(coords), _ptr(ao.ovlp), _ptr(ao.hcore), _ptr(e2),
                 _ptr(ao.ipovlp), _ptr(ao.hcore_deriv), _ptr(d2), _ptr(ao.e_nuc), _ptr(ao.grad_nuc),
                 _ptr(ws), ws.numel()))
        return ao

    def ao_pack8(self, eri=None, eri_ip1=None):
        """Device: full ``int2e (G, n, n, n, n)`` / ``int2e_ip1 (G, 3, n, n, n, n)`` -> packed
        ``(erip, eri_ip1p)`` (``evc_ao_pack8``); either may be ``None``."""
        src = eri if eri is not None else eri_ip1
        G, n = src.shape[0], src.shape[-1]
        shp = ao_shapes(G, n, 1, True)
        erip = self.empty(*shp["erip"]) if eri is not None else None
        ip1p = self.empty(*shp["eri_ip1p"]) if eri_ip1 is not None else None
        self._bind_stream()
        check(self.lib.evc_ao_pack8(self._ctx, G, n, _ptr(eri), _ptr(eri_ip1), _ptr(erip), _ptr(ip1p)))
        return erip, ip1p

    def energy_with_grad_coords(self, stack, sbasis, coords, ao=None, out=None, want_rdms=False):
        """The whole MD step from nuclear coordinates: K9 (integrals) then K3..K8.
        ``coords`` (G, natm, 3) bohr.  Same return value as :meth:`energy_with_grad`."""
        ao = self.ao_integrals(sbasis, coords, out=ao, packed=(sbasis.nao <= 13 and not want_rdms))
        return self.energy_with_grad(stack, ao, want_rdms=want_rdms, out=out)

    def energies(self, stack, ao, out=None):
        """Continuation energies only (K3, K4, K5, K6: ``approximate_ground_state_OAO`` for a batch): the packed
        step with ``grad = NULL``.  Returns ``(E[G], cvec[G, N])``."""
        G, n, natm, N = ao.nbatch, ao.nao, ao.natm, stack.ntrain
        if n != stack.norb:
            raise ValueError(f"mol.nao={n} does not match the stack's norb={stack.norb}")
        E, cvec = (self.empty(G), self.empty(G, N)) if out is None else out
        rh, rg = stack.packed()
        nbytes = C.c_size_t()
        check(self.lib.evc_energy_with_grad_packed_workspace_bytes(N, n, natm, G, C.byref(nbytes)))
        ws = self.workspace(nbytes.value)
        self._bind_stream()
        bundle = ao.bundle()
        check(self.lib.evc_energy_with_grad_packed(
            self._ctx, N, n, natm, _ptr(rh), _ptr(rg), _ptr(stack.linv), G, C.byref(bundle),
            _ptr(E), None, _ptr(cvec), _ptr(ws), ws.numel()))
        return E, cvec

    # -- device-resident velocity Verlet ------------------------------------------------------
    def md_positions(self, dt, x, v, a):
        G, natm = x.shape[0], x.shape[1]
        self._bind_stream()
        check(self.lib.evc_md_positions(self._ctx, G, natm, float(dt), _ptr(v), _ptr(a), _ptr(x)))

    def md_berendsen(self, dt, taut, temperature, ekin, v):
        G, natm = v.shape[0], v.shape[1]
        self._bind_stream()
        check(self.lib.evc_md_berendsen(self._ctx, G, natm, float(dt), float(taut), float(temperature),
                                        _ptr(ekin), _ptr(v)))

    def md_velocities(self, dt, first, inv_mass, mass, grad, x, epot, v, a, ekin, frame_idx, max_frames,
                      traj=None, epot_log=None, ekin_log=None):
        G, natm = x.shape[0], x.shape[1]
        self._bind_stream()
        check(self.lib.evc_md_velocities(self._ctx, G, natm, float(dt), 1 if first else 0, _ptr(inv_mass),
                                         _ptr(mass), _ptr(grad), _ptr(x), _ptr(epot), _ptr(v), _ptr(a),
                                         _ptr(ekin), _ptr(frame_idx), int(max_frames), _ptr(traj),
                                         _ptr(epot_log), _ptr(ekin_log)))

    # -- fused step ------------------------------------------------------------------------
    def energy_with_grad(self, stack, ao, want_rdms=False, out=None, packed=None):
        """One prediction step for a batch of geometries resident on the device.

        ``ao``: :class:`DeviceAO`.  Returns ``(E[G], grad[G,natm,3], gamma, Gamma, cvec)``
        (gamma/Gamma are ``None`` unless ``want_rdms``).  ``packed`` (default
        :data:`USE_PACKED`) selects the 8-fold-symmetric packed step
        (``evc_energy_with_grad_packed``); the full predicted RDMs only exist on the
        unpacked path, so ``want_rdms`` implies ``packed=False``.
        """
        G, n, natm, N = ao.nbatch, ao.nao, ao.natm, stack.ntrain
        if n != stack.norb:
            raise ValueError(f"mol.nao={n} does not match the stack's norb={stack.norb}")
        if out is None:
            E, grad, cvec = self.empty(G), self.empty(G, natm, 3), self.empty(G, N)
        else:
            E, grad, cvec = out
        packed = (USE_PACKED if packed is None else packed) and not want_rdms
        if getattr(ao, "packed", False) and not (packed and n <= 13):
            raise ValueError("packed AO arrays (erip / eri_ip1p) feed the packed step for n <= 13 only; "
                             "compute the full tensors for want_rdms / packed=False / larger n")
        if packed:
            rh, rg = stack.packed()
            nbytes = C.c_size_t()
            check(self.lib.evc_energy_with_grad_packed_workspace_bytes(N, n, natm, G, C.byref(nbytes)))
            ws = self.workspace(nbytes.value)
            self._bind_stream()
            bundle = ao.bundle()
            check(self.lib.evc_energy_with_grad_packed(
                self._ctx, N, n, natm, _ptr(rh), _ptr(rg), _ptr(stack.linv), G, C.byref(bundle),
                _ptr(E), _ptr(grad), _ptr(cvec), _ptr(ws), ws.numel()))
            return E, grad, None, None, cvec
        gamma = self.empty(G, n, n) if want_rdms else None
        Gamma = self.empty(G, n, n, n, n) if want_rdms else None
        nbytes = C.c_size_t()
        check(self.lib.evc_energy_with_grad_workspace_bytes(stack.layout, N, n, natm, G,
                                                            C.byref(nbytes)))
        ws = self.workspace(nbytes.value)
        self._bind_stream()
        bundle = ao.bundle()
        check(self.lib.evc_energy_with_grad(
            self._ctx, stack.layout, N, n, natm, _ptr(stack.one_rdm), _ptr(stack.two_rdm),
            _ptr(stack.linv), G, C.byref(bundle), _ptr(E), _ptr(grad), _ptr(gamma), _ptr(Gamma),
            _ptr(cvec), _ptr(ws), ws.numel()))
        return E, grad, gamma, Gamma, cvec

    def energy_with_grad_host(self, stack, host_ao, chunk=512, sync=True, packed=None):
        """The prediction step on HOST arrays (:class:`HostAO`): chunked host->device
        copies, kernels and read-back overlap on three streams.  Results land in
        ``host_ao.E`` / ``host_ao.grad`` (returned); with ``sync=False`` the caller
        synchronises the current stream before reading them."""
        G, n, natm, N = host_ao.nbatch, host_ao.nao, host_ao.natm, stack.ntrain
        if n != stack.norb:
            raise ValueError(f"mol.nao={n} does not match the stack's norb={stack.norb}")
        chunk = max(1, min(int(chunk), G))
        nbytes = C.c_size_t()
        if USE_PACKED if packed is None else packed:
            rh, rg = stack.packed()
            check(self.lib.evc_energy_with_grad_packed_host_workspace_bytes(N, n, natm, chunk,
                                                                            C.byref(nbytes)))
            ws = self.workspace(nbytes.value)
            self._bind_stream()
            bundle = host_ao.bundle()
            check(self.lib.evc_energy_with_grad_packed_host(
                self._ctx, N, n, natm, _ptr(rh), _ptr(rg), _ptr(stack.linv), G, C.byref(bundle),
                C.c_void_p(host_ao.E.data_ptr()), C.c_void_p(host_ao.grad.data_ptr()), chunk,
                _ptr(ws), ws.numel()))
            if sync:
                torch.cuda.current_stream(self.device).synchronize()
            return host_ao.E, host_ao.grad
        check(self.lib.evc_energy_with_grad_host_workspace_bytes(stack.layout, N, n, natm, chunk,
                                                                 C.byref(nbytes)))
        ws = self.workspace(nbytes.value)
        self._bind_stream()
        bundle = host_ao.bundle()
        check(self.lib.evc_energy_with_grad_host(
            self._ctx, stack.layout, N, n, natm, _ptr(stack.one_rdm), _ptr(stack.two_rdm),
            _ptr(stack.linv), G, C.byref(bundle), C.c_void_p(host_ao.E.data_ptr()),
            C.c_void_p(host_ao.grad.data_ptr()), chunk, _ptr(ws), ws.numel()))
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        return host_ao.E, host_ao.grad


class Workspace:
    """A scratch tensor owned by whoever holds this object.  ``freeze()`` pins the buffer: a later,
    larger request raises instead of re-allocating (a captured CUDA graph keeps the raw pointer)."""

    def __init__(self, device):
        self.device, self.tensor, self.frozen = device, None, False

    def get(self, nbytes):
        if self.tensor is None or self.tensor.numel() < nbytes:
            if self.frozen:
                raise RuntimeError(f"workspace frozen at {self.tensor.numel()} bytes (a CUDA graph references "
                                   f"it); {nbytes} bytes requested")
            self.tensor = None
            self.tensor = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=self.device)
        return self.tensor

    def freeze(self):
        self.frozen = True


class FCIHamiltonian:
    """``sigma = H c`` (``evc_fci_contract_2e``) and ``<K|H|K>`` (``evc_fci_hdiag``) for fixed
    integrals ``h1e (n,n)``, ``eri (n,n,n,n)`` in an orthonormal basis -- the two things a Davidson
    solver asks of the Hamiltonian (pyscf.fci.direct_spin1.contract_2e / make_hdiag)."""

    def __init__(self, engine, h1e, eri, norb, nelec):
        from . import cistring
        self.engine, self.norb = engine, int(norb)
        n = self.norb
        self.nelec = (int(nelec[0]), int(nelec[1]))
        self.na, self.nlink_a, self.link_a, _ = engine.link_tables(n, self.nelec[0])
        self.nb, self.nlink_b, self.link_b, _ = engine.link_tables(n, self.nelec[1])
        h1e = np.asarray(h1e, dtype=np.float64).reshape(n, n)
        eri = np.asarray(eri, dtype=np.float64).reshape(n, n, n, n)
        n2, n2p = n * n, (n * n + 1) & ~1
        h1eff = np.zeros(n2p)
        h1eff[:n2] = (h1e - 0.5 * np.einsum("pqqs->ps", eri)).reshape(-1)
        w2 = np.zeros((n2p, n2p))
        w2[:n2, :n2] = 0.5 * eri.reshape(n2, n2).T   # w2[(rs), (pq)] = 1/2 (pq|rs)
        self.h1eff, self.w2 = engine.to_device(h1eff), engine.to_device(w2)
        self.h1, self.eri = engine.to_device(h1e), engine.to_device(eri)
        self.strs_a = torch.from_numpy(cistring.make_strings(range(n), self.nelec[0])).to(engine.device)
        self.strs_b = torch.from_numpy(cistring.make_strings(range(n), self.nelec[1])).to(engine.device)
        nbytes = C.c_size_t()
        check(engine.lib.evc_fci_contract_workspace_bytes(n, self.na, self.nb, C.byref(nbytes)))
        self._ws_bytes = nbytes.value

    @property
    def ndet(self):
        return self.na * self.nb

    def hdiag(self):
        eng = self.engine
        out = eng.empty(self.ndet)
        eng._bind_stream()
        check(eng.lib.evc_fci_hdiag(eng._ctx, self.norb, self.na, self.nb, _ptr(self.strs_a), _ptr(self.strs_b),
                                    _ptr(self.h1), _ptr(self.eri), _ptr(out)))
        return out

    def contract(self, c, out=None):
        """``H c`` for a device vector ``c`` of ``ndet`` doubles."""
        eng = self.engine
        c = c.reshape(-1)
        out = eng.empty(self.ndet) if out is None else out
        ws = eng.workspace(self._ws_bytes)
        eng._bind_stream()
        check(eng.lib.evc_fci_contract_2e(eng._ctx, self.norb, self.na, self.nb, _ptr(self.link_a), self.nlink_a,
                                          _ptr(self.link_b), self.nlink_b, _ptr(self.h1eff), _ptr(self.w2),
                                          _ptr(c), _ptr(out), _ptr(ws), ws.numel()))
        return out


class AOTable:
    """Device copy of a basis description with nuclear charges and masses (``evc_aotable``): what the
    observables of the predicted density matrix need."""

    def __init__(self, engine, symbols, basis, masses):
        from .basis import has_p_shells, s_basis_tables, sp_basis_tables
        self.engine, self.symbols, self.basis = engine, tuple(symbols), basis.lower()
        general = has_p_shells(self.symbols, self.basis)
        t = sp_basis_tables(self.symbols, self.basis) if general else s_basis_tables(self.symbols, self.basis)
        self.tables, self.masses = t, np.ascontiguousarray(masses, dtype=np.float64)
        self.natm, self.nao = len(self.symbols), len(t["ao_atom"])
        handle = C.c_void_p()
        with torch.cuda.device(engine.device):
            check(engine.lib.evc_aotable_create(
                engine._ctx, self.natm, t["charges"].ctypes.data, self.masses.ctypes.data, self.nao,
                t["ao_atom"].ctypes.data, t["ao_pow"].ctypes.data if general else None, t["ao_nprim"].ctypes.data,
                t["prim_exp"].ctypes.data, t["prim_wt"].ctypes.data, C.byref(handle)))
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                self.engine.lib.evc_aotable_destroy(self.handle)
                self.handle = None
        except Exception:  # interpreter shutdown
            pass


class SBasis:
    """Device tables of a Gaussian basis for a fixed list of atoms: ``evc_sbasis`` (s shells only, the
    specialised kernel K9) or ``evc_gbasis`` (s and p shells, the general kernel K9g)."""

    def __init__(self, engine, symbols, basis):
        from .basis import has_p_shells, s_basis_tables, sp_basis_tables
        self.engine = engine
        self.symbols = tuple(s.capitalize() for s in symbols)
        self.basis = basis.lower()
        self.general = has_p_shells(self.symbols, self.basis)
        t = sp_basis_tables(self.symbols, self.basis) if self.general else s_basis_tables(self.symbols, self.basis)
        self.tables = t
        self.natm, self.nao = len(self.symbols), len(t["ao_atom"])
        handle = C.c_void_p()
        with torch.cuda.device(engine.device):
            if self.general:
                check(engine.lib.evc_gbasis_create(
                    engine._ctx, self.natm, t["charges"].ctypes.data, self.nao, t["ao_atom"].ctypes.data,
                    t["ao_pow"].ctypes.data, t["ao_nprim"].ctypes.data, t["prim_exp"].ctypes.data,
                    t["prim_wt"].ctypes.data, C.byref(handle)))
            else:
                check(engine.lib.evc_sbasis_create(
                    engine._ctx, self.natm, t["charges"].ctypes.data, self.nao, t["ao_atom"].ctypes.data,
                    t["ao_nprim"].ctypes.data, t["prim_exp"].ctypes.data, t["prim_wt"].ctypes.data,
                    C.byref(handle)))
        self.handle = handle
        sl = np.zeros((self.natm, 2), dtype=np.int32)
        for A in range(self.natm):
            idx = np.flatnonzero(t["ao_atom"] == A)
            sl[A] = (idx[0], idx[-1] + 1) if len(idx) else (0, 0)
        self.aoslices_host = sl
        self.charges = t["charges"]

    def __del__(self):
        try:
            if self.handle:
                if self.general:
                    self.engine.lib.evc_gbasis_destroy(self.handle)
                else:
                    self.engine.lib.evc_sbasis_destroy(self.handle)
                self.handle = None
        except Exception:  # interpreter shutdown
            pass


class HostAO:
    """AO arrays of a batch of geometries in (pinned) HOST memory -- the input of
    :meth:`Engine.energy_with_grad_host`.  Fields as in :class:`DeviceAO`."""

    FIELDS = ("ovlp", "hcore", "eri", "ipovlp", "hcore_deriv", "eri_ip1", "e_nuc", "grad_nuc")
    PACKED_FIELDS = ("ovlp", "hcore", "erip", "ipovlp", "hcore_deriv", "eri_ip1p", "e_nuc", "grad_nuc")

    def __init__(self, nbatch, nao, natm, aoslices, pin=True, packed=False):
        self.nbatch, self.nao, self.natm = nbatch, nao, natm
        self.packed = bool(packed)
        self.fields = self.PACKED_FIELDS if self.packed else self.FIELDS
        self.shapes = ao_shapes(nbatch, nao, natm, self.packed)
        pin = bool(pin and torch.cuda.is_available())
        self.eri = self.eri_ip1 = self.erip = self.eri_ip1p = None
        for k in self.fields:
            t = torch.empty(self.shapes[k], dtype=torch.float64, pin_memory=pin)
            if k == "erip":
                t.zero_()
            setattr(self, k, t)
        self.aoslices = torch.from_numpy(
            np.ascontiguousarray(aoslices, dtype=np.int32).reshape(natm, 2).copy())
        self.E = torch.empty(nbatch, dtype=torch.float64, pin_memory=pin)
        self.grad = torch.empty(nbatch, natm, 3, dtype=torch.float64, pin_memory=pin)

    @classmethod
    def from_bundles(cls, bundles, pin=True, packed=False):
        b0 = bundles[0]
        self = cls(len(bundles), b0["nao"], b0["natm"], b0["aoslices"], pin=pin, packed=packed)
        for g, b in enumerate(bundles):
            src = dict(b)
            if packed:
                src["erip"], src["eri_ip1p"] = pack_ao_host(np.asarray(b["eri"]), np.asarray(b["eri_ip1"]))
            for k in self.fields:
                getattr(self, k).numpy()[g] = np.asarray(src[k], dtype=np.float64).reshape(self.shapes[k][1:])
        return self

    def nbytes(self):
        return sum(getattr(self, k).numel() * 8 for k in self.fields)

    def bundle(self):
        b = AoBundle()
        for k in self.fields:
            setattr(b, k, getattr(self, k).data_ptr())
        b.aoslices = self.aoslices.data_ptr()
        return b


class DeviceStack:
    """The t-RDM stack resident in HBM, in one of the reference's four layouts.

    ``overlap (N,N)``, ``one_rdm (N,N,n,n)`` and ``two_rdm`` as accepted by
    evcont/ab_initio_eigenvector_continuation.py:12-90 (dispatch on ``two_rdm.ndim``:
    6 / 5 / 3 / 2).  The Cholesky factor of ``overlap`` is computed once here
    because S does not depend on the geometry.
    """

    def __init__(self, overlap, one_rdm, two_rdm, engine=None, norb=None):
        self.engine = engine or get_engine()
        eng = self.engine
        ndim = two_rdm.ndim
        if ndim not in (6, 5, 3, 2):
            raise AssertionError("two_RDM must have 2, 3, 5 or 6 dimensions")
        self.layout = ndim
        self.ntrain = int(overlap.shape[0])
        self.norb = int(one_rdm.shape[-1]) if norb is None else int(norb)
        self.overlap = eng.to_device(overlap)
        self.one_rdm = eng.to_device(one_rdm).reshape(self.ntrain * self.ntrain, -1)
        npairs = self.ntrain ** 2 if ndim in (6, 3) else self.ntrain * (self.ntrain + 1) // 2
        self.two_rdm = eng.to_device(two_rdm).reshape(npairs, -1)
        n2 = self.norb ** 2
        expect = n2 * n2 if ndim in (6, 5) else n2 * (n2 + 1) // 2
        if self.two_rdm.shape[1] != expect or self.one_rdm.shape[1] != n2:
            raise ValueError("t-RDM stack shapes are inconsistent with norb/ntrain")
        self.linv = eng.geneig_prepare(self.overlap)
        self._packed = None

    def packed(self):
        """``(RH, RG)``: the stack packed on the 8-fold integral symmetry
        (``evc_stack_pack8``), built on first use and kept in HBM."""
        if self._packed is None:
            eng = self.engine
            npairs = self.ntrain * (self.ntrain + 1) // 2
            row = int(eng.lib.evc_packed_row_len(self.norb))
            rh, rg = eng.empty(npairs, row), eng.empty(npairs, row)
            eng._bind_stream()
            check(eng.lib.evc_stack_pack8(eng._ctx, self.layout, self.ntrain, self.norb,
                                          _ptr(self.one_rdm), _ptr(self.two_rdm), _ptr(rh), _ptr(rg)))
            self._packed = (rh, rg)
        return self._packed

    @property
    def packed_nbytes(self):
        rh, rg = self.packed()
        return rh.numel() * 8 + rg.numel() * 8

    @property
    def nbytes(self):
        return self.two_rdm.numel() * 8 + self.one_rdm.numel() * 8


class DeviceAO:
    """AO arrays of a batch of geometries on the device (the ``evc_ao_bundle``).

    ``packed=True`` holds the two-electron arrays in the packed layouts of the bundle
    (``erip (G, np, pitch)``, ``eri_ip1p (G, 3, n, n, np)``, np = n(n+1)/2) instead of the full
    tensors ``eri`` / ``eri_ip1`` -- what ``evc_ao_integrals_s_packed`` emits and the packed
    prediction step reads directly (n <= 13)."""

    FIELDS = ("ovlp", "hcore", "eri", "ipovlp", "hcore_deriv", "eri_ip1", "e_nuc", "grad_nuc")
    PACKED_FIELDS = ("ovlp", "hcore", "erip", "ipovlp", "hcore_deriv", "eri_ip1p", "e_nuc", "grad_nuc")

    def __init__(self, engine, nbatch, nao, natm, aoslices, packed=False):
        self.engine, self.nbatch, self.nao, self.natm = engine, nbatch, nao, natm
        self.packed = bool(packed)
        self.shapes = ao_shapes(nbatch, nao, natm, self.packed)
        self.fields = self.PACKED_FIELDS if self.packed else self.FIELDS
        self.eri = self.eri_ip1 = self.erip = self.eri_ip1p = None
        for k in self.fields:
            setattr(self, k, engine.empty(*self.shapes[k]))
        self.aoslices = torch.from_numpy(
            np.ascontiguousarray(aoslices, dtype=np.int32).reshape(natm, 2)).to(engine.device)

    @classmethod
    def from_bundles(cls, engine, bundles):
        """Upload a list of host ``ao_bundle`` dicts (evcont_b200.mol.ao_bundle)."""
        b0 = bundles[0]
        self = cls(engine, len(bundles), b0["nao"], b0["natm"], b0["aoslices"])
        for k in cls.FIELDS:
            host = np.stack([np.asarray(b[k], dtype=np.float64) for b in bundles])
            getattr(self, k).copy_(torch.from_numpy(host.reshape(self.shapes[k])))
        return self

    def to_packed(self):
        """A packed copy (``evc_ao_pack8`` for the two-electron arrays; the small arrays are shared)."""
        if self.packed:
            return self
        out = DeviceAO.__new__(DeviceAO)
        out.engine, out.nbatch, out.nao, out.natm = self.engine, self.nbatch, self.nao, self.natm
        out.packed, out.fields = True, self.PACKED_FIELDS
        out.shapes = ao_shapes(self.nbatch, self.nao, self.natm, True)
        out.eri = out.eri_ip1 = None
        for k in ("ovlp", "hcore", "ipovlp", "hcore_deriv", "e_nuc", "grad_nuc"):
            setattr(out, k, getattr(self, k))
        out.aoslices = self.aoslices
        out.erip, out.eri_ip1p = self.engine.ao_pack8(self.eri, self.eri_ip1)
        return out

    def nbytes(self):
        return sum(getattr(self, k).numel() * 8 for k in self.fields)

    def bundle(self):
        b = AoBundle()
        for k in self.fields:
            setattr(b, k, getattr(self, k).data_ptr())
        b.aoslices = self.aoslices.data_ptr()
        return b


def ao_shapes(nbatch, n, natm, packed=False):
    """Shapes of the ``evc_ao_bundle`` arrays of ``nbatch`` geometries."""
    shp = dict(ovlp=(nbatch, n, n), hcore=(nbatch, n, n), ipovlp=(nbatch, 3, n, n),
               hcore_deriv=(nbatch, natm, 3, n, n), e_nuc=(nbatch,), grad_nuc=(nbatch, natm, 3))
    if packed:
        npair = n * (n + 1) // 2
        shp.update(erip=(nbatch, npair, int(_lib.lib().evc_erip_pitch(n))), eri_ip1p=(nbatch, 3, n, n, npair))
    else:
        shp.update(eri=(nbatch, n, n, n, n), eri_ip1=(nbatch, 3, n, n, n, n))
    return shp


def pack_ao_host(eri, eri_ip1):
    """numpy: full ``int2e (..., n, n, n, n)`` / ``int2e_ip1 (..., 3, n, n, n, n)`` -> the packed host arrays
    ``erip (..., np, pitch)`` (padding columns zero), ``eri_ip1p (..., 3, n, n, np)``.  Either may be ``None``."""
    erip = ip1p = None
    src = eri if eri is not None else eri_ip1
    n = src.shape[-1]
    ii, jj = np.tril_indices(n)
    if eri is not None:
        pitch = int(_lib.lib().evc_erip_pitch(n))
        pk = np.asarray(eri)[..., ii, jj, :, :][..., ii, jj]
        erip = np.zeros(pk.shape[:-1] + (pitch,))
        erip[..., : pk.shape[-1]] = pk
    if eri_ip1 is not None:
        ip1p = np.ascontiguousarray(np.asarray(eri_ip1)[..., ii, jj])
    return erip, ip1p
