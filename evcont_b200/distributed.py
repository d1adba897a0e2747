"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink).

Only two things on the EVCont path shard:

* **building the t-RDM stack** -- the N(N+1)/2 training-state pairs are
  independent units.  Every rank holds all CI vectors (they are small: 0.5 MB
  each at H10, 13 MB at H2O), computes the transition RDMs of its share of the
  pair list with the fused DMMA kernel, and ONE ``all_gather`` of the padded
  per-rank slabs assembles the full stack on every rank.  The reference runs
  this loop serially on one process (evcont/FCI_EVCont.py:117-127).
* **prediction** -- independent geometries / MD trajectories (the reference
  runs one OS process per trajectory, scripts/MD/Zundel_thermodynamics/
  continuation/04_Zundel_continuation_MD.py:32).  The stack is replicated, the
  geometries are dealt out; there is no per-step collective.

The data-path collective is called on device tensors with the NCCL backend; the
same code runs on CPU tensors with ``gloo`` (tests, world_size 2).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def _ptr(t):
    from .engine import _ptr as p
    return p(t)


def _check(rc):
    from .engine import check
    return check(rc)


_SCATTER_INDEX = {}   # (ntrain, world, device) -> (row, column) block indices of the gathered rows


def tril_pairs(ntrain):
    """Pairs (a, b) with a >= b in ``np.tril_indices`` order."""
    return [(a, b) for a in range(ntrain) for b in range(a + 1)]


def shard_range(nitems, rank, world):
    """Contiguous, balanced ``[lo, hi)`` share of ``nitems`` for ``rank``."""
    lo = nitems * rank // world
    hi = nitems * (rank + 1) // world
    return lo, hi


def slab_size(nitems, world):
    """Rows of the padded per-rank slab (equal on all ranks, as all_gather needs)."""
    return max(shard_range(nitems, r, world)[1] - shard_range(nitems, r, world)[0]
               for r in range(world))


def shard_geometries(ngeom, rank, world):
    """Indices of the geometries / trajectories rank ``rank`` predicts."""
    lo, hi = shard_range(ngeom, rank, world)
    return np.arange(lo, hi)


def _all_gather_rows(local, nitems, group=None):
    """all_gather of equally padded ``(slab, width)`` row blocks -> ``(nitems, width)``."""
    world = dist.get_world_size(group)
    slab = slab_size(nitems, world)
    width = local.shape[1]
    send = local.new_zeros((slab, width))
    send[: local.shape[0]] = local
    recv = local.new_empty((world * slab, width))
    dist.all_gather_into_tensor(recv, send, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(nitems, r, world)
        parts.append(recv[r * slab: r * slab + (hi - lo)])
    return torch.cat(parts, dim=0)


def stack_row_len(norb):
    """Doubles of one slab row ``[dm2 (n^4) | dm1 (n^2) | ovlp]`` (rounded up to even: 16-byte rows)."""
    n2 = norb * norb
    return (n2 * n2 + n2 + 1 + 1) & ~1


def build_stack_sharded(civecs, norb, nelec, pair_fn=None, group=None, device=None):
    """All-pairs t-RDM stack from ``civecs`` (N, na, nb), pairs sharded over the ranks, assembled with ONE
    ``all_gather`` of the per-rank slabs.

    Every pair's results form one contiguous row ``[dm2 | dm1 | ovlp]``; a rank's rows are its slab.  On the GPU
    the trans-RDM kernel writes the rows itself (``Engine.trans_rdm12_rows``) and one small kernel places the
    gathered rows into the reference's (N, N, ...) layout (``evc_stack_scatter_rows``), mirror blocks untransposed
    (evcont/FCI_EVCont.py:124-127).  ``pair_fn(civecs, pairs) -> (ovlp, dm1, dm2)`` replaces the kernel in the
    CPU (gloo) tests of this host logic.  Returns ``(overlap (N,N), one_rdm (N,N,n,n), two_rdm (N,N,n,n,n,n))``,
    identical on every rank and bit-identical to a single-GPU build: a pair's row does not depend on which rank
    computed it.
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if not isinstance(civecs, torch.Tensor):   # device tensors stay where they are (no re-upload per build)
        civecs = np.asarray(civecs, dtype=np.float64)
    N, n = civecs.shape[0], int(norb)
    n2 = n * n
    pairs = tril_pairs(N)
    lo, hi = shard_range(len(pairs), rank, world)
    # Every rank computes exactly `slab` pairs -- its share, padded with repeats of pairs[0] -- so that the kernel
    # output IS the equally sized send buffer of the all_gather; the padding rows rewrite block [0, 0] with the
    # value it already has.
    slab = slab_size(len(pairs), world)
    padded = pairs[lo:hi] + [pairs[0]] * (slab - (hi - lo))
    width = stack_row_len(n)
    eng = None
    if pair_fn is None:
        from .engine import get_engine
        eng = get_engine(device)
        rows = eng.trans_rdm12_rows(civecs, padded, n, nelec, plan_pairs=len(pairs))
    else:
        ovlp, dm1, dm2 = pair_fn(civecs, padded)
        rows = dm2.new_zeros((slab, width))
        rows[:, : n2 * n2] = dm2.reshape(slab, n2 * n2)
        rows[:, n2 * n2: n2 * n2 + n2] = dm1.reshape(slab, n2)
        rows[:, n2 * n2 + n2] = ovlp.reshape(slab)
    gathered = rows.new_empty((world * slab, width))
    dist.all_gather_into_tensor(gathered, rows.contiguous(), group=group)   # the one collective of the build
    key = (N, world, str(gathered.device))
    if key not in _SCATTER_INDEX:
        ab = []
        for r in range(world):
            rlo, rhi = shard_range(len(pairs), r, world)
            ab += pairs[rlo:rhi] + [pairs[0]] * (slab - (rhi - rlo))
        _SCATTER_INDEX[key] = torch.tensor(ab, dtype=torch.int32, device=gathered.device).reshape(-1, 2).contiguous()
    row_pairs = _SCATTER_INDEX[key]
    if eng is not None:
        return eng.stack_scatter_rows(gathered, row_pairs, N, n)
    # CPU tensors (gloo tests of the host logic): the same placement with index assignments
    ia, ib = row_pairs[:, 0].long(), row_pairs[:, 1].long()
    overlap = gathered.new_empty((N, N))
    one = gathered.new_empty((N, N, n, n))
    two = gathered.new_empty((N, N, n, n, n, n))
    g_d2 = gathered[:, : n2 * n2].reshape(-1, n, n, n, n)
    g_d1 = gathered[:, n2 * n2: n2 * n2 + n2].reshape(-1, n, n)
    g_ov = gathered[:, n2 * n2 + n2]
    overlap[ia, ib] = g_ov
    overlap[ib, ia] = g_ov
    one[ia, ib] = g_d1
    one[ib, ia] = g_d1
    two[ia, ib] = g_d2
    two[ib, ia] = g_d2
    return overlap, one, two


def build_stack_single(civecs, norb, nelec, device=None):
    """The same build on one GPU, through the same row kernel and placement kernel (no collective)."""
    from .engine import get_engine
    eng = get_engine(device)
    N, n = civecs.shape[0], int(norb)
    pairs = tril_pairs(N)
    rows = eng.trans_rdm12_rows(civecs, pairs, n, nelec)
    row_pairs = torch.tensor(pairs, dtype=torch.int32, device=rows.device).reshape(-1, 2).contiguous()
    return eng.stack_scatter_rows(rows, row_pairs, N, n)


def gather_predictions(E_local, grad_local, ngeom, group=None):
    """Optional final gather of sharded predictions: ``(E[ngeom], grad[ngeom, natm, 3])``."""
    natm = grad_local.shape[1]
    local = torch.cat([E_local.reshape(-1, 1), grad_local.reshape(-1, natm * 3)], dim=1)
    rows = _all_gather_rows(local.contiguous(), ngeom, group=group)
    return rows[:, 0].contiguous(), rows[:, 1:].reshape(ngeom, natm, 3).contiguous()


# ---- e3: one trajectory on a pair-sharded stack ------------------------------
class PairShardedStack:
    """One rank's slab of an exchange-compressed, lower-triangular stack (``two_rdm`` of shape
    ``(N (N + 1) / 2, n^2 (n^2 + 1) / 2)``, rows in ``np.tril_indices(N)`` order: the layout the reference's
    Zundel run assembles from its per-pair directories,
    scripts/MD/Zundel_thermodynamics/continuation/04_Zundel_continuation_MD.py:99-128).

    The N = 100 Zundel stack is 12.5 GB and one prediction streams it twice (K5, K7): 4 ms per MD step on one
    B200 for ONE trajectory.  Sharded over R ranks each rank streams 1/R of it; per step the ranks exchange
    N (N + 1) / 2 doubles (all_gather of the H entries) and one exchange-compressed two-body density matrix
    (all_reduce, 2.4 MB at n = 28).  Everything that does not touch the stack (integrals, Loewdin, AO->OAO
    transform, eigenproblem, gradient) is replicated: it is the same small work on every rank.

    ``overlap`` and ``one_rdm`` (N, N, n, n) are replicated (7.8 MB at N = 100, n = 28)."""

    def __init__(self, overlap, one_rdm, two_rdm_rows, lo, hi, engine=None):
        from .engine import get_engine
        self.engine = eng = engine or get_engine()
        self.overlap = eng.to_device(np.ascontiguousarray(overlap, dtype=np.float64) if not torch.is_tensor(overlap)
                                     else overlap)
        self.ntrain = int(self.overlap.shape[0])
        one = eng.to_device(one_rdm)
        self.norb = int(one.shape[-1])
        self.one_rdm = one.reshape(self.ntrain * self.ntrain, self.norb * self.norb).contiguous()
        self.lo, self.hi = int(lo), int(hi)
        self.rows = eng.to_device(two_rdm_rows).contiguous()
        n2 = self.norb * self.norb
        self.row_len = n2 * (n2 + 1) // 2
        if self.rows.shape != (self.hi - self.lo, self.row_len):
            raise ValueError(f"slab of shape {tuple(self.rows.shape)}; expected {(self.hi - self.lo, self.row_len)}")
        self.npairs = self.ntrain * (self.ntrain + 1) // 2
        il = np.tril_indices(self.ntrain)
        self._pa = torch.as_tensor(il[0], device=eng.device)
        self._pb = torch.as_tensor(il[1], device=eng.device)
        self.linv = eng.geneig_prepare(self.overlap)

    @classmethod
    def from_full(cls, overlap, one_rdm, two_rdm, rank, world, engine=None):
        """Slab ``rank`` of ``world`` of a stack held in full (tests, small stacks)."""
        lo, hi = shard_range(two_rdm.shape[0], rank, world)
        return cls(overlap, one_rdm, two_rdm[lo:hi], lo, hi, engine=engine)

    @classmethod
    def from_pair_dirs(cls, root, ntrain, rank, world, engine=None, pattern="MPS_cross_{}_{}"):
        """Every rank reads only ITS pairs' ``two_rdm.npy`` (and all of the small ``ovlp.npy`` / ``one_rdm.npy``)
        from the reference's per-pair directories (04_Zundel_continuation_MD.py:99-128)."""
        import os
        il = np.tril_indices(ntrain)
        npairs = len(il[0])
        lo, hi = shard_range(npairs, rank, world)
        first = np.load(os.path.join(root, pattern.format(il[0][0], il[1][0]), "one_rdm.npy"))
        n = first.shape[-1]
        overlap = np.zeros((ntrain, ntrain))
        one = np.zeros((ntrain, ntrain, n, n))
        for a, b in zip(*il):
            d = os.path.join(root, pattern.format(a, b))
            overlap[a, b] = overlap[b, a] = np.load(os.path.join(d, "ovlp.npy"))
            one[a, b] = np.load(os.path.join(d, "one_rdm.npy"))
            one[b, a] = one[a, b]          # the reference mirrors the untransposed block (:116-118)
        rows = np.stack([np.load(os.path.join(root, pattern.format(il[0][p], il[1][p]), "two_rdm.npy"))
                         for p in range(lo, hi)]) if hi > lo else np.zeros((0, n * n * (n * n + 1) // 2))
        return cls(overlap, one, rows, lo, hi, engine=engine)

    # -- the two stack passes of a step, on this rank's slab ---------------------------------------------
    def _ws(self, G):
        nbytes = C.c_size_t()
        _check(self.engine.lib.evc_stack_rows_workspace_bytes(self.row_len, max(1, self.hi - self.lo), G,
                                                              C.byref(nbytes)))
        return self.engine.workspace(nbytes.value)

    def partial_H(self, hv):
        """(G, hi - lo): two-body part of H for this rank's pairs, ``sum_l rows[p][l] hv[g][l]``."""
        eng, G, P = self.engine, hv.shape[0], self.hi - self.lo
        out = eng.empty(G, P)
        if P:
            ws = self._ws(G)
            eng._bind_stream()
            _check(eng.lib.evc_stack_rows_dot(eng._ctx, _ptr(self.rows), self.row_len, P, _ptr(hv), G, _ptr(out),
                                              _ptr(ws), ws.numel()))
        return out

    def partial_gamma2c(self, cvec):
        """(G, row_len): this rank's share of the exchange-compressed predicted two-body density matrix,
        ``sum_{p in slab} (2 - delta_ab) c_a c_b rows[p]``."""
        eng, G, P = self.engine, cvec.shape[0], self.hi - self.lo
        out = torch.zeros((G, self.row_len), dtype=torch.float64, device=eng.device)
        if P:
            a, b = self._pa[self.lo:self.hi], self._pb[self.lo:self.hi]
            w = (cvec[:, a] * cvec[:, b] * torch.where(a == b, 1.0, 2.0)).contiguous()
            ws = self._ws(G)
            eng._bind_stream()
            _check(eng.lib.evc_stack_rows_axpy(eng._ctx, _ptr(self.rows), self.row_len, P, _ptr(w), G, _ptr(out),
                                               _ptr(ws), ws.numel()))
        return out

    def assemble_H(self, h1, h_two):
        """(G, N, N), lower triangle: ``<one_rdm[a, b], h1> + h_two[p(a, b)]`` from the gathered entries."""
        G, N = h1.shape[0], self.ntrain
        H1 = (h1.reshape(G, -1) @ self.one_rdm.T).reshape(G, N, N)
        H = torch.zeros((G, N, N), dtype=torch.float64, device=h1.device)
        H[:, self._pa, self._pb] = H1[:, self._pa, self._pb] + h_two
        return H


def gather_pair_entries(local, npairs, group=None):
    """all_gather of the ranks' ``(slab_r, G)`` blocks of per-pair values into ``(npairs, G)`` in pair order
    (rank r owns ``shard_range(npairs, r, world)``; blocks are padded to the largest slab for the collective)."""
    world = dist.get_world_size(group)
    slab = slab_size(npairs, world)
    send = local.new_zeros((slab, local.shape[1]))
    send[: local.shape[0]] = local
    recv = local.new_empty((world * slab, local.shape[1]))
    dist.all_gather_into_tensor(recv, send, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(npairs, r, world)
        parts.append(recv[r * slab: r * slab + (hi - lo)])
    return torch.cat(parts, dim=0)


def sharded_energy_with_grad(shards, ao, group=None, want_rdms=False):
    """One prediction step (energy + forces) with the stack sharded by training pairs.

    ``shards``: this rank's :class:`PairShardedStack` -- or, without a process group, a LIST of slabs held by
    one process (their partial results are summed locally: the arithmetic of the sharded step without the
    collectives; tests and single-GPU use).  ``ao``: :class:`evcont_b200.engine.DeviceAO` (full, not packed
    two-electron arrays).  Returns ``(E (G,), grad (G, natm, 3))`` (+ gamma, Gamma with ``want_rdms``),
    identical on every rank.

    Per step: all_gather of ``slab`` H entries per rank, all_reduce of ``n^2 (n^2 + 1) / 2`` doubles."""
    local = shards if isinstance(shards, (list, tuple)) else [shards]
    sh0 = local[0]
    eng, n, N = sh0.engine, sh0.norb, sh0.ntrain
    G = ao.nbatch
    if ao.nao != n:
        raise ValueError(f"mol.nao={ao.nao} does not match the stack's norb={n}")
    distributed = group is not None or (dist.is_available() and dist.is_initialized() and len(local) == 1
                                        and not isinstance(shards, (list, tuple)))
    # replicated: Loewdin, AO -> OAO, exchange-compressed Hamiltonian
    x, evals, evecs = eng.loewdin(ao.ovlp)
    h1, h2, t3 = eng.ao2oao(ao.hcore, ao.eri, x, want_t3=True)
    hv = eng.empty(G, sh0.row_len)
    eng._bind_stream()
    _check(eng.lib.evc_exchange_compress(eng._ctx, n, G, _ptr(h2), _ptr(hv)))
    # K5 on the slabs -> all_gather of the entries
    if distributed:
        h_two = gather_pair_entries(sh0.partial_H(hv).T.contiguous(), sh0.npairs, group).T.contiguous()
    else:
        if [s.lo for s in local] != [0] + [s.hi for s in local[:-1]] or local[-1].hi != sh0.npairs:
            raise ValueError("the slabs of a local list must cover the pair list in order")
        h_two = torch.cat([s.partial_H(hv) for s in local], dim=1)
    H = sh0.assemble_H(h1, h_two)
    E, cv = eng.geneig(H, sh0.linv, 1)
    cvec = cv[:, 0, :].contiguous()
    # K7 on the slabs -> all_reduce of the compressed two-body density matrix
    if distributed:
        g2c = sh0.partial_gamma2c(cvec)
        dist.all_reduce(g2c, op=dist.ReduceOp.SUM, group=group)
    else:
        g2c = local[0].partial_gamma2c(cvec)
        for s in local[1:]:
            g2c += s.partial_gamma2c(cvec)
    Gamma = eng.empty(G, n, n, n, n)
    eng._bind_stream()
    _check(eng.lib.evc_exchange_restore(eng._ctx, n, G, _ptr(g2c), _ptr(Gamma)))
    cc = (cvec[:, :, None] * cvec[:, None, :]).reshape(G, N * N)
    gamma = (cc @ sh0.one_rdm).reshape(G, n, n).contiguous()
    # replicated: gradient
    g_el = eng.grad_elec(ao.aoslices, evals, evecs, x, ao.hcore, t3, gamma, Gamma, ao.ipovlp, ao.hcore_deriv,
                         ao.eri_ip1)
    Etot = E[:, 0] + ao.e_nuc
    grad = g_el + ao.grad_nuc
    return (Etot, grad, gamma, Gamma) if want_rdms else (Etot, grad)
