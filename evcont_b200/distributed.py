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
import numpy as np
import torch
import torch.distributed as dist


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
        rows = eng.trans_rdm12_rows(civecs, padded, n, nelec)
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
