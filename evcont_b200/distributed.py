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


def build_stack_sharded(civecs, norb, nelec, pair_fn=None, group=None, device=None):
    """All-pairs t-RDM stack from ``civecs`` (N, na, nb), pairs sharded over the ranks.

    ``pair_fn(civecs, pairs) -> (ovlp[np], dm1[np, n, n], dm2[np, n, n, n, n])`` as
    torch tensors computes this rank's pairs; the default is the GPU engine's
    batched trans-RDM kernel.  Returns tensors ``(overlap (N,N), one_rdm
    (N,N,n,n), two_rdm (N,N,n,n,n,n))`` identical on every rank, mirror blocks
    untransposed like the reference (evcont/FCI_EVCont.py:124-127).
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if not isinstance(civecs, torch.Tensor):   # device tensors stay where they are (no re-upload per build)
        civecs = np.asarray(civecs, dtype=np.float64)
    N, n = civecs.shape[0], int(norb)
    pairs = tril_pairs(N)
    lo, hi = shard_range(len(pairs), rank, world)
    if pair_fn is None:
        from .engine import get_engine
        eng = get_engine(device)

        def pair_fn(vecs, plist):
            return eng.trans_rdm12_batch(vecs, plist, n, nelec)
    # Every rank computes exactly `slab` pairs -- its share, padded with repeats of pairs[0] -- so that the kernel
    # outputs ARE the equally sized send buffers of the all_gather (no concatenation / zero-fill copies), and the
    # gathered rows scatter straight into the (N, N, ...) layout: the padding rows rewrite block [0, 0] with the
    # value it already has.
    slab = slab_size(len(pairs), world)
    padded = pairs[lo:hi] + [pairs[0]] * (slab - (hi - lo))
    ovlp, dm1, dm2 = pair_fn(civecs, padded)
    ovlp, dm1, dm2 = ovlp.reshape(slab).contiguous(), dm1.reshape(slab, n, n).contiguous(), \
        dm2.reshape(slab, n, n, n, n).contiguous()
    g_ov = ovlp.new_empty((world * slab,))
    g_d1 = dm1.new_empty((world * slab, n, n))
    g_d2 = dm2.new_empty((world * slab, n, n, n, n))
    dist.all_gather_into_tensor(g_d2, dm2, group=group)
    dist.all_gather_into_tensor(g_d1, dm1, group=group)
    dist.all_gather_into_tensor(g_ov, ovlp, group=group)
    key = (N, world, str(g_ov.device))
    if key not in _SCATTER_INDEX:
        ia, ib = [], []
        for r in range(world):
            rlo, rhi = shard_range(len(pairs), r, world)
            rows = pairs[rlo:rhi] + [pairs[0]] * (slab - (rhi - rlo))
            ia += [p[0] for p in rows]
            ib += [p[1] for p in rows]
        _SCATTER_INDEX[key] = (torch.tensor(ia, device=g_ov.device), torch.tensor(ib, device=g_ov.device))
    ia, ib = _SCATTER_INDEX[key]
    overlap = g_ov.new_empty((N, N))
    one = g_ov.new_empty((N, N, n, n))
    two = g_ov.new_empty((N, N, n, n, n, n))
    overlap[ia, ib] = g_ov
    overlap[ib, ia] = g_ov
    one[ia, ib] = g_d1
    one[ib, ia] = g_d1
    two[ia, ib] = g_d2
    two[ib, ia] = g_d2
    return overlap, one, two


def gather_predictions(E_local, grad_local, ngeom, group=None):
    """Optional final gather of sharded predictions: ``(E[ngeom], grad[ngeom, natm, 3])``."""
    natm = grad_local.shape[1]
    local = torch.cat([E_local.reshape(-1, 1), grad_local.reshape(-1, natm * 3)], dim=1)
    rows = _all_gather_rows(local.contiguous(), ngeom, group=group)
    return rows[:, 0].contiguous(), rows[:, 1:].reshape(ngeom, natm, 3).contiguous()
