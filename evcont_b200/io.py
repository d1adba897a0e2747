"""On-disk interchange of t-RDM stacks in the reference's own file formats.

* ``overlap[_i].npy / one_rdm[_i].npy / two_rdm[_i].npy`` -- what ``converge_EVCont_MD`` writes after
  every training step (evcont/MD_utils.py:176-184) and what the MD scripts read back to resume
  (scripts/MD/md_H30_evcont_from_DMRG.py:73-75).  Any of the four ``two_rdm`` layouts.
* per-pair directories ``MPS_cross_{a}_{b}/{ovlp,one_rdm,two_rdm}.npy`` for ``a >= b`` -- how the Zundel
  workflow stores the DMRG transition RDMs, read by
  scripts/MD/Zundel_thermodynamics/continuation/04_Zundel_continuation_MD.py:99-128 into the
  ``(N(N+1)/2, n^2(n^2+1)/2)`` layout with a mirrored ``one_rdm``.

Host-side file handling only; the arrays go to the device through the usual entry points.
"""
import os

import numpy as np


def _name(kind, index):
    return f"{kind}.npy" if index is None else f"{kind}_{index}.npy"


def save_stack(directory, overlap, one_rdm, two_rdm, index=None):
    os.makedirs(directory, exist_ok=True)
    np.save(os.path.join(directory, _name("overlap", index)), np.asarray(overlap))
    np.save(os.path.join(directory, _name("one_rdm", index)), np.asarray(one_rdm))
    np.save(os.path.join(directory, _name("two_rdm", index)), np.asarray(two_rdm))


def load_stack(directory, index=None, mmap=False):
    """``(overlap, one_rdm, two_rdm)``; ``mmap=True`` maps the (possibly multi-GB) two_rdm file."""
    mode = "r" if mmap else None
    overlap = np.load(os.path.join(directory, _name("overlap", index)))
    one_rdm = np.load(os.path.join(directory, _name("one_rdm", index)))
    two_rdm = np.load(os.path.join(directory, _name("two_rdm", index)), mmap_mode=mode)
    if two_rdm.ndim not in (2, 3, 5, 6):
        raise AssertionError("two_RDM must have 2, 3, 5 or 6 dimensions")
    return overlap, one_rdm, two_rdm


def load_pair_directories(root, ntrain, nao, pattern="MPS_cross_{}_{}"):
    """Assemble the stack from per-pair directories exactly as 04_Zundel_continuation_MD.py:99-128:
    ``overlap`` and ``one_rdm`` mirrored from the lower triangle (``one_rdm[b, a] = one_rdm[a, b]``,
    no transposition of the orbital indices), ``two_rdm`` as ``(N(N+1)/2, n^2(n^2+1)/2)`` rows in
    ``np.tril_indices(N)`` order."""
    ia, ib = np.tril_indices(ntrain)
    overlap = np.zeros((ntrain, ntrain))
    one_rdm = np.zeros((ntrain, ntrain, nao, nao))
    ncomp = nao * nao * (nao * nao + 1) // 2
    two_rdm = np.zeros((len(ia), ncomp))
    for k, (a, b) in enumerate(zip(ia, ib)):
        d = os.path.join(root, pattern.format(a, b))
        overlap[a, b] = overlap[b, a] = np.load(os.path.join(d, "ovlp.npy"))
        blk = np.load(os.path.join(d, "one_rdm.npy"))
        one_rdm[a, b] = blk
        one_rdm[b, a] = blk
        two_rdm[k] = np.load(os.path.join(d, "two_rdm.npy")).reshape(-1)
    return overlap, one_rdm, two_rdm


def save_pair_directories(root, overlap, one_rdm, two_rdm_tril_exch, pattern="MPS_cross_{}_{}"):
    """Inverse of :func:`load_pair_directories` (``two_rdm`` in the ``(N(N+1)/2, n^2(n^2+1)/2)`` layout)."""
    ntrain = overlap.shape[0]
    ia, ib = np.tril_indices(ntrain)
    for k, (a, b) in enumerate(zip(ia, ib)):
        d = os.path.join(root, pattern.format(a, b))
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, "ovlp.npy"), overlap[a, b])
        np.save(os.path.join(d, "one_rdm.npy"), one_rdm[a, b])
        np.save(os.path.join(d, "two_rdm.npy"), two_rdm_tril_exch[k])
