"""Keeps the t-RDM stack resident in HBM across calls of the numpy-facing API.

The reference passes ``one_RDM, two_RDM, S`` as numpy arrays on every call
(evcont/MD_utils.py:40-55 calls get_energy_with_grad once per MD step with the
same three arrays).  Uploading 32 MB - 12 GB per step would dominate, so the
device copy is cached per ``two_RDM`` array object and revalidated with a cheap
fingerprint (shapes + strided samples), which catches re-assignment and most
in-place edits; call :func:`invalidate` after editing a stack in place.
"""
import weakref

import numpy as np

from .engine import DeviceStack, get_engine

_cache = {}
_MAX_ENTRIES = 4


def _fingerprint(one_rdm, two_rdm, S):
    flat = two_rdm.reshape(-1)
    step = max(1, flat.size // 4096)
    f1 = one_rdm.reshape(-1)
    s1 = max(1, f1.size // 1024)
    return (two_rdm.shape, one_rdm.shape, float(flat[::step].sum()), float(f1[::s1].sum()),
            float(np.asarray(S).sum()), float(flat[-1]))


def invalidate():
    _cache.clear()


def as_device_stack(one_RDM, two_RDM, S, device=None):
    """Return a :class:`DeviceStack` for the given stack (cached for numpy inputs)."""
    if isinstance(two_RDM, DeviceStack):
        return two_RDM
    one_RDM = np.asarray(one_RDM)
    two_RDM = np.asarray(two_RDM)
    if two_RDM.ndim not in (2, 3, 5, 6):
        raise AssertionError("two_RDM must have 2, 3, 5 or 6 dimensions")
    eng = get_engine(device)
    key = (id(two_RDM), eng.device.index)
    fp = _fingerprint(one_RDM, two_RDM, S)
    hit = _cache.get(key)
    if hit is not None:
        ref, old_fp, stack = hit
        if ref() is two_RDM and old_fp == fp:
            return stack
    stack = DeviceStack(np.asarray(S), one_RDM, two_RDM, engine=eng, norb=one_RDM.shape[-1])
    if len(_cache) >= _MAX_ENTRIES:
        _cache.pop(next(iter(_cache)))
    try:
        ref = weakref.ref(two_RDM)
    except TypeError:  # views of some buffers cannot be weakly referenced
        ref = (lambda obj: (lambda: obj))(two_RDM)
    _cache[key] = (ref, fp, stack)
    return stack
