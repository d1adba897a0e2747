"""Subspace (eigenvector-continuation) problem and integral utilities -- CPU oracle.

numpy restatement of the reference's numpy code, function by function:

* :func:`get_loewdin_trafo`        <- evcont/electron_integral_utils.py:6-18
* :func:`transform_integrals`      <- evcont/electron_integral_utils.py:21-35
* :func:`compress_exchange` / :func:`restore_exchange`
                                   <- evcont/electron_integral_utils.py:38-88
* :func:`ao_to_oao`                <- evcont/electron_integral_utils.py:122-138 and the
                                      inline copy at ab_initio_gradients_loewdin.py:338-339
                                      (``ao2mo.kernel`` + ``restore(1, ...)`` written out
                                      as the four-index contraction it is)
* :func:`subspace_hamiltonian`     <- evcont/ab_initio_eigenvector_continuation.py:38-71
* :func:`approximate_ground_state` <- ...continuation.py:12-90
* :func:`approximate_multistate`   <- ...continuation.py:93-175

Pinned against the reference itself (imported under a pyscf stub) by
tests/golden/make_golden.py.  Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np
import scipy.linalg


def get_loewdin_trafo(s_ao):
    w, v = np.linalg.eigh(s_ao)
    f = np.zeros_like(w)
    keep = w > 1.0e-15
    f[keep] = 1.0 / np.sqrt(w[keep])
    return (v * f) @ v.conj().T


def transform_integrals(h1, h2, trafo):
    """Rotate with ``trafo[a, i]`` (new index first), batched over leading axes."""
    h1t = np.einsum("ai,...ij,bj->...ab", trafo, h1, trafo)
    t = np.einsum("ai,...ijkl->...ajkl", trafo, h2)
    t = np.einsum("bj,...ajkl->...abkl", trafo, t)
    t = np.einsum("ck,...abkl->...abcl", trafo, t)
    h2t = np.einsum("dl,...abcl->...abcd", trafo, t)
    return h1t, h2t


def ao_to_oao(hcore, eri, x):
    """h1 = X^T h X ; h2[abcd] = sum (ij|kl) X_ia X_jb X_kc X_ld."""
    h1 = x.T @ hcore @ x
    t = np.tensordot(x, eri, axes=(0, 0))                 # a j k l
    t = np.tensordot(x, t, axes=(0, 1))                   # b a k l
    t = np.tensordot(x, t, axes=(0, 2))                   # c b a l
    t = np.tensordot(x, t, axes=(0, 3))                   # d c b a
    return h1, np.ascontiguousarray(t.transpose(3, 2, 1, 0))


def compress_exchange(h2, diag_multiplier=1.0):
    n = h2.shape[0]
    assert h2.shape == (n, n, n, n)
    m = h2.reshape(n * n, n * n).copy()
    m[np.diag_indices(n * n)] *= diag_multiplier
    return m[np.tril_indices(n * n)]


def restore_exchange(h2c, norb):
    m = np.zeros((norb * norb, norb * norb))
    il = np.tril_indices(norb * norb)
    m[il] = h2c
    m[(il[1], il[0])] = h2c
    return m.reshape(norb, norb, norb, norb)


def subspace_hamiltonian(h1, h2, one_rdm, two_rdm, hermitian=True):
    """H for the four two_rdm layouts (ndim 6 / 5 / 3 / 2)."""
    H = np.tensordot(one_rdm, h1, axes=2)
    nd = two_rdm.ndim
    tril = np.tril_indices(H.shape[0])
    if nd == 6:
        H = H + 0.5 * np.tensordot(two_rdm, h2, axes=4)
    elif nd == 3:
        H = H + two_rdm @ compress_exchange(h2, 0.5)
    elif nd in (5, 2):
        if nd == 5:
            two = 0.5 * np.tensordot(two_rdm, h2, axes=4)
        else:
            two = two_rdm @ compress_exchange(h2, 0.5)
        H = H.copy()
        H[tril] += two
        if not hermitian:
            triu = np.triu_indices(H.shape[0])
            H[triu] = H.T.conj()[triu]
    else:
        raise AssertionError("two_rdm must have 2, 3, 5 or 6 dimensions")
    return H


def _solve(H, S, hermitian):
    if hermitian:
        return scipy.linalg.eigh(H, S)
    return scipy.linalg.eig(H, S)


def approximate_ground_state(h1, h2, one_rdm, two_rdm, S, hermitian=True):
    H = subspace_hamiltonian(h1, h2, one_rdm, two_rdm, hermitian)
    vals, vecs = _solve(H, S, hermitian)
    ok = np.abs(vals.imag) < 1.0e-5
    k = np.argmin(vals[ok].real)
    return vals[ok][k].real, vecs[:, ok][:, k].real


def approximate_multistate(h1, h2, one_rdm, two_rdm, S, nroots=1, hermitian=True):
    H = subspace_hamiltonian(h1, h2, one_rdm, two_rdm, hermitian)
    vals, vecs = _solve(H, S, hermitian)
    ok = np.abs(vals.imag) < 1.0e-5
    assert vals[ok].shape[0] >= nroots
    order = np.argsort(vals[ok].real)[:nroots]
    return vals[ok][order].real, vecs[:, ok][:, order].real.T
