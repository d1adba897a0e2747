"""K6 (lowest root of the subspace problem H c = E S c, scipy.linalg.eigh(H, S) + argmin in the reference,
evcont/ab_initio_eigenvector_continuation.py:75-88): the register-resident kernel (csrc/geneig_reg.cu, N <= 24) and
the shared-memory kernels (csrc/geneig.cu) against numpy, every template instance, and the cases a tridiagonal
eigensolver has to survive (already tridiagonal / diagonal H, degenerate lowest eigenvalue, badly scaled H)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref(H, S):
    L = np.linalg.cholesky(S)
    Li = np.linalg.inv(L)
    A = Li @ H @ Li.T
    w, V = np.linalg.eigh(0.5 * (A + A.transpose(0, 2, 1)))
    return w, np.einsum("ji,gjk->gik", Li, V)        # columns: S-orthonormal eigenvectors


def _solve(H, S):
    from evcont_b200.engine import get_engine
    eng = get_engine()
    linv = eng.geneig_prepare(eng.to_device(S))
    E, C = eng.geneig(eng.to_device(H), linv, 1)
    return E.cpu().numpy()[:, 0], C.cpu().numpy()[:, 0]


def _overlap(N, rng):
    b = rng.standard_normal((N, N))
    return np.eye(N) + 0.05 * (b + b.T)


@pytest.mark.parametrize("N", list(range(1, 26)) + [32, 40])
def test_every_size_against_numpy(N):
    rng = np.random.default_rng(100 + N)
    S = _overlap(N, rng)
    for G in (1, 5, 300):
        H = rng.standard_normal((G, N, N))
        H = H + H.transpose(0, 2, 1)
        E, C = _solve(H, S)
        w, V = _ref(H, S)
        scale = np.abs(w).max()
        assert np.abs(E - w[:, 0]).max() < 1e-12 * scale
        assert np.abs(np.einsum("gi,ij,gj->g", C, S, C) - 1.0).max() < 1e-12          # c^T S c = 1
        assert np.abs(np.einsum("gij,gj->gi", H, C) - E[:, None] * (C @ S)).max() < 1e-11 * scale
        if N > 1:
            gap = (w[:, 1] - w[:, 0]).min()
            if gap > 1e-3 * scale:
                assert np.abs(np.abs(np.einsum("gi,ij,gj->g", C, S, V[:, :, 0])) - 1.0).max() < 1e-9


@pytest.mark.parametrize("N", [2, 5, 12, 20, 24])
def test_structured_matrices(N):
    rng = np.random.default_rng(7)
    S = np.eye(N)
    cases = []
    d = rng.standard_normal(N)
    cases.append(np.diag(d))                                                   # diagonal: every sigma is zero
    T = np.diag(d) + np.diag(rng.standard_normal(N - 1), 1)
    cases.append(np.triu(T) + np.triu(T, 1).T)                                 # already tridiagonal
    Q, _ = np.linalg.qr(rng.standard_normal((N, N)))
    w = np.sort(rng.standard_normal(N)); w[1] = w[0]                           # degenerate lowest eigenvalue
    cases.append((Q * w) @ Q.T)
    big = rng.standard_normal((N, N)); big = (big + big.T) * 1e8                # badly scaled
    cases.append(big)
    small = rng.standard_normal((N, N)); small = (small + small.T) * 1e-9
    cases.append(small)
    H = np.stack([0.5 * (c + c.T) for c in cases])
    E, C = _solve(H, S)
    wref = np.linalg.eigvalsh(H)
    for g in range(len(cases)):
        scale = max(np.abs(wref[g]).max(), 1e-300)
        assert abs(E[g] - wref[g, 0]) < 1e-12 * scale, g
        assert abs(C[g] @ C[g] - 1.0) < 1e-12, g
        assert np.abs(H[g] @ C[g] - E[g] * C[g]).max() < 1e-10 * scale, g      # an eigenvector of the lowest eigenvalue


def test_position_independence():
    """A problem's result must not depend on its position in the batch (one problem per warp, four per CTA)."""
    rng = np.random.default_rng(3)
    N = 20
    S = _overlap(N, rng)
    H = rng.standard_normal((6, N, N)); H = H + H.transpose(0, 2, 1)
    big = np.stack([H[k % 6] for k in range(700)])
    E, C = _solve(big, S)
    for k in range(6, 700):
        assert E[k] == E[k % 6] and np.array_equal(C[k], C[k % 6])
