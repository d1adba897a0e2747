"""K3 (Loewdin transformation): the register-resident batched kernel (csrc/loewdin_reg.cu, n <= 16) and the
shared-memory kernel (csrc/dense.cu, n <= 32) against numpy's eigh -- the call the reference makes
(evcont/electron_integral_utils.py:6-18)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _overlaps(G, n, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((G, n, n)) * 0.3
    S = np.einsum("gij,gkj->gik", A, A) + np.eye(n)[None]
    d = 1.0 / np.sqrt(np.einsum("gii->gi", S))
    return S * d[:, :, None] * d[:, None, :]


def _check(eng, S, tol=2e-13):
    n = S.shape[-1]
    X, w, V = (t.cpu().numpy() for t in eng.loewdin(eng.to_device(S)))
    wr, Vr = np.linalg.eigh(S)
    f = np.where(wr > 1e-15, 1.0 / np.sqrt(np.abs(wr)), 0.0)
    Xr = np.einsum("gik,gk,gjk->gij", Vr, f, Vr)
    assert np.abs(w - wr).max() < tol
    assert np.all(np.diff(w, axis=1) >= 0)                      # ascending, as eigh returns them
    assert np.abs(np.einsum("gki,gkj->gij", V, V) - np.eye(n)).max() < tol
    assert np.abs(np.einsum("gik,gk,gjk->gij", V, w, V) - S).max() < tol
    return np.abs(X - Xr).max()


@pytest.mark.parametrize("n", list(range(1, 17)) + [20, 28])
def test_every_size_against_numpy(n):
    """Every template instance of the register kernel (2..16: two to sixteen matrices per warp, odd sizes with a
    bye in the tournament), batches that leave groups and whole warps partly empty, and the shared-memory kernel
    beyond 16."""
    from evcont_b200.engine import get_engine
    eng = get_engine()
    for G in (1, 2, 7, 333):
        assert _check(eng, _overlaps(G, n, 100 * n + G)) < 1e-12


def test_only_the_lower_triangle_is_read():
    """numpy.linalg.eigh reads the lower triangle: garbage above the diagonal must not change anything."""
    from evcont_b200.engine import get_engine
    eng = get_engine()
    S = _overlaps(50, 10, 3)
    X0 = eng.loewdin(eng.to_device(S))[0].cpu().numpy()
    Sg = S.copy()
    iu = np.triu_indices(10, 1)
    Sg[:, iu[0], iu[1]] = 7.25
    X1 = eng.loewdin(eng.to_device(Sg))[0].cpu().numpy()
    assert np.array_equal(X0, X1)


def test_position_in_the_batch_does_not_change_the_bits():
    """Three 10 x 10 matrices share a warp and converge after different numbers of sweeps: a matrix's result
    must not depend on its neighbours (a converged matrix stops rotating while the others go on)."""
    from evcont_b200.engine import get_engine
    eng = get_engine()
    S = _overlaps(7, 10, 11)
    S[3] = np.eye(10)                       # converged before the first sweep
    S[5] = 0.2 * S[5] + 0.8 * np.eye(10)    # nearly diagonal
    big = np.stack([S[k % 7] for k in range(1000)])
    X, w, V = (t.cpu().numpy() for t in eng.loewdin(eng.to_device(big)))
    for k in range(7, 1000):
        assert np.array_equal(X[k], X[k % 7]) and np.array_equal(w[k], w[k % 7]) and np.array_equal(V[k], V[k % 7])
    X1 = eng.loewdin(eng.to_device(S[2:3]))[0].cpu().numpy()   # alone in its warp
    assert np.array_equal(X1[0], X[2])


def test_degenerate_and_singular_overlaps():
    """Repeated eigenvalues (ties broken by index) and an exactly singular overlap: eigenvalues below the
    reference's 1e-15 cut-off contribute nothing to X."""
    from evcont_b200.engine import get_engine
    eng = get_engine()
    n = 8
    rng = np.random.default_rng(2)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    w = np.array([0.0, 0.5, 0.5, 0.5, 1.0, 1.0, 2.0, 2.5])
    S = (Q * w) @ Q.T
    S = 0.5 * (S + S.T)
    X, wd, V = (t.cpu().numpy() for t in eng.loewdin(eng.to_device(np.stack([S, S]))))
    assert np.abs(wd[0] - w).max() < 1e-14
    f = np.where(wd[0] > 1e-15, 1.0 / np.sqrt(np.abs(wd[0])), 0.0)
    assert np.abs(X[0] - (V[0] * f) @ V[0].T).max() < 1e-13
    keep = w > 1e-15
    Xr = (Q[:, keep] * w[keep] ** -0.5) @ Q[:, keep].T
    if wd[0][0] <= 1e-15:                   # the null direction was resolved below the cut-off
        assert np.abs(X[0] - Xr).max() < 1e-6
