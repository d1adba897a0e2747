"""K6 (lowest root of H c = E S c): register-resident kernel (csrc/geneig_reg.cu) against numpy, and timing against
the shared-memory kernels (EVC_GENEIG_REG=0 in the environment)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from evcont_b200.engine import get_engine

eng = get_engine()
rng = np.random.default_rng(7)


def problem(G, N, scale=1.0):
    b = rng.standard_normal((N, N))
    S = np.eye(N) + 0.05 * (b + b.T)
    H = rng.standard_normal((G, N, N)) * scale
    H = H + H.transpose(0, 2, 1)
    return H, S


def ref(H, S):
    L = np.linalg.cholesky(S)
    Li = np.linalg.inv(L)
    A = Li @ H @ Li.T
    w, V = np.linalg.eigh(A)
    return w[:, 0], (Li.T @ V[:, :, 0][..., None])[..., 0]


def check(N, G):
    H, S = problem(G, N)
    linv = eng.geneig_prepare(eng.to_device(S))
    E, C = eng.geneig(eng.to_device(H), linv, 1)
    E, C = E.cpu().numpy()[:, 0], C.cpu().numpy()[:, 0]
    Er, Cr = ref(H, S)
    sgn = np.sign(np.einsum('gi,ij,gj->g', C, S, Cr))
    res = np.abs(np.einsum('gij,gj->gi', H, C) - E[:, None] * (C @ S)).max()
    return np.abs(E - Er).max(), np.abs(C * sgn[:, None] - Cr).max(), res, np.abs(np.einsum('gi,ij,gj->g', C, S, C) - 1).max()


def timeit(N, G, reps=30):
    H, S = problem(G, N)
    linv = eng.geneig_prepare(eng.to_device(S))
    Hd = eng.to_device(H)
    for _ in range(5):
        eng.geneig(Hd, linv, 1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.geneig(Hd, linv, 1)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


if (sys.argv[1] if len(sys.argv) > 1 else 'check') == 'check':
    for N in range(2, 25):
        for G in (1, 3, 500):
            print('N', N, 'G', G, 'dE %.1e dC %.1e resid %.1e norm %.1e' % check(N, G))
else:
    for N in (10, 20, 24):
        for G in (1, 128, 1024, 4096):
            print('reg', os.environ.get('EVC_GENEIG_REG', '1'), 'N', N, 'G', G, 'ms %.4f' % timeit(N, G))
