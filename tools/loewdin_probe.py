"""Loewdin (K3): register-resident kernel against numpy and against the shared-memory kernel.
EVC_LOEWDIN_REG_MIN=<batch> in the environment selects the smallest batch that takes the register kernel."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from evcont_b200.engine import get_engine

eng = get_engine()
rng = np.random.default_rng(5)


def spd(G, n):
    A = rng.standard_normal((G, n, n)) * 0.3
    S = np.einsum('gij,gkj->gik', A, A) + np.eye(n)[None]
    d = 1.0 / np.sqrt(np.einsum('gii->gi', S))
    return S * d[:, :, None] * d[:, None, :]


def check(n, G):
    S = spd(G, n)
    X, w, V = eng.loewdin(eng.to_device(S))
    X, w, V = X.cpu().numpy(), w.cpu().numpy(), V.cpu().numpy()
    wr, Vr = np.linalg.eigh(S)
    Xr = np.einsum('gik,gk,gjk->gij', Vr, wr ** -0.5, Vr)
    rec = np.einsum('gik,gk,gjk->gij', V, w, V)
    return np.abs(X - Xr).max(), np.abs(w - wr).max(), np.abs(rec - S).max(), np.abs(np.einsum('gki,gkj->gij', V, V) - np.eye(n)).max()


def timeit(n, G, reps=50):
    S = eng.to_device(spd(G, n))
    for _ in range(5):
        eng.loewdin(S)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.loewdin(S)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


mode = sys.argv[1] if len(sys.argv) > 1 else 'check'
if mode == 'check':
    for n in range(2, 17):
        for G in (1, 5, 1000):
            print('n', n, 'G', G, 'dX %.1e dw %.1e rec %.1e orth %.1e' % check(n, G))
else:
    for n in (10, 13, 16):
        for G in (1, 16, 128, 1024, 4096):
            print('reg_min', os.environ.get('EVC_LOEWDIN_REG_MIN', '1'), 'n', n, 'G', G, 'ms %.4f' % timeit(n, G))
