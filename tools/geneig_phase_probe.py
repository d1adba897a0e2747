#!/usr/bin/env python
"""Phase clocks of the warp-team eigensolver (library built with EXTRA_NVFLAGS=-DEVC_PHASE_TIMING).  Development aid."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from evcont_b200.engine import get_engine

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
eng = get_engine()
rng = np.random.default_rng(0)
b = rng.standard_normal((N, N))
S = np.eye(N) + 0.01 * (b + b.T)
linv = eng.geneig_prepare(eng.to_device(S))
H = rng.standard_normal((G, N, N))
H = eng.to_device(H + H.transpose(0, 2, 1))
for _ in range(3):
    eng.geneig(H, linv)
torch.cuda.synchronize()
buf = (C.c_longlong * 16)()
eng.lib.evc_debug_geneig_clocks.argtypes = [C.c_void_p]
eng.lib.evc_debug_geneig_clocks(buf)
t = np.array(list(buf))[:7]
names = ["products", "householder", "sturm", "inverse iteration", "back-transform", "c = Linv^T y"]
print({n: int(t[i + 1] - t[i]) for i, n in enumerate(names)}, "total", int(t[6] - t[0]))
