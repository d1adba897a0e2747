#!/usr/bin/env python
"""Phase clocks of the per-geometry kernels (needs a library built with
EXTRA_NVFLAGS=-DEVC_PHASE_TIMING).  Development aid."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from evcont_b200.engine import DeviceAO, DeviceStack, get_engine
from evcont_b200.mol import ao_bundle, synthetic_mol
from tools.perf_probe import synthetic_stack_dev

n, natm, N = 10, 10, 20
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = get_engine()
S, one, two = synthetic_stack_dev(eng, torch, n, N, 6)
stack = DeviceStack(S, one, two, engine=eng, norb=n)
base = [ao_bundle(synthetic_mol(n, natm, seed=50 + k)) for k in range(8)]
ao = DeviceAO.from_bundles(eng, [base[g % 8] for g in range(G)])
for _ in range(3):
    eng.energy_with_grad(stack, ao)
torch.cuda.synchronize()
buf = (C.c_longlong * 96)()
eng.lib.evc_debug_phase_clocks.argtypes = [C.c_void_p]
eng.lib.evc_debug_phase_clocks(buf)
a = np.array(list(buf)).reshape(4, 24)
for slot in range(4):
    t = a[slot]
    marks = {int(i): int(t[i] - t[0]) for i in range(12, 24) if t[i]}
    t = t.copy(); t[12:] = 0
    nz = np.nonzero(t)[0]
    cum = {int(i): int(t[i] - t[0]) for i in nz}
    print("slot", slot, "cumulative:", cum, "marks:", marks)
    if len(nz) < 2:
        continue
    print("slot", slot, "phases (cycles):", {int(nz[i + 1]): int(t[nz[i + 1]] - t[nz[i]]) for i in range(len(nz) - 1)},
          "total", int(t[nz[-1]] - t[nz[0]]))
