#!/usr/bin/env python
"""clock64 stamps of CTA 0's three roles in the persistent K4p / K8a kernels (csrc/packed_pipe.cu,
evc_debug_pipe_clocks).  Development aid:  python tools/pipe_clock_probe.py [G]"""
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
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = get_engine()
S, one, two = synthetic_stack_dev(eng, torch, n, N, 6)
stack = DeviceStack(S, one, two, engine=eng, norb=n)
base = [ao_bundle(synthetic_mol(n, natm, seed=50 + k)) for k in range(8)]
ao = DeviceAO.from_bundles(eng, [base[g % 8] for g in range(G)]).to_packed()
for _ in range(3):
    eng.energy_with_grad(stack, ao)
torch.cuda.synchronize()
oa, og = C.c_int(), C.c_int()
eng.lib.evc_debug_pipe_occupancy(C.byref(oa), C.byref(og))
print("resident CTAs per SM: K4p", oa.value, "K8a", og.value)
fn = eng.lib.evc_debug_pipe_clocks
fn.argtypes = [C.c_int, C.c_void_p]
fn(1, None)
eng.energy_with_grad(stack, ao)
torch.cuda.synchronize()
buf = (C.c_longlong * (2 * 4 * 16 * 8))()
fn(0, buf)
a = np.array(list(buf), dtype=np.int64).reshape(2, 4, 16, 8)
names = {0: {0: ("MMA", ["top", "ready", "M1acc", "sync", "Tstored", "sync2", "M2acc", "hv"]),
             1: ("FRONT", ["top", "free", "smalls", "Qbuilt", "h1", "loadwait", "arrive", "-"])},
         1: {0: ("MMA", ["m1top", "m1ready", "u0arr", "m23top", "p0ok", "Rstored", "Wdone", "M3acc"]),
             1: ("FRONT", ["top", "free", "Gm", "sync", "chain1", "loadwait", "arrive", "-"]),
             2: ("MID", ["top", "u0ok", "Ydone", "P0arr", "Zdone", "free", "-", "-"]),
             3: ("MMA+", ["M1acc", "M1sync", "M2acc", "M2sync", "Rst", "-", "-", "-"])}}
for kern in (0, 1):
    t0 = a[kern][a[kern] > 0].min() if (a[kern] > 0).any() else 0
    print("kernel", "K4p" if kern == 0 else "K8a", "(cycles since the first stamp of CTA 0)")
    for role, (rname, evs) in names[kern].items():
        for it in range(12):
            row = a[kern, role, it]
            if not row.any():
                continue
            print(f"  {rname:5s} it={it:2d} " + " ".join(f"{e}={int(v - t0) if v else -1}" for e, v in zip(evs, row) if e != "-"))

cb = (C.c_longlong * (2 * 512 * 3))()
eng.lib.evc_debug_pipe_ctas.argtypes = [C.c_void_p]
eng.lib.evc_debug_pipe_ctas(cb)
ct = np.array(list(cb), dtype=np.int64).reshape(2, 512, 3)[0]
ct = ct[ct[:, 1] > 0]
t0 = ct[:, 1].min()
print("K4p CTAs:", len(ct), "span of the kernel (ns):", int(ct[:, 2].max() - t0))
print("  start offsets (ns): min %d median %d max %d" % (0, int(np.median(ct[:, 1] - t0)), int((ct[:, 1] - t0).max())))
dur = ct[:, 2] - ct[:, 1]
print("  CTA lifetimes (ns): min %d median %d max %d" % (int(dur.min()), int(np.median(dur)), int(dur.max())))
print("  end offsets (ns): min %d median %d max %d" % (int((ct[:, 2] - t0).min()), int(np.median(ct[:, 2] - t0)), int((ct[:, 2] - t0).max())))
per_sm = {}
for smid, a, b in ct:
    per_sm.setdefault(int(smid), []).append((int(a - t0), int(b - t0)))
cnt = np.bincount([len(v) for v in per_sm.values()])
print("  SMs by number of CTAs:", {k: int(v) for k, v in enumerate(cnt) if v})
