#!/usr/bin/env python
"""Per-kernel timing probes on one B200 (development tool, not the benchmark contract).

    python tools/perf_probe.py trdm  --norb 10 --nocc 5 --nvec 20
    python tools/perf_probe.py stack --norb 30 --ntrain 20 --layout 6 --batch 1
    python tools/perf_probe.py step  --norb 10 --natm 10 --ntrain 20 --batch 1024

Every number is CUDA-event time on the launching stream after warm-up; one JSON line
per probe.  Synthetic inputs as in SURVEY.md section 8(d).
"""
import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HBM_PEAK = 6549.8
FP64_PEAK = 36.0


def _peaks():
    global HBM_PEAK
    try:
        HBM_PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass


def timed(torch, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def probe_trdm(args):
    import torch
    from evcont_b200.engine import get_engine
    eng = get_engine()
    n, k = args.norb, args.nocc
    na = math.comb(n, k)
    vecs = []
    for v in range(args.nvec):
        c = np.random.default_rng(1000 + v).standard_normal((na, na))
        c = c + c.T
        vecs.append(c / np.linalg.norm(c))
    vecs_d = eng.to_device(np.stack(vecs))
    pairs = [(a, b) for a in range(args.nvec) for b in range(a + 1)]
    ms = timed(torch, lambda: eng.trans_rdm12_batch(vecs_d, pairs, n, (k, k)), args.reps)
    ndet = na * na
    alg = len(pairs) * (2.0 * n ** 4 * ndet + 2.0 * n * n * ndet)
    issued = eng.trans_rdm12_issued_flops()
    import hashlib
    out = eng.trans_rdm12_batch(vecs_d, pairs, n, (k, k))
    sha = hashlib.sha256(b"".join(np.ascontiguousarray(t.cpu().numpy()).tobytes() for t in out)).hexdigest()[:16]
    print(json.dumps({"probe": "trdm", "norb": n, "nocc": k, "ndet": ndet, "pairs": len(pairs), "ms": ms,
                      "sha256_16": sha, "pipe": os.environ.get("EVC_TRDM_PIPE", "default"),
                      "pairs_per_s": len(pairs) / ms * 1e3, "alg_tflops": alg / ms / 1e9,
                      "issued_tflops": issued / ms / 1e9, "frac_of_36": alg / ms / 1e9 / FP64_PEAK}), flush=True)


def synthetic_stack_dev(eng, torch, n, N, layout, seed=3):
    """Random device stack in ``layout`` (values irrelevant for timing; S is SPD)."""
    g = torch.Generator(device=eng.device)
    g.manual_seed(seed)
    n2 = n * n
    L = n2 * n2 if layout in (6, 5) else n2 * (n2 + 1) // 2
    P = N * N if layout in (6, 3) else N * (N + 1) // 2
    two = torch.randn(P, L, generator=g, dtype=torch.float64, device=eng.device)
    one = torch.randn(N, N, n, n, generator=g, dtype=torch.float64, device=eng.device)
    b = torch.randn(N, N, generator=g, dtype=torch.float64, device=eng.device)
    S = torch.eye(N, dtype=torch.float64, device=eng.device) + 0.01 * (b + b.T)
    shape = {6: (N, N, n, n, n, n), 5: (P, n, n, n, n), 3: (N, N, L), 2: (P, L)}[layout]
    return S, one, two.reshape(shape)


def probe_stack(args):
    import torch
    from evcont_b200.engine import DeviceStack, get_engine
    eng = get_engine()
    n, N, G = args.norb, args.ntrain, args.batch
    S, one, two = synthetic_stack_dev(eng, torch, n, N, args.layout)
    stack = DeviceStack(S, one, two, engine=eng, norb=n)
    h1 = torch.randn(G, n, n, dtype=torch.float64, device=eng.device)
    h2 = torch.randn(G, n, n, n, n, dtype=torch.float64, device=eng.device)
    cv = torch.randn(G, N, dtype=torch.float64, device=eng.device)
    ms_h = timed(torch, lambda: eng.subspace_H(stack, h1, h2), args.reps)
    ms_p = timed(torch, lambda: eng.predict_rdm(stack, cv), args.reps)
    nbytes = stack.two_rdm.numel() * 8
    flops = 2.0 * stack.two_rdm.numel() * G
    out = {"probe": "stack", "norb": n, "ntrain": N, "layout": args.layout, "batch": G,
           "stack_GB": nbytes / 1e9}
    for name, ms in (("subspace_H", ms_h), ("predict_rdm", ms_p)):
        out[name] = {"ms": ms, "GBps": nbytes / ms / 1e6, "hbm_frac": nbytes / ms / 1e6 / HBM_PEAK,
                     "tflops": flops / ms / 1e9}
    print(json.dumps(out), flush=True)


def probe_step(args):
    import torch
    from evcont_b200.engine import DeviceAO, DeviceStack, get_engine
    from evcont_b200.mol import ao_bundle, synthetic_mol
    eng = get_engine()
    n, N, G, natm = args.norb, args.ntrain, args.batch, args.natm
    S, one, two = synthetic_stack_dev(eng, torch, n, N, args.layout)
    stack = DeviceStack(S, one, two, engine=eng, norb=n)
    base = [ao_bundle(synthetic_mol(n, natm, seed=50 + k)) for k in range(min(G, 8))]
    ao = DeviceAO.from_bundles(eng, [base[g % len(base)] for g in range(G)])
    if args.packed:
        ao = ao.to_packed()
    out = (eng.empty(G), eng.empty(G, natm, 3), eng.empty(G, N))
    for _ in range(3):
        eng.energy_with_grad(stack, ao, out=out)
    torch.cuda.synchronize()
    eng.stage_timing(True)
    ms = timed(torch, lambda: eng.energy_with_grad(stack, ao, out=out), args.reps, warm=0)
    stage, calls = eng.stage_times()
    eng.stage_timing(False)
    print(json.dumps({"probe": "step", "packed_ao": bool(args.packed), "pipe": os.environ.get("EVC_PACKED_PIPE", "1"), "norb": n, "natm": natm, "ntrain": N, "layout": args.layout, "batch": G,
                      "ms": ms, "steps_per_s": G / ms * 1e3,
                      "stage_ms": {k: v / max(1, calls) for k, v in stage.items()}}), flush=True)


def probe_ints(args):
    """K9: device AO integrals of an H chain (s shells) + the whole step from coordinates."""
    import torch
    from evcont_b200.engine import DeviceAO, DeviceStack, get_engine
    eng = get_engine()
    n, N, G = args.norb, args.ntrain, args.batch
    rng = np.random.default_rng(1)
    co = np.zeros((G, n, 3))
    co[:, :, 0] = 1.78596 * np.arange(n)
    v = rng.standard_normal((G, n, 3))
    co += 0.3 * v / np.linalg.norm(v, axis=2)[..., None]
    sb = eng.sbasis(["H"] * n, args.basis)
    cd = eng.to_device(co)
    ao = DeviceAO(eng, G, sb.nao, n, sb.aoslices_host)
    ms_i = timed(torch, lambda: eng.ao_integrals(sb, cd, out=ao), args.reps)
    S, one, two = synthetic_stack_dev(eng, torch, sb.nao, N, args.layout)
    stack = DeviceStack(S, one, two, engine=eng, norb=sb.nao)
    out = (eng.empty(G), eng.empty(G, n, 3), eng.empty(G, N))
    ms_s = timed(torch, lambda: eng.energy_with_grad_coords(stack, sb, cd, ao=ao, out=out), args.reps)
    print(json.dumps({"probe": "ints", "natm": n, "nao": sb.nao, "basis": args.basis, "batch": G,
                      "integrals_ms": ms_i, "geoms_per_s": G / ms_i * 1e3,
                      "step_from_coords_ms": ms_s, "steps_per_s": G / ms_s * 1e3}), flush=True)


def probe_ints_sp(args):
    """K9g: device AO integrals of water (--natm 3) or the Zundel cation (--natm 7) in 6-31G + the step."""
    import torch
    from evcont_b200.engine import DeviceAO, DeviceStack, get_engine
    eng = get_engine()
    G, N = args.batch, args.ntrain
    rng = np.random.default_rng(1)
    ang = 1.0 / 0.52917721092
    if args.natm == 3:
        r, th = 0.9572 * ang, np.deg2rad(104.52)
        base = np.array([[0, 0, 0], [r * np.sin(th / 2), 0, r * np.cos(th / 2)], [-r * np.sin(th / 2), 0, r * np.cos(th / 2)]])
        sym = ["O", "H", "H"]
    else:
        base = np.array([[-2.25, 0, 0], [2.25, 0, 0], [0, 0.1, 0], [-2.9, 1.45, 0.3], [-2.9, -1.45, -0.3],
                         [2.9, 0.3, 1.45], [2.9, -0.3, -1.45]], dtype=float)
        sym = ["O", "O", "H", "H", "H", "H", "H"]
    co = base[None] + 0.05 * rng.standard_normal((G,) + base.shape)
    sb = eng.sbasis(sym, "6-31g")
    cd = eng.to_device(co)
    ao = DeviceAO(eng, G, sb.nao, len(sym), sb.aoslices_host)
    ms_i = timed(torch, lambda: eng.ao_integrals(sb, cd, out=ao), args.reps)
    S, one, two = synthetic_stack_dev(eng, torch, sb.nao, N, args.layout)
    stack = DeviceStack(S, one, two, engine=eng, norb=sb.nao)
    out = (eng.empty(G), eng.empty(G, len(sym), 3), eng.empty(G, N))
    ms_s = timed(torch, lambda: eng.energy_with_grad_coords(stack, sb, cd, ao=ao, out=out), args.reps)
    print(json.dumps({"probe": "ints_sp", "natm": len(sym), "nao": sb.nao, "batch": G, "ntrain": N,
                      "integrals_ms": ms_i, "geoms_per_s": G / ms_i * 1e3,
                      "step_from_coords_ms": ms_s, "steps_per_s": G / ms_s * 1e3}), flush=True)


def main():
    _peaks()
    ap = argparse.ArgumentParser()
    ap.add_argument("probe", choices=["trdm", "stack", "step", "ints", "ints_sp"])
    ap.add_argument("--norb", type=int, default=10)
    ap.add_argument("--nocc", type=int, default=5)
    ap.add_argument("--nvec", type=int, default=20)
    ap.add_argument("--natm", type=int, default=10)
    ap.add_argument("--ntrain", type=int, default=20)
    ap.add_argument("--layout", type=int, default=6)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--basis", default="sto-6g")
    ap.add_argument("--packed", action="store_true", help="step: hand the packed AO arrays (erip / eri_ip1p) in")
    args = ap.parse_args()
    {"trdm": probe_trdm, "stack": probe_stack, "step": probe_step, "ints": probe_ints, "ints_sp": probe_ints_sp}[args.probe](args)


if __name__ == "__main__":
    main()
