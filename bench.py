#!/usr/bin/env python
"""Benchmark of the EVCont FCI hot path on B200 (contract: see DESIGN.md, "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU path (oracle port)

Workload (BASELINE.json configs[1]): H10 chain / STO-6G sized FCI eigenvector
continuation -- norb 10, nelec (5,5), 63 504 determinants, 20 training states,
10 atoms.  Inputs are synthetic: random-normalised CI
vectors, seeded random AO arrays with the symmetries of real integrals.

One "step" = one pass of the prediction path (energy + nuclear gradient,
get_energy_with_grad of the reference) over one batch of G geometries, so
``value`` = predicted MD steps/s = G*K*ranks / (max over ranks of the device time).
The stack the steps predict from is built in the same run from the CI vectors by
the trans-RDM kernel (all 210 pairs a>=b), timed separately and reported as
``trans_rdm12`` (pairs/s + FP64 tensor roofline).

Multi-GPU (torchrun, one rank per GPU): the pair list of the stack build is sharded
over the ranks and assembled with one NCCL all_gather; prediction is weak scaling
(every rank predicts its own G geometries from the replicated stack, no per-step
collective).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NORB, NELEC, NATM = 10, (5, 5), 10
NA = NB = 252
WORKLOAD = "H10 chain STO-6G FCI EVCont: norb=10 nelec=(5,5) ndet=63504, 20 training states, E+F"
METRIC = "EVCont MD steps/s (E+F)"
UNIT = "steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="geometries per step per GPU")
    ap.add_argument("--ntrain", type=int, default=20)
    ap.add_argument("--chunk", type=int, default=512,
                    help="geometries per pipelined chunk (e2e); 128 / 256 / 512 / 1024: 266 k / 281 k / 286 k / 287 k steps/s")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of each cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dgemm-peak", action="store_true")
    ap.add_argument("--no-coords", action="store_true", help="skip the from-coordinates leg (K9)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the single-GPU diagnostics (latency, stack streaming, other trans-RDM sizes)")
    ap.add_argument("--no-settle", action="store_true",
                    help="skip the 0.4 s clock-settling loop (for runs under ncu)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------
def civecs(ntrain):
    out = []
    for k in range(ntrain):
        c = np.random.default_rng(1000 + k).standard_normal((NA, NB))
        c = c + c.T  # spin0-like
        # exactly rounded norm (no BLAS: its summation order follows the host's thread count and SIMD width, and
        # the stack checksum below is compared across runs with different OMP settings -- torchrun sets 1 thread)
        out.append(c / math.sqrt(math.fsum((c * c).ravel().tolist())))
    return np.stack(out)


def host_ao_batch(G, seed0, pin):
    """G seeded synthetic geometries as one pinned-host HostAO batch in the packed two-electron layouts
    (erip / eri_ip1p: what the device integral kernel emits; built here with pack_ao_host)."""
    import torch
    from evcont_b200.engine import HostAO, pack_ao_host
    from evcont_b200.mol import ao_bundle, synthetic_mol
    n, natm = NORB, NATM
    b0 = ao_bundle(synthetic_mol(n, natm, seed=seed0))
    hao = HostAO(G, n, natm, b0["aoslices"], pin=pin, packed=True)
    shapes = hao.shapes
    host = {k: getattr(hao, k) for k in hao.fields}
    # distinct geometries are cheap to draw but slow to draw by the thousand in
    # Python: draw up to 64 and tile them with a per-geometry scale so no two are equal
    base = []
    for k in range(min(G, 64)):
        b = dict(ao_bundle(synthetic_mol(n, natm, seed=seed0 + k)))
        b["erip"], b["eri_ip1p"] = pack_ao_host(np.asarray(b["eri"]), np.asarray(b["eri_ip1"]))
        base.append(b)
    for g in range(G):
        b = base[g % len(base)]
        f = 1.0 + 1.0e-3 * (g // len(base))
        for k in hao.fields:
            src = np.asarray(b[k], dtype=np.float64)
            if k in ("hcore", "hcore_deriv", "e_nuc", "grad_nuc"):
                src = src * f
            host[k][g].copy_(torch.from_numpy(np.ascontiguousarray(src).reshape(shapes[k][1:])))
    return hao, host, base[0]["aoslices"]


# ---------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------
class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    @staticmethod
    def _epoch(stamp):
        import datetime
        try:
            return datetime.datetime.strptime(stamp, "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, window=None):
        """Summary of the samples taken inside ``window`` = (t0, t1) wall-clock seconds
        (all samples if none fall inside)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=5)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r in self.rows if len(r) >= 7]
        if window is not None:
            inside = [r for r in rows if (self._epoch(r[0]) or 0) >= window[0] - 0.02
                      and (self._epoch(r[0]) or 0) <= window[1] + 0.02]
            if inside:
                rows = inside
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's numpy path; PySCF is not in the image)
# ---------------------------------------------------------------------------------
def cpu_predict_rate(stack_np, seconds, max_geoms=100000):
    from evcont_b200.mol import synthetic_mol
    from oracle import gradients as og
    S, one, two = stack_np
    mols = [synthetic_mol(NORB, NATM, seed=7000 + k) for k in range(16)]
    og.get_energy_with_grad(mols[0], one, two, S)  # warm-up
    t0 = time.perf_counter()
    done = 0
    while done < max_geoms:
        og.get_energy_with_grad(mols[done % len(mols)], one, two, S)
        done += 1
        if time.perf_counter() - t0 > seconds:
            break
    return done / (time.perf_counter() - t0), done


def cpu_trdm_rate(vecs, seconds):
    from oracle import trans_rdm as otr
    t0 = time.perf_counter()
    done = 0
    N = len(vecs)
    while True:
        a, b = done % N, (done * 7 + 1) % N
        otr.trans_rdm12(vecs[a], vecs[b], NORB, NELEC)
        done += 1
        if time.perf_counter() - t0 > seconds:
            break
    return done / (time.perf_counter() - t0), done


def synthetic_stack_np(ntrain, seed=9):
    """Random stack for the reference arm (the CPU oracle needs ~80 s to build the real
    one; the cost of a prediction step does not depend on the stack's values)."""
    rng = np.random.default_rng(seed)
    n = NORB
    b = rng.standard_normal((ntrain, ntrain))
    S = np.eye(ntrain) + 0.01 * (b + b.T)
    one = rng.standard_normal((ntrain, ntrain, n, n))
    one = one + one.transpose(1, 0, 2, 3)
    two = rng.standard_normal((ntrain, ntrain, n * n, n * n)) / n
    two = two + two.transpose(1, 0, 2, 3)
    two = two + two.transpose(0, 1, 3, 2)
    return S, one, two.reshape((ntrain, ntrain) + (n,) * 4)


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        nt = [p.get("num_threads", 1) for p in threadpool_info()]
        return max(nt) if nt else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:  # torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every host core
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    stack = synthetic_stack_np(args.ntrain)
    per_step_s = max(1.0, min(30.0, 120.0 / max(1, args.steps + args.warmup)))
    rates, total = [], 0
    for i in range(args.warmup + args.steps):
        r, done = cpu_predict_rate(stack, per_step_s)
        if i >= args.warmup:
            rates.append(r)
            total += done
    value = float(np.mean(rates))
    cores = host_threads()
    sample = (f"{args.steps} timed samples of ~{per_step_s:.0f} s ({total} geometries in all) of "
              "oracle.gradients.get_energy_with_grad (numpy port of evcont get_energy_with_grad), "
              "one geometry per call, BLAS threads = all host cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "ntrain": args.ntrain, "layout": "full (N,N,n,n,n,n)",
                   "geometries_per_step": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------
STAGE_KERNEL = {"loewdin": "loewdin_reg_kernel", "ao2oao": "ao2oao_pipe_kernel", "subspace_H": "dgemm_kernel (NT)",
                "geneig": "geneig_reg_kernel", "predict_rdm": "dgemm_kernel (NN)", "grad": "grad_pipe_kernel",
                "grad_stream": "grad_stream_kernel"}


def stage_work(n, natm, ntrain, G, pitch):
    """Algorithmic work of each stage of the packed prediction step for a batch of G
    geometries: flops on the FP64 tensor cores for the GEMM-shaped
    stages, bytes that have to cross HBM for the streaming ones.  The two-electron inputs
    are the packed arrays erip [np][pitch] and eri_ip1p [3][n][n][np]."""
    n2 = n * n
    npair = n * (n + 1) // 2
    L8 = n2 + npair * (npair + 1) // 2
    P = ntrain * (ntrain + 1) // 2
    stack_bytes = 8 * P * L8
    return {
        "loewdin": dict(bound="latency", flops=G * 30 * n ** 3, bytes=G * 8 * (4 * n2 + n)),
        # T = ERIp Q (np^3 MACs) + lower triangle of Q^T T; reads erip, writes T and hvec
        "ao2oao": dict(bound="tensor", flops=G * (2 * npair ** 3 + npair * npair * (npair + 1)),
                       bytes=G * 8 * (2 * npair * pitch + L8 + 2 * n2)),
        "subspace_H": dict(bound="tensor", flops=G * 2 * P * L8, bytes=stack_bytes + G * 8 * (L8 + P)),
        "geneig": dict(bound="latency", flops=G * 4 * ntrain ** 3, bytes=G * 8 * (P + ntrain)),
        "predict_rdm": dict(bound="tensor", flops=G * 2 * P * L8, bytes=stack_bytes + G * 8 * (L8 + P)),
        # U0 = T Gm, R = Gm P0^T, lower triangle of W = P0 R; reads T and out7, writes W
        "grad": dict(bound="tensor", flops=G * (4 * npair ** 3 + npair * npair * (npair + 1)),
                     bytes=G * 8 * (npair * pitch + L8 + npair * npair + 6 * n2)),
        # eri_ip1p + core-Hamiltonian derivative + int1e_ipovlp read once; the n rows (m, b) of W per AO m
        "grad_stream": dict(bound="hbm", flops=G * 2 * 3 * n2 * npair,
                            bytes=G * 8 * (3 * n2 * npair + natm * 3 * n2 + 3 * n2 + n * n * npair + 2 * n2)),
    }


# dram__bytes_read.sum + dram__bytes_write.sum per launch of 4096 geometries (H10 sizes, N = 20) from the
# `ncu --set full` capture of the kernels of THIS commit: profiles/r02_packed_step_ncu_full.txt
NCU_TRAFFIC_4096 = {"loewdin": 3.336e6, "ao2oao": 223.2e6, "subspace_H": 59.07e6, "geneig": 6.964e6,
                    "predict_rdm": 13.38e6, "grad": 227.5e6, "grad_stream": 737.8e6}


def h10_geometries(G, seed):
    """The test-geometry sampler of the reference's H10 script (examples/
    H10_continuation_3D_replacements.py:130-147): equidistant chain, d0 = 1.78596 bohr, every atom
    displaced by 0.3 bohr in a random direction."""
    rng = np.random.default_rng(seed)
    co = np.zeros((G, NATM, 3))
    co[:, :, 0] = 1.78596 * np.arange(NATM)
    v = rng.standard_normal((G, NATM, 3))
    return co + 0.3 * v / np.linalg.norm(v, axis=2)[..., None]


def screened_quartet_fraction(coords, basis="sto-6g", thr=1.0e-17):
    """Fraction of the primitive quartets of the contracted quartets (ab|cd), (ab) >= (cd), that pass the
    Schwarz screen of K9 (csrc/integrals.cu: the leading rectangle of primitive pairs whose bound times the
    partner's largest bound reaches kScreen = 1e-17), evaluated on the host for a few sample geometries."""
    from evcont_b200.basis import s_basis_tables
    fr = []
    for co in coords:
        natm = co.shape[0]
        t = s_basis_tables(["H"] * natm, basis)
        off = np.concatenate([[0], np.cumsum(t["ao_nprim"])])
        bounds = []
        for a in range(natm):
            for b in range(a + 1):
                ea, wa = t["prim_exp"][off[a]:off[a + 1]], t["prim_wt"][off[a]:off[a + 1]]
                eb, wb = t["prim_exp"][off[b]:off[b + 1]], t["prim_wt"][off[b]:off[b + 1]]
                r2 = float(((co[t["ao_atom"][a]] - co[t["ao_atom"][b]]) ** 2).sum())
                p = ea[:, None] + eb[None, :]
                kc = wa[:, None] * wb[None, :] * np.exp(-ea[:, None] * eb[None, :] / p * r2)
                bounds.append(np.sort((np.abs(kc) * np.sqrt(2.0 * np.pi ** 2.5 / (p * p * np.sqrt(2.0 * p)))).ravel())[::-1])
        kept = tot = 0
        for i, bi in enumerate(bounds):
            for bk in bounds[:i + 1]:
                kept += int((bi * bk[0] >= thr).sum()) * int((bk * bi[0] >= thr).sum())
                tot += bi.size * bk.size
        fr.append(kept / tot)
    return float(np.mean(fr))


def run_coords_leg(args, eng, stack, timed, world, rank, torch):
    """MD steps/s when only the nuclear coordinates come from the host: pinned coordinates in,
    AO integrals on the device (K9, s shells, STO-6G, packed two-electron output), prediction step,
    (E, grad) back in pinned host memory.  Also times the integral kernel alone for its FP64 roofline."""
    from evcont_b200.engine import DeviceAO
    G, K, W, N = args.batch, args.steps, args.warmup, args.ntrain
    sb = eng.sbasis(["H"] * NATM, "sto-6g")
    geoms = h10_geometries(G, 4000 + rank)
    co_h = torch.from_numpy(geoms).pin_memory()
    co_d = eng.empty(G, NATM, 3)
    ao = DeviceAO(eng, G, sb.nao, NATM, sb.aoslices_host, packed=True)
    out = (eng.empty(G), eng.empty(G, NATM, 3), eng.empty(G, N))
    E_h = torch.empty(G, dtype=torch.float64).pin_memory()
    g_h = torch.empty(G, NATM, 3, dtype=torch.float64).pin_memory()

    def step(_i):
        co_d.copy_(co_h, non_blocking=True)
        eng.energy_with_grad_coords(stack, sb, co_d, ao=ao, out=out)
        E_h.copy_(out[0], non_blocking=True)
        g_h.copy_(out[1], non_blocking=True)

    for i in range(W):
        step(i)
    launches0 = eng.launch_count()
    ms = timed(step, K)
    launches = eng.launch_count() - launches0
    ints_ms = timed(lambda i: eng.ao_integrals(sb, co_d, out=ao), K) / K
    # algorithmic work of K9: contracted quartets (ab|cd), (ab) >= (cd), times the primitive quartets of each
    # that pass the Schwarz screen, ~45 FMA per primitive quartet (PQ, T, Boys F0/F1, 9 accumulators)
    npc = NORB * (NORB + 1) // 2
    prim_quartets = npc * (npc + 1) // 2 * 6 ** 4
    kept = screened_quartet_fraction(geoms[:4])
    flops = G * prim_quartets * kept * 90.0
    peak_tf = 64 * 2 * eng.sm_count * 1.965e9 / 1e12  # 64 DFMA lanes/clk/SM (tools/fp64_latency.cu)
    ach = flops / (ints_ms * 1e-3) / 1e12
    return {"value": world * G * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K,
            "h2d_bytes_per_step": int(co_h.numel() * 8), "d2h_bytes_per_step": int((E_h.numel() + g_h.numel()) * 8),
            "basis": "STO-6G (s shells), integrals by evc_ao_integrals_s_packed", "gpu_launches": int(launches),
            "integrals_ms_per_step": ints_ms,
            "integrals_roofline": {"kernel": "sint_kernel", "bound": "fp64 fma pipe", "achieved": ach, "peak": peak_tf,
                                   "unit": "TFLOP/s", "frac": ach / peak_tf,
                                   "work": f"{prim_quartets} primitive quartets per geometry x {kept:.3f} that pass the "
                                           "Schwarz screen (host count on 4 sample geometries) x 90 flop",
                                   "screened_fraction": kept,
                                   "peak_source": "64 DFMA lanes/clk/SM x SMs x 1.965 GHz (tools/fp64_latency.cu)"}}


def run_latency_leg(args, eng, stack, torch):
    """Single-trajectory latency (one geometry per step): ms per predicted MD step for the
    H10 stack with the AO arrays resident and from nuclear coordinates."""
    from evcont_b200.engine import DeviceAO
    sb = eng.sbasis(["H"] * NATM, "sto-6g")
    co = eng.to_device(h10_geometries(1, 77))
    ao = DeviceAO(eng, 1, sb.nao, NATM, sb.aoslices_host, packed=True)
    out = (eng.empty(1), eng.empty(1, NATM, 3), eng.empty(1, args.ntrain))
    eng.ao_integrals(sb, co, out=ao)

    def t(fn, reps=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    res = t(lambda: eng.energy_with_grad(stack, ao, out=out))
    crd = t(lambda: eng.energy_with_grad_coords(stack, sb, co, ao=ao, out=out))
    # the same step replayed as a CUDA graph (what the device MD integrator does)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(eng.device)
    side.wait_stream(torch.cuda.current_stream(eng.device))
    from evcont_b200.engine import Workspace
    ws = Workspace(eng.device)
    with eng.using_workspace(ws):
        eng.energy_with_grad_coords(stack, sb, co, ao=ao, out=out)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                eng.energy_with_grad_coords(stack, sb, co, ao=ao, out=out)
    ws.freeze()
    torch.cuda.current_stream(eng.device).wait_stream(side)
    grf = t(g.replay)
    return {"workload": WORKLOAD, "geometries_per_step": 1, "resident_ms_per_step": res,
            "from_coordinates_ms_per_step": crd, "from_coordinates_cuda_graph_ms_per_step": grf,
            "from_coordinates_steps_per_s": 1000.0 / grf}


def run_stream_legs(args, eng, torch, hbm_peak):
    """The HBM-streaming regime of the stack contractions (north star subsystem 3): one geometry against
    a stack far larger than L2.  H30 / STO-6G sizes (configs[3]: n = 30, N = 20, layout (N,N,n,n,n,n),
    2.6 GB) and Zundel / 6-31G sizes (configs[4]: n = 28, N = 100, layout (N(N+1)/2, n^2(n^2+1)/2), 12.4 GB),
    synthetic stacks.  K5 = evc_subspace_H, K7 = evc_predict_rdm on the reference's own layouts; achieved
    GB/s = stack bytes / time, against the measured HBM copy bandwidth."""
    from evcont_b200.engine import DeviceStack
    out = {}
    for name, n, N, layout in (("h30_sto6g_N20_layout6", 30, 20, 6), ("zundel_631g_N100_layout2", 28, 100, 2)):
        gen = torch.Generator(device=eng.device)
        gen.manual_seed(3)
        n2 = n * n
        L = n2 * n2 if layout == 6 else n2 * (n2 + 1) // 2
        P = N * N if layout == 6 else N * (N + 1) // 2
        two = torch.randn(P, L, generator=gen, dtype=torch.float64, device=eng.device)
        one = torch.randn(N, N, n, n, generator=gen, dtype=torch.float64, device=eng.device)
        b = torch.randn(N, N, generator=gen, dtype=torch.float64, device=eng.device)
        S = torch.eye(N, dtype=torch.float64, device=eng.device) + 0.01 * (b + b.T)
        stack = DeviceStack(S, one, two.reshape((N, N, n, n, n, n) if layout == 6 else (P, L)), engine=eng, norb=n)
        h1 = torch.randn(1, n, n, generator=gen, dtype=torch.float64, device=eng.device)
        h2 = torch.randn(1, n, n, n, n, generator=gen, dtype=torch.float64, device=eng.device)
        cv = torch.randn(1, N, generator=gen, dtype=torch.float64, device=eng.device)
        nbytes = two.numel() * 8 + one.numel() * 8
        leg = {"norb": n, "ntrain": N, "layout": layout, "stack_GB": nbytes / 1e9, "geometries_per_step": 1}
        for key, fn in (("subspace_H", lambda: eng.subspace_H(stack, h1, h2)),
                        ("predict_rdm", lambda: eng.predict_rdm(stack, cv))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / reps
            leg[key] = {"ms": ms, "bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": nbytes / ms / 1e6 / hbm_peak}
        out[name] = leg
        del stack, two, one, h2
        torch.cuda.empty_cache()
    return out


def run_pair_sharded_leg(args, eng, torch, world, rank):
    """e3: ONE trajectory against a Zundel-size stack (n = 28, N = 100, 12.4 GB in
    the exchange-compressed lower-triangular layout) sharded by training pairs over the ranks: per step every rank
    streams its slab twice (K5, K7), the ranks all_gather 5050 H entries and all_reduce one compressed two-body
    density matrix (2.4 MB).  At N = 1 the same code runs on the whole stack (the latency the sharding cuts)."""
    from evcont_b200 import distributed as evd
    from evcont_b200.engine import DeviceAO
    from evcont_b200.mol import ao_bundle, synthetic_mol
    n, N, natm = 28, 100, 7
    n2 = n * n
    L, P = n2 * (n2 + 1) // 2, N * (N + 1) // 2
    lo, hi = evd.shard_range(P, rank, world)
    gen = torch.Generator(device=eng.device)
    gen.manual_seed(11 + rank)
    rows = torch.randn(hi - lo, L, generator=gen, dtype=torch.float64, device=eng.device) / n2
    g0 = torch.Generator(device=eng.device)
    g0.manual_seed(5)
    one = torch.randn(N, N, n, n, generator=g0, dtype=torch.float64, device=eng.device)
    one = one + one.transpose(0, 1)
    b = torch.randn(N, N, generator=g0, dtype=torch.float64, device=eng.device)
    S = torch.eye(N, dtype=torch.float64, device=eng.device) + 0.001 * (b + b.T)
    shard = evd.PairShardedStack(S, one, rows, lo, hi, engine=eng)
    ao = DeviceAO.from_bundles(eng, [ao_bundle(synthetic_mol(n, natm, seed=77))])
    step = (lambda: evd.sharded_energy_with_grad(shard, ao)) if world > 1 else \
           (lambda: evd.sharded_energy_with_grad([shard], ao))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=eng.device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    slab_bytes = rows.numel() * 8
    del shard, rows
    torch.cuda.empty_cache()
    return {"workload": "one trajectory, Zundel / 6-31G sizes: n = 28, N = 100, layout (N(N+1)/2, n^2(n^2+1)/2)",
            "stack_GB": P * L * 8 / 1e9, "slab_GB_per_gpu": slab_bytes / 1e9, "ms_per_step": ms,
            "steps_per_s": 1e3 / ms, "collectives_per_step": "all_gather of %d H entries + all_reduce of %d doubles"
            % (P, L) if world > 1 else "none (one slab)",
            "slab_stream_GBps": 2 * slab_bytes / (ms * 1e-3) / 1e9}


def run_trdm_sizes(args, eng, torch, dgemm_tf):
    """trans_rdm12 pairs/s at the other BASELINE sizes: H6 / STO-6G (configs[0]: 6 orbitals, 400 determinants,
    3 states -> 6 pairs) and H2O / 6-31G (configs[2]: 13 orbitals, 1 656 369 determinants, 4 states -> 10 pairs)."""
    import math
    out = {}
    for name, n, k, nvec, reps in (("h6_sto6g", 6, 3, 3, 20), ("h2o_631g", 13, 5, 4, 3)):
        na = math.comb(n, k)
        vecs = []
        for v in range(nvec):
            c = np.random.default_rng(1000 + v).standard_normal((na, na))
            c = c + c.T
            vecs.append(c / np.linalg.norm(c))
        vd = eng.to_device(np.stack(vecs))
        pairs = [(a, b) for a in range(nvec) for b in range(a + 1)]
        for _ in range(2):
            eng.trans_rdm12_batch(vd, pairs, n, (k, k))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.trans_rdm12_batch(vd, pairs, n, (k, k))
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        ndet = na * na
        alg = len(pairs) * (2.0 * n ** 4 * ndet + 2.0 * n * n * ndet)
        issued = eng.trans_rdm12_issued_flops()
        out[name] = {"norb": n, "ndet": ndet, "pairs": len(pairs), "ms": ms, "pairs_per_s": len(pairs) / ms * 1e3,
                     "roofline": {"bound": "tensor", "achieved": alg / ms / 1e9, "issued_tflops": issued / ms / 1e9,
                                  "peak": dgemm_tf, "unit": "TFLOP/s",
                                  "frac": (alg / ms / 1e9 / dgemm_tf) if dgemm_tf else None,
                                  "issued_frac": (issued / ms / 1e9 / dgemm_tf) if dgemm_tf else None}}
        del vd
    return out


def run_sp_legs(args, eng, timed, world, rank, torch):
    """The same from-coordinates MD step for the s+p configurations of BASELINE.json (configs[2] and
    configs[4]): water / 6-31G (13 orbitals, 10 training states, full layout) and the Zundel cation
    H5O2+ / 6-31G (28 orbitals, 20 training states, layout (N(N+1)/2, n^2(n^2+1)/2)), synthetic stacks,
    geometries = reference geometry + N(0, 0.05 bohr).  Integrals by evc_ao_integrals_sp (K9g)."""
    from evcont_b200.engine import DeviceAO, DeviceStack
    ang = 1.0 / 0.52917721092
    r, th = 0.9572 * ang, np.deg2rad(104.52)
    water = (["O", "H", "H"], np.array([[0, 0, 0], [r * np.sin(th / 2), 0, r * np.cos(th / 2)],
                                        [-r * np.sin(th / 2), 0, r * np.cos(th / 2)]]), 10, 6, 1024)
    zundel = (["O", "O", "H", "H", "H", "H", "H"],
              np.array([[-2.25, 0, 0], [2.25, 0, 0], [0, 0.1, 0], [-2.9, 1.45, 0.3], [-2.9, -1.45, -0.3],
                        [2.9, 0.3, 1.45], [2.9, -0.3, -1.45]], dtype=float), 20, 2, 64)
    out = {}
    K = max(2, min(args.steps, 10))
    for name, (sym, base, N, layout, G) in (("h2o_6-31g", water), ("zundel_6-31g", zundel)):
        sb = eng.sbasis(sym, "6-31g")
        natm, n = len(sym), sb.nao
        g = torch.Generator(device=eng.device)
        g.manual_seed(11)
        n2 = n * n
        L = n2 * n2 if layout == 6 else n2 * (n2 + 1) // 2
        P = N * N if layout == 6 else N * (N + 1) // 2
        two = torch.randn(P, L, generator=g, dtype=torch.float64, device=eng.device)
        one = torch.randn(N, N, n, n, generator=g, dtype=torch.float64, device=eng.device)
        b = torch.randn(N, N, generator=g, dtype=torch.float64, device=eng.device)
        S = torch.eye(N, dtype=torch.float64, device=eng.device) + 0.01 * (b + b.T)
        stack = DeviceStack(S, one, two.reshape((N, N, n, n, n, n) if layout == 6 else (P, L)), engine=eng, norb=n)
        rng = np.random.default_rng(7000 + rank)
        co_h = torch.from_numpy(base[None] + 0.05 * rng.standard_normal((G,) + base.shape)).pin_memory()
        co_d = eng.empty(G, natm, 3)
        ao = DeviceAO(eng, G, n, natm, sb.aoslices_host)
        res = (eng.empty(G), eng.empty(G, natm, 3), eng.empty(G, N))
        E_h = torch.empty(G, dtype=torch.float64).pin_memory()
        g_h = torch.empty(G, natm, 3, dtype=torch.float64).pin_memory()

        def step(_i):
            co_d.copy_(co_h, non_blocking=True)
            eng.energy_with_grad_coords(stack, sb, co_d, ao=ao, out=res)
            E_h.copy_(res[0], non_blocking=True)
            g_h.copy_(res[1], non_blocking=True)

        for i in range(3):
            step(i)
        ms = timed(step, K)
        ints_ms = timed(lambda i: eng.ao_integrals(sb, co_d, out=ao), K) / K
        out[name] = {"value": world * G * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K,
                     "geometries_per_step_per_gpu": G, "norb": n, "natm": natm, "ntrain": N, "layout": layout,
                     "integrals_ms_per_step": ints_ms, "integrals_geometries_per_s": G / (ints_ms * 1e-3),
                     "h2d_bytes_per_step": int(co_h.numel() * 8),
                     "d2h_bytes_per_step": int((E_h.numel() + g_h.numel()) * 8),
                     # K9g has no closed-form flop count (six shell-quartet classes, Hermite recursions of different
                     # depth, primitive screening): its roofline evidence is the measured pipe activity
                     "integrals_roofline": {"kernel": "gclass kernels (K9g, 16 two-electron + 4 one-electron launches)",
                                            "bound": "fp64 fma pipe", "achieved": None, "peak": None, "frac": None,
                                            "fp64_pipe_active_pct_ncu": {"two_electron_classes": [30, 39],
                                                                         "one_electron_classes": [19, 26]},
                                            "source": "profiles/r01e_gclass_ncu_full.txt (Zundel sizes, 64 geometries)"}}
        del stack, two, one, ao
        torch.cuda.empty_cache()
    return out


def measure_dgemm_peak(torch, dev):
    m = 6144
    a = torch.randn(m, m, dtype=torch.float64, device=dev)
    b = torch.randn(m, m, dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * m ** 3 / (best * 1e-3) / 1e12


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from evcont_b200 import distributed as evd
    from evcont_b200.engine import DeviceAO, DeviceStack, get_engine

    eng = get_engine(dev)
    G, K, W, N, n = args.batch, args.steps, args.warmup, args.ntrain, NORB
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    dgemm_tf = None if args.no_dgemm_peak else measure_dgemm_peak(torch, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        """max-over-ranks device time (ms) of ``reps`` calls of fn."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- phase 1: build the stack from the CI vectors (pairs sharded over ranks) -------
    vecs = civecs(N)
    vecs_d = eng.to_device(vecs)
    pairs = evd.tril_pairs(N)
    lo, hi = evd.shard_range(len(pairs), rank, world)
    my_pairs = pairs[lo:hi]

    # kernel-only timing of this rank's share of the pair list
    for _ in range(max(1, W)):
        eng.trans_rdm12_batch(vecs_d, my_pairs, n, NELEC)
    trdm_reps = max(3, K)
    launches0 = eng.launch_count()
    trdm_ms = timed(lambda i: eng.trans_rdm12_batch(vecs_d, my_pairs, n, NELEC), trdm_reps) / trdm_reps
    trdm_launches = (eng.launch_count() - launches0) // trdm_reps
    issued = eng.trans_rdm12_issued_flops()
    ndet = NA * NB
    trdm_alg_flops = len(my_pairs) * (2.0 * n ** 4 * ndet + 2.0 * n * n * ndet)
    # full build -> the stack used below: the same row kernel + placement kernel at every N, plus ONE all_gather
    # of the slabs for N > 1; timed identically
    def full_build(_i=0):
        if world > 1:
            return evd.build_stack_sharded(vecs_d, n, NELEC, group=None, device=dev)
        return evd.build_stack_single(vecs_d, n, NELEC, device=dev)

    S_d, one_d, two_d = full_build()
    build_ms = timed(full_build, 3) / 3
    # checksum of the assembled stack: must be identical at every N (bit-reproducible across GPU counts)
    import hashlib
    stack_sha = hashlib.sha256(two_d.cpu().numpy().tobytes() + one_d.cpu().numpy().tobytes()
                               + S_d.cpu().numpy().tobytes()).hexdigest()[:16]
    stack = DeviceStack(S_d, one_d, two_d, engine=eng, norb=n)
    pairs_per_s = len(pairs) / (build_ms * 1e-3)

    # ---- phase 2: prediction steps ------------------------------------------------------
    hao, host, aoslices = host_ao_batch(G, seed0=100 + 1000 * rank, pin=True)
    ao = DeviceAO(eng, G, n, NATM, aoslices, packed=True)
    for k in ao.fields:
        getattr(ao, k).copy_(host[k], non_blocking=True)
    torch.cuda.synchronize()
    E, grad, cvec = eng.empty(G), eng.empty(G, NATM, 3), eng.empty(G, N)
    out = (E, grad, cvec)
    h2d_bytes = sum(host[k].numel() * 8 for k in hao.fields)
    d2h_bytes = (hao.E.numel() + hao.grad.numel()) * 8

    def step_resident(_i):
        eng.energy_with_grad(stack, ao, out=out)

    def step_e2e(_i):
        # the public host-buffer call: pinned host AO arrays in, (E, grad) back on the host
        eng.energy_with_grad_host(stack, hao, chunk=args.chunk, sync=False)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_warm = time.time()
    for i in range(W):
        step_resident(i)
    torch.cuda.synchronize()
    while not args.no_settle and time.time() - t_warm < 0.4:  # let nvidia-smi come up and the clocks settle under load
        step_resident(0)
        torch.cuda.synchronize()
    launches0 = eng.launch_count()
    t0 = time.time()
    total_ms = timed(step_resident, K)
    step_launches = eng.launch_count() - launches0
    # per-stage CUDA events in a second pass of the same K steps: the event records between the stages cost
    # ~3 % of the step, so the headline pass runs without them
    eng.stage_timing(True)
    staged_ms = timed(step_resident, K)
    stage_ms, calls = eng.stage_times()
    eng.stage_timing(False)
    for i in range(W):
        step_e2e(i)
    e2e_ms = timed(step_e2e, K)
    # ---- the same step from nuclear coordinates only (K9: AO integrals on the device) ----
    coords_leg, sp_legs, latency, stream_legs, trdm_sizes = None, None, None, None, None
    if not args.no_coords:
        coords_leg = run_coords_leg(args, eng, stack, timed, world, rank, torch)
        try:   # extra legs (other BASELINE configs): never let them take the headline line down
            sp_legs = run_sp_legs(args, eng, timed, world, rank, torch)
        except Exception as exc:  # noqa: BLE001 -- reported in the JSON line
            if world > 1:
                raise          # ranks must stay in step inside the collectives of timed()
            sp_legs = {"error": f"{type(exc).__name__}: {exc}"}
    pair_sharded = None
    if not args.no_extra:
        try:
            pair_sharded = run_pair_sharded_leg(args, eng, torch, world, rank)
        except Exception as exc:  # noqa: BLE001 -- reported in the JSON line
            if world > 1:
                raise
            pair_sharded = {"error": f"{type(exc).__name__}: {exc}"}
    t1 = time.time()
    if world == 1 and not args.no_extra:   # single-GPU diagnostics (no collectives inside)
        for name, fn in (("latency", lambda: run_latency_leg(args, eng, stack, torch)),
                         ("stream", lambda: run_stream_legs(args, eng, torch, hbm_peak)),
                         ("trdm", lambda: run_trdm_sizes(args, eng, torch, dgemm_tf))):
            try:
                r = fn()
            except Exception as exc:  # noqa: BLE001 -- reported in the JSON line
                r = {"error": f"{type(exc).__name__}: {exc}"}
            if name == "latency":
                latency = r
            elif name == "stream":
                stream_legs = r
            else:
                trdm_sizes = r
    clocks = sampler.stop(window=(t0, t1)) if rank == 0 else None
    value = world * G * K / (total_ms * 1e-3)
    e2e_value = world * G * K / (e2e_ms * 1e-3)

    # ---- roofline of every stage, the dominant one as `roofline`; the trans-RDM kernel ----------------
    pitch = int(eng.lib.evc_erip_pitch(n))
    work = stage_work(n, NATM, N, G, pitch)
    per_call = {k: v / max(1, calls) for k, v in stage_ms.items()}
    fp64_src = "cuBLAS DGEMM 6144^3 measured in this run (no FP64 figure in MEASURED_PEAKS.json)"
    stages = []
    for name in eng.STAGES:
        wk, ms_k = work[name], per_call[name]
        entry = {"stage": name, "kernel": STAGE_KERNEL[name], "ms_per_step": ms_k, "bound": wk["bound"],
                 "share_of_step": ms_k / max(1e-12, sum(per_call.values()))}
        if wk["bound"] == "tensor" and dgemm_tf:
            ach = wk["flops"] / (ms_k * 1e-3) / 1e12
            entry.update(achieved=ach, peak=dgemm_tf, unit="TFLOP/s", frac=ach / dgemm_tf, peak_source=fp64_src)
        elif wk["bound"] == "hbm":
            ach = wk["bytes"] / (ms_k * 1e-3) / 1e9
            entry.update(achieved=ach, peak=hbm_peak, unit="GB/s", frac=ach / hbm_peak, peak_source=hbm_src)
        else:  # latency-bound stages (small dense eigenproblems): no roofline, the time is the figure of merit
            entry.update(achieved=None, peak=None, unit=None, frac=None,
                         note="latency / instruction bound: one small dense eigenproblem per geometry")
        entry["algorithmic_bytes"] = wk["bytes"]
        entry["algorithmic_flops"] = wk["flops"]
        tr = NCU_TRAFFIC_4096.get(name) if (n == 10 and N == 20) else None
        entry["traffic"] = tr * G / 4096.0 if tr else None
        stages.append(entry)
    rated = [e for e in stages if e["frac"] is not None]
    top = max(rated, key=lambda e: e["ms_per_step"])
    roofline = {"kernel": top["kernel"], "stage": top["stage"], "bound": top["bound"], "achieved": top["achieved"],
                "peak": top["peak"], "unit": top["unit"], "frac": top["frac"], "traffic": top["traffic"],
                "traffic_source": ("profiles/r02_packed_step_ncu_full.txt (dram bytes per launch of 4096 geometries, "
                                   "kernels of this commit)") if top["traffic"] else None,
                "algorithmic_bytes": top["algorithmic_bytes"], "peak_source": top["peak_source"],
                "share_of_step": top["share_of_step"],
                "stage_ms_per_step": per_call,
                "stage_pass": {"ms_per_step": staged_ms / K,
                               "note": "stage events recorded in a second pass of the same K steps on the same "
                                       "stream; the headline pass (value) runs without them"}}
    trdm = {"pairs_per_s": pairs_per_s, "pairs": len(pairs), "build_ms": build_ms, "stack_sha256_16": stack_sha,
            "build": "row kernel (evc_trans_rdm12_batch_strided) + one all_gather of the slabs (N > 1) + placement "
                     "kernel (evc_stack_scatter_rows)",
            "kernel_ms_this_rank": trdm_ms, "launches": int(trdm_launches),
            "roofline": {"bound": "tensor", "achieved": trdm_alg_flops / (trdm_ms * 1e-3) / 1e12,
                         "issued_tflops": issued / (trdm_ms * 1e-3) / 1e12,
                         "peak": dgemm_tf, "unit": "TFLOP/s",
                         "frac": (trdm_alg_flops / (trdm_ms * 1e-3) / 1e12 / dgemm_tf) if dgemm_tf else None,
                         "issued_frac": (issued / (trdm_ms * 1e-3) / 1e12 / dgemm_tf) if dgemm_tf else None,
                         "kernel": "trdm_pipe_kernel (norb = 10, 11) / trdm_fused_kernel + trdm_finalize_kernel",
                         # dram__bytes_read + dram__bytes_write of one trdm_pipe_kernel launch of 210 H10 pairs
                         # (10.6 + 52.0 MB: the CI vectors once, the split-K partials of the alpha slices)
                         "traffic": 62.6e6 * len(my_pairs) / 210.0 if (n == 10 and NELEC == (5, 5)) else None,
                         "traffic_source": "profiles/r02_trdm_pipe_ncu_full.txt"}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        stack_np = (S_d.cpu().numpy(), one_d.cpu().numpy(), two_d.cpu().numpy())
        rate, done = cpu_predict_rate(stack_np, args.cpu_seconds)
        trate, tdone = cpu_trdm_rate(vecs, args.cpu_seconds)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": host_threads(), "kind": "port",
                        "sample": f"{done} geometries of the same workload through oracle.gradients."
                                  "get_energy_with_grad (numpy port of the reference, one geometry per call)",
                        "trans_rdm12_pairs_per_s": trate,
                        "trans_rdm12_sample": f"{tdone} pairs through oracle.trans_rdm.trans_rdm12 (numpy)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "ntrain": N, "layout": "full (N,N,n,n,n,n), packed once on the 8-fold "
                   "integral symmetry (evc_stack_pack8)",
                   "ao_arrays": "two-electron arrays in the packed layouts erip / eri_ip1p (what the device integral "
                                "kernel emits; evc_ao_pack8 / pack_ao_host make them from full tensors)",
                   "geometries_per_step_per_gpu": G,
                   "l2_policy": f"inputs larger than L2: {h2d_bytes / 1e6:.0f} MB of AO arrays per step "
                                "(stack stays L2-resident as in production)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": e2e_ms / K},
        "gpu_launches": int(step_launches),
        "roofline": roofline, "stages": stages, "trans_rdm12": trdm, "cpu_baseline": cpu_baseline, "clocks": clocks,
        "fp64_dgemm_tflops": dgemm_tf,
        "from_coordinates": coords_leg,
        "from_coordinates_sp": sp_legs,
        "latency": latency,
        "stack_streaming": stream_legs,
        "pair_sharded_stack": pair_sharded,
        "trans_rdm12_sizes": trdm_sizes,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
