#!/usr/bin/env python
"""H10 / STO-6G eigenvector continuation end to end on one B200, without PySCF
(BASELINE.json configs[1]; the reference's scripts/PES_H_chain H10 workflow):

  1. FCI (63 504 determinants, device Davidson) at a few training geometries, in the OAO basis,
  2. transition-RDM stack between all training states (K1+K2),
  3. energies + forces at distorted test geometries against exact FCI,
  4. NVE molecular dynamics of many replicas on the continuation surface (device integrals,
     device velocity Verlet, CUDA graph).

    python examples/h10_evcont_md.py [--ntrain 5] [--replicas 256] [--steps 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def chain(n, d):
    from evcont_b200.mol import MolLite
    return MolLite([("H", (d * k, 0.0, 0.0)) for k in range(n)], basis="sto-6g", unit="Bohr")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--natm", type=int, default=10)
    ap.add_argument("--ntrain", type=int, default=5)
    ap.add_argument("--replicas", type=int, default=256)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--dt", type=float, default=5.0)
    args = ap.parse_args()
    import torch
    from evcont_b200.FCI_EVCont import FCI_EVCont_obj
    from evcont_b200.ab_initio_gradients_loewdin import get_energy_with_grad_coords
    from evcont_b200.electron_integral_utils import get_basis, get_integrals
    from evcont_b200.fci import B200FCISolver
    from evcont_b200.md import DeviceNVE

    n = args.natm
    d0 = 1.78596  # equilibrium spacing used by the reference's H10 scripts (bohr)
    stretches = np.linspace(-0.5, 1.0, args.ntrain) if args.ntrain > 1 else [0.0]
    out = {"natm": n, "ntrain": args.ntrain}

    cont = FCI_EVCont_obj(cibasis="OAO")
    t0 = time.perf_counter()
    for s in stretches:
        cont.append_to_rdms(chain(n, d0 + s))
    torch.cuda.synchronize()
    out["train_s"] = time.perf_counter() - t0
    out["train_energies"] = [float(e) for e in cont.ens]

    # test geometries: every atom displaced by 0.3 bohr in a random direction (H10 script sampler)
    rng = np.random.default_rng(1)
    mol = chain(n, d0)
    G = 8
    v = rng.standard_normal((G, n, 3))
    test = mol.atom_coords()[None] + 0.3 * v / np.linalg.norm(v, axis=2)[..., None]
    E, F = get_energy_with_grad_coords(mol, test, cont.one_rdm, cont.two_rdm, cont.overlap)
    solver = B200FCISolver()
    err = []
    for g in range(G):
        m = mol.copy().set_geom_(test[g])
        h1, h2 = get_integrals(m, get_basis(m, "OAO"))
        e_fci, _ = solver.kernel(h1, h2, n, m.nelec)
        err.append(E[g] - (e_fci + m.energy_nuc()))
    out["test_error_Ha"] = {"max": float(np.max(err)), "min": float(np.min(err)), "mean": float(np.mean(err))}

    B = args.replicas
    x0 = mol.atom_coords()[None] + 0.05 * rng.standard_normal((B, n, 3))
    nve = DeviceNVE(mol, cont.one_rdm, cont.two_rdm, cont.overlap, x0, None, dt=args.dt, max_frames=args.steps + 1)
    nve.run(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nve.run(args.steps - 3)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    _, epot, ekin = nve.frames()
    etot = epot + ekin
    out["md"] = {"replicas": B, "steps": args.steps, "md_steps_per_s": B * (args.steps - 3) / dt,
                 "max_energy_drift_Ha": float(np.abs(etot - etot[0]).max())}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
