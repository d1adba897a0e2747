#!/usr/bin/env python
"""H2O / 6-31G eigenvector continuation end to end on one B200, without PySCF
(BASELINE.json configs[2]; the reference's scripts/MD/H2O/md_H2O_6_31G_FCI.py workflow):

  1. at each training geometry: RHF (device Fock builds) -> FCI in the canonical basis (13 orbitals, 10
     electrons, 1 656 369 determinants, device Davidson) -> transform_ci to the OAO basis (device),
  2. transition-RDM stack between all training states (K1+K2, 94.6 GFLOP per pair),
  3. the continuation energy at every training geometry against its FCI energy (exactness),
  4. NVE molecular dynamics of many replicas from coordinates (s+p device integrals K9g, device velocity
     Verlet, CUDA graph), with the energy drift.

    python examples/h2o_evcont_md.py [--ntrain 3] [--replicas 64] [--steps 100]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ANG = 1.0 / 0.52917721092


def water(scale=1.0, angle=104.52):
    """Water with both O-H bonds stretched by ``scale`` (the reference trains along the stretched
    geometries of md_H2O_6_31G_FCI.py:23-37)."""
    from evcont_b200.mol import MolLite
    r, th = 0.9572 * ANG * scale, np.deg2rad(angle)
    return MolLite([("O", (0.0, 0.0, 0.0)), ("H", (r * np.sin(th / 2), 0.0, r * np.cos(th / 2))),
                    ("H", (-r * np.sin(th / 2), 0.0, r * np.cos(th / 2)))], basis="6-31g", unit="Bohr")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ntrain", type=int, default=3)
    ap.add_argument("--replicas", type=int, default=64)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--dt", type=float, default=5.0)
    ap.add_argument("--cibasis", default="canonical", choices=["canonical", "OAO"])
    args = ap.parse_args()
    import torch
    from evcont_b200.FCI_EVCont import FCI_EVCont_obj
    from evcont_b200.ab_initio_eigenvector_continuation import approximate_ground_state_OAO
    from evcont_b200.md import DeviceNVE

    out = {"ntrain": args.ntrain, "cibasis": args.cibasis}
    scales = np.linspace(1.0, 1.3, args.ntrain) if args.ntrain > 1 else [1.0]
    cont = FCI_EVCont_obj(cibasis=args.cibasis)
    t0 = time.perf_counter()
    for s in scales:
        cont.append_to_rdms(water(s))
    torch.cuda.synchronize()
    out["train_s"] = time.perf_counter() - t0
    out["train_energies"] = [float(e) for e in cont.ens]
    out["exactness_Ha"] = [float(approximate_ground_state_OAO(water(s), cont.one_rdm, cont.two_rdm, cont.overlap)[0] - e)
                           for s, e in zip(scales, cont.ens)]

    mol = water(1.2)   # the reference starts its trajectory from the 1.2x stretched molecule
    rng = np.random.default_rng(1)
    B = args.replicas
    x0 = mol.atom_coords()[None] + 0.02 * rng.standard_normal((B, 3, 3))
    nve = DeviceNVE(mol, cont.one_rdm, cont.two_rdm, cont.overlap, x0, None, dt=args.dt, max_frames=args.steps + 1)
    nve.run(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nve.run(args.steps - 3)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    _, epot, ekin = nve.frames()
    etot = epot + ekin
    out["md_steps_per_s"] = B * (args.steps - 3) / dt
    out["energy_drift_Ha"] = float(np.abs(etot - etot[0]).max())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
