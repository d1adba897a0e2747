#!/usr/bin/env python
"""On-the-fly training of the continuation along its own MD trajectories (evcont/MD_utils.py:128-502,
``converge_EVCont_MD``) for an H6 chain in STO-6G on one B200, without PySCF: every iteration runs one
NVE trajectory on the device with the current training set, re-evaluates all of its frames with the
previous training set (one batched launch), and adds the frame whose OAO Hamiltonian is farthest from
the training Hamiltonians (FCI solve + transform_ci + t-RDM growth on the device) until the largest
energy difference stayed below the threshold twice in a row.

    python examples/h6_active_learning.py [--natm 6] [--steps 100] [--dt 5] [--thresh 1e-4] [--workdir /tmp/h6_al]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--natm", type=int, default=6)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--dt", type=float, default=5.0)
    ap.add_argument("--thresh", type=float, default=1e-4)
    ap.add_argument("--spacing", type=float, default=1.5, help="initial H-H distance in bohr (compressed chain)")
    ap.add_argument("--workdir", default="/tmp/h6_active_learning")
    args = ap.parse_args()
    from evcont_b200.FCI_EVCont import FCI_EVCont_obj
    from evcont_b200.MD_utils import converge_EVCont_MD
    from evcont_b200.mol import MolLite

    os.makedirs(args.workdir, exist_ok=True)
    for f in os.listdir(args.workdir):      # a fresh run: the loop resumes from files it finds
        os.remove(os.path.join(args.workdir, f))
    xs = (np.arange(args.natm) - (args.natm - 1) / 2) * args.spacing
    mol = MolLite([("H", (x, 0.0, 0.0)) for x in xs], basis="sto-6g", unit="Bohr")
    cont = FCI_EVCont_obj()                 # the reference's default: canonical basis (device RHF + transform_ci)
    t0 = time.perf_counter()
    traj = converge_EVCont_MD(cont, mol, steps=args.steps, dt=args.dt, convergence_thresh=args.thresh,
                              workdir=args.workdir)
    wall = time.perf_counter() - t0
    n = len(cont.fcivecs)
    diffs = [float(np.max(np.atleast_1d(np.loadtxt(os.path.join(args.workdir, f"en_diff_{k}.txt"))))) for k in range(n)]
    print(json.dumps({"natm": args.natm, "training_points": n, "wall_s": wall,
                      "trn_times": [int(t) for t in np.loadtxt(os.path.join(args.workdir, "trn_times.txt"))],
                      "max_energy_difference_per_iteration_Ha": diffs,
                      "final_chain_length_bohr": float(traj[-1, -1, 0] - traj[-1, 0, 0])}))


if __name__ == "__main__":
    main()
