"""Golden vectors for the s+p integral kernels (K9g) on a molecule with TWO p centres and a hydrogen
(O, O, H in 6-31G: every two-centre class ppps / pppp / psps with p shells on different atoms appears),
from the CPU oracle oracle/integrals_sp.py (the oracle takes ~2 minutes on 20 AOs, too slow for the GPU
test run).  The small arrays are stored whole, the two-electron arrays as 6000 seeded random elements
each (index lists included).

    python tests/golden/make_integrals_sp_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import integrals_sp as osp  # noqa: E402

CO = np.array([[-2.10, 0.30, -0.20], [2.25, -0.15, 0.35], [0.20, 0.90, 0.55]])
SYM = ["O", "O", "H"]


def main():
    ref = osp.ao_arrays(osp.SPBasis([(s, c) for s, c in zip(SYM, CO)], "6-31g"))
    rng = np.random.default_rng(2026)
    n = ref["ovlp"].shape[0]
    idx_eri = rng.integers(0, n, size=(6000, 4))
    idx_ip1 = np.concatenate([rng.integers(0, 3, size=(6000, 1)), rng.integers(0, n, size=(6000, 4))], axis=1)
    out = dict(coords=CO, symbols=np.array(SYM), idx_eri=idx_eri, idx_ip1=idx_ip1,
               eri_vals=ref["eri"][tuple(idx_eri.T)], ip1_vals=ref["eri_ip1"][tuple(idx_ip1.T)],
               eri_sum=ref["eri"].sum(), eri_abs_sum=np.abs(ref["eri"]).sum(),
               ip1_abs_sum=np.abs(ref["eri_ip1"]).sum())
    for k in ("ovlp", "hcore", "ipovlp", "hcore_deriv", "e_nuc", "grad_nuc", "aoslices"):
        out[k] = ref[k]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "integrals_sp_OOH.npz"), **out)
    print("written", n, "AOs")


if __name__ == "__main__":
    main()
