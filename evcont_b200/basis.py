"""Gaussian basis-set data for the device integral engine (K9, ``csrc/integrals.cu``).

s shells (H and He in STO-3G / STO-6G / 6-31G: the hydrogen-chain configurations of the reference,
built there with ``basis="sto-6g"``) go through the specialised s kernel; molecules with p shells
(oxygen in 6-31G: scripts/MD/md_H2O_6_31G_FCI.py, the Zundel scripts) through the general s+p kernel
(``csrc/integrals_sp.cu``).  Exponents and contraction coefficients are the EMSL / Basis
Set Exchange values (coefficients refer to normalised primitives); like ``pyscf.gto``
the contracted function is renormalised to unit self-overlap.
"""
import math

import numpy as np

CHARGES = {"H": 1, "He": 2, "O": 8}

#: (element, basis) -> p shells (exponents, coefficients); "sp" shells of Pople bases are split
P_SHELLS = {
    ("O", "6-31g"): [
        ((15.539616, 3.5999336, 1.0137618), (0.0708743, 0.3397528, 0.7271586)),
        ((0.2700058,), (1.0,)),
    ],
}

#: (element, basis) -> list of s shells, each (exponents, contraction coefficients)
S_SHELLS = {
    ("H", "sto-6g"): [
        ((35.52322122, 6.513143725, 1.822142904, 0.625955266, 0.243076747, 0.100112428),
         (0.00916359628, 0.04936149294, 0.1685383049, 0.3705627997, 0.4164915298, 0.1303340841)),
    ],
    ("H", "sto-3g"): [
        ((3.42525091, 0.62391373, 0.16885540), (0.15432897, 0.53532814, 0.44463454)),
    ],
    ("H", "6-31g"): [
        ((18.7311370, 2.8253937, 0.6401217), (0.03349460, 0.23472695, 0.81375733)),
        ((0.1612778,), (1.0,)),
    ],
    ("He", "sto-3g"): [
        ((6.36242139, 1.15892300, 0.31364979), (0.15432897, 0.53532814, 0.44463454)),
    ],
    ("He", "sto-6g"): [
        ((65.98456824, 12.09819836, 3.384639924, 1.162715163, 0.451516322, 0.185959356),
         (0.00916359628, 0.04936149294, 0.1685383049, 0.3705627997, 0.4164915298, 0.1303340841)),
    ],
    ("He", "6-31g"): [
        ((38.4216340, 5.7780300, 1.2417740), (0.0237660, 0.1546790, 0.4696300)),
        ((0.2979640,), (1.0,)),
    ],
    ("O", "6-31g"): [
        ((5484.6717, 825.23495, 188.04696, 52.9645, 16.89757, 5.7996353),
         (0.0018311, 0.0139501, 0.0684451, 0.2327143, 0.470193, 0.3585209)),
        ((15.539616, 3.5999336, 1.0137618), (-0.1107775, -0.1480263, 1.130767)),
        ((0.2700058,), (1.0,)),
    ],
}


def normalised_s_shell(exps, coefs):
    """Weights ``c_k (2 a_k / pi)^(3/4) / sqrt(<phi|phi>)`` of a contracted s function."""
    e = np.asarray(exps, dtype=np.float64)
    c = np.asarray(coefs, dtype=np.float64) * (2.0 * e / np.pi) ** 0.75
    ss = (c[:, None] * c[None, :] * (np.pi / (e[:, None] + e[None, :])) ** 1.5).sum()
    return e, c / math.sqrt(ss)


def s_basis_tables(symbols, basis):
    """Host tables for ``evc_sbasis_create``: AOs grouped by atom in atom order.

    Returns ``dict(charges, ao_atom, ao_nprim, prim_exp, prim_wt)`` (numpy arrays).
    """
    key = basis.lower().replace("_", "-")
    charges, ao_atom, ao_nprim, prim_exp, prim_wt = [], [], [], [], []
    for ia, sym in enumerate(symbols):
        sym = sym.capitalize()
        if (sym, key) not in S_SHELLS:
            raise NotImplementedError(
                f"no s-shell data for element {sym!r} in basis {basis!r}: the device integral engine "
                f"covers {sorted(S_SHELLS)} (p shells are not implemented)")
        charges.append(CHARGES[sym])
        for exps, coefs in S_SHELLS[(sym, key)]:
            e, w = normalised_s_shell(exps, coefs)
            ao_atom.append(ia)
            ao_nprim.append(len(e))
            prim_exp.extend(e)
            prim_wt.extend(w)
    return dict(charges=np.asarray(charges, dtype=np.float64),
                ao_atom=np.asarray(ao_atom, dtype=np.int32),
                ao_nprim=np.asarray(ao_nprim, dtype=np.int32),
                prim_exp=np.asarray(prim_exp, dtype=np.float64),
                prim_wt=np.asarray(prim_wt, dtype=np.float64))


def normalised_p_shell(exps, coefs):
    """Weights of one Cartesian component of a contracted p function:
    ``c_k (2 a_k / pi)^(3/4) 2 sqrt(a_k) / sqrt(<phi|phi>)``."""
    e = np.asarray(exps, dtype=np.float64)
    c = np.asarray(coefs, dtype=np.float64) * (2.0 * e / np.pi) ** 0.75 * 2.0 * np.sqrt(e)
    pp = e[:, None] + e[None, :]
    ss = (c[:, None] * c[None, :] * (np.pi / pp) ** 1.5 / (2.0 * pp)).sum()
    return e, c / math.sqrt(ss)


def has_p_shells(symbols, basis):
    key = basis.lower().replace("_", "-")
    return any((s.capitalize(), key) in P_SHELLS for s in symbols)


def sp_basis_tables(symbols, basis):
    """Host tables for ``evc_gbasis_create``: contracted Cartesian AOs in pyscf.gto order (per atom the
    s shells, then the p shells with components x, y, z).  Returns ``dict(charges, ao_atom, ao_pow
    [nao, 3], ao_nprim, prim_exp, prim_wt)``."""
    key = basis.lower().replace("_", "-")
    charges, ao_atom, ao_pow, ao_nprim, prim_exp, prim_wt = [], [], [], [], [], []
    for ia, sym in enumerate(symbols):
        sym = sym.capitalize()
        if (sym, key) not in S_SHELLS:
            raise NotImplementedError(f"no basis data for element {sym!r} in basis {basis!r}")
        charges.append(CHARGES[sym])
        for exps, coefs in S_SHELLS[(sym, key)]:
            e, w = normalised_s_shell(exps, coefs)
            ao_atom.append(ia); ao_pow.append((0, 0, 0)); ao_nprim.append(len(e))
            prim_exp.extend(e); prim_wt.extend(w)
        for exps, coefs in P_SHELLS.get((sym, key), []):
            e, w = normalised_p_shell(exps, coefs)
            for comp in range(3):
                pw = [0, 0, 0]
                pw[comp] = 1
                ao_atom.append(ia); ao_pow.append(tuple(pw)); ao_nprim.append(len(e))
                prim_exp.extend(e); prim_wt.extend(w)
    return dict(charges=np.asarray(charges, dtype=np.float64), ao_atom=np.asarray(ao_atom, dtype=np.int32),
                ao_pow=np.ascontiguousarray(ao_pow, dtype=np.int32), ao_nprim=np.asarray(ao_nprim, dtype=np.int32),
                prim_exp=np.asarray(prim_exp, dtype=np.float64), prim_wt=np.asarray(prim_wt, dtype=np.float64))
