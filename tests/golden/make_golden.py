"""Generate tests/golden/*.npz by running the REFERENCE ITSELF on seeded inputs.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

The reference's numpy code only needs ``pyscf``/``mpi4py`` at *import* time, so
stub modules are injected into ``sys.modules`` (PySCF is not installed in this
image) and the few PySCF free functions the prediction path calls are routed to
the AO arrays carried by an ``evcont_b200.mol.ArrayMol``:

    scf.hf.get_hcore(mol)                     -> mol.get_hcore()
    ao2mo.kernel(mol, C) / ao2mo.restore(1,.) -> four-index einsum over mol.intor('int2e')
    grad.RHF(scf.RHF(mol)).hcore_generator()  -> mol.hcore_generator()
    grad.RHF(scf.RHF(mol)).grad_nuc()         -> mol.grad_nuc()

Everything else (approximate_ground_state / approximate_multistate for all four
two_RDM layouts, get_loewdin_trafo, transform_integrals, compress/restore,
loewdin_trafo_grad, get_overlap_grad, get_one_el_grad, two_el_grad,
get_grad_elec_OAO, get_energy_with_grad) runs verbatim from /root/reference.

The script also checks the repo's oracle (oracle/subspace.py, oracle/gradients.py)
against those outputs before writing them, so a drifting oracle is caught here.
``trans_rdm12`` is PySCF C code and cannot be run: its goldens
(``trans_rdm_*.npz``) come from the brute-force second-quantisation checker in
oracle/trans_rdm.py and are marked ``source='definition'``.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def install_pyscf_stub():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    pyscf = mod("pyscf")
    scf = mod("pyscf.scf")
    hf = mod("pyscf.scf.hf")
    lo = mod("pyscf.lo")
    ao2mo = mod("pyscf.ao2mo")
    grad = mod("pyscf.grad")
    fci = mod("pyscf.fci")
    addons = mod("pyscf.fci.addons")
    direct_spin0 = mod("pyscf.fci.direct_spin0")
    lib = mod("pyscf.lib")
    md = mod("pyscf.md")
    gto = mod("pyscf.gto")
    mpi4py = mod("mpi4py")

    class _Comm:
        def Get_rank(self):
            return 0

        def Get_size(self):
            return 1

    class _MPI:
        COMM_WORLD = _Comm()

    mpi4py.MPI = _MPI
    hf.get_hcore = lambda mol: mol.get_hcore()
    scf.hf = hf
    scf.RHF = lambda mol: mol

    class _Grad:
        def __init__(self, mol):
            self.mol = mol

        def hcore_generator(self):
            return self.mol.hcore_generator()

        def grad_nuc(self):
            return self.mol.grad_nuc()

    grad.RHF = _Grad

    def ao2mo_kernel(mol, c):
        return np.einsum("ijkl,ia,jb,kc,ld->abcd", mol.intor("int2e"), c, c, c, c,
                         optimize="optimal")

    ao2mo.kernel = ao2mo_kernel
    ao2mo.restore = lambda sym, eri, n: eri
    direct_spin0.FCI = lambda *a, **k: None
    fci.direct_spin0 = direct_spin0
    fci.addons = addons
    addons.transform_ci = None
    lib.GradScanner = object
    for name, m in (("scf", scf), ("lo", lo), ("ao2mo", ao2mo), ("grad", grad),
                    ("fci", fci), ("lib", lib), ("md", md), ("gto", gto)):
        setattr(pyscf, name, m)


def synthetic_stack(norb, ntrain, seed, layout):
    """Seeded random t-RDM stack, in one of the four layouts."""
    rng = np.random.default_rng(seed)
    b = rng.standard_normal((ntrain, ntrain))
    ovlp = np.eye(ntrain) + 0.01 * (b + b.T)
    one = rng.standard_normal((ntrain, ntrain, norb, norb))
    one = one + one.transpose(1, 0, 2, 3)
    two = rng.standard_normal((ntrain, ntrain, norb * norb, norb * norb)) / norb
    two = two + two.transpose(1, 0, 2, 3)
    two = two + two.transpose(0, 1, 3, 2)
    full = two.reshape((ntrain, ntrain) + (norb,) * 4)
    il = np.tril_indices(ntrain)
    if layout == 6:
        return ovlp, one, full
    if layout == 5:
        return ovlp, one, full[il]
    ic = np.tril_indices(norb * norb)
    comp = two[:, :, ic[0], ic[1]]
    if layout == 3:
        return ovlp, one, comp
    assert layout == 2
    return ovlp, one, comp[il]


def main():
    install_pyscf_stub()
    sys.path.insert(0, "/root/reference")
    from evcont import ab_initio_eigenvector_continuation as ref_evc
    from evcont import ab_initio_gradients_loewdin as ref_grad
    from evcont import electron_integral_utils as ref_int

    from evcont_b200.mol import synthetic_mol
    from oracle import gradients as o_grad
    from oracle import subspace as o_sub
    from oracle import trans_rdm as o_trdm
    from oracle import cistring as o_str

    def check(name, a, b, tol):
        err = float(np.max(np.abs(np.asarray(a) - np.asarray(b))))
        assert err <= tol, f"oracle drift in {name}: {err:.3e} > {tol:.1e}"
        return err

    # ---- prediction path: (norb, natm, ntrain) cases x four layouts ---------
    cases = [(4, 2, 3, 11), (6, 6, 3, 12), (7, 3, 5, 13), (10, 10, 6, 14)]
    for norb, natm, ntrain, seed in cases:
        mol = synthetic_mol(norb, natm, seed=seed)
        out = {"norb": norb, "natm": natm, "ntrain": ntrain, "seed": seed}
        s_ao = mol.intor("int1e_ovlp")
        x = ref_int.get_loewdin_trafo(s_ao)
        check("loewdin", x, o_sub.get_loewdin_trafo(s_ao), 1e-13)
        out["loewdin_X"] = x
        h1 = x.T @ mol.get_hcore() @ x
        h2 = sys.modules["pyscf.ao2mo"].kernel(mol, x)
        o_h1, o_h2 = o_sub.ao_to_oao(mol.get_hcore(), mol.intor("int2e"), x)
        check("ao_to_oao h1", h1, o_h1, 1e-12)
        check("ao_to_oao h2", h2, o_h2, 1e-12)
        out["h1"], out["h2"] = h1, h2
        # NOTE reference defect: transform_integrals (electron_integral_utils.py:33)
        # passes ONE trafo operand to a three-subscript einsum and raises
        # ValueError for every input, so the function cannot be called.  Its two
        # einsum expressions are evaluated here with the evidently intended
        # operands (trafo once per index) to pin the index convention trafo[a, i].
        try:
            ref_int.transform_integrals(h1, h2, x)
            raise SystemExit("reference transform_integrals no longer raises: update goldens")
        except ValueError:
            pass
        r1 = np.einsum("...ij,ai,bj->...ab", h1, x, x, optimize="optimal")
        r2 = np.einsum("...ijkl,ai,bj,ck,dl->...abcd", h2, x, x, x, x, optimize="optimal")
        q1, q2 = o_sub.transform_integrals(h1, h2, x)
        check("transform_integrals", r2, q2, 1e-11)
        check("transform_integrals h1", r1, q1, 1e-11)
        out["trafo_h1"], out["trafo_h2"] = r1, r2
        c = ref_int.compress_electron_exchange_symmetry(h2.copy(), 0.5)
        check("compress", c, o_sub.compress_exchange(h2, 0.5), 0.0)
        check("restore", ref_int.restore_electron_exchange_symmetry(c, norb),
              o_sub.restore_exchange(c, norb), 0.0)
        out["h2_compressed_half"] = c
        dxs = ref_grad.loewdin_trafo_grad(s_ao)
        check("loewdin_trafo_grad", dxs, o_grad.loewdin_trafo_grad(s_ao), 1e-11)
        dx = ref_grad.get_derivative_ao_mo_trafo(mol)
        check("dX/dR", dx, o_grad.get_derivative_ao_mo_trafo(mol), 1e-11)
        out["dX_dR"] = dx
        check("overlap_grad", ref_grad.get_overlap_grad(mol), o_grad.get_overlap_grad(mol), 0.0)
        h1j = ref_grad.get_one_el_grad(mol, ao_mo_trafo=x, ao_mo_trafo_grad=dx)
        check("one_el_grad", h1j, o_grad.get_one_el_grad(mol, x, dx), 1e-11)
        out["h1_jac"] = h1j
        for layout in (6, 5, 3, 2):
            ovlp, one, two = synthetic_stack(norb, ntrain, seed + 100, layout)
            e, v = ref_evc.approximate_ground_state(h1, h2, one, two, ovlp)
            oe, ov = o_sub.approximate_ground_state(h1, h2, one, two, ovlp)
            check(f"gs energy L{layout}", e, oe, 1e-11)
            check(f"gs vec L{layout}", v * np.sign(v[0]), ov * np.sign(ov[0]), 1e-9)
            nr = min(3, ntrain)
            em, vm = ref_evc.approximate_multistate(h1, h2, one, two, ovlp, nroots=nr)
            oem, _ = o_sub.approximate_multistate(h1, h2, one, two, ovlp, nroots=nr)
            check(f"ms energy L{layout}", em, oem, 1e-11)
            en, g, gam, Gam = ref_grad.get_energy_with_grad(
                mol, one, two, ovlp, return_density_matrices=True)
            oen, og, ogam, oGam = o_grad.get_energy_with_grad(
                mol, one, two, ovlp, return_density_matrices=True)
            check(f"E L{layout}", en, oen, 1e-11)
            check(f"grad L{layout}", g, og, 1e-10)
            check(f"gamma L{layout}", gam, ogam, 1e-10)
            check(f"Gamma L{layout}", Gam, oGam, 1e-10)
            out[f"L{layout}_E0"], out[f"L{layout}_c0"] = e, v
            out[f"L{layout}_Ems"], out[f"L{layout}_Cms"] = em, vm
            out[f"L{layout}_Etot"], out[f"L{layout}_grad"] = en, g
            out[f"L{layout}_gamma"] = gam
            if norb <= 7:
                out[f"L{layout}_Gamma"] = Gam
            if layout == 6:
                two_grad = ref_grad.two_el_grad(
                    mol.intor("int2e"), Gam, x, dx, mol.intor("int2e_ip1"),
                    tuple((s[2], s[3]) for s in mol.aoslice_by_atom()))
                out["two_el_grad"] = two_grad
        path = os.path.join(HERE, f"predict_n{norb}_a{natm}_N{ntrain}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "kB")

    # ---- trans_rdm12: definition-based goldens (PySCF cannot run here) -------
    rng = np.random.default_rng(7)
    for norb, nelec in [(2, (1, 1)), (3, (2, 1)), (4, (2, 2)), (4, (3, 1)), (5, (3, 2))]:
        na = o_str.num_strings(norb, nelec[0])
        nb = o_str.num_strings(norb, nelec[1])
        bra = rng.standard_normal((na, nb))
        ket = rng.standard_normal((na, nb))
        bra /= np.linalg.norm(bra)
        ket /= np.linalg.norm(ket)
        d1, d2 = o_trdm.brute_force_rdm12(bra, ket, norb, nelec)
        o1, o2 = o_trdm.trans_rdm12(bra, ket, norb, nelec)
        check("trans_rdm12 dm1", d1, o1, 1e-14)
        check("trans_rdm12 dm2", d2, o2, 1e-14)
        path = os.path.join(HERE, f"trans_rdm_n{norb}_e{nelec[0]}{nelec[1]}.npz")
        np.savez_compressed(path, norb=norb, nelec=np.array(nelec), bra=bra, ket=ket,
                            dm1=d1, dm2=d2, source="definition")
        print("wrote", path)
    # link-table golden: the hand-derived rows of SURVEY.md Appendix A.2
    np.savez_compressed(
        os.path.join(HERE, "linkindex_n4_k2.npz"),
        strings=np.array([3, 5, 6, 9, 10, 12]),
        row0=np.array([[0, 0, 0, 1], [1, 1, 0, 1], [2, 0, 2, -1], [3, 0, 4, -1],
                       [2, 1, 1, 1], [3, 1, 3, 1]], dtype=np.int32),
        source="SURVEY.md Appendix A.2 (hand-derived)")


if __name__ == "__main__":
    main()
