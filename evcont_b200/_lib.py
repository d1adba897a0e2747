"""ctypes binding of ``libevcont_b200.so`` (the C ABI in ``include/evcont_b200.h``).

The library is built in-tree by ``evcont_b200/csrc/Makefile`` (``python -m
evcont_b200.build`` or ``__graft_entry__.build()``).  There is no CPU fallback:
if the shared object is missing, :func:`lib` raises ``ImportError`` with the
build command; if it loads but no B200 is present, every compute entry point
fails with the library's own error message.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libevcont_b200.so")

c_double_p = C.c_void_p  # device pointers travel as integers
c_i64 = C.c_int64
c_sz_p = C.POINTER(C.c_size_t)


class AoBundle(C.Structure):
    """Mirror of ``evc_ao_bundle``."""
    _fields_ = [(name, C.c_void_p) for name in (
        "ovlp", "hcore", "eri", "ipovlp", "hcore_deriv", "eri_ip1", "e_nuc", "grad_nuc",
        "aoslices", "erip", "eri_ip1p")]


ABI_VERSION = 2  # EVC_ABI_VERSION of include/evcont_b200.h


class EvcError(RuntimeError):
    """An ``evc_*`` entry point returned a non-zero status."""


#: name -> (restype, argtypes); every symbol include/evcont_b200.h declares
SIGNATURES = {
    "evc_abi_version": (C.c_int, []),
    "evc_last_error": (C.c_char_p, []),
    "evc_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "evc_ctx_destroy": (C.c_int, [C.c_void_p]),
    "evc_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "evc_ctx_sm_count": (C.c_int, [C.c_void_p]),
    "evc_launch_count": (C.c_ulonglong, []),
    "evc_ctx_stage_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "evc_ctx_stage_times": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "evc_num_strings": (c_i64, [C.c_int, C.c_int]),
    "evc_num_links": (C.c_int, [C.c_int, C.c_int]),
    "evc_make_strings_host": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "evc_str2addr": (c_i64, [C.c_int, C.c_int, c_i64]),
    "evc_addr2str": (c_i64, [C.c_int, C.c_int, c_i64]),
    "evc_linkindex_build_host": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "evc_linkindex_pack_host": (C.c_int, [c_i64, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "evc_trans_rdm12_workspace_bytes": (C.c_int, [C.c_int, c_i64, c_i64, C.c_int, C.c_int, c_sz_p]),
    "evc_trans_rdm12_batch": (C.c_int, [
        C.c_void_p, C.c_int, c_i64, c_i64, c_double_p, c_i64, C.c_int, C.c_void_p, C.c_int,
        C.c_void_p, C.c_int, C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p,
        C.c_void_p, C.c_size_t]),
    "evc_trans_rdm12_batch_strided": (C.c_int, [
        C.c_void_p, C.c_int, c_i64, c_i64, c_double_p, c_i64, C.c_int, C.c_void_p, C.c_int,
        C.c_void_p, C.c_int, C.c_void_p, C.c_int, c_double_p, c_i64, c_double_p, c_i64, c_double_p, c_i64,
        C.c_void_p, C.c_size_t]),
    "evc_stack_row_len": (c_i64, [C.c_int]),
    "evc_stack_scatter_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_i64, C.c_int, C.c_void_p,
                                         c_double_p, c_double_p, c_double_p]),
    "evc_trans_rdm12_last_issued_flops": (C.c_double, [C.c_void_p]),
    "evc_loewdin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                              c_double_p]),
    "evc_loewdin_grad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p,
                                   c_double_p, c_double_p]),
    "evc_ao2oao": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                             C.c_int, c_double_p, c_double_p, c_double_p, C.c_void_p, C.c_size_t]),
    "evc_subspace_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_sz_p]),
    "evc_subspace_H": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p,
                                 C.c_int, c_double_p, c_double_p, c_double_p, C.c_void_p,
                                 C.c_size_t]),
    "evc_geneig_prepare": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.c_void_p]),
    "evc_geneig": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, C.c_int,
                             c_double_p, c_double_p]),
    "evc_predict_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_sz_p]),
    "evc_predict_rdm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p,
                                  C.c_int, c_double_p, c_i64, c_double_p, c_double_p, C.c_void_p,
                                  C.c_size_t]),
    "evc_grad_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, c_sz_p]),
    "evc_grad_elec": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p] +
                      [c_double_p] * 11 + [C.c_void_p, C.c_size_t]),
    "evc_energy_with_grad_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int,
                                                       C.c_int, c_sz_p]),
    "evc_energy_with_grad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p,
                                       c_double_p, c_double_p, C.c_int, C.POINTER(AoBundle),
                                       c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                       C.c_void_p, C.c_size_t]),
    "evc_energy_with_grad_host_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int,
                                                            C.c_int, c_sz_p]),
    "evc_energy_with_grad_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                            c_double_p, c_double_p, c_double_p, C.c_int,
                                            C.POINTER(AoBundle), C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_size_t]),
    "evc_packed_row_len": (c_i64, [C.c_int]),
    "evc_erip_pitch": (C.c_int, [C.c_int]),
    "evc_erip_len": (c_i64, [C.c_int]),
    "evc_eri_ip1p_len": (c_i64, [C.c_int]),
    "evc_ao_pack8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "evc_stack_pack8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p,
                                  c_double_p, c_double_p]),
    "evc_energy_with_grad_packed_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int,
                                                              c_sz_p]),
    "evc_energy_with_grad_packed": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p,
                                              c_double_p, c_double_p, C.c_int, C.POINTER(AoBundle),
                                              c_double_p, c_double_p, c_double_p, C.c_void_p,
                                              C.c_size_t]),
    "evc_energy_with_grad_packed_host_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int,
                                                                   C.c_int, c_sz_p]),
    "evc_energy_with_grad_packed_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                   c_double_p, c_double_p, c_double_p, C.c_int,
                                                   C.POINTER(AoBundle), C.c_void_p, C.c_void_p,
                                                   C.c_int, C.c_void_p, C.c_size_t]),
    "evc_sbasis_create": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "evc_sbasis_destroy": (C.c_int, [C.c_void_p]),
    "evc_sbasis_nao": (C.c_int, [C.c_void_p]),
    "evc_sbasis_natm": (C.c_int, [C.c_void_p]),
    "evc_sbasis_aoslices": (C.c_void_p, [C.c_void_p]),
    "evc_ao_integrals_s_workspace_bytes": (C.c_int, [C.c_void_p, C.c_int, c_sz_p]),
    "evc_ao_integrals_s": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int] + [c_double_p] * 9 +
                           [C.c_void_p, C.c_size_t]),
    "evc_ao_integrals_s_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int] + [c_double_p] * 9 +
                                  [C.c_void_p, C.c_size_t]),
    "evc_gbasis_create": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "evc_gbasis_destroy": (C.c_int, [C.c_void_p]),
    "evc_ao_integrals_sp_workspace_bytes": (C.c_int, [C.c_void_p, C.c_int, c_sz_p]),
    "evc_ao_integrals_sp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int] + [c_double_p] * 9 +
                            [C.c_void_p, C.c_size_t]),
    "evc_trans_rdm12_plan_pairs": (C.c_int, [C.c_void_p, C.c_int]),
    "evc_exchange_compress": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p]),
    "evc_exchange_restore": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p]),
    "evc_stack_rows_workspace_bytes": (C.c_int, [c_i64, C.c_int, C.c_int, c_sz_p]),
    "evc_stack_rows_dot": (C.c_int, [C.c_void_p, c_double_p, c_i64, C.c_int, c_double_p, C.c_int, c_double_p,
                                     C.c_void_p, C.c_size_t]),
    "evc_stack_rows_axpy": (C.c_int, [C.c_void_p, c_double_p, c_i64, C.c_int, c_double_p, C.c_int, c_double_p,
                                      C.c_void_p, C.c_size_t]),
    "evc_aotable_create": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "evc_aotable_destroy": (C.c_int, [C.c_void_p]),
    "evc_center_of_mass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_double_p, c_double_p]),
    "evc_int1e_r": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p]),
    "evc_rdm1_observables": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [c_double_p] * 9),
    "evc_fci_hdiag": (C.c_int, [C.c_void_p, C.c_int, c_i64, c_i64, C.c_void_p, C.c_void_p, c_double_p,
                                c_double_p, c_double_p]),
    "evc_fci_contract_workspace_bytes": (C.c_int, [C.c_int, c_i64, c_i64, c_sz_p]),
    "evc_fci_contract_2e": (C.c_int, [C.c_void_p, C.c_int, c_i64, c_i64, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, C.c_void_p,
                                      C.c_size_t]),
    "evc_transform_ci_workspace_bytes": (C.c_int, [C.c_int, c_i64, c_i64, c_sz_p]),
    "evc_transform_ci": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_i64, c_i64, C.c_void_p, C.c_void_p,
                                   c_double_p, c_double_p, c_double_p, C.c_void_p, C.c_size_t]),
    "evc_min_sqdist": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_i64, c_i64, c_double_p, c_double_p, c_double_p]),
    "evc_fock_rhf": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "evc_md_positions": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, c_double_p, c_double_p,
                                   c_double_p]),
    "evc_md_berendsen": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                   c_double_p, c_double_p]),
    "evc_md_velocities": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int] + [c_double_p] * 8 +
                          [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p]),
}

_lib = None


def lib():
    """The loaded shared library (loaded once); raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m evcont_b200.build` "
                "(nvcc, sm_100a).  evcont_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI drifted
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.evc_abi_version() != ABI_VERSION:
            raise ImportError("libevcont_b200.so ABI version mismatch; rebuild it")
        _lib = handle
    return _lib


def check(status):
    if status != 0:
        msg = lib().evc_last_error()
        raise EvcError(f"libevcont_b200 error {status}: {msg.decode() if msg else '?'}")
