/*
 * evcont_b200.h -- C ABI of libevcont_b200.so, the B200 (sm_100a) engine behind
 * evcont's FCI eigenvector-continuation path.
 *
 * The reference (BoothGroup/evcont) has no FFI of its own: its extension points
 * are duck-typed Python objects.  Each entry point below
 * names the reference interface (file:line under the reference tree) whose
 * arithmetic it replaces; the Python package `evcont_b200` binds them with
 * ctypes and mirrors the reference's function names on top.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; evc_last_error() gives
 *     the thread-local message of the last failure.
 *   - all array arguments are DEVICE pointers owned by the caller (torch tensors
 *     on the Python side) unless the parameter name ends in `_host`.
 *   - no hidden device allocation: scratch space is passed in as
 *     (workspace, workspace_bytes); `*_workspace_bytes()` tells how much.
 *   - all work is enqueued on the ctx stream and is asynchronous; the caller
 *     synchronises.  One ctx per GPU/thread; calls on one ctx are not re-entrant.
 *   - all floating point data is IEEE double, C order (row-major).
 *   - "batch" always means independent geometries (MD replicas / scan points).
 */
#ifndef EVCONT_B200_H
#define EVCONT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVC_ABI_VERSION 2

typedef struct evc_ctx evc_ctx;

/* layouts of the two-body t-RDM stack, named after the `two_RDM.ndim` dispatch of
 * evcont/ab_initio_eigenvector_continuation.py:41-71 */
enum {
  EVC_LAYOUT_FULL = 6,      /* (N, N, n, n, n, n)                        */
  EVC_LAYOUT_TRIL = 5,      /* (N(N+1)/2, n, n, n, n), np.tril_indices   */
  EVC_LAYOUT_FULL_EXCH = 3, /* (N, N, n^2(n^2+1)/2)                      */
  EVC_LAYOUT_TRIL_EXCH = 2  /* (N(N+1)/2, n^2(n^2+1)/2)                  */
};

/* ---- library / context ------------------------------------------------- */
int evc_abi_version(void);
const char *evc_last_error(void);
/* stream: a cudaStream_t (NULL = legacy default stream). */
int evc_ctx_create(int device, void *stream, evc_ctx **out);
int evc_ctx_destroy(evc_ctx *ctx);
int evc_ctx_set_stream(evc_ctx *ctx, void *stream);
int evc_ctx_sm_count(const evc_ctx *ctx);
/* kernel launches issued by this library so far (process-wide; for bench.py's
 * `gpu_launches`). */
unsigned long long evc_launch_count(void);
/* Per-stage device timing of evc_energy_with_grad (CUDA events on the ctx
 * stream; stages: 0 Loewdin, 1 AO->OAO, 2 subspace H, 3 eigensolve, 4 predicted
 * RDMs, 5 gradient (packed step: its per-geometry GEMM kernel), 6 streaming
 * contraction of the derivative integrals (packed step only).  evc_ctx_stage_timing(ctx, 1) resets and enables the
 * accumulators; evc_ctx_stage_times synchronises on the last call's events and
 * returns the accumulated milliseconds per stage and the number of calls. */
#define EVC_NUM_STAGES 7
int evc_ctx_stage_timing(evc_ctx *ctx, int enable);
int evc_ctx_stage_times(evc_ctx *ctx, double *ms /* [EVC_NUM_STAGES] */, int64_t *calls);

/* ---- K0: occupation strings and link tables (host, bit-exact) ---------
 * Replaces pyscf.fci.cistring.make_strings / str2addr / gen_linkstr_index, which
 * the reference reaches (and rebuilds on every call) through
 * cisolver.trans_rdm12 at evcont/FCI_EVCont.py:121.  Layout: SURVEY.md App. A.2. */
int64_t evc_num_strings(int norb, int nocc);
int evc_num_links(int norb, int nocc); /* nocc + nocc*(norb-nocc) */
int evc_make_strings_host(int norb, int nocc, int64_t *strings_host);
int64_t evc_str2addr(int norb, int nocc, int64_t string);
int64_t evc_addr2str(int norb, int nocc, int64_t addr);
/* out_host: int32 [nstr][nlink][4] = (cre a, des i, addr of a^+ i|str>, sign) */
int evc_linkindex_build_host(int norb, int nocc, int32_t *out_host);
/* pack the int32 table into the 8-byte device records the kernels read:
 * bits 0..31 target address, 32..39 a, 40..47 i, 48..55 sign (as int8).
 * link_major == 0: packed[str][link] (alpha role: one string per CTA);
 * link_major != 0: packed[link][str] (beta role: coalesced over strings). */
int evc_linkindex_pack_host(int64_t nstr, int nlink, const int32_t *tab_host,
                            int link_major, uint64_t *packed_host);

/* ---- K1+K2: transition 1-/2-RDMs --------------------------------------
 * Replaces cisolver.trans_rdm12(cibra, ciket, norb, nelec) as called at
 * evcont/FCI_EVCont.py:117-127 (PySCF: make_rdm12_spin1('FCItdm12kern_sf') +
 * reorder_rdm).  Conventions: dm1[p,q] = <bra|q^+ p|ket>,
 * dm2[p,q,r,s] = <bra|p^+ r^+ s q|ket>, spin-summed; ovlp = <bra|ket>.
 *
 * civecs: nvec vectors of shape (na, nb), vector v at civecs + v*vec_stride.
 * pairs : int32 [npairs][2] = (bra index, ket index), DEVICE pointer.
 * Outputs: ovlp[npairs], dm1[npairs][n][n], dm2[npairs][n][n][n][n].
 * The fused kernel gathers t1[K,(pq)] = <K|E_pq|v> for bra and ket through the
 * packed alpha/beta link tables into shared memory and contracts
 * t1_bra^T t1_ket on the FP64 tensor cores (DMMA); split-K partial sums are
 * reduced in a fixed order, so results are run-to-run bit-identical.
 * norb <= 13 in this version. */
int evc_trans_rdm12_workspace_bytes(int norb, int64_t na, int64_t nb, int npairs,
                                    int sm_count, size_t *bytes);
int evc_trans_rdm12_batch(evc_ctx *ctx, int norb, int64_t na, int64_t nb,
                          const double *civecs, int64_t vec_stride, int nvec,
                          const int32_t *pairs, int npairs,
                          const uint64_t *link_a /* string-major */, int nlink_a,
                          const uint64_t *link_b /* link-major */, int nlink_b,
                          double *ovlp, double *dm1, double *dm2,
                          void *workspace, size_t workspace_bytes);
/* The same with strided outputs (pair p: ovlp[p * ovlp_stride], dm1 + p * dm1_stride, dm2 + p * dm2_stride), so
 * that one pair's results can form ONE contiguous row [dm2 | dm1 | ovlp] of evc_stack_row_len(norb) doubles: the
 * send buffer of the single all_gather of the multi-GPU stack build. */
int evc_trans_rdm12_batch_strided(evc_ctx *ctx, int norb, int64_t na, int64_t nb,
                                  const double *civecs, int64_t vec_stride, int nvec,
                                  const int32_t *pairs, int npairs, const uint64_t *link_a,
                                  int nlink_a, const uint64_t *link_b, int nlink_b, double *ovlp,
                                  int64_t ovlp_stride, double *dm1, int64_t dm1_stride, double *dm2,
                                  int64_t dm2_stride, void *workspace, size_t workspace_bytes);
/* A pair's result is a sum over alpha-string slices whose number the kernel plans from the pair count of
 * the call.  A rank that computes only its share of a build (evcont_b200/distributed.py) announces the pair
 * count of the WHOLE build here (0 restores the default), so that every pair's bits are the same at any
 * number of ranks.  Workspace sizes computed for the call's own pair count remain sufficient. */
int evc_trans_rdm12_plan_pairs(evc_ctx *ctx, int total_pairs);
int64_t evc_stack_row_len(int norb); /* n^4 + n^2 + 1, rounded up to even */
/* rows [nrows][row_stride] of pairs row_pairs[nrows][2] = (a, b) (device int32) -> overlap (N,N), one_rdm
 * (N,N,n,n), two_rdm (N,N,n,n,n,n): block [a,b] and the untransposed copy at [b,a], exactly how
 * FCI_EVCont_obj.append_to_rdms stores them (evcont/FCI_EVCont.py:117-127). */
int evc_stack_scatter_rows(evc_ctx *ctx, int ntrain, int norb, const double *rows, int64_t row_stride,
                           int nrows, const int32_t *row_pairs, double *overlap, double *one_rdm,
                           double *two_rdm);
/* issued DMMA flop count of the last evc_trans_rdm12_batch call on this ctx
 * (for tensor-pipe utilisation; the algorithmic count is 2 n^4 ndet per pair) */
double evc_trans_rdm12_last_issued_flops(const evc_ctx *ctx);

/* ---- K3: Loewdin orthogonalisation -------------------------------------
 * Replaces get_loewdin_trafo (evcont/electron_integral_utils.py:6-18):
 * S = V diag(s) V^T (batched Jacobi), X = V diag(s>1e-15 ? s^-1/2 : 0) V^T.
 * s_ao, x, evecs: [nbatch][n][n]; evals: [nbatch][n] ascending; n <= 32.
 * evecs[b][i][k] = component i of eigenvector k (numpy eigh convention). */
int evc_loewdin(evc_ctx *ctx, int nbatch, int n, const double *s_ao, double *x,
                double *evals, double *evecs);
/* Replaces loewdin_trafo_grad + the einsum of get_derivative_ao_mo_trafo
 * (evcont/ab_initio_gradients_loewdin.py:41-134):
 * dX_xi = V (G o (V^T dS_xi V)) V^T with the exact divided differences
 * G_pq = -1/(sqrt(s_p) sqrt(s_q) (sqrt(s_p)+sqrt(s_q))).
 * dS, dX: [nbatch][nder][n][n]. */
int evc_loewdin_grad(evc_ctx *ctx, int nbatch, int n, int nder, const double *evals,
                     const double *evecs, const double *dS, double *dX);

/* ---- K4: AO -> orthogonal-basis integral transform ----------------------
 * Replaces get_integrals / the inline copy in get_energy_with_grad
 * (evcont/electron_integral_utils.py:122-138, ab_initio_gradients_loewdin.py:338-339;
 * PySCF ao2mo.kernel + restore(1)):  h1 = C^T hcore C,
 * h2[abcd] = sum (ij|kl) C_ia C_jb C_kc C_ld, four chained FP64 DMMA GEMM passes.
 * hcore, c, h1: [nbatch][n][n]; eri, h2: [nbatch][n^4]; t3 (optional, may be
 * NULL): [nbatch][n^4] receives the three-quarter transform
 * t3[l,a,b,c] = sum_ijk (ij|kl) C_ia C_jb C_kc needed by the gradient.
 * workspace: 2*nbatch*n^4 doubles.  `transpose_c` != 0 uses C^T (the
 * trafo[a,i] convention of transform_integrals, electron_integral_utils.py:21-35). */
int evc_ao2oao(evc_ctx *ctx, int nbatch, int n, const double *hcore, const double *eri,
               const double *c, int transpose_c, double *h1, double *h2, double *t3,
               void *workspace, size_t workspace_bytes);

/* ---- K5: subspace Hamiltonian from the stack ----------------------------
 * Replaces the H assembly of approximate_ground_state / approximate_multistate
 * (evcont/ab_initio_eigenvector_continuation.py:38-71) for all four layouts.
 * one_rdm: [N][N][n][n]; two_rdm: per `layout`; h1: [nbatch][n][n];
 * h2: [nbatch][n^4]; H: [nbatch][N][N] (lower triangle is what the eigensolver
 * reads; for the TRIL layouts the strict upper triangle holds the one-body part
 * only, exactly like the reference).
 * workspace: evc_subspace_workspace_bytes(). */
int evc_subspace_workspace_bytes(int layout, int ntrain, int n, int nbatch, size_t *bytes);
int evc_subspace_H(evc_ctx *ctx, int layout, int ntrain, int n, const double *one_rdm,
                   const double *two_rdm, int nbatch, const double *h1, const double *h2,
                   double *H, void *workspace, size_t workspace_bytes);

/* ---- K6: generalized symmetric-definite eigenproblem H c = E S c ----------
 * Replaces scipy.linalg.eigh(H, S) + root selection
 * (evcont/ab_initio_eigenvector_continuation.py:75-88, 157-173).
 * evc_geneig_prepare: S = L L^T once per stack (S does not depend on geometry);
 * writes Linv [N][N] (L^-1, lower) and info (device int: 0 ok, k>0 = leading
 * minor k not positive definite -- the caller raises like LAPACK does).
 * evc_geneig: per geometry A = L^-1 H L^-T (lower triangle of H), batched
 * two-sided Jacobi, c = L^-T y; returns the nroots lowest:
 * E: [nbatch][nroots], C: [nbatch][nroots][N] with c^T S c = 1. N <= 112. */
int evc_geneig_prepare(evc_ctx *ctx, int ntrain, const double *S, double *Linv, int *info);
int evc_geneig(evc_ctx *ctx, int nbatch, int ntrain, const double *H, const double *Linv,
               int nroots, double *E, double *C);

/* ---- K7: predicted RDMs (c (x) c) . stack ---------------------------------
 * Replaces evcont/ab_initio_gradients_loewdin.py:343-361.
 * C: [nbatch][c_stride] (first N entries = ground-state vector);
 * gamma: [nbatch][n][n]; Gamma: [nbatch][n^4] (exchange symmetry restored). */
int evc_predict_workspace_bytes(int layout, int ntrain, int n, int nbatch, size_t *bytes);
int evc_predict_rdm(evc_ctx *ctx, int layout, int ntrain, int n, const double *one_rdm,
                    const double *two_rdm, int nbatch, const double *C, int64_t c_stride,
                    double *gamma, double *Gamma, void *workspace, size_t workspace_bytes);

/* ---- e3: K5 / K7 on a slab of training pairs (pair-sharded stack, one trajectory) -------
 * for a stack too large or too slow to stream on one GPU
 * (Zundel, N = 100: 12.5 GB per step) every rank keeps a contiguous slab of the lower-triangular
 * pair list of the exchange-compressed layouts (np.tril_indices order, the per-pair directories of
 * scripts/MD/Zundel_thermodynamics/continuation/04_Zundel_continuation_MD.py:99-128), computes
 * its entries of H, and its share of the predicted two-body density matrix; the entries are
 * all-gathered, the shares all-reduced (evcont_b200/distributed.py).
 *   evc_exchange_compress   h2c [G][n2 (n2+1)/2] = lower triangle of h2 [G][n2][n2], diagonal halved
 *                           (compress_electron_exchange_symmetry(h2, 0.5), electron_integral_utils.py:38-66)
 *   evc_exchange_restore    Gamma [G][n^4] from its compressed form (restore_..., :69-88)
 *   evc_stack_rows_dot      out [G][nrows] = sum_l rows[p][l] hv[g][l]
 *   evc_stack_rows_axpy     out [G][row_len] = sum_p w[g][p] rows[p][l]
 * Few geometries stream the rows once at HBM speed; more than 16 run on the FP64 tensor cores. */
int evc_exchange_compress(evc_ctx *ctx, int n, int nbatch, const double *h2, double *h2c);
int evc_exchange_restore(evc_ctx *ctx, int n, int nbatch, const double *gamma2c, double *Gamma);
int evc_stack_rows_workspace_bytes(int64_t row_len, int nrows, int nbatch, size_t *bytes);
int evc_stack_rows_dot(evc_ctx *ctx, const double *rows, int64_t row_len, int nrows, const double *hv,
                       int nbatch, double *out, void *workspace, size_t workspace_bytes);
int evc_stack_rows_axpy(evc_ctx *ctx, const double *rows, int64_t row_len, int nrows, const double *w,
                        int nbatch, double *out, void *workspace, size_t workspace_bytes);

/* ---- K8: electronic gradient in the Loewdin basis -------------------------
 * Replaces get_grad_elec_OAO and its callees
 * (evcont/ab_initio_gradients_loewdin.py:13-305) in the adjoint form of
 * no dX/dR tensor is formed.
 * Inputs per geometry: evals/evecs/x from evc_loewdin, hcore [n][n], t3 from
 * evc_ao2oao, gamma/Gamma from evc_predict_rdm, ipovlp [3][n][n] (int1e_ipovlp),
 * hcore_deriv [natm][3][n][n] (hcore_generator per atom), eri_ip1 [3][n^4]
 * (int2e_ip1), aoslices int32 [natm][2] (shared by the batch).
 * Output grad_elec [nbatch][natm][3].  workspace: evc_grad_workspace_bytes().
 * Requires the 8-fold symmetric AO ERIs libcint produces (t3 is reused as the
 * three-quarter transform through (ij|kl) = (lk|ji)). */
int evc_grad_workspace_bytes(int n, int natm, int nbatch, size_t *bytes);
int evc_grad_elec(evc_ctx *ctx, int nbatch, int n, int natm, const int32_t *aoslices,
                  const double *evals, const double *evecs, const double *x,
                  const double *hcore, const double *t3, const double *gamma,
                  const double *Gamma, const double *ipovlp, const double *hcore_deriv,
                  const double *eri_ip1, double *grad_elec, void *workspace,
                  size_t workspace_bytes);

/* ---- fused prediction step -------------------------------------------------
 * Replaces get_energy_with_grad(mol, one_RDM, two_RDM, S)
 * (evcont/ab_initio_gradients_loewdin.py:308-379) for a batch of geometries whose
 * AO arrays are resident on the device: K3 -> K4 -> K5 -> K6 -> K7 -> K8 on the
 * ctx stream with no host round trip (CUDA-graph capturable).
 * e_nuc [nbatch] and grad_nuc [nbatch][natm][3] are added to the outputs.
 * E: [nbatch]; grad: [nbatch][natm][3]; gamma/Gamma may be NULL (then they live
 * in the workspace).  Linv from evc_geneig_prepare. */
typedef struct {
  const double *ovlp;        /* [nbatch][n][n]        int1e_ovlp            */
  const double *hcore;       /* [nbatch][n][n]        scf.hf.get_hcore      */
  const double *eri;         /* [nbatch][n^4]         int2e                 */
  const double *ipovlp;      /* [nbatch][3][n][n]     int1e_ipovlp          */
  const double *hcore_deriv; /* [nbatch][natm][3][n][n] hcore_generator     */
  const double *eri_ip1;     /* [nbatch][3][n^4]      int2e_ip1             */
  const double *e_nuc;       /* [nbatch]                                    */
  const double *grad_nuc;    /* [nbatch][natm][3]                           */
  const int32_t *aoslices;   /* [natm][2] (ao start, ao stop), shared       */
  /* Packed two-electron arrays (ABI 2) -- the fast inputs of the packed step (n <= 13); when
   * set, `eri` / `eri_ip1` may be NULL.  What the device integral kernels emit directly
   * (evc_ao_integrals_s_packed) and what evc_ao_pack8 makes from the full tensors:
   *   erip     [nbatch][np][evc_erip_pitch(n)]  erip[AB][CD] = (ab|cd), a >= b, c >= d,
   *            AB = a(a+1)/2 + b, np = n(n+1)/2; the columns [np, pitch) are padding (ignored)
   *   eri_ip1p [nbatch][3][n][n][np]            (d_x m b|c d), CD = c(c+1)/2 + d, c >= d
   * (int2e has the 8-fold symmetry, int2e_ip1 is symmetric in its last two indices:
   * evcont/ab_initio_gradients_loewdin.py:283-284.) */
  const double *erip;
  const double *eri_ip1p;
} evc_ao_bundle;

int evc_energy_with_grad_workspace_bytes(int layout, int ntrain, int n, int natm,
                                         int nbatch, size_t *bytes);
int evc_energy_with_grad(evc_ctx *ctx, int layout, int ntrain, int n, int natm,
                         const double *one_rdm, const double *two_rdm, const double *Linv,
                         int nbatch, const evc_ao_bundle *ao, double *E, double *grad,
                         double *gamma, double *Gamma, double *Cvec,
                         void *workspace, size_t workspace_bytes);

/* ---- packed prediction step (8-fold symmetric form) --------------------------
 * The same step as evc_energy_with_grad, restated on the permutational symmetry
 * (ij|kl) = (ji|kl) = (ij|lk) = (kl|ij) of the AO two-electron integrals that
 * mol.intor('int2e') returns (evcont/ab_initio_gradients_loewdin.py:283, 338-339):
 * H (ab_initio_eigenvector_continuation.py:38-71) and the gradient
 * (ab_initio_gradients_loewdin.py:190-305) only see the totally symmetric part of
 * each two-body t-RDM block, so the stack is packed once into N(N+1)/2 rows of
 * evc_packed_row_len(n) = n^2 + np(np+1)/2 doubles (np = n(n+1)/2, rounded up to
 * even): 12x fewer bytes and flops than the (N,N,n,n,n,n) layout at n=10, N=20.
 *
 * evc_stack_pack8: any of the four layouts -> RH (rows for H: blocks a >= b, what
 *   the eigensolver reads) and RG (rows for the predicted RDMs: (block[a,b] +
 *   block[b,a])/2, used with the weights 2 c_a c_b / c_a^2 of
 *   ab_initio_gradients_loewdin.py:345-353).  RH, RG: [N(N+1)/2][row_len] each.
 * evc_energy_with_grad_packed: Loewdin -> packed AO->OAO (two DMMA GEMMs per
 *   geometry in shared memory, n <= 13; the full-tensor kernels above for larger n)
 *   -> Hp = hvec . RH -> eigensolve -> out7 = w . RG -> gradient (three DMMA GEMMs
 *   per geometry in shared memory, int2e_ip1 streamed once).  Same outputs as
 *   evc_energy_with_grad (E, grad, optional Cvec); the full predicted RDMs are not
 *   formed -- use evc_energy_with_grad when they are wanted.  grad == NULL: energies (and
 *   Cvec) only, i.e. approximate_ground_state_OAO for a batch
 *   (ab_initio_eigenvector_continuation.py:178-211); the derivative arrays of the bundle
 *   are not read then. */
int64_t evc_packed_row_len(int n);
/* layout of the packed AO two-electron arrays of evc_ao_bundle: row pitch / doubles per geometry */
int evc_erip_pitch(int n);
int64_t evc_erip_len(int n);     /* np * pitch       */
int64_t evc_eri_ip1p_len(int n); /* 3 * n * n * np   */
/* int2e [nbatch][n^4] -> erip, int2e_ip1 [nbatch][3][n^4] -> eri_ip1p (either pair may be NULL) */
int evc_ao_pack8(evc_ctx *ctx, int nbatch, int n, const double *eri, const double *eri_ip1,
                 double *erip, double *eri_ip1p);
int evc_stack_pack8(evc_ctx *ctx, int layout, int ntrain, int n, const double *one_rdm,
                    const double *two_rdm, double *RH, double *RG);
int evc_energy_with_grad_packed_workspace_bytes(int ntrain, int n, int natm, int nbatch,
                                                size_t *bytes);
int evc_energy_with_grad_packed(evc_ctx *ctx, int ntrain, int n, int natm, const double *RH,
                                const double *RG, const double *Linv, int nbatch,
                                const evc_ao_bundle *ao, double *E, double *grad, double *Cvec,
                                void *workspace, size_t workspace_bytes);

/* The same step with HOST buffers on both sides -- the form the reference's call
 * site has (evcont/MD_utils.py:43: numpy arrays from libcint in, (E, grad) out).
 * Every pointer of `ao_host` (aoslices included), E_host and grad_host are HOST
 * pointers (pinned memory keeps the call asynchronous); one_rdm / two_rdm / Linv /
 * workspace are device pointers.  The batch is processed in chunks of `chunk`
 * geometries: host->device copies, the K3..K8 kernels and the device->host
 * read-back of consecutive chunks overlap on three streams; completion is ordered
 * on the ctx stream (synchronise it before reading E_host / grad_host). */
int evc_energy_with_grad_host_workspace_bytes(int layout, int ntrain, int n, int natm, int chunk,
                                              size_t *bytes);
int evc_energy_with_grad_host(evc_ctx *ctx, int layout, int ntrain, int n, int natm,
                              const double *one_rdm, const double *two_rdm, const double *Linv,
                              int nbatch, const evc_ao_bundle *ao_host, double *E_host,
                              double *grad_host, int chunk, void *workspace, size_t workspace_bytes);

/* Packed step with HOST buffers (same pipeline as evc_energy_with_grad_host). */
int evc_energy_with_grad_packed_host_workspace_bytes(int ntrain, int n, int natm, int chunk,
                                                     size_t *bytes);
int evc_energy_with_grad_packed_host(evc_ctx *ctx, int ntrain, int n, int natm, const double *RH,
                                     const double *RG, const double *Linv, int nbatch,
                                     const evc_ao_bundle *ao_host, double *E_host, double *grad_host,
                                     int chunk, void *workspace, size_t workspace_bytes);

/* ---- K9: AO integrals over contracted s-type Gaussians, on the device ------------
 * Replaces, for s shells (H / He: STO-nG, 6-31G), what the reference asks PySCF /
 * libcint for on every prediction step
 * (evcont/ab_initio_gradients_loewdin.py:25 int1e_ipovlp, :130/:338 int1e_ovlp,
 * :147 grad.RHF.hcore_generator [int1e_ipkin, int1e_ipnuc, int1e_iprinv],
 * :177/:338 scf.hf.get_hcore [int1e_kin + int1e_nuc], :283 int2e, :284 int2e_ip1,
 * :339 ao2mo.kernel, :370 grad_nuc, :378 energy_nuc): closed-form Gaussian
 * integrals with tabulated Boys functions, one warp per contracted quartet.
 *
 * evc_sbasis_create: host description of the basis (AOs grouped by atom, in atom
 *   order, as pyscf.gto orders them): charges_host[natm], ao_atom_host[nao],
 *   ao_nprim_host[nao], then the primitives of AO 0, AO 1, ... in
 *   prim_exp_host / prim_wt_host, where wt = contraction coefficient x primitive
 *   norm (2a/pi)^(3/4) x the factor that gives the contracted function unit
 *   self-overlap.  The handle owns small device tables (allocated here, never
 *   during evc_ao_integrals_s).  natm <= 64, nao <= 64, <= 16 primitives per AO.
 * evc_ao_integrals_s: coords [nbatch][natm][3] (bohr, device) -> the evc_ao_bundle
 *   arrays of every geometry, in the layouts documented at evc_ao_bundle;
 *   aoslices for the bundle come from evc_sbasis_aoslices (device, [natm][2]).  The
 *   primitive-pair tables of a geometry live in shared memory when they fit (H10/STO-6G:
 *   1980 pairs); larger systems (H30: 16 740 pairs) build them in `workspace` first
 *   (evc_ao_integrals_s_workspace_bytes; 256 bytes otherwise). */
typedef struct evc_sbasis evc_sbasis;
int evc_sbasis_create(evc_ctx *ctx, int natm, const double *charges_host, int nao,
                      const int32_t *ao_atom_host, const int32_t *ao_nprim_host,
                      const double *prim_exp_host, const double *prim_wt_host,
                      evc_sbasis **out);
int evc_sbasis_destroy(evc_sbasis *basis);
int evc_sbasis_nao(const evc_sbasis *basis);
int evc_sbasis_natm(const evc_sbasis *basis);
const int32_t *evc_sbasis_aoslices(const evc_sbasis *basis);
int evc_ao_integrals_s_workspace_bytes(const evc_sbasis *basis, int nbatch, size_t *bytes);
int evc_ao_integrals_s(evc_ctx *ctx, const evc_sbasis *basis, int nbatch, const double *coords,
                       double *ovlp, double *hcore, double *eri, double *ipovlp,
                       double *hcore_deriv, double *eri_ip1, double *e_nuc, double *grad_nuc,
                       void *workspace, size_t workspace_bytes);
/* The same with the two-electron arrays emitted directly in the packed layouts of evc_ao_bundle
 * (erip [nbatch][np][evc_erip_pitch(n)], eri_ip1p [nbatch][3][n][n][np]): every contracted quartet
 * is stored once per symmetry-distinct position instead of eight times, and the prediction step
 * reads 2.2x fewer bytes. */
int evc_ao_integrals_s_packed(evc_ctx *ctx, const evc_sbasis *basis, int nbatch,
                              const double *coords, double *ovlp, double *hcore, double *erip,
                              double *ipovlp, double *hcore_deriv, double *eri_ip1p, double *e_nuc,
                              double *grad_nuc, void *workspace, size_t workspace_bytes);

/* ---- device-resident velocity Verlet (batched over trajectories) ------------------
 * Replaces the host loop of pyscf.md.NVE that get_trajectory drives
 * (evcont/MD_utils.py:60-125; one scanner call per step): x, v, a: [nbatch][natm][3]
 * stay on the device.  One MD step is
 *   evc_md_positions  : x += dt v + dt^2/2 a
 *   (evc_ao_integrals_s + evc_energy_with_grad[_packed] at the new x)
 *   evc_md_velocities : a' = -grad / m;  v += dt/2 (a + a');  a = a';  E_kin; frame recording
 * `first` != 0 is the start-up call (frame 0 = initial geometry: only a is set).
 * frame_idx is a device counter (advanced by every evc_md_velocities call) so that a
 * captured CUDA graph of one step can be replayed; traj [max_frames][nbatch][natm][3],
 * epot_log / ekin_log [max_frames][nbatch] may be NULL. */
int evc_md_positions(evc_ctx *ctx, int nbatch, int natm, double dt, const double *v,
                     const double *a, double *x);
/* Berendsen velocity rescaling (pyscf.md.NVTBerendson, used by the reference's Zundel
 * scripts with T = 298.15 K, taut = 250 a.u.): called before evc_md_positions;
 * v *= clip(sqrt(1 + (T/T_inst - 1) dt/taut), 0.9, 1.1), T_inst from ekin [nbatch]. */
int evc_md_berendsen(evc_ctx *ctx, int nbatch, int natm, double dt, double taut,
                     double temperature, const double *ekin, double *v);
int evc_md_velocities(evc_ctx *ctx, int nbatch, int natm, double dt, int first,
                      const double *inv_mass /* [natm] */, const double *mass /* [natm] */,
                      const double *grad, const double *x, const double *epot, double *v,
                      double *a, double *ekin, int *frame_idx, int max_frames, double *traj,
                      double *epot_log, double *ekin_log);

/* ---- K9g: AO integrals over contracted Cartesian s AND p Gaussians ---------------------
 * The same arrays as evc_ao_integrals_s for molecules with p shells (6-31G oxygen: the H2O
 * and Zundel configurations of the reference, scripts/MD/md_H2O_6_31G_FCI.py,
 * scripts/MD/Zundel_thermodynamics), McMurchie-Davidson scheme.  AOs are contracted
 * Cartesian functions in pyscf.gto order (per atom: s shells, then p shells, components
 * x, y, z); ao_pow_host [nao][3] holds the Cartesian powers (all zero: s; one 1: p); the
 * three components of a p shell repeat its primitives.  prim_wt = contraction coefficient x
 * primitive norm ((2a/pi)^(3/4), times 2 sqrt(a) for p) x contracted normalisation.
 * natm <= 16, nao <= 64.  Workspace: evc_ao_integrals_sp_workspace_bytes. */
typedef struct evc_gbasis evc_gbasis;
int evc_gbasis_create(evc_ctx *ctx, int natm, const double *charges_host, int nao,
                      const int32_t *ao_atom_host, const int32_t *ao_pow_host,
                      const int32_t *ao_nprim_host, const double *prim_exp_host,
                      const double *prim_wt_host, evc_gbasis **out);
int evc_gbasis_destroy(evc_gbasis *basis);
int evc_ao_integrals_sp_workspace_bytes(const evc_gbasis *basis, int nbatch, size_t *bytes);
int evc_ao_integrals_sp(evc_ctx *ctx, const evc_gbasis *basis, int nbatch, const double *coords,
                        double *ovlp, double *hcore, double *eri, double *ipovlp,
                        double *hcore_deriv, double *eri_ip1, double *e_nuc, double *grad_nuc,
                        void *workspace, size_t workspace_bytes);

/* ---- f4: observables of the predicted one-body density matrix ---------------------------
 * What the reference's MD callback evaluates from scanner.base.predicted_one_rdm
 * (scripts/MD/Zundel_thermodynamics/continuation/04_Zundel_continuation_MD.py:71-92 dip_moment,
 * :140-159 callback), batched over geometries.  The AO table is the basis description of
 * evc_gbasis_create (ao_pow_host may be NULL: s functions only) plus the atomic masses.
 *   evc_center_of_mass     origin [G][3] from coords [G][natm][3]
 *   evc_int1e_r            r [G][3][n][n] = <i| r - origin |j>: mol.intor_symmetric("int1e_r", comp=3)
 *                          inside mol.with_common_orig(origin)
 *   evc_rdm1_observables   dm_ao = X gamma X^T (written when dm_ao != NULL), dipole [G][3] =
 *                          sum_A Z_A (R_A - origin) - sum_ij r_ij dm_ji (atomic units), atom_charges
 *                          [G][natm] = Z_A - sum_{mu on A} n_mu with method 0: Mulliken n_mu = (dm S)_mu,mu
 *                          (needs ovlp), 1: Loewdin n_mu = gamma_mu,mu (x symmetric: x = S^-1/2). */
typedef struct evc_aotable evc_aotable;
int evc_aotable_create(evc_ctx *ctx, int natm, const double *charges_host, const double *masses_host,
                       int nao, const int32_t *ao_atom_host, const int32_t *ao_pow_host,
                       const int32_t *ao_nprim_host, const double *prim_exp_host,
                       const double *prim_wt_host, evc_aotable **out);
int evc_aotable_destroy(evc_aotable *table);
int evc_center_of_mass(evc_ctx *ctx, const evc_aotable *table, int nbatch, const double *coords,
                       double *origin);
int evc_int1e_r(evc_ctx *ctx, const evc_aotable *table, int nbatch, const double *coords,
                const double *origin, double *r);
int evc_rdm1_observables(evc_ctx *ctx, const evc_aotable *table, int nbatch, int method,
                         const double *coords, const double *origin, const double *x,
                         const double *gamma, const double *ovlp, const double *rint, double *dm_ao,
                         double *dipole, double *atom_charges);

/* ---- FCI Hamiltonian action (training side) ------------------------------------------
 * What cisolver.kernel (evcont/FCI_EVCont.py:70; PySCF direct_spin0 Davidson:
 * contract_2e + make_hdiag) needs from the Hamiltonian, on the device with the K0 link
 * tables (string-major packing for BOTH spins here).
 * evc_fci_hdiag: <K|H|K> for all determinants from the occupation strings (int64 bit
 *   strings as evc_make_strings_host returns them, device copies), h1 [n][n], eri [n^4].
 * evc_fci_contract_2e: sigma = H c.  h1eff: [n2p], n2p = n^2 rounded up to even, holding
 *   h'_ps = h_ps - 1/2 sum_q (pq|qs) row-major, zero padded; w2: [n2p][n2p] with
 *   w2[(rs)][(pq)] = 1/2 (pq|rs); civec, sigma: [na*nb].  D[K,(rs)] = <K|E_rs|c> is
 *   gathered through the links, G = D w2 runs on the FP64 tensor cores, sigma gathers G
 *   back through the links; deterministic. */
int evc_fci_hdiag(evc_ctx *ctx, int norb, int64_t na, int64_t nb, const int64_t *strs_a,
                  const int64_t *strs_b, const double *h1, const double *eri, double *hdiag);
int evc_fci_contract_workspace_bytes(int norb, int64_t na, int64_t nb, size_t *bytes);
int evc_fci_contract_2e(evc_ctx *ctx, int norb, int64_t na, int64_t nb,
                        const uint64_t *link_a /* string-major */, int nlink_a,
                        const uint64_t *link_b /* string-major */, int nlink_b,
                        const double *h1eff, const double *w2, const double *civec,
                        double *sigma, void *workspace, size_t workspace_bytes);

/* ---- transform_ci: CI vector in a rotated one-particle basis --------------------------
 * Replaces pyscf.fci.addons.transform_ci(ci, nelec, u), called at evcont/FCI_EVCont.py:79-85
 * with u = basis^T S basis_oao (rows: old orbitals, columns: new orbitals, [norb][norb]
 * row-major, device):  ci_new = Ta^T ci Tb,  T[I][J] = det(u[occ(I), occ(J)]) over the
 * occupation strings of each spin (device int64, ascending, as evc_make_strings_host
 * returns them).  One thread per minor (partial-pivoting elimination), the two products on
 * the FP64 tensor cores.  ci_in / ci_out: [na][nb] contiguous; at most 12 electrons per spin. */
int evc_transform_ci_workspace_bytes(int norb, int64_t na, int64_t nb, size_t *bytes);
int evc_transform_ci(evc_ctx *ctx, int norb, int nelec_a, int nelec_b, int64_t na, int64_t nb,
                     const int64_t *strs_a, const int64_t *strs_b, const double *u,
                     const double *ci_in, double *ci_out, void *workspace,
                     size_t workspace_bytes);

/* ---- closed-shell Fock matrix ------------------------------------------------------------
 * The device part of the RHF behind get_basis(mol, "canonical")
 * (evcont/electron_integral_utils.py:103-106: scf.RHF(mol).scf() -> mo_coeff; the reference's
 * default cibasis):  F = hcore + J - K/2 with J[p,q] = sum_rs (pq|rs) D[r,s] and
 * K[p,s] = sum_qr (pq|rs) D[q,r]; hcore, dm, fock: [n][n]; eri: [n^4] chemists' notation. */
int evc_fock_rhf(evc_ctx *ctx, int n, const double *hcore, const double *eri, const double *dm,
                 double *fock);

/* ---- farthest-point selection in Hamiltonian space ----------------------------------------
 * converge_EVCont_MD(..., data_addition="farthest_point_ham") (evcont/MD_utils.py:363-405):
 * dmin[g] = min_t ( |h1_g - h1_t|^2 + 1/2 |h2_g - h2_t|^2 ) over the training rows;
 * frames: [nframes][len_total], train: [ntrain][len_total], each row h1 (len_one doubles)
 * followed by h2, OAO basis, device. */
int evc_min_sqdist(evc_ctx *ctx, int nframes, int ntrain, int64_t len_one, int64_t len_total,
                   const double *frames, const double *train, double *dmin);

#ifdef __cplusplus
}
#endif
#endif /* EVCONT_B200_H */
