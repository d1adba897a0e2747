// Internal helpers shared by the translation units of libevcont_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "evcont_b200.h"

enum {
  EVC_STAGE_LOEWDIN = 0,
  EVC_STAGE_AO2OAO,
  EVC_STAGE_SUBSPACE_H,
  EVC_STAGE_GENEIG,
  EVC_STAGE_PREDICT,
  EVC_STAGE_GRAD,         // K8 (full path) / K8a: per-geometry GEMMs of the gradient (packed path)
  EVC_STAGE_GRAD_STREAM,  // K8b: streaming contraction of the derivative integrals (packed path)
  EVC_NSTAGE
};

struct evc_ctx {
  int device;
  cudaStream_t stream;
  int sm_count;
  size_t smem_optin;
  double last_trdm_flops;
  int trdm_plan_pairs;  // evc_trans_rdm12_plan_pairs: pair count the alpha-slice plan is made for (0: the call's own)
  // optional per-stage timing of evc_energy_with_grad (evc_ctx_stage_timing)
  int stage_timing;
  cudaEvent_t stage_ev[EVC_NSTAGE + 1];
  double stage_ms[EVC_NSTAGE];
  long long stage_calls;
  int stage_pending;
  // copy/compute pipeline of evc_energy_with_grad_host (created on first use)
  int pipe_ready;
  cudaStream_t h2d_stream, d2h_stream;
  cudaEvent_t ev_h2d[2], ev_compute[2], ev_d2h[2], ev_start;
};

// number of kernel launches issued by this library (all contexts of the process)
extern unsigned long long g_evc_launches;

void evc_set_error(const char* fmt, ...);
extern "C" int evc_stage_mark(evc_ctx* ctx, int stage);

// internal cross-file helpers (not part of the C ABI)
int evc_launch_rot_pass(cudaStream_t st, int nbatch, int n, const double* in, const double* M,
                        int transpose_m, double* out);
int evc_launch_geneig(evc_ctx* ctx, int nbatch, int N, int packed_lower, const double* H,
                      const double* Linv, int nroots, double* E, double* C);
int evc_launch_geneig_lowest(evc_ctx* ctx, int nbatch, int N, int packed_lower, const double* H,
                             const double* Linv, double* E, double* C);
size_t evc_rows_dot_ws_bytes(int64_t L, int P, int G);
int evc_rows_dot(evc_ctx* ctx, const double* rows, int64_t L, int P, const double* hv, int G, double* out,
                 void* workspace, size_t workspace_bytes);
size_t evc_rows_axpy_ws_bytes(int64_t L, int P, int G);
int evc_rows_axpy(evc_ctx* ctx, const double* rows, int64_t L, int P, const double* w, int G, double* out,
                  void* workspace, size_t workspace_bytes);
// packed (8-fold symmetric) prediction step pieces, packed.cu
constexpr int kPackedMaxNorb = 13;      // per-geometry shared-memory kernels up to this many orbitals
constexpr int kPackedPipeMaxNorb = 10;  // persistent warp-specialised kernels (packed_pipe.cu) up to this many
// erip: [nbatch][np][pA] packed (ab|cd); Tout: same layout; eri_ip1p: [nbatch][3][n][n][np]
int evc_packed_ao2oao(evc_ctx* ctx, int nbatch, int n, const double* x, const double* hcore,
                      const double* erip, double* hvec, double* Tout);
int evc_packed_grad(evc_ctx* ctx, int nbatch, int n, int natm, const int32_t* aoslices, const double* x,
                    const double* evals, const double* evecs, const double* hcore, const double* Tin,
                    const double* out7, const double* ipovlp, const double* hcore_deriv,
                    const double* eri_ip1p, const double* grad_nuc, double* Wg, double* OmS, double* Pao,
                    double* grad);
// packed_pipe.cu
bool evc_packed_pipe_supported(int n);
// geneig_reg.cu: K6 (lowest root) with the eigensolver of one problem in the registers of one warp (2 <= N <= 24)
bool evc_geneig_reg_supported(int N);
int evc_geneig_reg(evc_ctx* ctx, int nbatch, int N, int packed_lower, const double* H, const double* Linv, double* E,
                   double* C);
// loewdin_reg.cu: K3 with the Jacobi eigensolver in registers (2 <= n <= 16), several matrices per warp
bool evc_loewdin_reg_supported(int n);
int evc_loewdin_reg_min_batch();
int evc_loewdin_reg(evc_ctx* ctx, int nbatch, int n, const double* s_ao, double* x, double* evals, double* evecs);
int evc_packed_ao2oao_pipe(evc_ctx* ctx, int nbatch, int n, const double* x, const double* hcore,
                           const double* erip, double* hvec, double* Tout);
int evc_packed_grad_pipe(evc_ctx* ctx, int nbatch, int n, const double* x, const double* evals,
                         const double* evecs, const double* hcore, const double* Timg, const double* out7,
                         double* Wg, double* OmS, double* Pao);
// K8 on full (n^4) arrays with the nuclear gradient added (grad.cu); sym8 != 0: Gamma carries the 8-fold
// permutational symmetry of the integrals (packed step), which turns the symmetrising gathers into streams
int evc_grad_elec_full(evc_ctx* ctx, int nbatch, int n, int natm, const int32_t* aoslices,
                       const double* evals, const double* evecs, const double* x, const double* hcore,
                       const double* t3, const double* gamma, const double* Gamma, const double* ipovlp,
                       const double* hcore_deriv, const double* eri_ip1, const double* grad_nuc,
                       double* grad, void* workspace, size_t workspace_bytes, int sym8);
int evc_packed_hvec_from_full(evc_ctx* ctx, int nbatch, int n, const double* h1, const double* h2, double* hvec);
int evc_packed_unpack_rdms(evc_ctx* ctx, int nbatch, int n, const double* out7, double* gamma, double* Gamma8);
int evc_packed_pair_weights(evc_ctx* ctx, int nbatch, int N, const double* C, int64_t c_stride, double* w);

#define EVC_CHECK_CUDA(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      evc_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,              \
                    cudaGetErrorString(_e));                                        \
      return -2;                                                                    \
    }                                                                               \
  } while (0)

#define EVC_REQUIRE(cond, ...)                                                      \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      evc_set_error(__VA_ARGS__);                                                   \
      return -1;                                                                    \
    }                                                                               \
  } while (0)

#define EVC_CHECK_LAUNCH()              \
  do {                                  \
    ++g_evc_launches;                   \
    EVC_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

static inline size_t evc_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller-provided workspace
struct evc_arena {
  char* base;
  size_t size;
  size_t used;
  evc_arena(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = evc_align_up(count * sizeof(T), 256);
    if (used + bytes > size) return nullptr;
    T* r = reinterpret_cast<T*>(base + used);
    used += bytes;
    return r;
  }
};

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
