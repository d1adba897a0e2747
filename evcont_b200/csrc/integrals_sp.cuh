// Shared declarations of the s+p AO-integral engine K9g (integrals_sp.cu: tables, one-electron part, C ABI;
// gclass.cu: the two-electron class kernels, compiled in several parts).
#pragma once
#include "common.cuh"

namespace evc_gint {

constexpr int kGTop = 11;            // Boys table: F_11(T0) and exp(-T0) on the grid T0 = i / 64
constexpr int kGPerUnit = 64;
constexpr int kGTmax = 32;
constexpr int kGBoysN = kGTmax * kGPerUnit + 1;
constexpr int kGThreads = 256;
constexpr int kGMaxAtoms = 16;
constexpr int kGMaxL = 6;            // highest Boys order: (pp|pp) with one derivative = 5 (+1 spare)

constexpr int kGAux = 7;       // side streams of a basis (small batches: concurrent class kernels)
constexpr int kGClasses = 6;   // ssss, psss, ppss, psps, ppps, pppp (canonical shell-quartet classes)

struct GView {
  int natm, nao;
  const int32_t *ao_atom, *ao_pow, *ao_poff;
  const int32_t *sh_atom, *sh_ao0, *sh_p0, *sh_np;
  const double *prim_exp, *prim_wt, *charges, *boys;
};


struct GOut {
  double *ovlp, *hcore, *eri, *ipovlp, *vtmp, *eri_ip1, *e_nuc, *grad_nuc;
};

// Every launcher takes a ring of `nst` streams and puts its i-th kernel on sts[(k + i) % nst] (nst = 1: one stream).
// launches the two-electron class kernels of compilation part `part` (0: ssss, psss, ppss, psps; 1: ppps;
// 2..4: pppp with the component of the second function fixed) for the unit lists of the basis
int launch_gclass_part0(const cudaStream_t* sts, int nst, int k, int sm_count, int nbatch, const GView& v, const int32_t* cq, const int32_t* cunits,
                        const int* cunit_off, const double* coords, const GOut& o);
int launch_g1e(const cudaStream_t* sts, int nst, int k, int sm_count, int nbatch, const GView& v, const int32_t* plist, const int* p_off,
               const double* coords, const GOut& o);
int launch_gclass_part1(const cudaStream_t* sts, int nst, int k, int sm_count, int nbatch, const GView& v, const int32_t* cq, const int32_t* cunits,
                        const int* cunit_off, const double* coords, const GOut& o);
int launch_gclass_part2(const cudaStream_t* sts, int nst, int k, int sm_count, int nbatch, const GView& v, const int32_t* cq, const int32_t* cunits,
                        const int* cunit_off, const double* coords, const GOut& o);
int launch_gclass_part3(const cudaStream_t* sts, int nst, int k, int sm_count, int nbatch, const GView& v, const int32_t* cq, const int32_t* cunits,
                        const int* cunit_off, const double* coords, const GOut& o);
int launch_gclass_part4(const cudaStream_t* sts, int nst, int k, int sm_count, int nbatch, const GView& v, const int32_t* cq, const int32_t* cunits,
                        const int* cunit_off, const double* coords, const GOut& o);

}  // namespace evc_gint

struct evc_gbasis {
  int natm, nao, nprim, nshell;
  int32_t *ao_atom, *ao_pow, *ao_poff, *aoslices;  // ao_pow: [nao][3]
  int32_t *sh_atom, *sh_ao0, *sh_p0, *sh_np;       // shells (s: one AO, p: three consecutive AOs)
  int32_t *cq, *cunits;                            // shell-quartet work lists of the class kernels
  int32_t *plist;                                  // ordered shell pairs of the one-electron classes
  int p_off[5];
  cudaStream_t aux[evc_gint::kGAux];               // side streams: the class kernels of a SMALL batch run concurrently
  cudaEvent_t ev_fork, ev_join[evc_gint::kGAux];
  int cq_off[evc_gint::kGClasses + 1], cunit_off[evc_gint::kGClasses + 1];
  double *prim_exp, *prim_wt, *charges, *boys;
};

