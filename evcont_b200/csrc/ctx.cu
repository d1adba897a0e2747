// Context, error reporting.
#include "common.cuh"

static thread_local char g_err[1024] = "";
unsigned long long g_evc_launches = 0;

void evc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

int evc_abi_version(void) { return EVC_ABI_VERSION; }
const char* evc_last_error(void) { return g_err; }

int evc_ctx_create(int device, void* stream, evc_ctx** out) {
  EVC_REQUIRE(out != nullptr, "evc_ctx_create: out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    evc_set_error("evc_ctx_create: no CUDA device available (%s); libevcont_b200 has no CPU fallback",
                  cudaGetErrorString(e));
    return -3;
  }
  EVC_REQUIRE(device >= 0 && device < ndev, "evc_ctx_create: device %d out of range [0,%d)", device, ndev);
  EVC_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  EVC_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  EVC_REQUIRE(prop.major >= 10,
              "evc_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
              device, prop.major, prop.minor);
  evc_ctx* c = new evc_ctx;
  c->device = device;
  c->stream = static_cast<cudaStream_t>(stream);
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->last_trdm_flops = 0.0;
  c->trdm_plan_pairs = 0;
  c->stage_timing = 0;
  c->stage_calls = 0;
  c->stage_pending = 0;
  c->pipe_ready = 0;
  for (int k = 0; k < EVC_NSTAGE; ++k) c->stage_ms[k] = 0.0;
  for (int k = 0; k <= EVC_NSTAGE; ++k) {
    cudaError_t ee = cudaEventCreate(&c->stage_ev[k]);
    if (ee != cudaSuccess) {
      evc_set_error("evc_ctx_create: cudaEventCreate failed: %s", cudaGetErrorString(ee));
      delete c;
      return -2;
    }
  }
  *out = c;
  return 0;
}

int evc_ctx_destroy(evc_ctx* ctx) {
  if (ctx) {
    for (int k = 0; k <= EVC_NSTAGE; ++k) cudaEventDestroy(ctx->stage_ev[k]);
    if (ctx->pipe_ready) {
      cudaStreamDestroy(ctx->h2d_stream);
      cudaStreamDestroy(ctx->d2h_stream);
      for (int k = 0; k < 2; ++k) {
        cudaEventDestroy(ctx->ev_h2d[k]);
        cudaEventDestroy(ctx->ev_compute[k]);
        cudaEventDestroy(ctx->ev_d2h[k]);
      }
      cudaEventDestroy(ctx->ev_start);
    }
  }
  delete ctx;
  return 0;
}

int evc_ctx_set_stream(evc_ctx* ctx, void* stream) {
  EVC_REQUIRE(ctx != nullptr, "evc_ctx_set_stream: ctx is NULL");
  ctx->stream = static_cast<cudaStream_t>(stream);
  return 0;
}

int evc_ctx_sm_count(const evc_ctx* ctx) { return ctx ? ctx->sm_count : -1; }

unsigned long long evc_launch_count(void) { return g_evc_launches; }

// fold the events of the last timed evc_energy_with_grad call into stage_ms
static int evc_collect_stage_times(evc_ctx* ctx) {
  if (!ctx->stage_pending) return 0;
  EVC_CHECK_CUDA(cudaEventSynchronize(ctx->stage_ev[EVC_NSTAGE]));
  for (int k = 0; k < EVC_NSTAGE; ++k) {
    float ms = 0.f;
    EVC_CHECK_CUDA(cudaEventElapsedTime(&ms, ctx->stage_ev[k], ctx->stage_ev[k + 1]));
    ctx->stage_ms[k] += ms;
  }
  ctx->stage_calls += 1;
  ctx->stage_pending = 0;
  return 0;
}

int evc_stage_mark(evc_ctx* ctx, int stage) {
  if (!ctx->stage_timing) return 0;
  if (stage == 0) {
    int rc = evc_collect_stage_times(ctx);
    if (rc) return rc;
  }
  EVC_CHECK_CUDA(cudaEventRecord(ctx->stage_ev[stage], ctx->stream));
  if (stage == EVC_NSTAGE) ctx->stage_pending = 1;
  return 0;
}

int evc_ctx_stage_timing(evc_ctx* ctx, int enable) {
  EVC_REQUIRE(ctx != nullptr, "evc_ctx_stage_timing: ctx is NULL");
  if (ctx->stage_timing) {
    int rc = evc_collect_stage_times(ctx);
    if (rc) return rc;
  }
  ctx->stage_timing = enable ? 1 : 0;
  if (enable) {
    for (int k = 0; k < EVC_NSTAGE; ++k) ctx->stage_ms[k] = 0.0;
    ctx->stage_calls = 0;
  }
  return 0;
}

int evc_ctx_stage_times(evc_ctx* ctx, double* ms, int64_t* calls) {
  EVC_REQUIRE(ctx && ms && calls, "evc_ctx_stage_times: NULL argument");
  int rc = evc_collect_stage_times(ctx);
  if (rc) return rc;
  for (int k = 0; k < EVC_NSTAGE; ++k) ms[k] = ctx->stage_ms[k];
  *calls = ctx->stage_calls;
  return 0;
}

}  // extern "C"
