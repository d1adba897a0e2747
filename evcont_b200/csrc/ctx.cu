// Context, error reporting.
#include "common.cuh"

static thread_local char g_err[1024] = "";

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
  *out = c;
  return 0;
}

int evc_ctx_destroy(evc_ctx* ctx) {
  delete ctx;
  return 0;
}

int evc_ctx_set_stream(evc_ctx* ctx, void* stream) {
  EVC_REQUIRE(ctx != nullptr, "evc_ctx_set_stream: ctx is NULL");
  ctx->stream = static_cast<cudaStream_t>(stream);
  return 0;
}

int evc_ctx_sm_count(const evc_ctx* ctx) { return ctx ? ctx->sm_count : -1; }

}  // extern "C"
