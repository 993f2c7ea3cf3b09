// vimure_b200 -- the extern "C" entry points of include/vimure_b200.h.
// The kernels are instantiated per K in several translation units (csrc/vm_kernels.cu compiled once per TU by
// vimure_b200/build.py); every entry point routes to the TU that owns ctx->K through its function table.
#include <stdio.h>

#include "vm_common.cuh"
#include "vm_tu_list.h"  // generated: declarations of vm_tu_api_<id>() and vm_tu_for_K()

static const vm_tu_api* tu(const vm_ctx* c) {
  if (!c || c->K < 2 || c->K > VM_MAX_K) return nullptr;
  return vm_tu_for_K((int)c->K);
}
#define VM_ROUTE(call)              \
  const vm_tu_api* t = tu(c);       \
  if (!t) return VM_EINVAL;         \
  return t->call

extern "C" int64_t vm_ctx_size(void) { return (int64_t)sizeof(vm_ctx); }
extern "C" int64_t vm_abi_version(void) { return VM_ABI_VERSION; }
extern "C" int64_t vm_dense_tile_w(int64_t K) { return vm_dense_tile_w_host(K); }

extern "C" int vm_materialize_prior(const vm_ctx* c, void* stream) { VM_ROUTE(materialize_prior(c, stream)); }
extern "C" int vm_refresh_cache(const vm_ctx* c, void* stream) { VM_ROUTE(refresh_cache(c, stream)); }
extern "C" int vm_init_stats(const vm_ctx* c, void* stream) { VM_ROUTE(init_stats(c, stream)); }
extern "C" int vm_phase_gamma(const vm_ctx* c, void* stream) { VM_ROUTE(phase_gamma(c, stream)); }
extern "C" int vm_phase_phi(const vm_ctx* c, void* stream) { VM_ROUTE(phase_phi(c, stream)); }
extern "C" int vm_phase_rho(const vm_ctx* c, int flags, void* stream) { VM_ROUTE(phase_rho(c, flags, stream)); }
extern "C" int vm_dense_only(const vm_ctx* c, int flags, void* stream) { VM_ROUTE(dense_only(c, flags, stream)); }
extern "C" int vm_phase_finish(const vm_ctx* c, int flags, void* stream) { VM_ROUTE(phase_finish(c, flags, stream)); }
extern "C" int vm_iteration(const vm_ctx* c, int flags, void* stream) { VM_ROUTE(iteration(c, flags, stream)); }
extern "C" int vm_run(const vm_ctx* c, int n_iter, int flags, int last_flags, void* stream) {
  VM_ROUTE(run(c, n_iter, flags, last_flags, stream));
}
extern "C" int vm_infer(const vm_ctx* c, int mode, double threshold, uint8_t* out, void* stream) {
  VM_ROUTE(infer(c, mode, threshold, out, stream));
}

__global__ void k_test_special(const double* x, double* dg, double* lg, int64_t n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  dg[t] = vm_digamma(x[t]);
  lg[t] = lgamma(x[t]);
}
extern "C" int vm_test_special(const double* x, double* dg, double* lg, int64_t n, void* stream) {
  if (n <= 0) return 0;
  k_test_special<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, dg, lg, n);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
