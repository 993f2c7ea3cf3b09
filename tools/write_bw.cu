// Pure-write bandwidth probe: how fast can a B200 write a 3.2 GB fp32 buffer? (calibration for the dense kernel)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

template <int MODE>
__global__ void k_write(float4* p, size_t n4, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float4 val = make_float4(v, v, v, v);
  for (; i < n4; i += stride) {
    if (MODE == 0) p[i] = val;
    if (MODE == 1) __stcs(&p[i], val);
    if (MODE == 2) __stwt(&p[i], val);
    if (MODE == 3) __stcg(&p[i], val);
  }
}
// each thread writes 2 consecutive float4 (32 B), like the dense kernel with K=2
__global__ void k_write32(float4* p, size_t n4, float v) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 2;
  const float4 val = make_float4(v, v, v, v);
  for (; i + 1 < n4; i += stride) {
    p[i] = val;
    p[i + 1] = val;
  }
}
__global__ void k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) b[i] = a[i];
}

int main() {
  const size_t bytes = 3200000000ull;
  const size_t n4 = bytes / 16;
  float4 *p, *q;
  cudaMalloc(&p, bytes);
  cudaMalloc(&q, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
#define TIME(name, stmt, nbytes)                                         \
  for (int w = 0; w < 2; ++w) { stmt; }                                   \
  cudaEventRecord(e0);                                                    \
  for (int r = 0; r < 10; ++r) { stmt; }                                  \
  cudaEventRecord(e1);                                                    \
  cudaEventSynchronize(e1);                                               \
  cudaEventElapsedTime(&ms, e0, e1);                                      \
  printf("%-40s %8.3f ms  %8.1f GB/s\n", name, ms / 10, (nbytes) / (ms / 10 * 1e-3) / 1e9);
  TIME("cudaMemsetAsync", cudaMemsetAsync(p, 0, bytes), (double)bytes);
  int grids[] = {148 * 4, 148 * 8, 148 * 16, 148 * 32, 148 * 64};
  for (int g : grids) {
    char nm[64];
    snprintf(nm, 64, "st.global      grid=%d x256", g);
    TIME(nm, (k_write<0><<<g, 256>>>(p, n4, 1.f)), (double)bytes);
  }
  TIME("st.global.cs   grid=148*16", (k_write<1><<<148 * 16, 256>>>(p, n4, 1.f)), (double)bytes);
  TIME("st.global.wt   grid=148*16", (k_write<2><<<148 * 16, 256>>>(p, n4, 1.f)), (double)bytes);
  TIME("st.global.cg   grid=148*16", (k_write<3><<<148 * 16, 256>>>(p, n4, 1.f)), (double)bytes);
  TIME("2x float4 / thread grid=148*16", (k_write32<<<148 * 16, 256>>>(p, n4, 1.f)), (double)bytes);
  TIME("copy (r+w)     grid=148*16", (k_copy<<<148 * 16, 256>>>(p, q, n4)), 2.0 * bytes);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
