// Store-pattern probe for the dense kernel (K = 2 geometry of config 3: 20000 rows of 160000 bytes, 3.2 GB):
// which lane/warp/CTA -> address mapping of a TILED writer reaches the bandwidth of a linear one, and does a TMA bulk
// store (cp.async.bulk.global.shared::cta) from a shared-memory stage beat STG.128 from registers?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu && ./store_patterns
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ROWS 20000
#define ROWB 160000ull  // bytes per row
#define THREADS 256

// linear grid-stride writer
__global__ void __launch_bounds__(THREADS, 4) k_linear(float4* p, size_t n4, float v) {
  const float4 val = make_float4(v, v, v, v);
  for (size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x; i < n4; i += (size_t)gridDim.x * THREADS) p[i] = val;
}

// tiled writer: CTA = tile of TH rows x (NWC * SEG) bytes; the CTA's 8 warps are arranged NWC across the columns and
// 8/NWC down the rows; a warp writes SEG contiguous bytes of a row (STG.128, 512 contiguous bytes per instruction),
// then goes to its next row.
template <int SEG, int NWC, int TH>
__global__ void __launch_bounds__(THREADS, 4) k_tiled(char* p, int nct, float v) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wc = warp % NWC, wr = warp / NWC, NWR = 8 / NWC;
  const int ct = blockIdx.x % nct, rt = blockIdx.x / nct;
  const size_t col0 = ((size_t)ct * NWC + wc) * SEG;
  if (col0 + SEG > ROWB) return;
  const float4 val = make_float4(v, v, v, v);
  for (int r = wr; r < TH; r += NWR) {
    const int row = rt * TH + r;
    if (row >= ROWS) break;
    char* dst = p + (size_t)row * ROWB + col0;
#pragma unroll
    for (int b = 0; b < SEG; b += 512) *reinterpret_cast<float4*>(dst + b + lane * 16) = val;
  }
}

// TMA bulk-store writer: same tiling with NWC = 1; a warp fills a SEG-byte stage in shared memory (two stages per warp)
// and one lane issues cp.async.bulk.global.shared::cta of the whole segment.
template <int SEG, int TH>
__global__ void __launch_bounds__(THREADS, 4) k_tma(char* p, int nct, float v) {
  extern __shared__ __align__(128) char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ct = blockIdx.x % nct, rt = blockIdx.x / nct;
  const size_t col0 = (size_t)ct * SEG;
  if (col0 + SEG > ROWB) return;
  const float4 val = make_float4(v, v, v, v);
  char* stage = smem + (size_t)warp * 2 * SEG;
  int buf = 0;
  for (int r = warp; r < TH; r += 8) {
    const int row = rt * TH + r;
    if (row >= ROWS) break;
    char* s = stage + buf * SEG;
    // the bulk store that last read this stage must have finished reading it
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int b = 0; b < SEG; b += 512) *reinterpret_cast<float4*>(s + b + lane * 16) = val;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      char* dst = p + (size_t)row * ROWB + col0;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                   "r"((uint32_t)__cvta_generic_to_shared(s)), "r"(SEG)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    buf ^= 1;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const size_t bytes = (size_t)ROWS * ROWB;
  char* p;
  cudaMalloc(&p, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
#define TIME(name, stmt)                                                                                  \
  for (int w = 0; w < 2; ++w) { stmt; }                                                                    \
  cudaEventRecord(e0);                                                                                     \
  for (int r = 0; r < 10; ++r) { stmt; }                                                                   \
  cudaEventRecord(e1);                                                                                     \
  cudaEventSynchronize(e1);                                                                                \
  cudaEventElapsedTime(&ms, e0, e1);                                                                       \
  printf("%-58s %8.3f ms  %8.1f GB/s  (%s)\n", name, ms / 10, bytes / (ms / 10 * 1e-3) / 1e9,             \
         cudaGetErrorString(cudaGetLastError()));
  TIME("cudaMemsetAsync", cudaMemsetAsync(p, 0, bytes));
  TIME("linear STG.128 grid=148*16", (k_linear<<<148 * 16, THREADS>>>((float4*)p, bytes / 16, 1.f)));
  TIME("linear STG.128 grid=148*4", (k_linear<<<148 * 4, THREADS>>>((float4*)p, bytes / 16, 1.f)));
#define TILED(SEG, NWC, TH)                                                                  \
  {                                                                                          \
    const int nct = (int)((ROWB + (size_t)SEG * NWC - 1) / ((size_t)SEG * NWC));             \
    const int nrt = (ROWS + TH - 1) / TH;                                                    \
    TIME("tiled SEG=" #SEG " warps across=" #NWC " rows/tile=" #TH, (k_tiled<SEG, NWC, TH><<<nct * nrt, THREADS>>>(p, nct, 1.f))); \
  }
  TILED(4096, 1, 128)   // the dense kernel today
  TILED(4096, 1, 64)
  TILED(4096, 1, 256)
  TILED(4096, 2, 64)
  TILED(4096, 4, 32)
  TILED(4096, 8, 16)
  TILED(4096, 8, 32)
  TILED(4096, 8, 64)
  TILED(8192, 1, 64)
  TILED(8192, 1, 128)
  TILED(16384, 1, 32)
  TILED(16384, 1, 64)
  TILED(2048, 1, 128)
  TILED(2048, 8, 32)
  TILED(1024, 8, 64)
#define TMA(SEG, TH)                                                                                        \
  {                                                                                                         \
    const int nct = (int)((ROWB + SEG - 1) / SEG);                                                          \
    const int nrt = (ROWS + TH - 1) / TH;                                                                   \
    cudaFuncSetAttribute(k_tma<SEG, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * SEG);         \
    TIME("TMA bulk store SEG=" #SEG " rows/tile=" #TH, (k_tma<SEG, TH><<<nct * nrt, THREADS, 8 * 2 * SEG>>>(p, nct, 1.f))); \
  }
  TMA(4096, 128)
  TMA(4096, 64)
  TMA(2048, 128)
  TMA(1024, 128)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
