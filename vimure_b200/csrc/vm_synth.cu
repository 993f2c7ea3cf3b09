// vimure_b200 -- device-side synthetic reports.
//
// Samples the observed network X of the reference's `_build_X` under the self-reporter (ego) mask
// (`synthetic.py:138-209`, mask `synthetic.py:1184-1204`) for ONE node-row block, sparsely and without a sequential RNG:
// every draw is a pure function of (seed, layer, reporter m, partner n), so any rank can evaluate any (reporter, pair) and
// all ranks that do agree bit for bit.  A rank therefore generates exactly the entries it needs -- X[l,i,j,m] with i in its
// rows, plus (optionally) the reciprocal entries X[l,j,i,m] whose row j lives on another rank -- with nothing exchanged.
//
// Law (per layer l, reporter m with reliability theta_lm, partner node n != m; lambda_ab = Y_lab if Y_lab > 0 else 0.01):
//   a fair coin picks the first direction (a->b); x_first ~ Poisson((theta lambda_ab + eta theta lambda_ba)/(1-eta^2)),
//   x_second ~ Poisson(theta lambda_ba + eta x_first)                                  (synthetic.py:170-190)
//   the self tie (m,m): the same two-step draw with both directions equal to (m,m) -- the reference's loop visits it with
//   i == j, so the second assignment overwrites the first: x ~ Poisson(theta 0.01 + eta Poisson(theta 0.01/(1-eta))).
// Pairs without a true tie in either direction have the tiny base mean theta*0.01 both ways: for them a single uniform
// decides "both counts are zero" (98-99 % of the candidates) before anything else is computed.
//
// The stream is NOT numpy's: parity is always checked on identical inputs, never on regenerated ones; the law is
// checked by moments (tests/test_gpu_synth.py) against the reference generator's.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "vimure_b200.h"

namespace {

__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
// uniform on (0,1) from 32 random bits
__device__ __forceinline__ double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }

// Poisson(mu) by inversion of one uniform (means here are O(1): a handful of steps); conditioned on >= 1 if `zt`
__device__ __forceinline__ int poisson_inv(double mu, double u, bool zt) {
  if (mu <= 0.0) return zt ? 1 : 0;
  double p = exp(-mu);
  if (zt) u = p + u * (1.0 - p);  // uniform on (P(0), 1)
  double F = p;
  int k = 0;
  while (u > F && k < 4096) {
    ++k;
    p *= mu / (double)k;
    F += p;
    if (p < 1e-300) break;
  }
  return (zt && k == 0) ? 1 : k;
}

// lambda of the directed tie (l,a,b): its true value Y_lab, or 0 when there is no true tie
__device__ __forceinline__ int y_lookup(const int64_t* __restrict__ yk, const int32_t* __restrict__ yv, int64_t nY,
                                        int64_t key) {
  int64_t lo = 0, hi = nY;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (yk[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return (lo < nY && yk[lo] == key) ? yv[lo] : 0;
}

// average interactions of a tie whose ground truth is category y: the caller's table, or the reference generator's
// default (0.01 for "no tie", y itself otherwise; synthetic.py:140-157)
__device__ __forceinline__ double lam_of(const vm_synth& s, int l, int y) {
  if (s.lam) return s.lam[(int64_t)l * s.K + (y < (int)s.K ? y : (int)s.K - 1)];
  return y > 0 ? (double)y : 0.01;
}

struct Emit {
  const vm_synth s;
  __device__ __forceinline__ void operator()(int l, int i, int j, int m, int x) const {
    if (x <= 0) return;
    const bool own = i >= s.row0 && i < s.row0 + s.nloc;
    const bool tr = s.emit_transposed && !own && j >= s.row0 && j < s.row0 + s.nloc;
    if (!own && !tr) return;
    const unsigned long long p = atomicAdd(reinterpret_cast<unsigned long long*>(s.counter), 1ull);
    if ((int64_t)p >= s.cap) return;  // counted, not stored: the caller sees counter > cap and retries with more room
    s.o_l[p] = l;
    s.o_i[p] = i;
    s.o_j[p] = j;
    s.o_m[p] = m;
    s.o_x[p] = x;
  }
};

// counts of reporter m on the pair {m, n}, neither direction a true tie: (x_{m->n}, x_{n->m})
__device__ __forceinline__ bool base_pair(const vm_synth& s, int l, int m, int n, double th, int& x_mn, int& x_nm,
                                          bool& maybe) {
  const uint2 key = make_uint2((uint32_t)s.seed ^ 0x243F6A88u, (uint32_t)(s.seed >> 32) + (uint32_t)l * 0x9E3779B9u);
  const uint4 r = philox4x32(make_uint4((uint32_t)m, (uint32_t)n, 0u, 0x5eedu), key);
  const double eta = s.eta, mu = th * lam_of(s, l, 0), mm = mu / (1.0 - eta);
  const double a = -expm1(-mm), b = exp(-mm) * (-expm1(-mu));  // P(first > 0), P(first = 0, second > 0)
  const double u0 = u01(r.x);
  maybe = u0 < a + b;
  if (!maybe) return false;
  const bool first_pos = u0 < a;
  const int x1 = first_pos ? poisson_inv(mm, u01(r.y), true) : 0;
  const int x2 = first_pos ? poisson_inv(mu + eta * (double)x1, u01(r.z), false) : poisson_inv(mu, u01(r.z), true);
  const bool coin = (r.w & 1u) != 0;  // first direction is m->n
  x_mn = coin ? x1 : x2;
  x_nm = coin ? x2 : x1;
  return true;
}

// ---- kernel A: every owned tie (l, i, j), reporters m in {i, j}; pairs WITHOUT a true tie, and the self ties -------
__global__ void __launch_bounds__(256) k_synth_base(const __grid_constant__ vm_synth s) {
  const Emit emit{s};
  const int64_t T = s.L * s.nloc * s.N;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lrow = t / s.N;
    const int j = (int)(t - lrow * s.N), l = (int)(lrow / s.nloc), i = (int)(lrow - (int64_t)l * s.nloc + s.row0);
    if (i == j) {
      if (i >= s.M) continue;
      const double th = s.theta[(int64_t)l * s.M + i], mu = th * lam_of(s, l, 0), mm = mu / (1.0 - s.eta);
      const uint2 key = make_uint2((uint32_t)s.seed ^ 0x243F6A88u, (uint32_t)(s.seed >> 32) + (uint32_t)l * 0x9E3779B9u);
      const uint4 r = philox4x32(make_uint4((uint32_t)i, (uint32_t)i, 2u, 0x5eedu), key);
      const int y = poisson_inv(mm, u01(r.x), false);
      emit(l, i, i, i, poisson_inv(mu + s.eta * (double)y, u01(r.y), false));
      continue;
    }
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const int m = side == 0 ? i : j, n = side == 0 ? j : i;
      if (m >= s.M) continue;
      int x_mn = 0, x_nm = 0;
      bool maybe;
      if (!base_pair(s, l, m, n, s.theta[(int64_t)l * s.M + m], x_mn, x_nm, maybe)) continue;
      // a pair with a true tie is sampled by kernel B instead
      const int64_t kmn = ((int64_t)l * s.N + m) * s.N + n, knm = ((int64_t)l * s.N + n) * s.N + m;
      if (y_lookup(s.y_key, s.y_val, s.nY, kmn) > 0 || y_lookup(s.y_key, s.y_val, s.nY, knm) > 0) continue;
      // this thread's tie is i->j: x = count in that direction, x^T the other one
      const int x = side == 0 ? x_mn : x_nm, xT = side == 0 ? x_nm : x_mn;
      emit(l, i, j, m, x);
      if (s.emit_transposed && !(j >= s.row0 && j < s.row0 + s.nloc)) emit(l, j, i, m, xT);
    }
  }
}

// ---- kernel B: pairs WITH a true tie in either direction, one thread per (true tie, reporter side) ---------------------
__global__ void __launch_bounds__(256) k_synth_edges(const __grid_constant__ vm_synth s) {
  const Emit emit{s};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < 2 * s.nY; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = e >> 1;
    const int side = (int)(e & 1);
    const int64_t key = s.y_key[q];
    const int b_ = (int)(key % s.N), a_ = (int)((key / s.N) % s.N), l = (int)(key / (s.N * s.N));
    if (a_ == b_) continue;  // the generators never produce a true self tie
    const int64_t krev = ((int64_t)l * s.N + b_) * s.N + a_;
    const int yab = s.y_val[q], yba = y_lookup(s.y_key, s.y_val, s.nY, krev);
    if (yba > 0 && a_ > b_) continue;  // both directions are true ties: the pair is handled from its (min,max) entry
    const int m = side == 0 ? a_ : b_, n = side == 0 ? b_ : a_;
    if (m >= s.M) continue;
    const bool mine = (m >= s.row0 && m < s.row0 + s.nloc) || (n >= s.row0 && n < s.row0 + s.nloc);
    if (!mine) continue;
    const double th = s.theta[(int64_t)l * s.M + m], eta = s.eta;
    const double lam_mn = lam_of(s, l, side == 0 ? yab : yba);
    const double lam_nm = lam_of(s, l, side == 0 ? yba : yab);
    const uint2 pk = make_uint2((uint32_t)s.seed ^ 0x243F6A88u, (uint32_t)(s.seed >> 32) + (uint32_t)l * 0x9E3779B9u);
    const uint4 r = philox4x32(make_uint4((uint32_t)m, (uint32_t)n, 1u, 0x5eedu), pk);
    const bool coin = (r.w & 1u) != 0;  // first direction is m->n
    const double l1 = coin ? lam_mn : lam_nm, l2 = coin ? lam_nm : lam_mn;
    const int x1 = poisson_inv((th * l1 + eta * th * l2) / (1.0 - eta * eta), u01(r.x), false);
    const int x2 = poisson_inv(th * l2 + eta * (double)x1, u01(r.y), false);
    emit(l, m, n, m, coin ? x1 : x2);
    emit(l, n, m, m, coin ? x2 : x1);
  }
}

}  // namespace

extern "C" int64_t vm_synth_size(void) { return (int64_t)sizeof(vm_synth); }

extern "C" int vm_synth_ego(const vm_synth* s, void* stream) {
  if (!s || s->L < 1 || s->N < 1 || s->M < 1 || s->M > s->N || s->nloc < 0 || s->row0 < 0 || s->row0 + s->nloc > s->N ||
      s->eta < 0.0 || s->eta >= 1.0 || s->cap < 0 || !s->counter || !s->theta || (s->nY > 0 && (!s->y_key || !s->y_val)))
    return VM_EINVAL;
  if (s->N >= (int64_t)1 << 31) return VM_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(s->counter, 0, sizeof(int64_t), st);
  if (e != cudaSuccess) return (int)e;
  const int64_t T = s->L * s->nloc * s->N;
  if (T > 0) {
    int64_t nb = (T + 255) / 256;
    if (nb > 148 * 64) nb = 148 * 64;
    k_synth_base<<<(unsigned)nb, 256, 0, st>>>(*s);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  }
  if (s->nY > 0) {
    int64_t nb = (2 * s->nY + 255) / 256;
    if (nb > 148 * 64) nb = 148 * 64;
    k_synth_edges<<<(unsigned)nb, 256, 0, st>>>(*s);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  }
  return 0;
}
