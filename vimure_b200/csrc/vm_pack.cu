// vimure_b200 -- device-side data packing behind the C ABI (`vm_pack`, include/vimure_b200.h).
//
// Replaces, for the hot path, the data preparation of the reference's `__check_fit_params` (model.py:147-176): `data_T` /
// `data_T_vals` -- an O(nnz^2) python lookup (utils.py:73-84) -- become ONE radix sort of the reports by (l,i,j,m) plus a
// binary search that pre-pairs every report X[l,i,j,m] with its reciprocal X[l,j,i,m]; the union-of-ties DataFrame merges
// of `_set_rho_prior` (model.py:509-556) become head flags + a prefix sum over the sorted list.  Everything runs on the
// device with no host synchronisation inside: launches are sized by the input's upper bounds and read the actual counts
// from device memory; the caller reads `counts` once at the end.
//
// Structured reporter masks only (ego / all-reporter: the BASELINE configurations and the edgelist parser's output); a
// general COO mask is packed by the host-side packer (vimure_b200/_packing.py).
//
// Sorting and prefix sums are CUB's (library plumbing); the packing kernels are this file's.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vm_common.cuh"

namespace {

constexpr int PT = 256;  // threads per block of the packing kernels
constexpr int GAMMA_CHUNK = 256;

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline unsigned nblk(int64_t n) { return (unsigned)(n > 0 ? cdiv64(n, PT) : 1); }
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

__device__ __forceinline__ int64_t lower_bound64(const int64_t* a, int64_t n, int64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// device-side view of the work arrays (carved from the caller's workspace)
struct Work {
  int64_t *key, *key2;        // [n] sort keys (in / out)
  int32_t *idx, *idx2;        // [n] entry index (in / out)
  int32_t *own, *own_pos;     // [n] owned flag of sorted position p, its exclusive prefix sum
  int32_t *head, *head_pos;   // [n] (owned-entry order) first entry of its tie, prefix sum
  int32_t *e1, *e1_pos;       // [n] (owned-entry order) entry visits the gamma/phi passes, prefix sum
  int32_t *tr, *tr_pos;       // [n] sorted position p belongs to the transposed list, prefix sum
  float* xT;                  // [n] reciprocal count of sorted position p
  int64_t* tkey;              // [n] local tie key of owned entry e
  int64_t* uk_e;              // [n] unique tie keys among the owned entries
  int64_t* ukeys;             // [cap_u] tie key of every special tie
  int32_t *dflag, *dpre;      // [L*nloc+1] diagonal tie of the row carries no entry (ego), prefix sum
  int32_t *cx, *cx_pos;       // [cap_u] special tie takes no shortcut, prefix sum
  int32_t *lm, *lm2;          // [n] reporter key of E1 entry (in / out)
  int32_t *gi, *gi2;          // [n] E1 index (in / out)
  int32_t *nch;               // [L*M+1] gamma chunks per reporter
  int64_t* g0i;               // [L*M] integer sum of x over the E0 entries of a reporter
  void* cub_tmp;
  size_t cub_bytes;
};

// counts[] slots (device int64[16], read by the caller after the stream has drained)
enum { C_U = 0, C_I, C_I1, C_IT, C_NCX, C_NGCHUNK, C_MAXUL, C_MAXLAY, C_MAXCXL, C_DUP, C_OOB, C_SUMX, C_BALL, C_NNEED, C_NE };

// ---- 1. keys: (l,i,j,m) of the entries a row-block shard needs (own rows, or reciprocal of an own row) ------------
__device__ __forceinline__ void warp_add_counter(int64_t* counter, long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v != 0) atomicAdd(reinterpret_cast<unsigned long long*>(counter), (unsigned long long)v);
}

__global__ void __launch_bounds__(PT) k_pack_keys(const vm_pack_args p, Work w, int64_t sentinel) {
  const int64_t e = (int64_t)blockIdx.x * PT + threadIdx.x;  // (no early return: the warp reductions need every lane)
  int64_t key = sentinel;  // L*N*N*M: sorts behind every valid key
  int need = 0;
  long long xsum = 0;
  if (e < p.n_in) {
    const int64_t l = p.x_l[e], i = p.x_i[e], j = p.x_j[e], m = p.x_m[e];
    if (l < 0 || l >= p.L || i < 0 || i >= p.N || j < 0 || j >= p.N || m < 0 || m >= p.M) {
      p.counts[C_OOB] = 1;
    } else {
      const bool own = i >= p.row0 && i < p.row0 + p.nloc, ownT = j >= p.row0 && j < p.row0 + p.nloc;
      if (own || ownT) {
        key = ((l * p.N + i) * p.N + j) * p.M + m;
        need = 1;
        if (own) xsum = p.x_v[e];
      }
    }
    w.key[e] = key;
    w.idx[e] = (int32_t)e;
  }
  // one atomic per warp and counter (a single address hit by every thread serialises)
  warp_add_counter(p.counts + C_NNEED, (long long)need);
  warp_add_counter(p.counts + C_SUMX, xsum);
}

// reporter mask multiplicity of (l,i,j,m) for the structured masks
__device__ __forceinline__ int mask_mult(const vm_pack_args& p, int64_t l, int64_t i, int64_t j, int64_t m) {
  if (p.r_mode == VM_R_ALL) return 1;
  if (!p.rep[l * p.M + m]) return 0;
  if (i == j) return (p.ego_diag && m == i) ? 1 : 0;
  return (m == i || m == j) ? 1 : 0;
}

// ---- 2. per sorted position: duplicate check, reciprocal count, ownership, membership of the transposed list ------
__global__ void __launch_bounds__(PT) k_pack_pair(const vm_pack_args p, Work w) {
  const int64_t q = (int64_t)blockIdx.x * PT + threadIdx.x;
  if (q >= p.n_in) return;
  const int64_t n_need = p.counts[C_NNEED];
  int own = 0, tr = 0;
  float xT = 0.f;
  if (q < n_need) {
    const int64_t key = w.key2[q];
    if (q > 0 && w.key2[q - 1] == key) p.counts[C_DUP] = 1;
    const int64_t e = w.idx2[q];
    const int64_t l = p.x_l[e], i = p.x_i[e], j = p.x_j[e], m = p.x_m[e];
    const int64_t keyT = ((l * p.N + j) * p.N + i) * p.M + m;
    const int64_t pos = lower_bound64(w.key2, n_need, keyT);
    if (pos < n_need && w.key2[pos] == keyT) xT = (float)p.x_v[w.idx2[pos]];
    own = (i >= p.row0 && i < p.row0 + p.nloc) ? 1 : 0;
    // X entry (l,i,j,m) is the "X_T" value of the mask entry (l,j,i,m), owned by the rank of row j (model.py:1269-1290)
    tr = (j >= p.row0 && j < p.row0 + p.nloc && mask_mult(p, l, j, i, m) > 0) ? 1 : 0;
  }
  w.own[q] = own;
  w.tr[q] = tr;
  w.xT[q] = xT;
}

// ---- 3. owned entries in tie order: e_m, e_x, e_xT, e_flags, tie keys, E1 flags, g0, transposed list ----------------
__global__ void __launch_bounds__(PT) k_pack_entries(const vm_pack_args p, Work w) {
  const int64_t q = (int64_t)blockIdx.x * PT + threadIdx.x;  // (no early return: the warp reduction needs every lane)
  const int64_t n_need = p.counts[C_NNEED];
  long long ball = 0;
  if (q < p.n_in) {
    if (q == p.n_in - 1) {  // totals of the two compactions
      p.counts[C_I] = w.own_pos[q] + w.own[q];
      p.counts[C_IT] = w.tr_pos[q] + w.tr[q];
    }
    if (q < n_need) {
      const int64_t e = w.idx2[q];
      const int64_t l = p.x_l[e], i = p.x_i[e], j = p.x_j[e], m = p.x_m[e];
      const float x = (float)p.x_v[e], xT = w.xT[q];
      if (w.own[q]) {
        const int64_t o = w.own_pos[q];
        p.e_m[o] = (int32_t)m;
        p.e_src[o] = (int32_t)e;
        p.e_x[o] = x;
        p.e_xT[o] = xT;
        p.e_flags[o] = mask_mult(p, l, i, j, m) > 0 ? 1 : 0;
        w.tkey[o] = (l * p.nloc + (i - p.row0)) * p.N + j;
        // E1 = the entries the gamma / phi passes visit (reciprocal report present); E0 = constant allocation dz1 = x
        const int is1 = !p.split_e0 ? 1 : (p.mutuality ? (xT != 0.f ? 1 : 0) : 0);
        w.e1[o] = is1;
        if (!is1) atomicAdd(reinterpret_cast<unsigned long long*>(w.g0i) + l * p.M + m, (unsigned long long)(long long)p.x_v[e]);
      }
      if (w.tr[q]) {
        const int64_t t = w.tr_pos[q];
        p.t_lrow[t] = (int32_t)(l * p.nloc + (j - p.row0));
        p.t_col[t] = (int32_t)i;
        const int mult = mask_mult(p, l, j, i, m);
        p.t_x[t] = x * (float)mult;
        ball = (long long)p.x_v[e] * mult;
      }
    }
  }
  warp_add_counter(p.counts + C_BALL, ball);
}

// head flags of the owned entries (first entry of its tie)
__global__ void __launch_bounds__(PT) k_pack_heads(const vm_pack_args p, Work w) {
  const int64_t o = (int64_t)blockIdx.x * PT + threadIdx.x;
  if (o >= p.n_in) return;
  const int64_t I = p.counts[C_I];
  int h = 0;
  if (o < I) h = (o == 0 || w.tkey[o - 1] != w.tkey[o]) ? 1 : 0;
  w.head[o] = h;
  if (o >= I) w.e1[o] = 0;  // (the prefix sums run over the input's upper bound)
}

// unique tie keys of the owned entries
__global__ void __launch_bounds__(PT) k_pack_unique(const vm_pack_args p, Work w) {
  const int64_t o = (int64_t)blockIdx.x * PT + threadIdx.x;
  if (o >= p.n_in) return;
  const int64_t I = p.counts[C_I];
  if (o < I && w.head[o]) w.uk_e[w.head_pos[o]] = w.tkey[o];
  if (o == I - 1) p.counts[C_NE] = w.head_pos[o] + w.head[o];  // number of distinct ties among the owned entries
}

// ego mask: the diagonal tie of every owned row is a special tie; flag the rows whose diagonal tie carries no entry
__global__ void __launch_bounds__(PT) k_pack_diag(const vm_pack_args p, Work w) {
  const int64_t r = (int64_t)blockIdx.x * PT + threadIdx.x;  // local row over all layers
  const int64_t R = p.L * p.nloc;
  if (r > R) return;
  int f = 0;
  if (r < R && p.r_mode == VM_R_EGO) {
    const int64_t n_e = p.counts[C_NE];
    const int64_t i_glob = r % p.nloc + p.row0;
    const int64_t dk = r * p.N + i_glob;
    const int64_t pos = lower_bound64(w.uk_e, n_e, dk);
    f = (pos < n_e && w.uk_e[pos] == dk) ? 0 : 1;
  }
  w.dflag[r] = f;  // dflag[R] = 0: the scan's last element gives the total
}

// final position of every special tie: entry ties shifted by the diagonal-only ties before them, and the diagonal-only ties
__global__ void __launch_bounds__(PT) k_pack_ties(const vm_pack_args p, Work w) {
  const int64_t t = (int64_t)blockIdx.x * PT + threadIdx.x;
  const int64_t n_e = p.counts[C_NE], R = p.L * p.nloc;
  const int64_t n_d = w.dpre[R];
  if (t == 0) p.counts[C_U] = n_e + n_d;
  if (t < n_e) {
    const int64_t key = w.uk_e[t], lrow = key / p.N, j = key - lrow * p.N, i_glob = lrow % p.nloc + p.row0;
    const int64_t pos = t + w.dpre[lrow] + ((w.dflag[lrow] && i_glob < j) ? 1 : 0);
    w.ukeys[pos] = key;
  }
  if (t < R && w.dflag[t]) {
    const int64_t i_glob = t % p.nloc + p.row0, key = t * p.N + i_glob;
    const int64_t pos = w.dpre[t] + lower_bound64(w.uk_e, n_e, key);
    w.ukeys[pos] = key;
  }
}

// per special tie: coordinates, entry range, first entry inline, E0 sum, flags, class, patch constants
__global__ void __launch_bounds__(PT) k_pack_tie_data(const vm_pack_args p, Work w, int tile_w) {
  const int64_t u = (int64_t)blockIdx.x * PT + threadIdx.x;
  const int64_t U = p.counts[C_U], I = p.counts[C_I];
  if (u >= p.cap_u + 1) return;
  if (u >= U) {
    w.cx[u] = 0;  // (the prefix sum runs over the capacity)
    if (u == U) p.u_ptr[u] = I;
    return;
  }
  const int64_t key = w.ukeys[u], lrow = key / p.N, col = key - lrow * p.N, l = lrow / p.nloc, i = lrow - l * p.nloc + p.row0;
  const int64_t e0 = lower_bound64(w.tkey, I, key);
  int64_t e1 = e0;
  while (e1 < I && w.tkey[e1] == key) ++e1;  // entries of one tie: at most its reporters (2 for an ego mask)
  const int cnt = (int)(e1 - e0);
  p.u_lrow[u] = (int32_t)lrow;
  p.u_col[u] = (int32_t)col;
  p.u_ptr[u] = e0;
  p.u_cnt[u] = cnt;
  p.u_gflat[u] = (l * p.N + i) * p.N + col;
  float x0s = 0.f;
  int has_e1 = 0;
  for (int64_t e = e0; e < e1; ++e) {
    p.e_u[e] = (int32_t)u;
    if (w.e1[e]) has_e1 = 1;
    else x0s += p.e_x[e];
  }
  const int m0 = cnt ? p.e_m[e0] : 0;
  const float x0 = cnt ? p.e_x[e0] : 0.f, xT0 = cnt ? p.e_xT[e0] : 0.f;
  p.u_m0[u] = m0;
  p.u_x0[u] = x0;
  p.u_xT0[u] = xT0;
  p.u_x0sum[u] = x0s;
  p.u_has_x[u] = cnt > 0;
  int reported = 1;
  if (p.r_mode == VM_R_EGO) {
    const bool ri = i < p.M && p.rep[l * p.M + i], rj = col < p.M && p.rep[l * p.M + col];
    reported = (i == col) ? (p.ego_diag && ri) : (ri || rj);
  }
  p.u_reported[u] = (uint8_t)reported;
  // shortcut ties (vm_ctx.simple_mode): off the diagonal, in a full column tile, SIMPLE (no E1 entry) or SINGLE (one
  // entry, E1, reported by the row or the column node)
  const bool inside = p.simple && i != col && col < (p.N / tile_w) * tile_w;
  const bool simple = inside && cnt > 0 && !has_e1;
  const bool single = inside && p.single && p.mutuality && p.split_e0 && cnt == 1 && has_e1 && (m0 == i || m0 == col);
  p.u_px[u] = simple ? x0s : (single ? x0 : 0.f);
  p.u_pxt[u] = single ? (m0 == i ? xT0 : -xT0) : 0.f;
  w.cx[u] = (simple || single) ? 0 : 1;
}

// dense-tile pointers: first special tie of every (local row, column tile); per-layer maxima
__global__ void __launch_bounds__(PT) k_pack_tiles(const vm_pack_args p, Work w, int tile_w, int64_t nct) {
  const int64_t t = (int64_t)blockIdx.x * PT + threadIdx.x;
  const int64_t nt = p.L * p.nloc * nct, U = p.counts[C_U];
  if (t > nt) return;
  if (t == nt) {
    p.utile_ptr[t] = (int32_t)U;
    return;
  }
  const int64_t lrow = t / nct, ct = t - lrow * nct;
  p.utile_ptr[t] = (int32_t)lower_bound64(w.ukeys, U, lrow * p.N + ct * tile_w);
}

// compaction of the special ties that take no shortcut + their compacted per-tie arrays; E1 compaction of the entries
__global__ void __launch_bounds__(PT) k_pack_compact(const vm_pack_args p, Work w) {
  const int64_t t = (int64_t)blockIdx.x * PT + threadIdx.x;
  const int64_t U = p.counts[C_U], I = p.counts[C_I];
  if (t < U && w.cx[t]) {
    const int64_t c = w.cx_pos[t];
    p.cx_idx[c] = (int32_t)t;
    p.cx_lrow[c] = p.u_lrow[t];
    p.cx_col[c] = p.u_col[t];
    p.cx_cnt[c] = p.u_cnt[t];
    p.cx_m0[c] = p.u_m0[t];
    p.cx_x0[c] = p.u_x0[t];
    p.cx_xT0[c] = p.u_xT0[t];
    p.cx_x0sum[c] = p.u_x0sum[t];
  }
  if (U > 0 && t == U - 1) p.counts[C_NCX] = w.cx_pos[t] + w.cx[t];
  if (t < I && w.e1[t]) {
    const int64_t f = w.e1_pos[t];
    p.f_u[f] = p.e_u[t];
    p.f_m[f] = p.e_m[t];
    p.f_x[f] = p.e_x[t];
    p.f_xT[f] = p.e_xT[t];
    const int64_t l = w.tkey[t] / (p.nloc * p.N);
    w.lm[f] = (int32_t)(l * p.M + p.e_m[t]);
    w.gi[f] = (int32_t)f;
  }
  if (I > 0 && t == I - 1) p.counts[C_I1] = w.e1_pos[t] + w.e1[t];
}

// sentinel reporter keys beyond I1 (the second sort runs over the input's upper bound)
__global__ void __launch_bounds__(PT) k_pack_lm_pad(const vm_pack_args p, Work w) {
  const int64_t t = (int64_t)blockIdx.x * PT + threadIdx.x;
  if (t >= p.n_in) return;
  if (t >= p.counts[C_I1]) {
    w.lm[t] = INT32_MAX;
    w.gi[t] = 0;
  }
}

// per-layer ranges (cx list, E1 entries, special ties) and their maxima; transposed list's special-tie indices
__global__ void __launch_bounds__(PT) k_pack_layers(const vm_pack_args p, Work w, int64_t nct) {
  const int64_t t = (int64_t)blockIdx.x * PT + threadIdx.x;
  const int64_t U = p.counts[C_U], I1 = p.counts[C_I1], ncx = p.counts[C_NCX], IT = p.counts[C_IT];
  if (t <= p.L) {
    // first cx entry / E1 entry of layer t (both lists are sorted by layer)
    int64_t lo = 0, hi = ncx;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (p.cx_lrow[mid] / p.nloc < t) lo = mid + 1;
      else hi = mid;
    }
    p.cx_ptr[t] = lo;
    lo = 0, hi = I1;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (p.u_lrow[p.f_u[mid]] / p.nloc < t) lo = mid + 1;
      else hi = mid;
    }
    p.lay_eptr[t] = lo;
  }
  if (t < IT) {
    const int64_t key = (int64_t)p.t_lrow[t] * p.N + p.t_col[t];
    const int64_t pos = lower_bound64(w.ukeys, U, key);
    p.t_u[t] = (pos < U && w.ukeys[pos] == key) ? (int32_t)pos : -1;
  }
  (void)nct;
}
__global__ void k_pack_layer_max(const vm_pack_args p, int64_t nct) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int64_t mu = 0, ml = 0, mc = 0;
  for (int64_t l = 0; l < p.L; ++l) {
    const int64_t nu = (int64_t)p.utile_ptr[(l + 1) * p.nloc * nct] - p.utile_ptr[l * p.nloc * nct];
    mu = nu > mu ? nu : mu;
    const int64_t nl = p.lay_eptr[l + 1] - p.lay_eptr[l];
    ml = nl > ml ? nl : ml;
    const int64_t nc = p.cx_ptr[l + 1] - p.cx_ptr[l];
    mc = nc > mc ? nc : mc;
  }
  p.counts[C_MAXUL] = mu;
  p.counts[C_MAXLAY] = ml;
  p.counts[C_MAXCXL] = mc;
}

// reporter-sorted E1 entries (gamma pass) and the per-reporter chunk counts
__global__ void __launch_bounds__(PT) k_pack_gamma(const vm_pack_args p, Work w) {
  const int64_t t = (int64_t)blockIdx.x * PT + threadIdx.x;
  const int64_t I1 = p.counts[C_I1], LM = p.L * p.M;
  if (t < I1) {
    const int32_t f = w.gi2[t];
    p.g_u[t] = p.f_u[f];
    p.g_x[t] = p.f_x[f];
    p.g_xT[t] = p.f_xT[f];
  }
  if (t <= LM) {
    int n = 0;
    if (t < LM) {
      // entries of reporter t: [lower_bound(t), lower_bound(t+1)) in the sorted reporter keys
      int64_t lo = 0, hi = I1;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (w.lm2[mid] < t) lo = mid + 1;
        else hi = mid;
      }
      int64_t a = lo;
      hi = I1;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (w.lm2[mid] < t + 1) lo = mid + 1;
        else hi = mid;
      }
      n = (int)((lo - a + GAMMA_CHUNK - 1) / GAMMA_CHUNK);
      w.key[t] = a;  // (the first sort's key buffer is free by now) start of the reporter's entries
      p.g0[t] = (double)w.g0i[t];
    }
    w.nch[t] = n;
  }
}
__global__ void __launch_bounds__(PT) k_pack_chunks(const vm_pack_args p, Work w, const int32_t* cptr) {
  const int64_t t = (int64_t)blockIdx.x * PT + threadIdx.x;
  const int64_t LM = p.L * p.M;
  if (t > LM) return;
  p.g_lm_cptr[t] = cptr[t];
  if (t == LM) {
    p.counts[C_NGCHUNK] = cptr[t];
    if (cptr[t] <= p.cap_g) p.g_chunk_ptr[cptr[t]] = p.counts[C_I1];
    return;
  }
  const int n = w.nch[t];
  const int64_t a = w.key[t];
  for (int q = 0; q < n; ++q) {
    const int64_t c = (int64_t)cptr[t] + q;
    if (c < p.cap_g) {
      p.g_chunk_lm[c] = (int32_t)t;
      p.g_chunk_ptr[c] = a + (int64_t)q * GAMMA_CHUNK;
    }
  }
}

static size_t carve(Work& w, const vm_pack_args* p, void* base, size_t cub_bytes) {
  const size_t n = (size_t)(p->n_in > 0 ? p->n_in : 1), cu = (size_t)p->cap_u + 1, R = (size_t)(p->L * p->nloc) + 1;
  const size_t LM = (size_t)(p->L * p->M) + 1;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* q = base ? (void*)((char*)base + off) : nullptr;
    off += al256(bytes);
    return q;
  };
  w.key = (int64_t*)take(n * 8 > LM * 8 ? n * 8 : LM * 8);
  w.key2 = (int64_t*)take(n * 8);
  w.idx = (int32_t*)take(n * 4);
  w.idx2 = (int32_t*)take(n * 4);
  w.own = (int32_t*)take(n * 4);
  w.own_pos = (int32_t*)take(n * 4);
  w.head = (int32_t*)take(n * 4);
  w.head_pos = (int32_t*)take(n * 4);
  w.e1 = (int32_t*)take(n * 4);
  w.e1_pos = (int32_t*)take(n * 4);
  w.tr = (int32_t*)take(n * 4);
  w.tr_pos = (int32_t*)take(n * 4);
  w.xT = (float*)take(n * 4);
  w.tkey = (int64_t*)take(n * 8);
  w.uk_e = (int64_t*)take(n * 8);
  w.ukeys = (int64_t*)take(cu * 8);
  w.dflag = (int32_t*)take(R * 4);
  w.dpre = (int32_t*)take(R * 4);
  w.cx = (int32_t*)take(cu * 4);
  w.cx_pos = (int32_t*)take(cu * 4);
  w.lm = (int32_t*)take(n * 4);
  w.lm2 = (int32_t*)take(n * 4);
  w.gi = (int32_t*)take(n * 4);
  w.gi2 = (int32_t*)take(n * 4);
  w.nch = (int32_t*)take(LM * 4 * 2);  // counts + their prefix sum
  w.g0i = (int64_t*)take(LM * 8);
  w.cub_tmp = take(cub_bytes);
  w.cub_bytes = cub_bytes;
  return off;
}

static size_t cub_temp_bytes(const vm_pack_args* p) {
  const int n = (int)(p->n_in > 0 ? p->n_in : 1);
  const int cu = (int)p->cap_u + 1;
  size_t a = 0, b = 0, c = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const int64_t*)nullptr, (int64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, n, 0, 64);
  cub::DeviceRadixSort::SortPairs(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, n, 0, 32);
  cub::DeviceScan::ExclusiveSum(nullptr, c, (const int32_t*)nullptr, (int32_t*)nullptr, n > cu ? n : cu);
  size_t m = a > b ? a : b;
  return (m > c ? m : c) + 256;
}

static int bits_for(int64_t v) {
  int b = 1;
  while (b < 63 && ((int64_t)1 << b) <= v) ++b;
  return b;
}

}  // namespace

extern "C" int64_t vm_pack_size(void) { return (int64_t)sizeof(vm_pack_args); }

extern "C" int64_t vm_pack_workspace_bytes(const vm_pack_args* p) {
  if (!p || p->n_in < 0 || p->cap_u < 0 || p->L < 1 || p->nloc < 0 || p->M < 1) return VM_EINVAL;
  if (p->n_in >= ((int64_t)1 << 31) - 2 || p->cap_u >= ((int64_t)1 << 31) - 2) return VM_EINVAL;
  Work w;
  return (int64_t)carve(w, p, nullptr, cub_temp_bytes(p));
}

#define PK_CHECK()                                   \
  do {                                               \
    const cudaError_t e__ = cudaGetLastError();      \
    if (e__ != cudaSuccess) return (int)e__;         \
  } while (0)

extern "C" int vm_pack(const vm_pack_args* p, void* stream) {
  if (!p || p->L < 1 || p->N < 1 || p->M < 1 || p->K < 2 || p->K > VM_MAX_K || p->n_in < 0 || p->nloc < 0 || p->row0 < 0 ||
      p->row0 + p->nloc > p->N || p->tile_h < 1 || !p->counts || !p->workspace)
    return VM_EINVAL;
  if (p->r_mode != VM_R_EGO && p->r_mode != VM_R_ALL) return VM_ENOTSUP;  // general masks: host-side packer
  if (p->r_mode == VM_R_EGO && !p->rep) return VM_EINVAL;
  if (p->n_in >= ((int64_t)1 << 31) - 2 || p->cap_u >= ((int64_t)1 << 31) - 2) return VM_EINVAL;
  if (p->cap_e < p->n_in || p->cap_u < p->n_in + (p->r_mode == VM_R_EGO ? p->L * p->nloc : 0) ||
      p->cap_g < p->n_in / GAMMA_CHUNK + p->L * p->M + 1)
    return VM_EINVAL;
  // L*N*N*M must fit the 63-bit sort key
  const double keyspace = (double)p->L * (double)p->N * (double)p->N * (double)p->M;
  if (keyspace >= 9.0e18) return VM_EINVAL;
  const int64_t tile_w = vm_dense_tile_w_host(p->K);
  if (tile_w <= 0) return VM_EINVAL;
  const int64_t nct = cdiv64(p->N, tile_w);
  cudaStream_t st = (cudaStream_t)stream;
  Work w;
  const size_t cub_bytes = cub_temp_bytes(p);
  const size_t need = carve(w, p, p->workspace, cub_bytes);
  if ((int64_t)need > p->workspace_bytes) return VM_EINVAL;
  const int64_t n = p->n_in, R = p->L * p->nloc, LM = p->L * p->M;
  cudaMemsetAsync(p->counts, 0, 16 * sizeof(int64_t), st);
  cudaMemsetAsync(w.g0i, 0, (size_t)(LM + 1) * 8, st);
  PK_CHECK();
  const int n32 = (int)(n > 0 ? n : 1);
  if (n > 0) {
    size_t tb = w.cub_bytes;
    const int64_t sentinel = p->L * p->N * p->N * p->M;  // one more than the largest key
    const int end_bit = bits_for(sentinel);              // only the bits a key can have are sorted
    k_pack_keys<<<nblk(n), PT, 0, st>>>(*p, w, sentinel);
    PK_CHECK();
    cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.key, w.key2, w.idx, w.idx2, n32, 0, end_bit, st);
    PK_CHECK();
    k_pack_pair<<<nblk(n), PT, 0, st>>>(*p, w);
    PK_CHECK();
    tb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.own, w.own_pos, n32, st);
    tb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.tr, w.tr_pos, n32, st);
    PK_CHECK();
    k_pack_entries<<<nblk(n), PT, 0, st>>>(*p, w);
    PK_CHECK();
    k_pack_heads<<<nblk(n), PT, 0, st>>>(*p, w);
    PK_CHECK();
    tb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.head, w.head_pos, n32, st);
    tb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.e1, w.e1_pos, n32, st);
    PK_CHECK();
    k_pack_unique<<<nblk(n), PT, 0, st>>>(*p, w);
    PK_CHECK();
  }
  // special ties = entry ties + (ego) the diagonal ties that carry no entry
  k_pack_diag<<<nblk(R + 1), PT, 0, st>>>(*p, w);
  PK_CHECK();
  {
    size_t tb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.dflag, w.dpre, (int)(R + 1), st);
    PK_CHECK();
  }
  const int64_t cap_ties = p->cap_u + 1;
  k_pack_ties<<<nblk(n > R ? n : R), PT, 0, st>>>(*p, w);
  PK_CHECK();
  k_pack_tie_data<<<nblk(cap_ties), PT, 0, st>>>(*p, w, (int)tile_w);
  PK_CHECK();
  k_pack_tiles<<<nblk(R * nct + 1), PT, 0, st>>>(*p, w, (int)tile_w, nct);
  PK_CHECK();
  {
    size_t tb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.cx, w.cx_pos, (int)cap_ties, st);
    PK_CHECK();
  }
  k_pack_compact<<<nblk(cap_ties > n ? cap_ties : n), PT, 0, st>>>(*p, w);
  PK_CHECK();
  k_pack_layers<<<nblk((n > p->L + 1 ? n : p->L + 1)), PT, 0, st>>>(*p, w, nct);
  PK_CHECK();
  k_pack_layer_max<<<1, 32, 0, st>>>(*p, nct);
  PK_CHECK();
  // reporter order of the E1 entries (gamma pass): stable radix sort by l*M+m
  if (n > 0) {
    k_pack_lm_pad<<<nblk(n), PT, 0, st>>>(*p, w);
    PK_CHECK();
    size_t tb = w.cub_bytes;
    cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.lm, w.lm2, w.gi, w.gi2, n32, 0, 32, st);
    PK_CHECK();
  }
  k_pack_gamma<<<nblk((n > LM + 1 ? n : LM + 1)), PT, 0, st>>>(*p, w);
  PK_CHECK();
  {
    size_t tb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.nch, w.nch + (LM + 1), (int)(LM + 1), st);
    PK_CHECK();
  }
  k_pack_chunks<<<nblk(LM + 1), PT, 0, st>>>(*p, w, w.nch + (LM + 1));
  PK_CHECK();
  return 0;
}
