// K2 — deterministic embedding backward with in-place sparse optimizers.
//
// Replaces IndexedSlices -> UnsortedSegmentSum (atomics, non-deterministic) -> dense
// ResourceApplyAdam that the reference's model.compile(optimizer=Adam) lowers to
// (src/ctr/fm/train.py:49-50).
//
// Pipeline, all on the caller's stream, no host sync:
//   1. keys[p] = table(field(p)) << row_bits | id      (p = lookup position, ascending b,l,f)
//   2. stable LSD radix sort of (key, p), 8 bits per pass over the significant bits only
//   3. head flags + scan -> segment starts (one segment per touched row); a second scan over the
//      segments emits self-contained 16-byte WORK ITEMS: one per short segment
//      {key, start, len, first lookup position} and one per RTF_SEG_CHUNK-row chunk of a long
//      segment {key, start, len, long-segment slot}.  Steps 1-3 need only the ids and run on a
//      side stream behind the dense MLP (rtf_embed_bwd_prepare).
//   4. ONE persistent launch (seg_apply): a lane-group per item, the next item's record
//      prefetched while the current one is processed, so an item costs one round of
//      independent loads (gradient rows + W/m/v) instead of a 4-5 deep dependent chain.
//      Short items add their gradient rows in ascending p and update the row in place.  Chunk
//      items write their partial sum; the last chunk of a segment to arrive (atomic ticket)
//      combines the partials IN CHUNK ORDER — the summation tree depends only on the segment,
//      never on the arrival order, so results are reproducible bit for bit — and updates the row.
// HBM-bound: per lookup one gradient row read, per touched row W/m/v read+write.
#include <cstdlib>

#include "rtf_common.cuh"

namespace rtf {

// ------------------------------------------------------------------ parameters
struct BwdParams {
  float* w[RTF_MAX_FIELDS];
  float* s1[RTF_MAX_FIELDS];
  float* s2[RTF_MAX_FIELDS];
  long long rows[RTF_MAX_FIELDS];
  int dim[RTF_MAX_FIELDS];        // per table
  int field_table[RTF_MAX_FIELDS];
  int field_off[RTF_MAX_FIELDS];  // column offset of a field in one grad row
  int n_tables, n_fields, sumD, row_bits, L, pool;
  long long grad_sb;
  rtf_opt opt;
};

// ------------------------------------------------------------------ 1. keys
template <typename IdT>
__global__ void __launch_bounds__(256)
make_keys(const __grid_constant__ BwdParams P, const IdT* __restrict__ ids, long long n,
          long long sb, long long sf, long long sl, uint32_t* __restrict__ keys) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int F = P.n_fields;
  const long long LF = (long long)P.L * F;
  const long long b = p / LF;
  const int r = (int)(p - b * LF);
  const int l = r / F;
  const int f = r - l * F;
  const int t = P.field_table[f];
  const long long id = (long long)__ldg(ids + b * sb + (long long)f * sf + (long long)l * sl);
  uint32_t key;
  if (id < 0 || id >= P.rows[t])
    key = (uint32_t)P.n_tables << P.row_bits;  // sentinel "table": sorts last, never applied
  else
    key = ((uint32_t)t << P.row_bits) | (uint32_t)id;
  keys[p] = key;
}

// ------------------------------------------------------------------ scan utility
// exclusive scan of f(i), i in [0,n): tile sums -> scan of tile sums -> apply
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total) {
  // 256 threads; returns exclusive prefix of v across the block
  __shared__ T warp_tot[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  T base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; ++w) {
    const T t = warp_tot[w];
    if (w < wid) base += t;
    tot += t;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

template <typename T, typename In>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_reduce(In in, long long n, T* __restrict__ tile_sums) {
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  T s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) s += in(base + k);
  T tot;
  block_exclusive_scan<T>(s, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

template <typename T>
__global__ void __launch_bounds__(1024) scan_tile_sums(T* __restrict__ tile_sums, int nt) {
  __shared__ T wsum[32];
  __shared__ T carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < nt; base += 1024) {
    const int i = base + threadIdx.x;
    const T v = i < nt ? tile_sums[i] : 0;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      T y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      T w = wsum[lane], winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        T y = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += y;
      }
      wsum[lane] = winc - w;  // exclusive warp offsets
    }
    __syncthreads();
    const T carry = carry_s;
    const T excl = carry + wsum[wid] + inc - v;
    if (i < nt) tile_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
}

template <typename T, typename In, typename Out>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_apply(In in, Out out, long long n, const T* __restrict__ tile_sums) {
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  T v[SCAN_ITEMS];
  T s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = base + k < n ? in(base + k) : 0;
    s += v[k];
  }
  T tot;
  T ex = block_exclusive_scan<T>(s, &tot) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out(base + k, ex, v[k]);
    ex += v[k];
  }
}

struct InArray {
  const uint32_t* a;
  __device__ uint32_t operator()(long long i) const { return a[i]; }
};
struct OutArray {
  uint32_t* a;
  __device__ void operator()(long long i, uint32_t ex, uint32_t) const { a[i] = ex; }
};
struct InHeadFlag {  // 1 where a new (table,row) segment starts in the sorted keys
  const uint32_t* keys;
  __device__ uint32_t operator()(long long i) const {
    return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
  }
};
struct OutSegments {
  const uint32_t* keys;
  uint32_t* seg_start;  // [n+1]
  int32_t* counters;    // [0] = n_seg, [1] = n_valid_seg
  int32_t* d_num_uniq;  // optional
  long long n;
  uint32_t sentinel_table;
  int row_bits;
  __device__ void operator()(long long i, uint32_t ex, uint32_t flag) const {
    if (flag) seg_start[ex] = (uint32_t)i;
    if (i == n - 1) {
      const uint32_t nseg = ex + flag;
      seg_start[nseg] = (uint32_t)n;
      const int valid = (int)nseg - ((keys[i] >> row_bits) == sentinel_table ? 1 : 0);
      counters[0] = (int)nseg;
      counters[1] = valid;
      if (d_num_uniq) *d_num_uniq = valid;
    }
  }
};

// Work items of step 4.  One uint4 each:
//   short segment (len <= RTF_SEG_CHUNK): {key, start, len, vals[start]}      at items[cap_chunk + j]
//   chunk c of a long segment:            {key, start + c*CH, len_c, slot0}   at items[slot0 + c]
// slot0 = first chunk slot of the segment; long_seg[slot0] = {segment index, #chunks} and
// arrive[slot0] is the segment's arrival counter (arrive[slot0 + 1 + g]: chunk group g's, for
// segments of more than RTF_SEG_GROUP chunks).
struct InSegKinds {  // low 32 bits: 1 for a short segment; high 32 bits: #chunks of a long one
  const uint32_t* seg_start;
  const int32_t* counters;
  __device__ unsigned long long operator()(long long i) const {
    if (i >= counters[1]) return 0ull;
    const uint32_t len = seg_start[i + 1] - seg_start[i];
    if (len <= RTF_SEG_CHUNK) return 1ull;
    return (unsigned long long)((len + RTF_SEG_CHUNK - 1) / RTF_SEG_CHUNK) << 32;
  }
};
struct OutItems {
  const uint32_t* keys;
  const uint32_t* vals;
  const uint32_t* seg_start;
  int32_t* counters;  // [2] = #short items, [3] = #chunk items
  uint4* items;
  uint32_t* short_seg;
  uint2* long_seg;
  uint32_t* arrive;
  uint32_t cap_chunk;
  __device__ void operator()(long long i, unsigned long long ex, unsigned long long v) const {
    const int n_valid = counters[1];
    if (i >= n_valid) return;
    const uint32_t start = seg_start[i], len = seg_start[i + 1] - start;
    const uint32_t key = keys[start];
    if (v & 0xffffffffull) {
      const uint32_t j = (uint32_t)ex;
      items[cap_chunk + j] = make_uint4(key, start, len, vals[start]);
      short_seg[j] = (uint32_t)i;
    } else {
      const uint32_t nch = (uint32_t)(v >> 32), slot0 = (uint32_t)(ex >> 32);
      long_seg[slot0] = make_uint2((uint32_t)i, nch);
      for (uint32_t c = 0; c < nch; ++c) {
        const uint32_t s = start + c * RTF_SEG_CHUNK;
        arrive[slot0 + c] = 0u;   // [slot0]: the segment's counter, [slot0 + 1 + g]: group g's
        items[slot0 + c] = make_uint4(key, s, min((uint32_t)RTF_SEG_CHUNK, start + len - s), slot0);
      }
    }
    if (i == n_valid - 1) {
      const unsigned long long tot = ex + v;
      counters[2] = (int)(uint32_t)tot;
      counters[3] = (int)(uint32_t)(tot >> 32);
    }
  }
};

template <typename T, typename In, typename Out>
static int exclusive_scan(In in, Out out, long long n, T* tile_sums, cudaStream_t st) {
  const long long nt = (n + SCAN_TILE - 1) / SCAN_TILE;
  scan_tile_reduce<T, In><<<(unsigned)nt, SCAN_THREADS, 0, st>>>(in, n, tile_sums);
  scan_tile_sums<T><<<1, 1024, 0, st>>>(tile_sums, (int)nt);
  scan_tile_apply<T, In, Out><<<(unsigned)nt, SCAN_THREADS, 0, st>>>(in, out, n, tile_sums);
  RTF_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------ 2. radix sort
// LSD, stable.  The digit width is chosen per call (<= SORT_MAX_BITS) so that the significant
// key bits take the fewest passes: 29 bits (Criteo: 24 row bits + 5 table bits) = 3 passes of 10.
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 keys per CTA; a warp owns 512 contiguous
constexpr int SORT_MAX_BITS = 10;
constexpr int SORT_MAX_BINS = 1 << SORT_MAX_BITS;

__global__ void __launch_bounds__(SORT_THREADS)
sort_hist(const uint32_t* __restrict__ keys, long long n, int shift, int bits,
          uint32_t* __restrict__ hist, int nblk) {
  __shared__ uint32_t sh[SORT_MAX_BINS];
  const int bins = 1 << bits;
  const uint32_t mask = (uint32_t)bins - 1u;
  for (int i = threadIdx.x; i < bins; i += SORT_THREADS) sh[i] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * SORT_TILE;
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    const long long i = base + r * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&sh[(keys[i] >> shift) & mask], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < bins; d += SORT_THREADS)
    hist[(long long)d * nblk + blockIdx.x] = sh[d];  // digit-major
}

__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
             uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, long long n,
             int shift, int bits, const uint32_t* __restrict__ hist_scanned, int nblk,
             int iota_vals) {
  __shared__ uint32_t whist[SORT_WARPS][SORT_MAX_BINS];
  const int bins = 1 << bits;
  const uint32_t mask = (uint32_t)bins - 1u;
  for (int i = threadIdx.x; i < SORT_WARPS * SORT_MAX_BINS; i += SORT_THREADS)
    (&whist[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const long long wbase = (long long)blockIdx.x * SORT_TILE + (long long)wid * (32 * SORT_ITEMS);
  uint32_t k[SORT_ITEMS], v[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    const long long i = wbase + r * 32 + lane;
    const bool valid = i < n;
    k[r] = valid ? keys_in[i] : 0xffffffffu;
    v[r] = valid ? (iota_vals ? (uint32_t)i : vals_in[i]) : 0u;
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
    rank[r] = 0;
    if (valid) {
      const uint32_t d = (k[r] >> shift) & mask;
      const uint32_t m = __match_any_sync(vmask, d);
      const uint32_t prior = whist[wid][d];
      __syncwarp(vmask);
      if ((m & lt_mask) == 0) whist[wid][d] = prior + __popc(m);  // lowest lane of the match set
      __syncwarp(vmask);
      rank[r] = prior + __popc(m & lt_mask);
    }
  }
  __syncthreads();
  // per-warp counts -> global bases (warps in index order => stable)
  for (int d = threadIdx.x; d < bins; d += SORT_THREADS) {
    uint32_t run = hist_scanned[(long long)d * nblk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      const uint32_t c = whist[w][d];
      whist[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    const long long i = wbase + r * 32 + lane;
    if (i < n) {
      const uint32_t d = (k[r] >> shift) & mask;
      const uint32_t pos = whist[wid][d] + rank[r];
      keys_out[pos] = k[r];
      vals_out[pos] = v[r];
    }
  }
}

// ------------------------------------------------------------------ 4. segment reduce
template <int VEC>
struct Vec {
  float v[VEC];
};
template <int VEC>
__device__ __forceinline__ Vec<VEC> vload(const float* p) {
  Vec<VEC> r;
  if constexpr (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    r.v[0] = *p;
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ Vec<VEC> vload_stream(const float* p) {
  Vec<VEC> r;
  if constexpr (VEC == 4) {
    const float4 t = ldg_nc_f4(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void vstore(float* p, const Vec<VEC>& a) {
  if constexpr (VEC == 4)
    *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
  else
    *p = a.v[0];
}

// pointer to the gradient row of lookup position p
__device__ __forceinline__ const float* grad_row(const BwdParams& P, const float* grad,
                                                 uint32_t p) {
  const int F = P.n_fields;
  const uint32_t LF = (uint32_t)P.L * (uint32_t)F;
  const uint32_t b = p / LF;
  const uint32_t r = p - b * LF;
  const uint32_t l = r / (uint32_t)F;
  const uint32_t f = r - l * (uint32_t)F;
  const long long lo = P.pool == RTF_POOL_NONE ? (long long)l * P.sumD : 0;
  return grad + (long long)b * P.grad_sb + lo + P.field_off[f];
}

// acc[k] (+)= gradient rows of lookup positions vals[start..end) in ascending order, U loads
// in flight.  `first_pos` (when have_first) is vals[start], already known from the work item.
template <int VEC, int G, int VPL>
__device__ __forceinline__ void sum_rows(const BwdParams& P, const float* __restrict__ grad,
                                         const uint32_t* __restrict__ vals, uint32_t start,
                                         uint32_t end, bool have_first, uint32_t first_pos, int nv,
                                         int lg, float scale, Vec<VEC> (&acc)[VPL]) {
  constexpr int U = 4;
#pragma unroll 1
  for (uint32_t j0 = start; j0 < end; j0 += U) {
    const float* src[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      src[u] = nullptr;
      if (j0 + u < end) {
        const uint32_t p = (have_first && j0 + u == start) ? first_pos : __ldg(vals + j0 + u);
        src[u] = grad_row(P, grad, p);
      }
    }
    Vec<VEC> x[U][VPL];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vi = lg + k * G;
        if (src[u] && vi < nv) x[u][k] = vload_stream<VEC>(src[u] + VEC * vi);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (src[u]) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const int vi = lg + k * G;
          if (vi < nv) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              float g = x[u][k].v[e];
              if (P.pool == RTF_POOL_MEAN) g = __fmul_rn(g, scale);
              acc[k].v[e] = __fadd_rn(acc[k].v[e], g);
            }
          }
        }
      }
  }
}

// W / m / v values of one touched row
template <int VEC, int VPL>
struct RowState {
  Vec<VEC> w[VPL], a[VPL], b[VPL];
};

template <int VEC, int G, int VPL>
__device__ __forceinline__ void load_row_state(const BwdParams& P, uint32_t key, int nv, int lg,
                                               RowState<VEC, VPL>& st) {
  const rtf_opt& o = P.opt;
  if (o.kind == RTF_OPT_NONE) return;
  const uint32_t t = key >> P.row_bits;
  const long long off = (long long)(key & ((1u << P.row_bits) - 1u)) * P.dim[t];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = lg + k * G;
    if (vi >= nv) continue;
    st.w[k] = vload<VEC>(P.w[t] + off + VEC * vi);
    if (o.kind >= RTF_OPT_ADAGRAD) st.a[k] = vload<VEC>(P.s1[t] + off + VEC * vi);
    if (o.kind == RTF_OPT_ADAM) st.b[k] = vload<VEC>(P.s2[t] + off + VEC * vi);
  }
}

// apply the optimizer to the row of `key` with summed gradient acc; optional copies out
template <int VEC, int G, int VPL>
__device__ __forceinline__ void finish_row(const BwdParams& P, uint32_t key, uint32_t seg,
                                           int nv, int lg, Vec<VEC> (&acc)[VPL],
                                           RowState<VEC, VPL>& st, uint32_t* uniq_key,
                                           float* uniq_grad, int dim_max) {
  const uint32_t t = key >> P.row_bits;
  const long long row = (long long)(key & ((1u << P.row_bits) - 1u));
  const int dim = P.dim[t];
  if (uniq_key && lg == 0) uniq_key[seg] = key;
  if (uniq_grad) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lg + k * G;
      if (vi < nv) vstore<VEC>(uniq_grad + (long long)seg * dim_max + VEC * vi, acc[k]);
    }
  }
  const rtf_opt& o = P.opt;
  if (o.kind == RTF_OPT_NONE) return;
  const float lr = o.lr_dev ? __ldg(o.lr_dev) : o.lr;   // device scalar under CUDA-graph replay
  float* w = P.w[t] + row * dim;
  float* s1 = P.s1[t] ? P.s1[t] + row * dim : nullptr;
  float* s2 = P.s2[t] ? P.s2[t] + row * dim : nullptr;
  const float two_l2 = __fmul_rn(2.0f, o.l2);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = lg + k * G;
    if (vi >= nv) continue;
    Vec<VEC>& wv = st.w[k];
    Vec<VEC>& a = st.a[k];
    Vec<VEC>& b = st.b[k];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float g = acc[k].v[e];
      if (o.l2 > 0.f) g = __fadd_rn(g, __fmul_rn(two_l2, wv.v[e]));
      if (o.kind == RTF_OPT_SGD) {
        wv.v[e] = __fsub_rn(wv.v[e], __fmul_rn(lr, g));
      } else if (o.kind == RTF_OPT_ADAGRAD) {
        a.v[e] = __fadd_rn(a.v[e], __fmul_rn(g, g));
        wv.v[e] = __fsub_rn(wv.v[e], __fdiv_rn(__fmul_rn(lr, g),
                                               __fadd_rn(__fsqrt_rn(a.v[e]), o.eps)));
      } else {  // Adam (Keras form; lr already carries the bias corrections)
        a.v[e] = __fadd_rn(__fmul_rn(o.beta1, a.v[e]), __fmul_rn(__fsub_rn(1.0f, o.beta1), g));
        b.v[e] = __fadd_rn(__fmul_rn(o.beta2, b.v[e]),
                           __fmul_rn(__fsub_rn(1.0f, o.beta2), __fmul_rn(g, g)));
        wv.v[e] = __fsub_rn(wv.v[e], __fdiv_rn(__fmul_rn(lr, a.v[e]),
                                               __fadd_rn(__fsqrt_rn(b.v[e]), o.eps)));
      }
    }
    vstore<VEC>(w + VEC * vi, wv);
    if (o.kind >= RTF_OPT_ADAGRAD) vstore<VEC>(s1 + VEC * vi, a);
    if (o.kind == RTF_OPT_ADAM) vstore<VEC>(s2 + VEC * vi, b);
  }
}

struct SegWork {
  const uint32_t* vals;       // sorted lookup positions
  const int32_t* counters;    // [0] n_seg [1] n_valid [2] short items [3] chunk items
  const uint4* items;         // [0, cap_chunk): chunk items; [cap_chunk, ...): short items
  const uint32_t* short_seg;  // segment index of short item j (unique-row outputs only)
  const uint2* long_seg;      // [slot0] = (segment index, #chunks)
  uint32_t* arrive;           // [slot0] arrival counter of a long segment's chunks
  float* partials;            // [chunk slot][dim_max]
  uint32_t* uniq_key;
  float* uniq_grad;
  uint32_t cap_chunk;
  int dim_max;
};

// acc = partials[first] + partials[first + stride] + ... (count terms), IN ORDER
template <int VEC, int G, int VPL>
__device__ __forceinline__ void sum_partials(const SegWork& S, uint32_t first, uint32_t count,
                                             uint32_t stride, int nv, int lg, Vec<VEC> (&acc)[VPL]) {
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[k].v[e] = 0.f;
  constexpr int U = 4;  // partials are L2-resident; 4 loads in flight
#pragma unroll 1
  for (uint32_t c0 = 0; c0 < count; c0 += U) {
    Vec<VEC> x[U][VPL];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vi = lg + k * G;
        if (c0 + u < count && vi < nv) {
          const float* src = S.partials + (long long)(first + (c0 + u) * stride) * S.dim_max + VEC * vi;
          if constexpr (VEC == 4) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(src));
            x[u][k].v[0] = t.x; x[u][k].v[1] = t.y; x[u][k].v[2] = t.z; x[u][k].v[3] = t.w;
          } else {
            x[u][k].v[0] = __ldcg(src);
          }
        }
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (c0 + u < count) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const int vi = lg + k * G;
          if (vi < nv) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[k].v[e] = __fadd_rn(acc[k].v[e], x[u][k].v[e]);
          }
        }
      }
  }
}

// Combination of a long segment's chunk partials.  The summation tree depends only on the
// segment's length, never on which thread runs it: chunks form GROUPS of RTF_SEG_GROUP; the last
// chunk of a group to arrive adds the group's partials in chunk order and leaves the group sum in
// the group's first slot; the last GROUP to finish adds the group sums in group order and updates
// the row.  (One level was enough for the Criteo tables; a padding id that fills half of a
// behaviour-sequence batch is a 200 000-row segment = 3 200 partials, which one lane group added
// serially in 0.24 ms — half of the DIN step.)  Kept out of line: rare, and it must not cost the
// streaming path registers.
template <int VEC, int G, int VPL>
__device__ __noinline__ void combine_long(const BwdParams& P, const SegWork& S, uint32_t key,
                                          uint32_t slot0, uint32_t chunk, uint2 ls, int nv, int lg,
                                          uint32_t gmask, int g0) {
  const uint32_t nch = ls.y;
  const uint32_t ngrp = (nch + RTF_SEG_GROUP - 1) / RTF_SEG_GROUP;
  const uint32_t grp = chunk / RTF_SEG_GROUP;
  const uint32_t gsz = min((uint32_t)RTF_SEG_GROUP, nch - grp * RTF_SEG_GROUP);
  Vec<VEC> acc[VPL];
  if (ngrp > 1) {
    // arrive[slot0 + 1 + g] counts the chunks of group g (ngrp + 1 <= nch slots belong to us)
    uint32_t t1 = 0;
    if (lg == 0) t1 = atomicAdd(S.arrive + slot0 + 1 + grp, 1u);
    t1 = __shfl_sync(gmask, t1, g0);
    if (t1 != gsz - 1) return;
    __threadfence();
    if (lg == 0) S.arrive[slot0 + 1 + grp] = 0u;   // the prepared work list can be applied again
    sum_partials<VEC, G, VPL>(S, slot0 + grp * RTF_SEG_GROUP, gsz, 1u, nv, lg, acc);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lg + k * G;
      if (vi < nv)
        vstore<VEC>(S.partials + (long long)(slot0 + grp * RTF_SEG_GROUP) * S.dim_max + VEC * vi, acc[k]);
    }
    __threadfence();
    __syncwarp(gmask);
  }
  if (ngrp > 1) {       // (one group: the caller already holds the last ticket)
    uint32_t t0 = 0;
    if (lg == 0) t0 = atomicAdd(S.arrive + slot0, 1u);
    t0 = __shfl_sync(gmask, t0, g0);
    if (t0 != ngrp - 1) return;
  }
  __threadfence();
  if (lg == 0) S.arrive[slot0] = 0u;
  if (ngrp > 1) sum_partials<VEC, G, VPL>(S, slot0, ngrp, (uint32_t)RTF_SEG_GROUP, nv, lg, acc);
  else sum_partials<VEC, G, VPL>(S, slot0, nch, 1u, nv, lg, acc);
  RowState<VEC, VPL> st;
  load_row_state<VEC, G, VPL>(P, key, nv, lg, st);
  finish_row<VEC, G, VPL>(P, key, ls.x, nv, lg, acc, st, S.uniq_key, S.uniq_grad, S.dim_max);
}

// One launch, one lane-group of G lanes per work item (chunk items first, then short segments).
// The item record carries everything the group needs, so an item costs one round of independent
// loads: for the common 1-2 row segment the gradient rows and the row's W/m/v are all in flight
// together; longer segments stream their gradient rows 4 at a time and load W/m/v afterwards
// (fewer live registers; measured: the kernel lives on resident warps, spills are fatal).
template <int VEC, int G, int VPL, int MINB>
__global__ void __launch_bounds__(256, MINB)
seg_apply(const __grid_constant__ BwdParams P, const __grid_constant__ SegWork S,
          const float* __restrict__ grad) {
  const uint32_t i = (uint32_t)(((long long)blockIdx.x * 256 + threadIdx.x) / G);
  const uint32_t n_chunk = (uint32_t)S.counters[3];
  if (i >= n_chunk + (uint32_t)S.counters[2]) return;
  const int lg = (int)(threadIdx.x % G);
  const uint4 rec = __ldg(S.items + (i < n_chunk ? i : S.cap_chunk + (i - n_chunk)));
  const uint32_t key = rec.x, start = rec.y, len = rec.z;
  const int nv = P.dim[key >> P.row_bits] / VEC;
  const float scale = __fdiv_rn(1.0f, (float)P.L);
  Vec<VEC> acc[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[k].v[e] = 0.f;
  if (i >= n_chunk) {
    // ---- short segment
    RowState<VEC, VPL> st;
    if (len <= 2) {
      // everything in one round: W/m/v and up to two gradient rows
      load_row_state<VEC, G, VPL>(P, key, nv, lg, st);
      const float* r0 = grad_row(P, grad, rec.w);
      const float* r1 = len == 2 ? grad_row(P, grad, __ldg(S.vals + start + 1)) : nullptr;
      Vec<VEC> x0[VPL], x1[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vi = lg + k * G;
        if (vi < nv) {
          x0[k] = vload_stream<VEC>(r0 + VEC * vi);
          if (r1) x1[k] = vload_stream<VEC>(r1 + VEC * vi);
        }
      }
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vi = lg + k * G;
        if (vi < nv) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            float g = x0[k].v[e];
            if (P.pool == RTF_POOL_MEAN) g = __fmul_rn(g, scale);
            acc[k].v[e] = __fadd_rn(acc[k].v[e], g);
            if (r1) {
              float h = x1[k].v[e];
              if (P.pool == RTF_POOL_MEAN) h = __fmul_rn(h, scale);
              acc[k].v[e] = __fadd_rn(acc[k].v[e], h);
            }
          }
        }
      }
    } else {
      sum_rows<VEC, G, VPL>(P, grad, S.vals, start, start + len, true, rec.w, nv, lg, scale, acc);
      load_row_state<VEC, G, VPL>(P, key, nv, lg, st);
    }
    uint32_t seg = 0;
    if (S.uniq_key || S.uniq_grad) seg = __ldg(S.short_seg + (i - n_chunk));
    finish_row<VEC, G, VPL>(P, key, seg, nv, lg, acc, st, S.uniq_key, S.uniq_grad, S.dim_max);
    return;
  }
  // ---- chunk of a long segment: partial sum; the last chunk to arrive combines in chunk order
  const int g0 = (int)(threadIdx.x & 31) / G * G;  // first lane of this group inside its warp
  const uint32_t gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << g0);
  const uint32_t slot0 = rec.w;
  sum_rows<VEC, G, VPL>(P, grad, S.vals, start, start + len, false, 0u, nv, lg, scale, acc);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = lg + k * G;
    if (vi < nv) vstore<VEC>(S.partials + (long long)i * S.dim_max + VEC * vi, acc[k]);
  }
  __threadfence();
  __syncwarp(gmask);
  const uint2 ls = __ldg(S.long_seg + slot0);  // (segment index, #chunks)
  if (ls.y <= RTF_SEG_GROUP) {   // one level: the ticket is taken here, the rare winner goes out of line
    uint32_t ticket = 0;
    if (lg == 0) ticket = atomicAdd(S.arrive + slot0, 1u);
    ticket = __shfl_sync(gmask, ticket, g0);
    if (ticket != ls.y - 1) return;
  }
  combine_long<VEC, G, VPL>(P, S, key, slot0, i - slot0, ls, nv, lg, gmask, g0);
}

template <int VEC, int G, int VPL, int MINB>
static int launch_apply(const BwdParams& P, const SegWork& S, const float* grad,
                        long long max_items, cudaStream_t st) {
  const long long blocks = (max_items * G + 255) / 256;
  seg_apply<VEC, G, VPL, MINB><<<(unsigned)blocks, 256, 0, st>>>(P, S, grad);
  RTF_CHECK_LAUNCH();
  return 0;
}

// Tuning record (B200, DLRM cfg: 26 tables, D = 128, B = 65 536, uniform ids; CUDA events):
//   round 1 (seg_short + seg_partial + seg_combine, 4-5 dependent loads per segment)  0.67 ms
//   persistent grid-stride groups with the next record prefetched, W/m/v early:
//     48 regs 0.94 ms, 64 regs 0.84 ms — 150-400 B of spills per thread go to L2 and sit on
//     every item's critical path; occupancy 32-40 warps/SM
//   this kernel, one item per group, resident blocks/SM 4 / 5 / 6 / 8 (56 / 48 / 40 / 32 regs):
//     0.539 / 0.498 / 0.498 / 0.472 ms  -> 8 blocks (64 warps/SM, ~110 B spill) = 5.3 TB/s,
//     0.81 of the measured HBM copy peak.
template <int VEC, int G, int VPL>
static int launch_segments(const BwdParams& P, const SegWork& S, const float* grad,
                           long long max_items, cudaStream_t st) {
  return launch_apply<VEC, G, VPL, (VPL == 1 ? 8 : 4)>(P, S, grad, max_items, st);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WsLayout {
  size_t keys0, keys1, vals0, vals1, hist, tile_sums, seg_start, counters, items, short_seg,
      long_seg, arrive, partials, total;
  size_t cap_chunk;
};
static WsLayout ws_layout(long long n, int dim_max) {
  WsLayout w;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  const size_t nblk = (nn + SORT_TILE - 1) / SORT_TILE;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  w.keys0 = take(nn * 4);
  w.keys1 = take(nn * 4);
  w.vals0 = take(nn * 4);
  w.vals1 = take(nn * 4);
  w.hist = take((size_t)SORT_MAX_BINS * nblk * 4);
  const size_t scan_n = nn > (size_t)SORT_MAX_BINS * nblk ? nn : (size_t)SORT_MAX_BINS * nblk;
  w.tile_sums = take(((scan_n + SCAN_TILE - 1) / SCAN_TILE + 1) * 8);
  w.seg_start = take((nn + 1) * 4);
  w.counters = take(16);
  // a long segment of len > CH rows has ceil(len/CH) <= 2 len/CH chunks: at most 2n/CH slots
  w.cap_chunk = 2 * nn / RTF_SEG_CHUNK + 4;
  w.items = take((w.cap_chunk + nn) * 16);
  w.short_seg = take(nn * 4);
  w.long_seg = take(w.cap_chunk * 8);
  w.arrive = take(w.cap_chunk * 4);
  w.partials = take(w.cap_chunk * (size_t)dim_max * 4);
  w.total = o;
  return w;
}

}  // namespace rtf

extern "C" int rtf_embed_bwd_workspace(int64_t n_lookups, int dim_max, size_t* bytes) {
  if (!bytes || n_lookups < 0 || dim_max <= 0) return RTF_E_ARG;
  *bytes = rtf::ws_layout(n_lookups, dim_max).total;
  return 0;
}

// phase bit 0: keys + sort + segments (needs only the ids); bit 1: segment reduce + optimizer
static int embed_bwd_impl(int phase, float* const* weights, float* const* state1,
                          float* const* state2, const int64_t* rows, const int32_t* dims,
                          int n_tables, const int32_t* field_table, int n_fields,
                          const void* d_ids, int ids_i64, int64_t B, int L, int64_t ids_sb,
                          int64_t ids_sf, int64_t ids_sl, int pool, const float* d_grad,
                          int64_t grad_sb, const rtf_opt* opt, uint32_t* d_uniq_key,
                          float* d_uniq_grad, int32_t* d_num_uniq, int* row_bits_out,
                          void* d_workspace, size_t workspace_bytes, void* stream) {
  using namespace rtf;
  static const rtf_opt kNoOpt = {RTF_OPT_NONE, 0.f, 0.f, 0.f, 0.f, 0.f, nullptr};
  static float* const kNoPtrs[RTF_MAX_FIELDS] = {};
  if (!(phase & 2)) {
    opt = &kNoOpt;
    weights = kNoPtrs;
    state1 = state2 = nullptr;
  }
  if (!weights || !rows || !dims || !field_table || !opt) return RTF_E_ARG;
  if (n_tables <= 0 || n_fields <= 0 || B < 0 || L <= 0) return RTF_E_ARG;
  if (B > 0 && !d_workspace) return RTF_E_ARG;
  if (B > 0 && (phase & 1) && !d_ids) return RTF_E_ARG;
  if (B > 0 && (phase & 2) && !d_grad) return RTF_E_ARG;
  if (n_tables > RTF_MAX_FIELDS || n_fields > RTF_MAX_FIELDS) return RTF_E_RANGE;
  if (pool < RTF_POOL_NONE || pool > RTF_POOL_MEAN) return RTF_E_ARG;
  if (opt->kind < RTF_OPT_NONE || opt->kind > RTF_OPT_ADAM) return RTF_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;

  BwdParams P;
  long long rows_max = 1;
  int dim_max = 0;
  bool vec_ok = ((uintptr_t)d_grad % 16 == 0) && (grad_sb % 4 == 0) &&
                (!d_uniq_grad || (uintptr_t)d_uniq_grad % 16 == 0);
  for (int t = 0; t < n_tables; ++t) {
    if (rows[t] <= 0 || dims[t] <= 0) return RTF_E_ARG;
    P.w[t] = weights[t];
    P.s1[t] = state1 ? state1[t] : nullptr;
    P.s2[t] = state2 ? state2[t] : nullptr;
    P.rows[t] = rows[t];
    P.dim[t] = dims[t];
    if (opt->kind != RTF_OPT_NONE && !P.w[t]) return RTF_E_ARG;
    if (opt->kind >= RTF_OPT_ADAGRAD && !P.s1[t]) return RTF_E_ARG;
    if (opt->kind == RTF_OPT_ADAM && !P.s2[t]) return RTF_E_ARG;
    if (dims[t] % 4 || (uintptr_t)P.w[t] % 16 || (uintptr_t)P.s1[t] % 16 ||
        (uintptr_t)P.s2[t] % 16)
      vec_ok = false;
    if (rows[t] > rows_max) rows_max = rows[t];
    if (dims[t] > dim_max) dim_max = dims[t];
  }
  int off = 0;
  for (int f = 0; f < n_fields; ++f) {
    const int t = field_table[f];
    if (t < 0 || t >= n_tables) return RTF_E_ARG;
    P.field_table[f] = t;
    P.field_off[f] = off;
    if (off % 4) vec_ok = false;
    off += dims[t];
  }
  P.sumD = off;
  if (off % 4) vec_ok = false;
  int row_bits = 1;
  while (((long long)1 << row_bits) < rows_max) ++row_bits;
  int table_bits = 1;
  while ((1 << table_bits) < n_tables + 1) ++table_bits;  // +1: sentinel table for bad ids
  if (row_bits + table_bits > 32) return RTF_E_RANGE;
  if (row_bits_out) *row_bits_out = row_bits;
  P.n_tables = n_tables;
  P.n_fields = n_fields;
  P.row_bits = row_bits;
  P.L = L;
  P.pool = pool;
  P.grad_sb = grad_sb;
  P.opt = *opt;

  const long long n = B * (long long)L * n_fields;
  if (n >= 0x7fffffffLL) return RTF_E_RANGE;
  if (n == 0) {
    if (d_num_uniq) cudaMemsetAsync(d_num_uniq, 0, 4, st);
    return 0;
  }
  if (dim_max > (vec_ok ? 512 : 128)) return RTF_E_RANGE;
  const WsLayout W = ws_layout(n, dim_max);
  if (workspace_bytes < W.total) return RTF_E_WORKSPACE;
  if ((uintptr_t)d_workspace % 256) return RTF_E_ALIGN;
  char* ws = (char*)d_workspace;
  uint32_t* keys[2] = {(uint32_t*)(ws + W.keys0), (uint32_t*)(ws + W.keys1)};
  uint32_t* vals[2] = {(uint32_t*)(ws + W.vals0), (uint32_t*)(ws + W.vals1)};
  uint32_t* hist = (uint32_t*)(ws + W.hist);
  uint32_t* tile_sums = (uint32_t*)(ws + W.tile_sums);
  uint32_t* seg_start = (uint32_t*)(ws + W.seg_start);
  int32_t* counters = (int32_t*)(ws + W.counters);

  const int nblk = (int)((n + SORT_TILE - 1) / SORT_TILE);
  const int total_bits = row_bits + table_bits;
  const int npass = (total_bits + SORT_MAX_BITS - 1) / SORT_MAX_BITS;
  const int bits = (total_bits + npass - 1) / npass;
  int cur = npass & 1;  // buffer holding the sorted keys/vals after npass ping-pong passes
  if (phase & 1) {
  cur = 0;
  // 1. keys
  const unsigned kb = (unsigned)((n + 255) / 256);
  if (ids_i64)
    make_keys<int64_t><<<kb, 256, 0, st>>>(P, (const int64_t*)d_ids, n, ids_sb, ids_sf, ids_sl,
                                           keys[0]);
  else
    make_keys<int32_t><<<kb, 256, 0, st>>>(P, (const int32_t*)d_ids, n, ids_sb, ids_sf, ids_sl,
                                           keys[0]);
  RTF_CHECK_LAUNCH();

  // 2. LSD radix sort over the significant bits
  for (int shift = 0, pass = 0; pass < npass; shift += bits, ++pass) {
    sort_hist<<<nblk, SORT_THREADS, 0, st>>>(keys[cur], n, shift, bits, hist, nblk);
    int rc = exclusive_scan<uint32_t>(InArray{hist}, OutArray{hist}, (long long)(1 << bits) * nblk,
                                      tile_sums, st);
    if (rc) return rc;
    sort_scatter<<<nblk, SORT_THREADS, 0, st>>>(keys[cur], vals[cur], keys[cur ^ 1],
                                                vals[cur ^ 1], n, shift, bits, hist, nblk,
                                                pass == 0);
    RTF_CHECK_LAUNCH();
    cur ^= 1;
  }

  // 3. segments
  cudaError_t ce = cudaMemsetAsync(counters, 0, 16, st);
  if (ce != cudaSuccess) return (int)ce;
  {
    OutSegments outseg{keys[cur], seg_start, counters, d_num_uniq, n, (uint32_t)n_tables,
                       row_bits};
    int rc = exclusive_scan<uint32_t>(InHeadFlag{keys[cur]}, outseg, n, tile_sums, st);
    if (rc) return rc;
  }
  // 3b. work items: one per short segment, one per chunk of a long segment
  {
    OutItems outit{keys[cur], vals[cur], seg_start, counters, (uint4*)(ws + W.items),
                   (uint32_t*)(ws + W.short_seg), (uint2*)(ws + W.long_seg),
                   (uint32_t*)(ws + W.arrive), (uint32_t)W.cap_chunk};
    int rc = exclusive_scan<unsigned long long>(InSegKinds{seg_start, counters}, outit, n,
                                                (unsigned long long*)tile_sums, st);
    if (rc) return rc;
  }
  }  // phase & 1
  if (!(phase & 2)) return 0;

  // 4. segment reduce + optimizer
  SegWork S;
  S.vals = vals[cur];
  S.counters = counters;
  S.items = (const uint4*)(ws + W.items);
  S.short_seg = (const uint32_t*)(ws + W.short_seg);
  S.long_seg = (const uint2*)(ws + W.long_seg);
  S.arrive = (uint32_t*)(ws + W.arrive);
  S.partials = (float*)(ws + W.partials);
  S.uniq_key = d_uniq_key;
  S.uniq_grad = d_uniq_grad;
  S.cap_chunk = (uint32_t)W.cap_chunk;
  S.dim_max = dim_max;

  // upper bound on the work items: a table with n_t lookups and R_t rows has at most
  // min(n_t, R_t) segments and n_t/CH + min(R_t, n_t/CH) chunks of long segments
  long long max_items = 0;
  {
    long long per_table[RTF_MAX_FIELDS] = {};
    for (int f = 0; f < n_fields; ++f) per_table[field_table[f]] += B * (long long)L;
    for (int t = 0; t < n_tables; ++t) {
      const long long nt = per_table[t];
      const long long b = rows[t] + 2 * (nt / RTF_SEG_CHUNK) + 2;
      max_items += nt < b ? nt : b;
    }
    if (max_items > n) max_items = n;
    if (max_items < 1) max_items = 1;
  }
  if (vec_ok) {
    const int nv = dim_max / 4;
    int G = 1;
    while (G < nv && G < 32) G <<= 1;
    const int vpl = (nv + G - 1) / G;
    if (vpl == 1) {
      switch (G) {
        case 1: return launch_segments<4, 1, 1>(P, S, d_grad, max_items, st);
        case 2: return launch_segments<4, 2, 1>(P, S, d_grad, max_items, st);
        case 4: return launch_segments<4, 4, 1>(P, S, d_grad, max_items, st);
        case 8: return launch_segments<4, 8, 1>(P, S, d_grad, max_items, st);
        case 16: return launch_segments<4, 16, 1>(P, S, d_grad, max_items, st);
        default: return launch_segments<4, 32, 1>(P, S, d_grad, max_items, st);
      }
    }
    if (vpl == 2) return launch_segments<4, 32, 2>(P, S, d_grad, max_items, st);
    return launch_segments<4, 32, 4>(P, S, d_grad, max_items, st);
  }
  {
    int G = 1;
    while (G < dim_max && G < 32) G <<= 1;
    const int vpl = (dim_max + G - 1) / G;
    if (vpl == 1) {
      switch (G) {
        case 1: return launch_segments<1, 1, 1>(P, S, d_grad, max_items, st);
        case 2: return launch_segments<1, 2, 1>(P, S, d_grad, max_items, st);
        case 4: return launch_segments<1, 4, 1>(P, S, d_grad, max_items, st);
        case 8: return launch_segments<1, 8, 1>(P, S, d_grad, max_items, st);
        case 16: return launch_segments<1, 16, 1>(P, S, d_grad, max_items, st);
        default: return launch_segments<1, 32, 1>(P, S, d_grad, max_items, st);
      }
    }
    if (vpl == 2) return launch_segments<1, 32, 2>(P, S, d_grad, max_items, st);
    return launch_segments<1, 32, 4>(P, S, d_grad, max_items, st);
  }
}

extern "C" int rtf_embed_bwd(float* const* weights, float* const* state1, float* const* state2,
                             const int64_t* rows, const int32_t* dims, int n_tables,
                             const int32_t* field_table, int n_fields, const void* d_ids,
                             int ids_i64, int64_t B, int L, int64_t ids_sb, int64_t ids_sf,
                             int64_t ids_sl, int pool, const float* d_grad, int64_t grad_sb,
                             const rtf_opt* opt, uint32_t* d_uniq_key, float* d_uniq_grad,
                             int32_t* d_num_uniq, int* row_bits_out, void* d_workspace,
                             size_t workspace_bytes, void* stream) {
  return embed_bwd_impl(3, weights, state1, state2, rows, dims, n_tables, field_table, n_fields,
                        d_ids, ids_i64, B, L, ids_sb, ids_sf, ids_sl, pool, d_grad, grad_sb, opt,
                        d_uniq_key, d_uniq_grad, d_num_uniq, row_bits_out, d_workspace,
                        workspace_bytes, stream);
}

extern "C" int rtf_embed_bwd_prepare(const int64_t* rows, const int32_t* dims, int n_tables,
                                     const int32_t* field_table, int n_fields, const void* d_ids,
                                     int ids_i64, int64_t B, int L, int64_t ids_sb, int64_t ids_sf,
                                     int64_t ids_sl, int32_t* d_num_uniq, int* row_bits_out,
                                     void* d_workspace, size_t workspace_bytes, void* stream) {
  return embed_bwd_impl(1, nullptr, nullptr, nullptr, rows, dims, n_tables, field_table, n_fields,
                        d_ids, ids_i64, B, L, ids_sb, ids_sf, ids_sl, RTF_POOL_NONE, nullptr, 0,
                        nullptr, nullptr, nullptr, d_num_uniq, row_bits_out, d_workspace,
                        workspace_bytes, stream);
}

extern "C" int rtf_embed_bwd_apply(float* const* weights, float* const* state1,
                                   float* const* state2, const int64_t* rows, const int32_t* dims,
                                   int n_tables, const int32_t* field_table, int n_fields,
                                   int64_t B, int L, int pool, const float* d_grad,
                                   int64_t grad_sb, const rtf_opt* opt, uint32_t* d_uniq_key,
                                   float* d_uniq_grad, void* d_workspace, size_t workspace_bytes,
                                   void* stream) {
  return embed_bwd_impl(2, weights, state1, state2, rows, dims, n_tables, field_table, n_fields,
                        nullptr, 0, B, L, 0, 0, 0, pool, d_grad, grad_sb, opt, d_uniq_key,
                        d_uniq_grad, nullptr, nullptr, d_workspace, workspace_bytes, stream);
}
