// K1 — fused multi-table embedding gather (+ sum/mean pooling over the length axis).
//
// Replaces the F strided-slices + F gathers + concat the reference issues per step
// (src/ctr/dlrm/model.py:45-46 and siblings) with one launch
// that reads each table row once with 128-bit loads and writes the concatenated
// (B, [L,] sumD) output once.
//
// Mapping: a group of G lanes (G = pow2 >= dim/4, <= 32) owns one output row; each
// lane moves VPL float4.  Every group keeps U independent rows in flight, so a
// 256-thread CTA at D=128 has 8 warps x 4 x 512 B = 16 KB of gather loads
// outstanding — enough to cover HBM latency at full occupancy.  HBM-bound:
// algorithmic bytes per lookup = id + row read + row write.
#include "rtf_common.cuh"

namespace rtf {

struct FwdFields {
  const float* table[RTF_MAX_FIELDS];
  long long rows[RTF_MAX_FIELDS];
  int dim[RTF_MAX_FIELDS];
  int off[RTF_MAX_FIELDS];  // column offset of the field inside one (sumD) output row
  int n_fields;             // fields in this launch
  int field0;               // index of this launch's first field in the full field list
  int sumD;                 // width of one full output row (all fields)
  int skip_invalid;         // 1: an out-of-range id leaves its output row untouched (no zeros, no
                            // error flag) — the owner-gather exchange skips foreign lookups this way
};

// Register cap: with one float4 per lane the kernel is pure latency hiding, so 32 registers
// (64 resident warps/SM, ~60 B of spill) beat 40 and 64 registers: 240 vs 252 vs 290 us on the
// DLRM configuration (ncu gpu__time_duration, profiles/README.md).
template <typename IdT, int G, int VPL>
__global__ void __launch_bounds__(256, VPL == 1 ? 8 : 4)
embed_fwd_vec(const __grid_constant__ FwdFields P, const IdT* __restrict__ ids, long long B,
              int L, long long sb, long long sf, long long sl, int pool,
              float* __restrict__ out, long long out_sb, int32_t* err) {
  constexpr int U = 4;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gid = tid / G;
  const int lg = (int)(threadIdx.x % G);
  const long long n_groups = (long long)gridDim.x * blockDim.x / G;
  const int F = P.n_fields;

  if (pool == RTF_POOL_NONE) {
    // a group owns U consecutive (b, l, f) items: one decode (a 32-bit division whenever the
    // item count allows — the emulated 64-bit one cost ~100 issue slots per row), then steps
    const long long LF = (long long)L * F;
    const long long n_items = B * LF;
    const long long it0 = gid * U;
    if (it0 >= n_items) return;
    long long b;
    int l, f;
    if (n_items < 0x7fffffffLL) {
      const unsigned q = (unsigned)it0 / (unsigned)LF;
      const unsigned r = (unsigned)it0 - q * (unsigned)LF;
      b = q;
      l = (int)(r / (unsigned)F);
      f = (int)(r - (unsigned)l * (unsigned)F);
    } else {
      b = it0 / LF;
      const int r = (int)(it0 - b * LF);
      l = r / F;
      f = r - l * F;
    }
    const float* src[U];
    float* dst[U];
    int nv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      src[u] = nullptr;
      dst[u] = nullptr;
      nv[u] = 0;
      if (it0 + u < n_items) {
        const long long id = load_id(ids, b * sb + (long long)(P.field0 + f) * sf + l * sl,
                                     P.rows[f], P.skip_invalid ? nullptr : err);
        nv[u] = (id < 0 && P.skip_invalid) ? 0 : (P.dim[f] >> 2);
        if (id >= 0) src[u] = P.table[f] + id * P.dim[f];
        dst[u] = out + b * out_sb + (long long)l * P.sumD + P.off[f];
        if (++f == F) {
          f = 0;
          if (++l == L) {
            l = 0;
            ++b;
          }
        }
      }
    }
    float4 v[U][VPL];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vi = lg + k * G;
        v[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vi < nv[u] && src[u]) v[u][k] = ldg_nc_f4(src[u] + 4 * vi);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vi = lg + k * G;
        if (vi < nv[u]) stg_cs_f4(dst[u] + 4 * vi, v[u][k]);
      }
  } else {
    // pooled: one group per (b, f); rows are added in ascending l (fixed fp32 order)
    const long long n_items = B * F;
    for (long long it = gid; it < n_items; it += n_groups) {
      const long long b = it / F;
      const int f = (int)(it - b * F);
      const int dim = P.dim[f];
      const int nv = dim >> 2;
      const long long rows = P.rows[f];
      const float* table = P.table[f];
      const long long idbase = b * sb + (long long)(P.field0 + f) * sf;
      float4 acc[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int l0 = 0; l0 < L; l0 += U) {
        const float* src[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          src[u] = nullptr;
          if (l0 + u < L) {
            const long long id = load_id(ids, idbase + (long long)(l0 + u) * sl, rows, err);
            if (id >= 0) src[u] = table + id * dim;
          }
        }
        float4 v[U][VPL];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int k = 0; k < VPL; ++k) {
            const int vi = lg + k * G;
            v[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vi < nv && src[u]) v[u][k] = ldg_nc_f4(src[u] + 4 * vi);
          }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (l0 + u < L) {
#pragma unroll
            for (int k = 0; k < VPL; ++k) acc[k] = f4_add(acc[k], v[u][k]);
          }
      }
      float* dst = out + b * out_sb + P.off[f];
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vi = lg + k * G;
        if (vi < nv) {
          float4 a = acc[k];
          if (pool == RTF_POOL_MEAN) {
            const float fl = (float)L;
            a = make_float4(__fdiv_rn(a.x, fl), __fdiv_rn(a.y, fl), __fdiv_rn(a.z, fl),
                            __fdiv_rn(a.w, fl));
          }
          stg_cs_f4(dst + 4 * vi, a);
        }
      }
    }
  }
}

// Scalar path for dims that are not a multiple of 4 or misaligned buffers:
// one thread per output element (still a single fused launch).
template <typename IdT>
__global__ void __launch_bounds__(256)
embed_fwd_scalar(const __grid_constant__ FwdFields P, const IdT* __restrict__ ids, long long B,
                 int L, long long sb, long long sf, long long sl, int pool,
                 float* __restrict__ out, long long out_sb, int32_t* err, int sumD_launch) {
  const int Lo = pool == RTF_POOL_NONE ? L : 1;
  const long long total = B * Lo * sumD_launch;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % sumD_launch);
    const long long bl = e / sumD_launch;
    const int lo = (int)(bl % Lo);
    const long long b = bl / Lo;
    // locate the field holding launch-local column c
    int f = 0;
    const int c0 = P.off[0];
    while (f + 1 < P.n_fields && P.off[f + 1] - c0 <= c) ++f;
    const int d = c - (P.off[f] - c0);
    const long long idbase = b * sb + (long long)(P.field0 + f) * sf;
    float acc = 0.f;
    if (pool == RTF_POOL_NONE) {
      const long long id = load_id(ids, idbase + (long long)lo * sl, P.rows[f], err);
      if (id >= 0) acc = __ldg(P.table[f] + id * P.dim[f] + d);
    } else {
      for (int l = 0; l < L; ++l) {
        const long long id = load_id(ids, idbase + (long long)l * sl, P.rows[f], err);
        const float x = id >= 0 ? __ldg(P.table[f] + id * P.dim[f] + d) : 0.f;
        acc = __fadd_rn(acc, x);
      }
      if (pool == RTF_POOL_MEAN) acc = __fdiv_rn(acc, (float)L);
    }
    out[b * out_sb + (long long)lo * P.sumD + P.off[f] + d] = acc;
  }
}

// skip-invalid gather (multi-GPU owner gather over the GLOBAL batch: 7 of 8 lookups of a row-wise
// table belong to another rank and are skipped).  A lane tests ONE (b, l, f) item, the warp then
// copies only the valid rows, up to 4 in flight — the generic kernel spends a whole lane-group
// (decode + redundant id loads) on every skipped item: 0.48 ms for 2.4 M items at 8 GPUs of which
// 0.35 M are valid.
template <typename IdT>
__global__ void __launch_bounds__(256)
embed_fwd_skip(const __grid_constant__ FwdFields P, const IdT* __restrict__ ids, long long B, int L,
               long long sb, long long sf, long long sl, float* __restrict__ out, long long out_sb) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int F = P.n_fields;
  const long long LF = (long long)L * F, n_items = B * LF;
  const long long it = warp * 32 + lane;
  const float* src = nullptr;
  float* dst = nullptr;
  int nv = 0;
  if (it < n_items) {
    const long long b = it / LF;
    const int r = (int)(it - b * LF);
    const int l = r / F, f = r - l * F;
    const long long id = (long long)__ldg(ids + b * sb + (long long)(P.field0 + f) * sf + l * sl);
    if (id >= 0 && id < P.rows[f]) {
      src = P.table[f] + id * P.dim[f];
      dst = out + b * out_sb + (long long)l * P.sumD + P.off[f];
      nv = P.dim[f] >> 2;
    }
  }
  unsigned m = __ballot_sync(0xffffffffu, src != nullptr);
  while (m) {
    const float* s[4];
    float* d[4];
    int n[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int l0 = m ? __ffs(m) - 1 : 0;
      const bool on = m != 0;
      if (on) m &= m - 1;
      s[u] = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)src, l0));
      d[u] = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, (unsigned long long)dst, l0));
      n[u] = on ? __shfl_sync(0xffffffffu, nv, l0) : 0;
      if (!on) n[u] = 0;
    }
    for (int v0 = 0; v0 < 128; v0 += 32) {   // dim <= 512: at most 4 float4 per lane and row
      const int vi = v0 + lane;
      float4 v[4];
      bool any = false;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vi < n[u]) {
          v[u] = ldg_nc_f4(s[u] + 4 * vi);
          any = true;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (vi < n[u]) stg_cs_f4(d[u] + 4 * vi, v[u]);
      if (!__any_sync(0xffffffffu, any)) break;
    }
  }
}

template <typename IdT>
static int launch_fwd(const FwdFields& P, int dim_max, bool vec_ok, const void* d_ids,
                      long long B, int L, long long sb, long long sf, long long sl, int pool,
                      float* out, long long out_sb, int32_t* err, cudaStream_t st) {
  const IdT* ids = (const IdT*)d_ids;
  const int Lo = pool == RTF_POOL_NONE ? L : 1;
  if (!vec_ok) {
    int sumD_launch = 0;
    for (int f = 0; f < P.n_fields; ++f) sumD_launch += P.dim[f];
    const long long total = B * Lo * sumD_launch;
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 64) blocks = kNumSMs * 64;
    embed_fwd_scalar<IdT><<<(unsigned)blocks, 256, 0, st>>>(P, ids, B, L, sb, sf, sl, pool, out,
                                                            out_sb, err, sumD_launch);
    RTF_CHECK_LAUNCH();
    return 0;
  }
  if (P.skip_invalid && pool == RTF_POOL_NONE) {
    const long long n_it = B * (long long)L * P.n_fields;
    const long long blocks = (n_it + 255) / 256;   // one lane per item
    if (blocks > 0x7fffffffLL) return RTF_E_RANGE;
    embed_fwd_skip<IdT><<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(P, ids, B, L, sb, sf, sl,
                                                                            out, out_sb);
    RTF_CHECK_LAUNCH();
    return 0;
  }
  const int nv = dim_max / 4;
  int G = 1;
  while (G < nv && G < 32) G <<= 1;
  const int vpl = (nv + G - 1) / G;
  const long long n_items = B * Lo * P.n_fields;
  long long n_groups = pool == RTF_POOL_NONE ? (n_items + 3) / 4 : n_items;
  long long blocks = (n_groups * G + 255) / 256;
  if (pool != RTF_POOL_NONE && blocks > kNumSMs * 32) blocks = kNumSMs * 32;  // grid-stride
  if (blocks < 1) blocks = 1;
  if (blocks > 0x7fffffffLL) return RTF_E_RANGE;
#define RTF_FWD_CASE(GG, VV)                                                                 \
  embed_fwd_vec<IdT, GG, VV><<<(unsigned)blocks, 256, 0, st>>>(P, ids, B, L, sb, sf, sl, pool, \
                                                               out, out_sb, err)
  if (vpl == 1) {
    switch (G) {
      case 1: RTF_FWD_CASE(1, 1); break;
      case 2: RTF_FWD_CASE(2, 1); break;
      case 4: RTF_FWD_CASE(4, 1); break;
      case 8: RTF_FWD_CASE(8, 1); break;
      case 16: RTF_FWD_CASE(16, 1); break;
      default: RTF_FWD_CASE(32, 1); break;
    }
  } else if (vpl == 2) {
    RTF_FWD_CASE(32, 2);
  } else if (vpl <= 4) {
    RTF_FWD_CASE(32, 4);
  } else {
    return RTF_E_RANGE;  // dim > 512
  }
#undef RTF_FWD_CASE
  RTF_CHECK_LAUNCH();
  return 0;
}

}  // namespace rtf

extern "C" int rtf_embed_fwd(const float* const* tables, const int64_t* rows,
                             const int32_t* dims, int n_fields, const void* d_ids, int ids_i64,
                             int64_t B, int L, int64_t ids_sb, int64_t ids_sf, int64_t ids_sl,
                             int pool, float* d_out, int64_t out_sb, int32_t* d_err,
                             void* stream) {
  using namespace rtf;
  if (!tables || !rows || !dims) return RTF_E_ARG;
  if (n_fields <= 0 || B < 0 || L <= 0) return RTF_E_ARG;
  const int skip_invalid = (pool & RTF_POOL_SKIP_INVALID) ? 1 : 0;
  pool &= ~RTF_POOL_SKIP_INVALID;
  if (pool < RTF_POOL_NONE || pool > RTF_POOL_MEAN) return RTF_E_ARG;
  if (skip_invalid && pool != RTF_POOL_NONE) return RTF_E_ARG;
  if (B == 0) return 0;  // empty batch: nothing to read or write
  if (!d_ids || !d_out) return RTF_E_ARG;
  long long sumD = 0;
  for (int f = 0; f < n_fields; ++f) {
    if (!tables[f] || rows[f] <= 0 || dims[f] <= 0) return RTF_E_ARG;
    sumD += dims[f];
  }
  if (sumD > 0x7fffffffLL) return RTF_E_RANGE;
  const long long need = (pool == RTF_POOL_NONE ? (long long)L : 1LL) * sumD;
  if (out_sb < need) return RTF_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;

  int off = 0;
  for (int f0 = 0; f0 < n_fields; f0 += RTF_MAX_FIELDS) {
    FwdFields P;
    P.n_fields = n_fields - f0 < RTF_MAX_FIELDS ? n_fields - f0 : RTF_MAX_FIELDS;
    P.field0 = f0;
    P.sumD = (int)sumD;
    P.skip_invalid = skip_invalid;
    int dim_max = 0;
    bool vec_ok = ((uintptr_t)d_out % 16 == 0) && (out_sb % 4 == 0) && (sumD % 4 == 0);
    for (int f = 0; f < P.n_fields; ++f) {
      P.table[f] = tables[f0 + f];
      P.rows[f] = rows[f0 + f];
      P.dim[f] = dims[f0 + f];
      P.off[f] = off;
      if (off % 4 || P.dim[f] % 4 || (uintptr_t)P.table[f] % 16) vec_ok = false;
      if (P.dim[f] > dim_max) dim_max = P.dim[f];
      off += P.dim[f];
    }
    if (dim_max > 512) vec_ok = false;
    int rc = ids_i64 ? launch_fwd<int64_t>(P, dim_max, vec_ok, d_ids, B, L, ids_sb, ids_sf,
                                           ids_sl, pool, d_out, out_sb, d_err, st)
                     : launch_fwd<int32_t>(P, dim_max, vec_ok, d_ids, B, L, ids_sb, ids_sf,
                                           ids_sl, pool, d_out, out_sb, d_err, st);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int rtf_version(int* sm_arch) {
  if (sm_arch) *sm_arch = 100;
  return 1;
}
