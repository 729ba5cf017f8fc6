// K7 (shared core of K6) — fused short-sequence attention: QK^T -> mask -> softmax -> .V
//
// Replaces the materialised (B,H,L,L) logits / paddings / probabilities of
//   match scaled_dot_product_attention   src/match/layers/modules.py:76-96   (SASRec)
//   ctr   _scaled_dot_product_attention  src/ctr/layers/modules.py:222-240  (AutoInt / DIN)
// Layout: q/k/v are the Dense outputs (B, L, H*hs) — head h is columns [h*hs,(h+1)*hs) — so
// split_heads / the merge transpose (modules.py:63-74,130) never touch memory; the output is
// already the merged (B, Lq, H*hs).
//
// Mask semantics follow the source: masked logits are REPLACED by -2^32+1 before the softmax,
// so a fully masked row is exactly uniform.  row_mask (B,Lq) is the match-side quirk — the
// (B,L,1) mask broadcasts over the key axis and blanks whole QUERY rows (modules.py:90-91);
// key_mask (B,Lk) / causal are the conventional variants.
//
// One CTA per (sample, head): the K and V tiles live in shared memory (row stride hs+4 floats
// so 128-bit loads of consecutive rows hit distinct banks); a warp owns R=4 query rows and
// KPL keys per lane, i.e. an R x KPL register tile per lane for QK^T, a warp-shuffle softmax,
// then P (staged transposed in shared memory) times V with lanes over head columns.
// fp32 FFMA throughout (1e-5 parity rules out single-pass TF32).  FMA-bound, not HBM-bound.
#include "rtf_common.cuh"

namespace rtf {

constexpr int ATT_R = 4;          // query rows (or key rows in dkv) per warp step
// warps per CTA, chosen per kernel from its register need (ptxas: fwd 96, dq 159, dkv 155
// registers at 8 keys per lane): one CTA owns a (sample, head) and its K/V (or Q/dO) tiles fill
// most of the shared memory, so resident warps per SM = warps per CTA — the kernels are bound by
// shared-memory / FMA latency, and more warps is what hides it.
constexpr int ATT_FWD_THREADS = 512;
constexpr int ATT_DQ_THREADS = 384;
constexpr int ATT_DKV_THREADS = 384;

struct AttnParams {
  const float* q; const float* k; const float* v;
  long long q_sb, q_sl, k_sb, k_sl, v_sb, v_sl;  // element strides: sample, sequence position
  const float* row_mask; long long rm_sb;         // (B, Lq) or null; 0 => row fully padded
  const float* key_mask; long long km_sb;         // (B, Lk) or null; 0 => key padded
  int causal;
  int B, H, Lq, Lk, hs;
  float scale;
  float* out; long long o_sb, o_sl;               // (B, Lq, H*hs)
  float* stat_m; float* stat_il;                  // (B, H, Lq): row max and 1/row-sum
  // backward
  const float* dout; long long do_sb, do_sl;
  float* delta;                                   // (B, H, Lq): rowsum(dO * O)
  float* dq; long long dq_sb, dq_sl;
  float* dk; long long dk_sb, dk_sl;
  float* dv; long long dv_sb, dv_sl;
};

__device__ __forceinline__ void load_tile(float* dst, int RS, const float* src, long long sl, int L,
                                          int hs) {
  const int nv = hs >> 2;
  for (int e = threadIdx.x; e < L * nv; e += blockDim.x) {
    const int r = e / nv, c = e - r * nv;
    const float4 t = *reinterpret_cast<const float4*>(src + (long long)r * sl + 4 * c);
    *reinterpret_cast<float4*>(dst + r * RS + 4 * c) = t;
  }
}

// logit after masking, given the raw scaled score
__device__ __forceinline__ float mask_logit(float s, bool row_ok, bool key_ok, bool causal_ok) {
  return (row_ok && key_ok && causal_ok) ? s : kPadLogit;
}

// ----------------------------------------------------------------------------- forward
template <int KPL>
__global__ void __launch_bounds__(ATT_FWD_THREADS, 1) attn_fwd_kernel(const __grid_constant__ AttnParams P) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x / P.H, h = blockIdx.x - b * P.H;
  const int hs = P.hs, Lq = P.Lq, Lk = P.Lk, RS = hs + 4;
  float* Ks = smem;                       // [Lk][RS]
  float* Vs = Ks + Lk * RS;               // [Lk][RS]
  float* Qw = Vs + Lk * RS + warp * (ATT_R * hs + Lk * ATT_R);  // per warp: Q rows [R][hs]
  float* Pw = Qw + ATT_R * hs;                                   // per warp: P^T [Lk][R]
  load_tile(Ks, RS, P.k + (long long)b * P.k_sb + h * hs, P.k_sl, Lk, hs);
  load_tile(Vs, RS, P.v + (long long)b * P.v_sb + h * hs, P.v_sl, Lk, hs);
  __syncthreads();
  const float* qb = P.q + (long long)b * P.q_sb + h * hs;
  const int CT = (hs + 31) >> 5;
  for (int i0 = warp * ATT_R; i0 < Lq; i0 += (blockDim.x >> 5) * ATT_R) {
    for (int e = lane; e < ATT_R * hs; e += 32) {
      const int r = e / hs, c = e - r * hs;
      Qw[e] = (i0 + r < Lq) ? qb[(long long)(i0 + r) * P.q_sl + c] : 0.f;
    }
    __syncwarp();
    float acc[ATT_R][KPL];
#pragma unroll
    for (int r = 0; r < ATT_R; ++r)
#pragma unroll
      for (int t = 0; t < KPL; ++t) acc[r][t] = 0.f;
    for (int d = 0; d < hs; d += 4) {
      float4 qv[ATT_R];
#pragma unroll
      for (int r = 0; r < ATT_R; ++r) qv[r] = *reinterpret_cast<const float4*>(Qw + r * hs + d);
#pragma unroll
      for (int t = 0; t < KPL; ++t) {
        const int j = lane + 32 * t;
        if (j < Lk) {
          const float4 kv = *reinterpret_cast<const float4*>(Ks + j * RS + d);
#pragma unroll
          for (int r = 0; r < ATT_R; ++r) {
            acc[r][t] = fmaf(qv[r].x, kv.x, acc[r][t]);
            acc[r][t] = fmaf(qv[r].y, kv.y, acc[r][t]);
            acc[r][t] = fmaf(qv[r].z, kv.z, acc[r][t]);
            acc[r][t] = fmaf(qv[r].w, kv.w, acc[r][t]);
          }
        }
      }
    }
    // masks + softmax per row
#pragma unroll
    for (int r = 0; r < ATT_R; ++r) {
      const int i = i0 + r;
      const bool row_ok = !P.row_mask || (i < Lq && P.row_mask[(long long)b * P.rm_sb + i] != 0.f);
      float m = -INFINITY;
#pragma unroll
      for (int t = 0; t < KPL; ++t) {
        const int j = lane + 32 * t;
        if (j < Lk) {
          const bool key_ok = !P.key_mask || P.key_mask[(long long)b * P.km_sb + j] != 0.f;
          acc[r][t] = mask_logit(acc[r][t] * P.scale, row_ok, key_ok, !P.causal || j <= i);
          m = fmaxf(m, acc[r][t]);
        }
      }
      m = warp_max(m);
      float l = 0.f;
#pragma unroll
      for (int t = 0; t < KPL; ++t) {
        const int j = lane + 32 * t;
        if (j < Lk) {
          acc[r][t] = expf(acc[r][t] - m);
          l += acc[r][t];
        }
      }
      l = warp_sum(l);
      const float il = 1.f / l;
#pragma unroll
      for (int t = 0; t < KPL; ++t) {
        const int j = lane + 32 * t;
        if (j < Lk) Pw[j * ATT_R + r] = acc[r][t] * il;
      }
      if (lane == 0 && i < Lq && P.stat_m) {
        const long long si = ((long long)b * P.H + h) * Lq + i;
        P.stat_m[si] = m;
        P.stat_il[si] = il;
      }
    }
    __syncwarp();
    // O = P V, lanes over head columns
    float o[ATT_R][4];
#pragma unroll
    for (int r = 0; r < ATT_R; ++r)
#pragma unroll
      for (int t = 0; t < 4; ++t) o[r][t] = 0.f;
    for (int j = 0; j < Lk; ++j) {
      const float4 p = *reinterpret_cast<const float4*>(Pw + j * ATT_R);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int c = lane + 32 * t;
        if (t < CT && c < hs) {
          const float vv = Vs[j * RS + c];
          o[0][t] = fmaf(p.x, vv, o[0][t]);
          o[1][t] = fmaf(p.y, vv, o[1][t]);
          o[2][t] = fmaf(p.z, vv, o[2][t]);
          o[3][t] = fmaf(p.w, vv, o[3][t]);
        }
      }
    }
    float* ob = P.out + (long long)b * P.o_sb + h * hs;
#pragma unroll
    for (int r = 0; r < ATT_R; ++r)
      if (i0 + r < Lq) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int c = lane + 32 * t;
          if (t < CT && c < hs) ob[(long long)(i0 + r) * P.o_sl + c] = o[r][t];
        }
      }
    __syncwarp();
  }
}

// Forward for hs % 8 == 0, hs >= 32 (the SASRec shape: L = 200, hs = 64): bigger register tiles.
// On the CUDA cores a shared-memory load costs one wavefront per 128 bytes delivered to the warp —
// broadcast or not — and the SM moves one wavefront per clock against four FFMA warp
// instructions, so an r x c register tile (r + c words per r*c FMAs) has to be large.  The kernel
// above runs QK^T on 4 x 7 tiles (0.39 wavefronts per FMA) and P.V on 4 x 2 (0.75): it is bound by
// operand delivery at ~0.2 of the FMA peak.  Here a warp owns EIGHT query rows: QK^T on 8 x KPL
// tiles (0.27), and P.V with lane = (column group cg, key group jg): a lane accumulates 8 rows x 8
// columns over the keys j = jg, jg + JG, ... (16 words per 64 FMAs: 0.25) and the JG partial tiles
// are added by xor-shuffles (fixed order).  The lane's 8 columns are two float4 groups hs/2 apart,
// so a quarter-warp reads 128 contiguous bytes of a V row (conflict-free).
constexpr int ATT_R8 = 8;
template <int KPL, int CG>
__global__ void __launch_bounds__(448, 1) attn_fwd8_kernel(const __grid_constant__ AttnParams P) {
  extern __shared__ __align__(16) float smem[];
  constexpr int R = ATT_R8, JG = 32 / CG;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x / P.H, h = blockIdx.x - b * P.H;
  const int hs = P.hs, Lq = P.Lq, Lk = P.Lk, RS = hs + 4;
  float* Ks = smem;                       // [Lk][RS]
  float* Vs = Ks + Lk * RS;               // [Lk][RS]
  float* Qw = Vs + Lk * RS + warp * (R * hs + Lk * R);  // per warp: Q rows [R][hs]
  float* Pw = Qw + R * hs;                               // per warp: P^T [Lk][R]
  load_tile(Ks, RS, P.k + (long long)b * P.k_sb + h * hs, P.k_sl, Lk, hs);
  load_tile(Vs, RS, P.v + (long long)b * P.v_sb + h * hs, P.v_sl, Lk, hs);
  __syncthreads();
  const float* qb = P.q + (long long)b * P.q_sb + h * hs;
  const int cg = lane & (CG - 1), jg = lane / CG;
  const int c0 = 4 * cg, c1 = (hs >> 1) + 4 * cg;   // the lane's two column groups
  const bool col_ok = c0 < (hs >> 1);
  for (int i0 = warp * R; i0 < Lq; i0 += (blockDim.x >> 5) * R) {
    for (int e = lane; e < R * hs; e += 32) {
      const int r = e / hs, c = e - r * hs;
      Qw[e] = (i0 + r < Lq) ? qb[(long long)(i0 + r) * P.q_sl + c] : 0.f;
    }
    __syncwarp();
    float acc[R][KPL];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int t = 0; t < KPL; ++t) acc[r][t] = 0.f;
    for (int d = 0; d < hs; d += 4) {
      float4 qv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) qv[r] = *reinterpret_cast<const float4*>(Qw + r * hs + d);
#pragma unroll
      for (int t = 0; t < KPL; ++t) {
        const int j = lane + 32 * t;
        if (j < Lk) {
          const float4 kv = *reinterpret_cast<const float4*>(Ks + j * RS + d);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            acc[r][t] = fmaf(qv[r].x, kv.x, acc[r][t]);
            acc[r][t] = fmaf(qv[r].y, kv.y, acc[r][t]);
            acc[r][t] = fmaf(qv[r].z, kv.z, acc[r][t]);
            acc[r][t] = fmaf(qv[r].w, kv.w, acc[r][t]);
          }
        }
      }
    }
    // masks + softmax per row (same arithmetic and order as attn_fwd_kernel)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = i0 + r;
      const bool row_ok = !P.row_mask || (i < Lq && P.row_mask[(long long)b * P.rm_sb + i] != 0.f);
      float m = -INFINITY;
#pragma unroll
      for (int t = 0; t < KPL; ++t) {
        const int j = lane + 32 * t;
        if (j < Lk) {
          const bool key_ok = !P.key_mask || P.key_mask[(long long)b * P.km_sb + j] != 0.f;
          acc[r][t] = mask_logit(acc[r][t] * P.scale, row_ok, key_ok, !P.causal || j <= i);
          m = fmaxf(m, acc[r][t]);
        }
      }
      m = warp_max(m);
      float l = 0.f;
#pragma unroll
      for (int t = 0; t < KPL; ++t) {
        const int j = lane + 32 * t;
        if (j < Lk) {
          acc[r][t] = expf(acc[r][t] - m);
          l += acc[r][t];
        }
      }
      l = warp_sum(l);
      const float il = 1.f / l;
#pragma unroll
      for (int t = 0; t < KPL; ++t) {
        const int j = lane + 32 * t;
        if (j < Lk) Pw[j * R + r] = acc[r][t] * il;
      }
      if (lane == 0 && i < Lq && P.stat_m) {
        const long long si = ((long long)b * P.H + h) * Lq + i;
        P.stat_m[si] = m;
        P.stat_il[si] = il;
      }
    }
    __syncwarp();
    // O = P V: 8 rows x 8 columns per lane over the keys of its key group
    float o[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) o[r][c] = 0.f;
    if (col_ok) {
#pragma unroll 2
      for (int j = jg; j < Lk; j += JG) {
        const float4 p0 = *reinterpret_cast<const float4*>(Pw + j * R);
        const float4 p1 = *reinterpret_cast<const float4*>(Pw + j * R + 4);
        const float4 v0 = *reinterpret_cast<const float4*>(Vs + j * RS + c0);
        const float4 v1 = *reinterpret_cast<const float4*>(Vs + j * RS + c1);
        const float pr[R] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        const float vc[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) o[r][c] = fmaf(pr[r], vc[c], o[r][c]);
      }
    }
#pragma unroll
    for (int off = CG; off < 32; off <<= 1)
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) o[r][c] += __shfl_xor_sync(0xffffffffu, o[r][c], off);
    if (jg == 0 && col_ok) {
      float* ob = P.out + (long long)b * P.o_sb + h * hs;
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (i0 + r < Lq) {
          float* orow = ob + (long long)(i0 + r) * P.o_sl;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            orow[c0 + c] = o[r][c];
            orow[c1 + c] = o[r][4 + c];
          }
        }
    }
    __syncwarp();
  }
}

// ---- "P.V"-shaped products on 4 x 8 lane tiles (backward kernels, hs % 8 == 0, hs >= 32) ------
// o[r][c] += sum_j A[j][r] * T[j][col(c)] over the lane's key group j = jg, jg + JG, ...; A is a
// per-warp [L][4] stage (P^T or dS^T), T a [L][RS] tile in shared memory; the lane's 8 columns are
// the float4 groups at c0 and c1 = hs/2 + c0 (a quarter-warp reads 128 contiguous bytes).  12
// words per 32 FMAs instead of the column-per-lane form's 6 per 8 (see attn_fwd8_kernel).
template <int CG>
__device__ __forceinline__ void pv_tile4(const float* __restrict__ A, const float* __restrict__ T,
                                         int L, int RS, int c0, int c1, int jg, float (&o)[4][8]) {
  constexpr int JG = 32 / CG;
#pragma unroll 2
  for (int j = jg; j < L; j += JG) {
    const float4 a = *reinterpret_cast<const float4*>(A + j * 4);
    const float4 t0 = *reinterpret_cast<const float4*>(T + j * RS + c0);
    const float4 t1 = *reinterpret_cast<const float4*>(T + j * RS + c1);
    const float ar[4] = {a.x, a.y, a.z, a.w};
    const float tc[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) o[r][c] = fmaf(ar[r], tc[c], o[r][c]);
  }
}
// add the key groups' partial tiles (xor-shuffles: fixed order) and let key group 0 store rows
// row0 .. row0+3 (< nrows) of the (.., hs) output
template <int CG>
__device__ __forceinline__ void pv_reduce_store4(float (&o)[4][8], float* __restrict__ dst,
                                                 long long stride, int row0, int nrows, int c0,
                                                 int c1, bool writer) {
#pragma unroll
  for (int off = CG; off < 32; off <<= 1)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) o[r][c] += __shfl_xor_sync(0xffffffffu, o[r][c], off);
  if (writer) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (row0 + r < nrows) {
        float* d = dst + (long long)(row0 + r) * stride;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          d[c0 + c] = o[r][c];
          d[c1 + c] = o[r][4 + c];
        }
      }
  }
}

// ----------------------------------------------------------------------------- backward: dQ
// rows owned by warps (as forward): recompute P, dP = dO V^T, dS = P (dP - delta), dQ = scale dS K
template <int KPL, int CG>   // CG > 0: hs / 8 column groups (4 x 8 lane tiles for dQ = dS K)
__global__ void __launch_bounds__(ATT_DQ_THREADS, 1) attn_bwd_dq_kernel(const __grid_constant__ AttnParams P) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x / P.H, h = blockIdx.x - b * P.H;
  const int hs = P.hs, Lq = P.Lq, Lk = P.Lk, RS = hs + 4;
  float* Ks = smem;
  float* Vs = Ks + Lk * RS;
  float* Qw = Vs + Lk * RS + warp * (2 * ATT_R * hs + Lk * ATT_R);  // Q rows, then dO rows
  float* Dw = Qw + ATT_R * hs;
  float* Sw = Dw + ATT_R * hs;  // dS^T [Lk][R]
  load_tile(Ks, RS, P.k + (long long)b * P.k_sb + h * hs, P.k_sl, Lk, hs);
  load_tile(Vs, RS, P.v + (long long)b * P.v_sb + h * hs, P.v_sl, Lk, hs);
  __syncthreads();
  const float* qb = P.q + (long long)b * P.q_sb + h * hs;
  const float* dob = P.dout + (long long)b * P.do_sb + h * hs;
  const float* ob = P.out + (long long)b * P.o_sb + h * hs;
  const int CT = (hs + 31) >> 5;
  for (int i0 = warp * ATT_R; i0 < Lq; i0 += (blockDim.x >> 5) * ATT_R) {
    float dl[ATT_R];
#pragma unroll
    for (int r = 0; r < ATT_R; ++r) dl[r] = 0.f;
    for (int e = lane; e < ATT_R * hs; e += 32) {
      const int r = e / hs, c = e - r * hs;
      const bool ok = i0 + r < Lq;
      Qw[e] = ok ? qb[(long long)(i0 + r) * P.q_sl + c] : 0.f;
      Dw[e] = ok ? dob[(long long)(i0 + r) * P.do_sl + c] : 0.f;
    }
    __syncwarp();
    // delta_i = sum_c dO[i][c] * O[i][c]
#pragma unroll
    for (int r = 0; r < ATT_R; ++r) {
      float t = 0.f;
      if (i0 + r < Lq)
        for (int c = lane; c < hs; c += 32) t = fmaf(Dw[r * hs + c], ob[(long long)(i0 + r) * P.o_sl + c], t);
      dl[r] = warp_sum(t);
      if (lane == 0 && i0 + r < Lq) P.delta[((long long)b * P.H + h) * Lq + i0 + r] = dl[r];
    }
    float s[ATT_R][KPL], dp[ATT_R][KPL];
#pragma unroll
    for (int r = 0; r < ATT_R; ++r)
#pragma unroll
      for (int t = 0; t < KPL; ++t) s[r][t] = dp[r][t] = 0.f;
    for (int d = 0; d < hs; d += 4) {
      float4 qv[ATT_R], gv[ATT_R];
#pragma unroll
      for (int r = 0; r < ATT_R; ++r) {
        qv[r] = *reinterpret_cast<const float4*>(Qw + r * hs + d);
        gv[r] = *reinterpret_cast<const float4*>(Dw + r * hs + d);
      }
#pragma unroll
      for (int t = 0; t < KPL; ++t) {
        const int j = lane + 32 * t;
        if (j < Lk) {
          const float4 kv = *reinterpret_cast<const float4*>(Ks + j * RS + d);
          const float4 vv = *reinterpret_cast<const float4*>(Vs + j * RS + d);
#pragma unroll
          for (int r = 0; r < ATT_R; ++r) {
            s[r][t] = fmaf(qv[r].x, kv.x, s[r][t]);
            s[r][t] = fmaf(qv[r].y, kv.y, s[r][t]);
            s[r][t] = fmaf(qv[r].z, kv.z, s[r][t]);
            s[r][t] = fmaf(qv[r].w, kv.w, s[r][t]);
            dp[r][t] = fmaf(gv[r].x, vv.x, dp[r][t]);
            dp[r][t] = fmaf(gv[r].y, vv.y, dp[r][t]);
            dp[r][t] = fmaf(gv[r].z, vv.z, dp[r][t]);
            dp[r][t] = fmaf(gv[r].w, vv.w, dp[r][t]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ATT_R; ++r) {
      const int i = i0 + r;
      const bool in = i < Lq;
      const bool row_ok = !P.row_mask || (in && P.row_mask[(long long)b * P.rm_sb + i] != 0.f);
      const long long si = ((long long)b * P.H + h) * Lq + (in ? i : 0);
      const float m = P.stat_m[si], il = P.stat_il[si];
#pragma unroll
      for (int t = 0; t < KPL; ++t) {
        const int j = lane + 32 * t;
        if (j < Lk) {
          const bool key_ok = !P.key_mask || P.key_mask[(long long)b * P.km_sb + j] != 0.f;
          const bool live = row_ok && key_ok && (!P.causal || j <= i);
          const float lg = live ? s[r][t] * P.scale : kPadLogit;
          const float p = expf(lg - m) * il;
          // a padded logit is a constant: no gradient flows through it
          Sw[j * ATT_R + r] = (live && in) ? p * (dp[r][t] - dl[r]) * P.scale : 0.f;
        }
      }
    }
    __syncwarp();
    float* gq = P.dq + (long long)b * P.dq_sb + h * hs;
    if constexpr (CG > 0) {
      const int cg = lane & (CG - 1), jg = lane / CG;
      const int c0 = 4 * cg, c1 = (hs >> 1) + 4 * cg;
      const bool col_ok = c0 < (hs >> 1);
      float o8[4][8];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) o8[r][c] = 0.f;
      if (col_ok) pv_tile4<CG>(Sw, Ks, Lk, RS, c0, c1, jg, o8);
      pv_reduce_store4<CG>(o8, gq, P.dq_sl, i0, Lq, c0, c1, jg == 0 && col_ok);
    } else {
      float o[ATT_R][4];
#pragma unroll
      for (int r = 0; r < ATT_R; ++r)
#pragma unroll
        for (int t = 0; t < 4; ++t) o[r][t] = 0.f;
      for (int j = 0; j < Lk; ++j) {
        const float4 p = *reinterpret_cast<const float4*>(Sw + j * ATT_R);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int c = lane + 32 * t;
          if (t < CT && c < hs) {
            const float kk = Ks[j * RS + c];
            o[0][t] = fmaf(p.x, kk, o[0][t]);
            o[1][t] = fmaf(p.y, kk, o[1][t]);
            o[2][t] = fmaf(p.z, kk, o[2][t]);
            o[3][t] = fmaf(p.w, kk, o[3][t]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < ATT_R; ++r)
        if (i0 + r < Lq) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int c = lane + 32 * t;
            if (t < CT && c < hs) gq[(long long)(i0 + r) * P.dq_sl + c] = o[r][t];
          }
        }
    }
    __syncwarp();
  }
}

// ----------------------------------------------------------------------------- backward: dK, dV
// keys owned by warps; the Q and dO tiles live in shared memory; lanes own queries
template <int QPL, int CG>   // CG > 0: 4 x 8 lane tiles for dV = P^T dO and dK = dS^T Q
__global__ void __launch_bounds__(ATT_DKV_THREADS, 1) attn_bwd_dkv_kernel(const __grid_constant__ AttnParams P) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x / P.H, h = blockIdx.x - b * P.H;
  const int hs = P.hs, Lq = P.Lq, Lk = P.Lk, RS = hs + 4;
  float* Qs = smem;                 // [Lq][RS]
  float* Ds = Qs + Lq * RS;         // [Lq][RS] dO
  float* stat = Ds + Lq * RS;       // m[Lq], il[Lq], delta[Lq], rowok[Lq]
  float* Kw = stat + 4 * Lq + warp * (2 * ATT_R * hs + 2 * Lq * ATT_R);  // K rows, V rows
  float* Vw = Kw + ATT_R * hs;
  float* Pw = Vw + ATT_R * hs;      // P^T  [Lq][R]
  float* Sw = Pw + Lq * ATT_R;      // dS^T [Lq][R]
  load_tile(Qs, RS, P.q + (long long)b * P.q_sb + h * hs, P.q_sl, Lq, hs);
  load_tile(Ds, RS, P.dout + (long long)b * P.do_sb + h * hs, P.do_sl, Lq, hs);
  for (int i = threadIdx.x; i < Lq; i += blockDim.x) {
    const long long si = ((long long)b * P.H + h) * Lq + i;
    stat[i] = P.stat_m[si];
    stat[Lq + i] = P.stat_il[si];
    stat[2 * Lq + i] = P.delta[si];
    stat[3 * Lq + i] = (!P.row_mask || P.row_mask[(long long)b * P.rm_sb + i] != 0.f) ? 1.f : 0.f;
  }
  __syncthreads();
  const float* kb = P.k + (long long)b * P.k_sb + h * hs;
  const float* vb = P.v + (long long)b * P.v_sb + h * hs;
  const int CT = (hs + 31) >> 5;
  for (int j0 = warp * ATT_R; j0 < Lk; j0 += (blockDim.x >> 5) * ATT_R) {
    for (int e = lane; e < ATT_R * hs; e += 32) {
      const int r = e / hs, c = e - r * hs;
      const bool ok = j0 + r < Lk;
      Kw[e] = ok ? kb[(long long)(j0 + r) * P.k_sl + c] : 0.f;
      Vw[e] = ok ? vb[(long long)(j0 + r) * P.v_sl + c] : 0.f;
    }
    __syncwarp();
    float s[ATT_R][QPL], dp[ATT_R][QPL];
#pragma unroll
    for (int r = 0; r < ATT_R; ++r)
#pragma unroll
      for (int t = 0; t < QPL; ++t) s[r][t] = dp[r][t] = 0.f;
    for (int d = 0; d < hs; d += 4) {
      float4 kv[ATT_R], vv[ATT_R];
#pragma unroll
      for (int r = 0; r < ATT_R; ++r) {
        kv[r] = *reinterpret_cast<const float4*>(Kw + r * hs + d);
        vv[r] = *reinterpret_cast<const float4*>(Vw + r * hs + d);
      }
#pragma unroll
      for (int t = 0; t < QPL; ++t) {
        const int i = lane + 32 * t;
        if (i < Lq) {
          const float4 qv = *reinterpret_cast<const float4*>(Qs + i * RS + d);
          const float4 gv = *reinterpret_cast<const float4*>(Ds + i * RS + d);
#pragma unroll
          for (int r = 0; r < ATT_R; ++r) {
            s[r][t] = fmaf(qv.x, kv[r].x, s[r][t]);
            s[r][t] = fmaf(qv.y, kv[r].y, s[r][t]);
            s[r][t] = fmaf(qv.z, kv[r].z, s[r][t]);
            s[r][t] = fmaf(qv.w, kv[r].w, s[r][t]);
            dp[r][t] = fmaf(gv.x, vv[r].x, dp[r][t]);
            dp[r][t] = fmaf(gv.y, vv[r].y, dp[r][t]);
            dp[r][t] = fmaf(gv.z, vv[r].z, dp[r][t]);
            dp[r][t] = fmaf(gv.w, vv[r].w, dp[r][t]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < QPL; ++t) {
      const int i = lane + 32 * t;
      if (i < Lq) {
        const float m = stat[i], il = stat[Lq + i], dl = stat[2 * Lq + i];
        const bool row_ok = stat[3 * Lq + i] != 0.f;
        float4 pp, ss;
        float* ppv = &pp.x;
        float* ssv = &ss.x;
#pragma unroll
        for (int r = 0; r < ATT_R; ++r) {
          const int j = j0 + r;
          const bool in = j < Lk;
          const bool key_ok = !P.key_mask || (in && P.key_mask[(long long)b * P.km_sb + j] != 0.f);
          const bool live = row_ok && key_ok && (!P.causal || j <= i);
          const float lg = live ? s[r][t] * P.scale : kPadLogit;
          const float p = in ? expf(lg - m) * il : 0.f;
          ppv[r] = p;
          ssv[r] = (live && in) ? p * (dp[r][t] - dl) * P.scale : 0.f;
        }
        *reinterpret_cast<float4*>(Pw + i * ATT_R) = pp;
        *reinterpret_cast<float4*>(Sw + i * ATT_R) = ss;
      }
    }
    __syncwarp();
    float* gkb = P.dk + (long long)b * P.dk_sb + h * hs;
    float* gvb = P.dv + (long long)b * P.dv_sb + h * hs;
    if constexpr (CG > 0) {
      const int cg = lane & (CG - 1), ig = lane / CG;
      const int c0 = 4 * cg, c1 = (hs >> 1) + 4 * cg;
      const bool col_ok = c0 < (hs >> 1);
      float o8[4][8];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) o8[r][c] = 0.f;
      if (col_ok) pv_tile4<CG>(Pw, Ds, Lq, RS, c0, c1, ig, o8);       // dV = P^T dO
      pv_reduce_store4<CG>(o8, gvb, P.dv_sl, j0, Lk, c0, c1, ig == 0 && col_ok);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) o8[r][c] = 0.f;
      if (col_ok) pv_tile4<CG>(Sw, Qs, Lq, RS, c0, c1, ig, o8);       // dK = dS^T Q
      pv_reduce_store4<CG>(o8, gkb, P.dk_sl, j0, Lk, c0, c1, ig == 0 && col_ok);
    } else {
      float gk[ATT_R][4], gv[ATT_R][4];
#pragma unroll
      for (int r = 0; r < ATT_R; ++r)
#pragma unroll
        for (int t = 0; t < 4; ++t) gk[r][t] = gv[r][t] = 0.f;
      for (int i = 0; i < Lq; ++i) {
        const float4 p = *reinterpret_cast<const float4*>(Pw + i * ATT_R);
        const float4 ds = *reinterpret_cast<const float4*>(Sw + i * ATT_R);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int c = lane + 32 * t;
          if (t < CT && c < hs) {
            const float qq = Qs[i * RS + c], gg = Ds[i * RS + c];
            gv[0][t] = fmaf(p.x, gg, gv[0][t]);
            gv[1][t] = fmaf(p.y, gg, gv[1][t]);
            gv[2][t] = fmaf(p.z, gg, gv[2][t]);
            gv[3][t] = fmaf(p.w, gg, gv[3][t]);
            gk[0][t] = fmaf(ds.x, qq, gk[0][t]);
            gk[1][t] = fmaf(ds.y, qq, gk[1][t]);
            gk[2][t] = fmaf(ds.z, qq, gk[2][t]);
            gk[3][t] = fmaf(ds.w, qq, gk[3][t]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < ATT_R; ++r)
        if (j0 + r < Lk) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int c = lane + 32 * t;
            if (t < CT && c < hs) {
              gkb[(long long)(j0 + r) * P.dk_sl + c] = gk[r][t];
              gvb[(long long)(j0 + r) * P.dv_sl + c] = gv[r][t];
            }
          }
        }
    }
    __syncwarp();
  }
}

template <typename Kern>
static int attn_launch(Kern kern, const AttnParams& P, size_t smem_bytes, int threads,
                       cudaStream_t st) {
  if (smem_bytes > 227 * 1024) return RTF_E_RANGE;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int)e;
  kern<<<(unsigned)(P.B * P.H), threads, smem_bytes, st>>>(P);
  RTF_CHECK_LAUNCH();
  return 0;
}

static int attn_check(const AttnParams& P) {
  if (P.B < 0 || P.H <= 0 || P.Lq <= 0 || P.Lk <= 0 || P.hs <= 0) return RTF_E_ARG;
  if (P.hs % 4 || P.hs > 128 || P.Lq > 256 || P.Lk > 256) return RTF_E_RANGE;
  if ((long long)P.B * P.H > 0x7fffffffLL) return RTF_E_RANGE;
  const long long str[] = {P.q_sb, P.q_sl, P.k_sb, P.k_sl, P.v_sb, P.v_sl};
  for (long long s : str)
    if (s % 4) return RTF_E_ALIGN;
  if ((uintptr_t)P.q % 16 || (uintptr_t)P.k % 16 || (uintptr_t)P.v % 16) return RTF_E_ALIGN;
  return 0;
}

}  // namespace rtf

using namespace rtf;

extern "C" int rtf_attn_fwd(const float* d_q, int64_t q_sb, int64_t q_sl, const float* d_k,
                            int64_t k_sb, int64_t k_sl, const float* d_v, int64_t v_sb,
                            int64_t v_sl, const float* d_row_mask, int64_t rm_sb,
                            const float* d_key_mask, int64_t km_sb, int causal, int B, int H, int Lq,
                            int Lk, int hs, float scale, float* d_out, int64_t o_sb, int64_t o_sl,
                            float* d_stat_m, float* d_stat_il, void* stream) {
  AttnParams P = {};
  P.q = d_q; P.k = d_k; P.v = d_v; P.q_sb = q_sb; P.q_sl = q_sl; P.k_sb = k_sb; P.k_sl = k_sl;
  P.v_sb = v_sb; P.v_sl = v_sl; P.row_mask = d_row_mask; P.rm_sb = rm_sb; P.key_mask = d_key_mask;
  P.km_sb = km_sb; P.causal = causal; P.B = B; P.H = H; P.Lq = Lq; P.Lk = Lk; P.hs = hs;
  P.scale = scale; P.out = d_out; P.o_sb = o_sb; P.o_sl = o_sl; P.stat_m = d_stat_m;
  P.stat_il = d_stat_il;
  int rc = attn_check(P);
  if (rc) return rc;
  if (B == 0) return 0;
  if (!d_q || !d_k || !d_v || !d_out) return RTF_E_ARG;
  if ((d_stat_m == nullptr) != (d_stat_il == nullptr)) return RTF_E_ARG;
  const int RS = hs + 4;
  // as many warps as the per-warp staging (Q rows + P^T) leaves room for, up to the bound
  int warps = ATT_FWD_THREADS / 32;
  auto smem_for = [&](int w) { return ((size_t)2 * Lk * RS + (size_t)w * (ATT_R * hs + Lk * ATT_R)) * 4; };
  while (warps > 4 && smem_for(warps) > 227 * 1024) warps -= 4;
  const int rows_w = (Lq + ATT_R - 1) / ATT_R;       // no more warps than there are row groups
  while (warps > 4 && warps - 4 >= rows_w) warps -= 4;
  const size_t smem = smem_for(warps);
  const int thr = warps * 32;
  cudaStream_t st = (cudaStream_t)stream;
  const int kpl = (Lk + 31) / 32;
  if (hs % 8 == 0 && hs >= 32 && Lq >= 64) {
    // 8 query rows per warp (see attn_fwd8_kernel); warps = what the per-warp staging leaves room
    // for, at most 14 (448 threads x 144 registers), no more than there are row blocks
    auto smem8 = [&](int w) { return ((size_t)2 * Lk * RS + (size_t)w * (ATT_R8 * hs + Lk * ATT_R8)) * 4; };
    int w8 = 14;
    while (w8 > 2 && smem8(w8) > 227 * 1024) --w8;
    const int blocks8 = (Lq + ATT_R8 - 1) / ATT_R8;
    if (w8 > blocks8) w8 = blocks8;
    // even out the last round: 25 row blocks over 14 warps = 2 rounds -> 13 warps do the same
    const int rounds = (blocks8 + w8 - 1) / w8;
    w8 = (blocks8 + rounds - 1) / rounds;
    if (smem8(w8) <= 227 * 1024) {
      const size_t sm8 = smem8(w8);
      const int cgn = hs / 8;   // column groups: 4, 8 or 16 (hs = 32..128)
#define RTF_FWD8(K)                                                                              \
  (cgn <= 4 ? attn_launch(attn_fwd8_kernel<K, 4>, P, sm8, w8 * 32, st)                           \
            : cgn <= 8 ? attn_launch(attn_fwd8_kernel<K, 8>, P, sm8, w8 * 32, st)                \
                       : attn_launch(attn_fwd8_kernel<K, 16>, P, sm8, w8 * 32, st))
      if (kpl <= 2) return RTF_FWD8(2);
      if (kpl <= 4) return RTF_FWD8(4);
      return RTF_FWD8(8);
#undef RTF_FWD8
    }
  }
  if (kpl <= 1) return attn_launch(attn_fwd_kernel<1>, P, smem, thr, st);
  if (kpl <= 2) return attn_launch(attn_fwd_kernel<2>, P, smem, thr, st);
  if (kpl <= 4) return attn_launch(attn_fwd_kernel<4>, P, smem, thr, st);
  return attn_launch(attn_fwd_kernel<8>, P, smem, thr, st);
}

extern "C" int rtf_attn_bwd(const float* d_q, int64_t q_sb, int64_t q_sl, const float* d_k,
                            int64_t k_sb, int64_t k_sl, const float* d_v, int64_t v_sb,
                            int64_t v_sl, const float* d_row_mask, int64_t rm_sb,
                            const float* d_key_mask, int64_t km_sb, int causal, int B, int H, int Lq,
                            int Lk, int hs, float scale, const float* d_out, int64_t o_sb,
                            int64_t o_sl, const float* d_stat_m, const float* d_stat_il,
                            const float* d_dout, int64_t do_sb, int64_t do_sl, float* d_delta,
                            float* d_dq, int64_t dq_sb, int64_t dq_sl, float* d_dk, int64_t dk_sb,
                            int64_t dk_sl, float* d_dv, int64_t dv_sb, int64_t dv_sl, void* stream) {
  AttnParams P = {};
  P.q = d_q; P.k = d_k; P.v = d_v; P.q_sb = q_sb; P.q_sl = q_sl; P.k_sb = k_sb; P.k_sl = k_sl;
  P.v_sb = v_sb; P.v_sl = v_sl; P.row_mask = d_row_mask; P.rm_sb = rm_sb; P.key_mask = d_key_mask;
  P.km_sb = km_sb; P.causal = causal; P.B = B; P.H = H; P.Lq = Lq; P.Lk = Lk; P.hs = hs;
  P.scale = scale; P.out = const_cast<float*>(d_out); P.o_sb = o_sb; P.o_sl = o_sl;
  P.stat_m = const_cast<float*>(d_stat_m); P.stat_il = const_cast<float*>(d_stat_il);
  P.dout = d_dout; P.do_sb = do_sb; P.do_sl = do_sl; P.delta = d_delta; P.dq = d_dq;
  P.dq_sb = dq_sb; P.dq_sl = dq_sl; P.dk = d_dk; P.dk_sb = dk_sb; P.dk_sl = dk_sl; P.dv = d_dv;
  P.dv_sb = dv_sb; P.dv_sl = dv_sl;
  int rc = attn_check(P);
  if (rc) return rc;
  if (B == 0) return 0;
  if (!d_q || !d_k || !d_v || !d_out || !d_stat_m || !d_stat_il || !d_dout || !d_delta || !d_dq ||
      !d_dk || !d_dv)
    return RTF_E_ARG;
  if (do_sb % 4 || do_sl % 4 || (uintptr_t)d_dout % 16) return RTF_E_ALIGN;
  const int RS = hs + 4;
  cudaStream_t st = (cudaStream_t)stream;
  {
    int warps = ATT_DQ_THREADS / 32;
    auto smem_for = [&](int w) {
      return ((size_t)2 * Lk * RS + (size_t)w * (2 * ATT_R * hs + Lk * ATT_R)) * 4;
    };
    while (warps > 4 && smem_for(warps) > 227 * 1024) warps -= 2;
    const size_t smem = smem_for(warps);
    const int kpl = (Lk + 31) / 32;
    // 4 x 8 lane tiles for the P.V-shaped products when the head is wide enough (see pv_tile4)
    const int cgn = (hs % 8 == 0 && hs >= 32) ? (hs <= 32 ? 4 : hs <= 64 ? 8 : 16) : 0;
#define RTF_DQ(K)                                                                            \
  (cgn == 0 ? attn_launch(attn_bwd_dq_kernel<K, 0>, P, smem, warps * 32, st)                 \
   : cgn == 4 ? attn_launch(attn_bwd_dq_kernel<K, 4>, P, smem, warps * 32, st)               \
   : cgn == 8 ? attn_launch(attn_bwd_dq_kernel<K, 8>, P, smem, warps * 32, st)               \
              : attn_launch(attn_bwd_dq_kernel<K, 16>, P, smem, warps * 32, st))
    if (kpl <= 1) rc = RTF_DQ(1);
    else if (kpl <= 2) rc = RTF_DQ(2);
    else if (kpl <= 4) rc = RTF_DQ(4);
    else rc = RTF_DQ(8);
#undef RTF_DQ
    if (rc) return rc;
  }
  {
    int warps = ATT_DKV_THREADS / 32;
    auto smem_for = [&](int w) {
      return ((size_t)2 * Lq * RS + 4 * (size_t)Lq + (size_t)w * (2 * ATT_R * hs + 2 * Lq * ATT_R)) * 4;
    };
    while (warps > 4 && smem_for(warps) > 227 * 1024) warps -= 2;
    const size_t smem = smem_for(warps);
    const int qpl = (Lq + 31) / 32;
    const int cgn = (hs % 8 == 0 && hs >= 32) ? (hs <= 32 ? 4 : hs <= 64 ? 8 : 16) : 0;
#define RTF_DKV(K)                                                                           \
  (cgn == 0 ? attn_launch(attn_bwd_dkv_kernel<K, 0>, P, smem, warps * 32, st)                \
   : cgn == 4 ? attn_launch(attn_bwd_dkv_kernel<K, 4>, P, smem, warps * 32, st)              \
   : cgn == 8 ? attn_launch(attn_bwd_dkv_kernel<K, 8>, P, smem, warps * 32, st)              \
              : attn_launch(attn_bwd_dkv_kernel<K, 16>, P, smem, warps * 32, st))
    if (qpl <= 1) rc = RTF_DKV(1);
    else if (qpl <= 2) rc = RTF_DKV(2);
    else if (qpl <= 4) rc = RTF_DKV(4);
    else rc = RTF_DKV(8);
#undef RTF_DKV
  }
  return rc;
}
