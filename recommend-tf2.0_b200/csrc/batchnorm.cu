// BatchNormalization over the feature axis of a (B, C) activation — the layer the reference's DNN
// block puts in front of its Dense stack (src/ctr/layers/modules.py:129-135; Keras defaults:
// momentum 0.99, epsilon 1e-3, batch statistics in training, biased variance everywhere — App. A9).
//
// Training forward:  mean_c, var_c over the B rows;  y = (x - mean) * gamma / sqrt(var + eps) + beta;
//                    moving = moving * momentum + batch * (1 - momentum)
// Backward:          dbeta = sum dy;  dgamma = invstd * sum dy (x - mean);
//                    dx = (dy - dbeta / B - (x - mean) invstd^2 sum dy (x - mean) / B) gamma invstd
//
// Both directions are one column reduction over a tall matrix + one streaming pass: HBM-bound.
// The reduction is what a generic library kernel does badly on these shapes (65536 x 13: 198 us for
// 3.4 MB; 65536 x 480: 76 us for 126 MB, 1.7 TB/s): here a CTA owns a (row chunk x column tile),
// threads keep per-column running sums in registers over coalesced row reads, row lanes are
// combined through shared memory, and a second tiny launch adds the chunk partials in double, in
// chunk order (deterministic, no atomics).  The variance uses sums shifted by the first row
// (sum (x - x0), sum (x - x0)^2): one pass over x, no catastrophic cancellation for |mean| >> std.
#include "rtf_common.cuh"

namespace rtf {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_CHUNKS = 1024;

struct BnTile {
  int cx;      // columns (or float4 column groups) per CTA
  int ry;      // row lanes per CTA
  int chunks;  // row chunks (gridDim.y)
  long long rows_per_chunk;
};

// column tile = min(width, 256) threads wide, the other threads of the CTA are row lanes;
// chunks sized for ~4 CTAs per SM
static BnTile bn_tile(long long B, int width) {
  BnTile t;
  t.cx = width < BN_THREADS ? width : BN_THREADS;
  t.ry = BN_THREADS / t.cx;
  const int col_tiles = (width + t.cx - 1) / t.cx;
  long long chunks = (kNumSMs * 4 + col_tiles - 1) / col_tiles;
  const long long max_by_rows = (B + (long long)t.ry * 8 - 1) / ((long long)t.ry * 8);  // >= 8 rows per lane
  if (chunks > max_by_rows) chunks = max_by_rows;
  if (chunks > BN_MAX_CHUNKS) chunks = BN_MAX_CHUNKS;
  if (chunks < 1) chunks = 1;
  t.rows_per_chunk = (B + chunks - 1) / chunks;
  t.chunks = (int)((B + t.rows_per_chunk - 1) / t.rows_per_chunk);
  return t;
}

template <int V>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

// Stage 1 of both reductions.  MODE 0 (forward):  p0 = sum (x - x0), p1 = sum (x - x0)^2
//                              MODE 1 (backward): p0 = sum dy,       p1 = sum dy (x - mean)
// partial layout: [chunk][2][C]
template <int V, int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_reduce_stage1(const float* __restrict__ x, long long ldx, const float* __restrict__ dy,
                 long long lddy, const float* __restrict__ mean, long long B, int C, int cx, int ry,
                 long long rows_per_chunk, float* __restrict__ partial) {
  __shared__ float red[2][BN_THREADS * V];
  const int tx = threadIdx.x % cx, ty = threadIdx.x / cx;
  const int col = (blockIdx.x * cx + tx) * V;
  const bool live = col < C && ty < ry;
  float a0[V], a1[V], ref[V];
#pragma unroll
  for (int e = 0; e < V; ++e) a0[e] = a1[e] = ref[e] = 0.f;
  if (live) {
    Vec<V> r;
    r.load(MODE == 0 ? x + col : mean + col);   // shift = first row (fwd) / batch mean (bwd)
#pragma unroll
    for (int e = 0; e < V; ++e) ref[e] = r.v[e];
    const long long b0 = (long long)blockIdx.y * rows_per_chunk;
    const long long b1 = min(b0 + rows_per_chunk, B);
    long long b = b0 + ty;
    for (; b + 3LL * ry < b1; b += 4LL * ry) {      // four rows in flight per thread
      Vec<V> xv[4], gv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xv[k].load(x + (b + (long long)k * ry) * ldx + col);
        if (MODE == 1) gv[k].load(dy + (b + (long long)k * ry) * lddy + col);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float d = __fsub_rn(xv[k].v[e], ref[e]);
          if (MODE == 0) {
            a0[e] = __fadd_rn(a0[e], d);
            a1[e] = __fmaf_rn(d, d, a1[e]);
          } else {
            a0[e] = __fadd_rn(a0[e], gv[k].v[e]);
            a1[e] = __fmaf_rn(gv[k].v[e], d, a1[e]);
          }
        }
    }
    for (; b < b1; b += ry) {
      Vec<V> xv, gv;
      xv.load(x + b * ldx + col);
      if (MODE == 1) gv.load(dy + b * lddy + col);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float d = __fsub_rn(xv.v[e], ref[e]);
        if (MODE == 0) {
          a0[e] = __fadd_rn(a0[e], d);
          a1[e] = __fmaf_rn(d, d, a1[e]);
        } else {
          a0[e] = __fadd_rn(a0[e], gv.v[e]);
          a1[e] = __fmaf_rn(gv.v[e], d, a1[e]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < V; ++e) {
    red[0][threadIdx.x * V + e] = a0[e];
    red[1][threadIdx.x * V + e] = a1[e];
  }
  __syncthreads();
  if (live && ty == 0) {                            // row lanes added in lane order
    float* p = partial + (long long)blockIdx.y * 2 * C;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      float s0 = a0[e], s1 = a1[e];
      for (int r = 1; r < ry; ++r) {
        s0 = __fadd_rn(s0, red[0][(r * cx + tx) * V + e]);
        s1 = __fadd_rn(s1, red[1][(r * cx + tx) * V + e]);
      }
      p[col + e] = s0;
      p[C + col + e] = s1;
    }
  }
}

// Stage 2: 32 columns per CTA, 32 warps; warp w adds chunks w, w + 32, ... of its column in double
// (ascending, 16 loads in flight), the warps' sums are then added in warp order — a fixed tree:
// reproducible.  (8 warps with 8 loads in flight took 12-20 us on 432-591 chunks: load latency.)
constexpr int BN2_WARPS = 32;
__device__ __forceinline__ void bn_stage2_sums(const float* __restrict__ partial, int chunks, int C,
                                               int c, double (*red)[2][32], double& s0, double& s1) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double a0 = 0.0, a1 = 0.0;
  if (c < C) {
    int k = w;
    for (; k + 7 * BN2_WARPS < chunks; k += 8 * BN2_WARPS) {   // 16 independent loads in flight
      float t0[8], t1[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        t0[u] = partial[(long long)(k + BN2_WARPS * u) * 2 * C + c];
        t1[u] = partial[(long long)(k + BN2_WARPS * u) * 2 * C + C + c];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a0 += (double)t0[u];
        a1 += (double)t1[u];
      }
    }
    for (; k < chunks; k += BN2_WARPS) {
      a0 += (double)partial[(long long)k * 2 * C + c];
      a1 += (double)partial[(long long)k * 2 * C + C + c];
    }
  }
  red[w][0][lane] = a0;
  red[w][1][lane] = a1;
  __syncthreads();
  s0 = red[0][0][lane];
  s1 = red[0][1][lane];
#pragma unroll
  for (int g = 1; g < BN2_WARPS; ++g) {
    s0 += red[g][0][lane];
    s1 += red[g][1][lane];
  }
}

// forward stage 2: chunk partials -> mean, invstd (and the Keras moving statistics)
__global__ void __launch_bounds__(BN2_WARPS * 32)
bn_stats_stage2(const float* __restrict__ partial, int chunks, const float* __restrict__ x,
                long long B, int C, float eps, float momentum, float* __restrict__ mean,
                float* __restrict__ invstd, float* __restrict__ moving_mean,
                float* __restrict__ moving_var) {
  __shared__ double red[BN2_WARPS][2][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double s0, s1;
  bn_stage2_sums(partial, chunks, C, c, red, s0, s1);
  if (c >= C || threadIdx.x >= 32) return;
  const double n = (double)B;
  const double dm = s0 / n;
  double var = s1 / n - dm * dm;        // biased (population) variance, as Keras normalises with
  if (var < 0.0) var = 0.0;
  const float m = (float)((double)x[c] + dm);
  const float v = (float)var;
  mean[c] = m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (moving_mean) moving_mean[c] = __fmaf_rn(moving_mean[c], momentum, __fmul_rn(m, 1.0f - momentum));
  if (moving_var) moving_var[c] = __fmaf_rn(moving_var[c], momentum, __fmul_rn(v, 1.0f - momentum));
}

// backward stage 2: -> sum dy (= dbeta), sum dy (x - mean), dgamma
__global__ void __launch_bounds__(BN2_WARPS * 32)
bn_bwd_stage2(const float* __restrict__ partial, int chunks, int C, const float* __restrict__ invstd,
              float* __restrict__ sum_dy, float* __restrict__ sum_dy_xmu, float* __restrict__ dgamma,
              float* __restrict__ dbeta) {
  __shared__ double red[BN2_WARPS][2][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double s0, s1;
  bn_stage2_sums(partial, chunks, C, c, red, s0, s1);
  if (c >= C || threadIdx.x >= 32) return;
  sum_dy[c] = (float)s0;
  sum_dy_xmu[c] = (float)s1;
  if (dbeta) dbeta[c] = (float)s0;
  if (dgamma) dgamma[c] = (float)(s1 * (double)invstd[c]);
}

// y = (x - mean) * (gamma * invstd) + beta
template <int V>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, long long ldx, long long B, int C,
                const float* __restrict__ mean, const float* __restrict__ invstd,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ y, long long ldy) {
  const int cw = C / V;
  const long long total = B * cw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / cw;
    const int c = (int)(i - b * cw) * V;
    Vec<V> xv, mv, sv, gv, bv, o;
    xv.load(x + b * ldx + c);
    mv.load(mean + c);
    sv.load(invstd + c);
#pragma unroll
    for (int e = 0; e < V; ++e) { gv.v[e] = 1.f; bv.v[e] = 0.f; }
    if (gamma) gv.load(gamma + c);
    if (beta) bv.load(beta + c);
#pragma unroll
    for (int e = 0; e < V; ++e)
      o.v[e] = __fmaf_rn(__fsub_rn(xv.v[e], mv.v[e]), __fmul_rn(gv.v[e], sv.v[e]), bv.v[e]);
    o.store(y + b * ldy + c);
  }
}

// dx = (dy - sum_dy / B - (x - mean) * invstd^2 * sum_dy_xmu / B) * gamma * invstd
template <int V>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ x,
                    long long ldx, long long B, int C, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ sum_dy, const float* __restrict__ sum_dy_xmu,
                    float* __restrict__ dx, long long lddx) {
  const int cw = C / V;
  const long long total = B * cw;
  const float inv_n = 1.0f / (float)B;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / cw;
    const int c = (int)(i - b * cw) * V;
    Vec<V> dv, xv, mv, sv, gv, s0, s1, o;
    dv.load(dy + b * lddy + c);
    xv.load(x + b * ldx + c);
    mv.load(mean + c);
    sv.load(invstd + c);
    s0.load(sum_dy + c);
    s1.load(sum_dy_xmu + c);
#pragma unroll
    for (int e = 0; e < V; ++e) gv.v[e] = 1.f;
    if (gamma) gv.load(gamma + c);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float k = __fmul_rn(__fmul_rn(__fmul_rn(sv.v[e], sv.v[e]), s1.v[e]), inv_n);
      const float t = __fsub_rn(__fsub_rn(dv.v[e], __fmul_rn(s0.v[e], inv_n)),
                                __fmul_rn(__fsub_rn(xv.v[e], mv.v[e]), k));
      o.v[e] = __fmul_rn(t, __fmul_rn(gv.v[e], sv.v[e]));
    }
    o.store(dx + b * lddx + c);
  }
}

static bool bn_vec4(int C, std::initializer_list<long long> lds, std::initializer_list<const void*> ptrs) {
  if (C % 4) return false;
  for (long long l : lds)
    if (l % 4) return false;
  for (const void* p : ptrs)
    if ((uintptr_t)p % 16) return false;
  return true;
}

static unsigned bn_apply_grid(long long total) {
  long long g = (total + 255) / 256;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace rtf

using namespace rtf;

extern "C" int rtf_bn_workspace(int64_t B, int C, size_t* bytes) {
  if (!bytes || B < 0 || C <= 0) return RTF_E_ARG;
  // [chunk][2][C] partials + sum_dy, sum_dy_xmu
  *bytes = ((size_t)BN_MAX_CHUNKS * 2 + 2) * (size_t)C * 4 + 64;
  return 0;
}

// Training-mode forward.  d_gamma / d_beta may be NULL (scale=False / center=False), d_y may be NULL
// (statistics only), d_moving_* may be NULL.  d_mean / d_invstd (C floats each) are kept for the
// backward.  Rows of x / y are ldx / ldy floats apart.
extern "C" int rtf_bn_fwd(const float* d_x, int64_t ldx, int64_t B, int C, const float* d_gamma,
                          const float* d_beta, float eps, float momentum, float* d_y, int64_t ldy,
                          float* d_mean, float* d_invstd, float* d_moving_mean, float* d_moving_var,
                          void* d_ws, size_t ws_bytes, void* stream) {
  if (B <= 0 || C <= 0 || !d_x || !d_mean || !d_invstd || !d_ws || ldx < C || (d_y && ldy < C))
    return RTF_E_ARG;
  size_t need = 0;
  rtf_bn_workspace(B, C, &need);
  if (ws_bytes < need) return RTF_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)d_ws;
  const bool v4 = bn_vec4(C, {ldx}, {d_x, d_ws});
  const int width = v4 ? C / 4 : C;
  const BnTile t = bn_tile(B, width);
  dim3 g((width + t.cx - 1) / t.cx, t.chunks);
  if (v4)
    bn_reduce_stage1<4, 0><<<g, BN_THREADS, 0, st>>>(d_x, ldx, nullptr, 0, nullptr, B, C, t.cx, t.ry,
                                                     t.rows_per_chunk, partial);
  else
    bn_reduce_stage1<1, 0><<<g, BN_THREADS, 0, st>>>(d_x, ldx, nullptr, 0, nullptr, B, C, t.cx, t.ry,
                                                     t.rows_per_chunk, partial);
  bn_stats_stage2<<<(C + 31) / 32, BN2_WARPS * 32, 0, st>>>(partial, t.chunks, d_x, B, C, eps, momentum, d_mean,
                                                   d_invstd, d_moving_mean, d_moving_var);
  if (d_y) {
    if (bn_vec4(C, {ldx, ldy}, {d_x, d_y, d_mean, d_invstd, d_gamma, d_beta}))
      bn_apply_kernel<4><<<bn_apply_grid(B * (C / 4)), 256, 0, st>>>(d_x, ldx, B, C, d_mean, d_invstd,
                                                                     d_gamma, d_beta, d_y, ldy);
    else
      bn_apply_kernel<1><<<bn_apply_grid(B * C), 256, 0, st>>>(d_x, ldx, B, C, d_mean, d_invstd,
                                                               d_gamma, d_beta, d_y, ldy);
  }
  RTF_CHECK_LAUNCH();
  return 0;
}

// Backward of the training-mode forward.  d_dx may be NULL (the input needs no gradient: the
// bottom MLP's BatchNormalization sits on the raw dense features); d_dgamma / d_dbeta may be NULL.
extern "C" int rtf_bn_bwd(const float* d_dy, int64_t lddy, const float* d_x, int64_t ldx, int64_t B,
                          int C, const float* d_mean, const float* d_invstd, const float* d_gamma,
                          float* d_dx, int64_t lddx, float* d_dgamma, float* d_dbeta, void* d_ws,
                          size_t ws_bytes, void* stream) {
  if (B <= 0 || C <= 0 || !d_dy || !d_x || !d_mean || !d_invstd || !d_ws || ldx < C || lddy < C ||
      (d_dx && lddx < C))
    return RTF_E_ARG;
  size_t need = 0;
  rtf_bn_workspace(B, C, &need);
  if (ws_bytes < need) return RTF_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)d_ws;
  float* sum_dy = partial + (size_t)BN_MAX_CHUNKS * 2 * C;
  float* sum_dy_xmu = sum_dy + C;
  const bool v4 = bn_vec4(C, {ldx, lddy}, {d_x, d_dy, d_mean, d_ws});
  const int width = v4 ? C / 4 : C;
  const BnTile t = bn_tile(B, width);
  dim3 g((width + t.cx - 1) / t.cx, t.chunks);
  if (v4)
    bn_reduce_stage1<4, 1><<<g, BN_THREADS, 0, st>>>(d_x, ldx, d_dy, lddy, d_mean, B, C, t.cx, t.ry,
                                                     t.rows_per_chunk, partial);
  else
    bn_reduce_stage1<1, 1><<<g, BN_THREADS, 0, st>>>(d_x, ldx, d_dy, lddy, d_mean, B, C, t.cx, t.ry,
                                                     t.rows_per_chunk, partial);
  bn_bwd_stage2<<<(C + 31) / 32, BN2_WARPS * 32, 0, st>>>(partial, t.chunks, C, d_invstd, sum_dy, sum_dy_xmu,
                                                 d_dgamma, d_dbeta);
  if (d_dx) {
    if (bn_vec4(C, {ldx, lddy, lddx}, {d_x, d_dy, d_dx, d_mean, d_invstd, d_gamma, d_ws}))
      bn_bwd_apply_kernel<4><<<bn_apply_grid(B * (C / 4)), 256, 0, st>>>(
          d_dy, lddy, d_x, ldx, B, C, d_mean, d_invstd, d_gamma, sum_dy, sum_dy_xmu, d_dx, lddx);
    else
      bn_bwd_apply_kernel<1><<<bn_apply_grid(B * C), 256, 0, st>>>(
          d_dy, lddy, d_x, ldx, B, C, d_mean, d_invstd, d_gamma, sum_dy, sum_dy_xmu, d_dx, lddx);
  }
  RTF_CHECK_LAUNCH();
  return 0;
}
