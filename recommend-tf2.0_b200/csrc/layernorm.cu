// LayerNormalization over the last axis of (rows, C) activations — the two normalisations of the
// reference's TransformerEncoder block (src/match/layers/modules.py:173-185; Keras
// LayerNormalization(epsilon): biased variance, gamma / beta per feature — App. A8) and its
// gradient.  SASRec normalises 204 800 rows of 64 floats four times per step: the framework's
// kernel runs one tiny CTA per row (315 us for 52 MB in + 52 MB out, 0.33 TB/s); here a warp owns
// a row (lanes hold C/32 values in registers, two shuffle reductions, mean first then the centred
// sum of squares — the two-pass form, exact to fp32 rounding), 8 rows per CTA, and the kernel is
// the streaming pass it should be.  Backward: the per-row dx in the same shape, dgamma / dbeta as
// deterministic two-stage column sums (CTA partials over row chunks, added in chunk order).
#include "rtf_common.cuh"

namespace rtf {

constexpr int LN_MAXV = 8;          // values per lane: C <= 256 on the warp-per-row path
constexpr int LN_WARPS = 8;
constexpr int LN_MAX_CHUNKS = 592;  // backward: CTAs (4 per SM); each leaves one partial row

template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const float* __restrict__ x, long long rows, int C, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, float* __restrict__ y,
              float* __restrict__ mean, float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * C;
  float v[NV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane + 32 * k;
    v[k] = c < C ? __ldg(xr + c) : 0.f;
    s += v[k];
  }
  const float mu = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane + 32 * k;
    const float d = c < C ? v[k] - mu : 0.f;
    q = fmaf(d, d, q);
  }
  const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
  float* yr = y + row * C;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane + 32 * k;
    if (c < C) {
      const float g = gamma ? __ldg(gamma + c) : 1.f, b = beta ? __ldg(beta + c) : 0.f;
      yr[c] = fmaf((v[k] - mu) * rs, g, b);
    }
  }
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)),  g = dy * gamma;  per-CTA partial sums of
// dy * xhat (-> dgamma) and dy (-> dbeta) over the CTA's rows: warp w adds its 8 rows in row order,
// the 8 warps are added in warp order.
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, long long rows, int C,
              int rows_per_cta, float* __restrict__ dx, float* __restrict__ partial) {
  __shared__ float red[2][LN_WARPS][32 * NV];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float gw[NV], pg[NV], pb[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = lane + 32 * k;
    gw[k] = (gamma && c < C) ? __ldg(gamma + c) : 1.f;
    pg[k] = pb[k] = 0.f;
  }
  const long long r0 = (long long)blockIdx.x * rows_per_cta + (long long)w * (rows_per_cta / LN_WARPS);
  const long long r1 = min(r0 + rows_per_cta / LN_WARPS, rows);
  // the next row's dy / x are fetched while this one is reduced (a warp has one row in flight
  // otherwise, and every row is two dependent shuffle reductions behind its loads)
  float dn[NV], xn[NV];
  auto fetch = [&](long long row) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      const bool ok = row < r1 && c < C;
      dn[k] = ok ? __ldg(dy + row * C + c) : 0.f;
      xn[k] = ok ? __ldg(x + row * C + c) : 0.f;
    }
  };
  fetch(r0);
  for (long long row = r0; row < r1; ++row) {
    const float mu = mean[row], rs = rstd[row];
    float xh[NV], g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      const float d = dn[k];
      xh[k] = c < C ? (xn[k] - mu) * rs : 0.f;
      g[k] = d * gw[k];
      s1 += g[k];
      s2 = fmaf(g[k], xh[k], s2);
      pg[k] = fmaf(d, xh[k], pg[k]);
      pb[k] += d;
    }
    fetch(row + 1);
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    if (dx) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane + 32 * k;
        if (c < C) dx[row * C + c] = rs * (g[k] - s1 - xh[k] * s2);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    red[0][w][lane + 32 * k] = pg[k];
    red[1][w][lane + 32 * k] = pb[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = red[0][0][c], b = red[1][0][c];
#pragma unroll
    for (int u = 1; u < LN_WARPS; ++u) {
      a = __fadd_rn(a, red[0][u][c]);
      b = __fadd_rn(b, red[1][u][c]);
    }
    partial[(long long)blockIdx.x * 2 * C + c] = a;
    partial[(long long)blockIdx.x * 2 * C + C + c] = b;
  }
}

// chunk partials -> dgamma, dbeta: 32 columns per CTA, 32 warps; warp w adds chunks w, w+32, ... in
// double (ascending, 16 loads in flight), the warps' sums are added in warp order: a fixed tree
// (8 warps with 8 loads in flight took 15 us per call: load latency)
constexpr int LN2_WARPS = 32;
__global__ void __launch_bounds__(LN2_WARPS * 32)
ln_bwd_stage2(const float* __restrict__ partial, long long chunks, int C, float* __restrict__ dgamma,
              float* __restrict__ dbeta) {
  __shared__ double red[LN2_WARPS][2][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double a0 = 0.0, a1 = 0.0;
  if (c < C) {
    long long k = w;
    for (; k + 7 * LN2_WARPS < chunks; k += 8 * LN2_WARPS) {
      float t0[8], t1[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        t0[u] = partial[(k + LN2_WARPS * u) * 2 * C + c];
        t1[u] = partial[(k + LN2_WARPS * u) * 2 * C + C + c];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a0 += (double)t0[u];
        a1 += (double)t1[u];
      }
    }
    for (; k < chunks; k += LN2_WARPS) {
      a0 += (double)partial[k * 2 * C + c];
      a1 += (double)partial[k * 2 * C + C + c];
    }
  }
  red[w][0][lane] = a0;
  red[w][1][lane] = a1;
  __syncthreads();
  if (w == 0 && c < C) {
    double s0 = red[0][0][lane], s1 = red[0][1][lane];
#pragma unroll
    for (int g = 1; g < LN2_WARPS; ++g) {
      s0 += red[g][0][lane];
      s1 += red[g][1][lane];
    }
    if (dgamma) dgamma[c] = (float)s0;
    if (dbeta) dbeta[c] = (float)s1;
  }
}

}  // namespace rtf

using namespace rtf;

// rows per CTA of the backward: a multiple of 8 (one block of rows per warp), >= 64, and large
// enough that at most LN_MAX_CHUNKS partial rows are left for stage 2
static int ln_rows_per_cta(int64_t rows) {
  int64_t r = (rows + LN_MAX_CHUNKS - 1) / LN_MAX_CHUNKS;
  if (r < 64) r = 64;
  return (int)((r + 7) / 8 * 8);
}

extern "C" int rtf_layernorm_workspace(int64_t rows, int C, size_t* bytes) {
  if (!bytes || rows < 0 || C <= 0) return RTF_E_ARG;
  *bytes = (size_t)(LN_MAX_CHUNKS + 1) * 2 * (size_t)C * 4;
  return 0;
}

// y = (x - mean_row) / sqrt(var_row + eps) * gamma + beta over the last axis of contiguous (rows, C);
// d_mean / d_rstd (rows floats each) are kept for the backward.  gamma / beta may be NULL.
extern "C" int rtf_layernorm_fwd(const float* d_x, int64_t rows, int C, const float* d_gamma,
                                 const float* d_beta, float eps, float* d_y, float* d_mean,
                                 float* d_rstd, void* stream) {
  if (rows < 0 || C <= 0) return RTF_E_ARG;
  if (C > 32 * LN_MAXV) return RTF_E_RANGE;
  if (rows == 0) return 0;
  if (!d_x || !d_y || !d_mean || !d_rstd) return RTF_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((rows + LN_WARPS - 1) / LN_WARPS);
  const int nv = (C + 31) / 32;
#define RTF_LN_FWD(NV)                                                                          \
  ln_fwd_kernel<NV><<<grid, LN_WARPS * 32, 0, st>>>(d_x, rows, C, d_gamma, d_beta, eps, d_y, d_mean, \
                                                    d_rstd)
  if (nv <= 1) RTF_LN_FWD(1);
  else if (nv <= 2) RTF_LN_FWD(2);
  else if (nv <= 4) RTF_LN_FWD(4);
  else RTF_LN_FWD(8);
#undef RTF_LN_FWD
  RTF_CHECK_LAUNCH();
  return 0;
}

// d_dx may be NULL (input needs no gradient); d_dgamma / d_dbeta may be NULL.
extern "C" int rtf_layernorm_bwd(const float* d_dy, const float* d_x, const float* d_mean,
                                 const float* d_rstd, const float* d_gamma, int64_t rows, int C,
                                 float* d_dx, float* d_dgamma, float* d_dbeta, void* d_ws,
                                 size_t ws_bytes, void* stream) {
  if (rows < 0 || C <= 0) return RTF_E_ARG;
  if (C > 32 * LN_MAXV) return RTF_E_RANGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    if (d_dgamma) cudaMemsetAsync(d_dgamma, 0, (size_t)C * 4, st);
    if (d_dbeta) cudaMemsetAsync(d_dbeta, 0, (size_t)C * 4, st);
    return 0;
  }
  if (!d_dy || !d_x || !d_mean || !d_rstd || !d_ws) return RTF_E_ARG;
  size_t need = 0;
  rtf_layernorm_workspace(rows, C, &need);
  if (ws_bytes < need) return RTF_E_WORKSPACE;
  const int rpc = ln_rows_per_cta(rows);
  const long long chunks = (rows + rpc - 1) / rpc;
  const int nv = (C + 31) / 32;
#define RTF_LN_BWD(NV)                                                                       \
  ln_bwd_kernel<NV><<<(unsigned)chunks, LN_WARPS * 32, 0, st>>>(d_dy, d_x, d_mean, d_rstd, d_gamma, \
                                                                rows, C, rpc, d_dx, (float*)d_ws)
  if (nv <= 1) RTF_LN_BWD(1);
  else if (nv <= 2) RTF_LN_BWD(2);
  else if (nv <= 4) RTF_LN_BWD(4);
  else RTF_LN_BWD(8);
#undef RTF_LN_BWD
  if (d_dgamma || d_dbeta)
    ln_bwd_stage2<<<(C + 31) / 32, LN2_WARPS * 32, 0, st>>>((const float*)d_ws, chunks, C, d_dgamma, d_dbeta);
  RTF_CHECK_LAUNCH();
  return 0;
}
