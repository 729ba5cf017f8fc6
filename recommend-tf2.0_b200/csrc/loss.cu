// Keras binary_crossentropy on probabilities — the loss every CTR training script of the
// reference compiles its model with (src/ctr/fm/train.py:49, src/ctr/deep_fm/train.py:50,
// src/ctr/din/train.py:103; semantics SURVEY.md App. A11):
//   pc = clip(p, 1e-7, 1 - 1e-7),  loss = -mean(y log(pc + 1e-7) + (1 - y) log(1 - pc + 1e-7))
// Composed from framework elementwise ops this is ~12 launches forward and ~18 backward over a
// (B,) vector — 30 launches of 2-3 us each, a third of the FM step and 0.1 ms of the DLRM step.
// Here: one pass writes the per-chunk partial sums AND d loss / d p (clip's gradient included:
// zero outside [1e-7, 1 - 1e-7]); a one-warp second stage adds the partials in chunk order
// (double, fixed tree: reproducible).
#include "rtf_common.cuh"

namespace rtf {

constexpr int BCE_THREADS = 256, BCE_PER = 4, BCE_CHUNK = BCE_THREADS * BCE_PER;

__global__ void __launch_bounds__(BCE_THREADS)
bce_stage1(const float* __restrict__ y, const float* __restrict__ p, long long n, float inv_n,
           float* __restrict__ dp, float* __restrict__ partial) {
  __shared__ float wsum[BCE_THREADS / 32];
  const float eps = 1e-7f, hi = (float)(1.0 - 1e-7);
  const long long base = (long long)blockIdx.x * BCE_CHUNK + threadIdx.x;
  float pv[BCE_PER], yv[BCE_PER];
#pragma unroll
  for (int e = 0; e < BCE_PER; ++e) {
    const long long i = base + (long long)e * BCE_THREADS;
    pv[e] = i < n ? p[i] : 0.5f;
    yv[e] = i < n ? y[i] : 0.f;
  }
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < BCE_PER; ++e) {
    const long long i = base + (long long)e * BCE_THREADS;
    const float pc = fminf(fmaxf(pv[e], eps), hi);
    const float a = __fadd_rn(pc, eps), b = __fadd_rn(__fsub_rn(1.0f, pc), eps);
    const float yy = yv[e], ny = __fsub_rn(1.0f, yy);
    const float term = __fadd_rn(__fmul_rn(yy, logf(a)), __fmul_rn(ny, logf(b)));
    if (i < n) {
      acc = __fadd_rn(acc, term);
      if (dp) {
        const bool inside = pv[e] >= eps && pv[e] <= hi;   // clip passes the gradient on [lo, hi]
        const float d = __fsub_rn(__fdiv_rn(ny, b), __fdiv_rn(yy, a));
        dp[i] = inside ? __fmul_rn(d, inv_n) : 0.f;
      }
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = wsum[0];
#pragma unroll
    for (int w = 1; w < BCE_THREADS / 32; ++w) t = __fadd_rn(t, wsum[w]);
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(32)
bce_stage2(const float* __restrict__ partial, int nchunks, double inv_n, float* __restrict__ loss) {
  double a = 0.0;
  for (int k = threadIdx.x; k < nchunks; k += 32) a += (double)partial[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (threadIdx.x == 0) *loss = (float)(-a * inv_n);
}

}  // namespace rtf

using namespace rtf;

extern "C" int rtf_bce_workspace(int64_t n, size_t* bytes) {
  if (!bytes || n < 0) return RTF_E_ARG;
  *bytes = (size_t)((n + BCE_CHUNK - 1) / BCE_CHUNK + 1) * 4;
  return 0;
}

extern "C" int rtf_bce_fwd(const float* d_y, const float* d_p, int64_t n, float* d_loss, float* d_dp,
                           void* d_ws, void* stream) {
  if (n <= 0 || !d_y || !d_p || !d_loss || !d_ws) return RTF_E_ARG;
  if (n > ((int64_t)1 << 40)) return RTF_E_RANGE;
  cudaStream_t st = (cudaStream_t)stream;
  const long long nchunks = (n + BCE_CHUNK - 1) / BCE_CHUNK;
  bce_stage1<<<(unsigned)nchunks, BCE_THREADS, 0, st>>>(d_y, d_p, n, (float)(1.0 / (double)n), d_dp,
                                                        (float*)d_ws);
  bce_stage2<<<1, 32, 0, st>>>((const float*)d_ws, (int)nchunks, 1.0 / (double)n, d_loss);
  RTF_CHECK_LAUNCH();
  return 0;
}
