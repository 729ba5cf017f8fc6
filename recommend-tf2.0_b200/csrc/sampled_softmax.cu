// K8 — sampled softmax: tf.nn.sampled_softmax_loss as SampledSoftmaxLayer.call uses it
// (src/match/layers/modules.py:54-60; semantics restated in SURVEY.md App. A13/A14).
//
//   sampled, true_exp, samp_exp = log_uniform_candidate_sampler(unique=True)  — ONE draw per batch
//   true_logit[b]  = x_b . W[label_b] + bias[label_b]              - log(true_exp[b])
//   samp_logit[b,j]= x_b . W[s_j]     + bias[s_j]  (+ -FLT_MAX if s_j == label_b) - log(samp_exp[j])
//   loss[b] = logsumexp([true | sampled]) - true_logit[b]          (soft label 1 on column 0)
//
// TF runs the sampler and the accidental-hit search as CPU ops (a host sync per step); here
// the sampler is a device kernel (counter-based RNG, rejection to S unique ids) and the hit
// mask is a compare inside the logits loop.  The S sampled rows are gathered once into a
// compact (S, D) buffer that every sample's warp then streams from L2; a warp keeps x_b in
// registers (lanes over D), reduces each logit with shuffles and folds it into a running
// log-sum-exp.  TF's Philox stream cannot be reproduced: parity is on injected samples.
#include <float.h>

#include "rtf_common.cuh"

namespace rtf {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// P(c) = log((c+2)/(c+1)) / log(range_max+1); draw = floor(exp(u * log(range_max+1))) - 1
//
// One warp, 32 tries per round: try t0 + lane is drawn by lane (counter-based RNG, so the sample
// sequence is a pure function of the seed), duplicates inside the round are resolved with
// match_any (lowest lane = earliest try wins), duplicates against earlier rounds by probing an
// open-addressing hash set, the survivors are appended IN TRY ORDER (ballot ranks) and inserted
// with atomicCAS.  The result — sampled ids, their order and num_tries — is exactly what the
// sequential rejection loop of App. A14 yields for this RNG stream; the set lives in shared
// memory when it fits (global-memory probes on one thread cost 465 us for S = 1024, this 25 us).
__global__ void __launch_bounds__(32)
log_uniform_sample_kernel(uint64_t seed, const uint64_t* __restrict__ seed_dev, int S,
                          long long range_max, long long* __restrict__ out,
                          int32_t* __restrict__ num_tries, long long* __restrict__ gtable, int cap,
                          int use_smem) {
  extern __shared__ long long stab[];
  if (seed_dev) seed = *seed_dev;     // per-step seed of a CUDA-graph replay lives in memory
  long long* table = use_smem ? stab : gtable;
  const int lane = threadIdx.x;
  for (int i = lane; i < cap; i += 32) table[i] = -1;
  __syncwarp();
  const double log_range = log((double)range_max + 1.0);
  const uint32_t lt = (1u << lane) - 1u;
  int got = 0, t0 = 0;
  while (true) {
    const uint64_t r = splitmix64(seed ^ splitmix64((uint64_t)(t0 + lane)));
    const double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);  // [0,1)
    long long c = (long long)(exp(u * log_range)) - 1;
    if (c < 0) c = 0;
    if (c >= range_max) c = range_max - 1;
    // earliest try of every distinct value in this round
    const uint32_t same = __match_any_sync(0xffffffffu, c);
    bool fresh = (same & lt) == 0;
    uint32_t hsh = (uint32_t)(splitmix64((uint64_t)c) & (uint64_t)(cap - 1));
    if (fresh) {   // seen in an earlier round?
      while (true) {
        const long long e = reinterpret_cast<volatile long long*>(table)[hsh];
        if (e == -1) break;
        if (e == c) { fresh = false; break; }
        hsh = (hsh + 1) & (uint32_t)(cap - 1);
      }
    }
    const uint32_t fm = __ballot_sync(0xffffffffu, fresh);
    const int rank = __popc(fm & lt);
    const bool keep = fresh && got + rank < S;
    if (keep) {
      out[got + rank] = c;
      uint32_t h2 = hsh;   // the first free slot seen; another lane of this round may take it
      while (atomicCAS(reinterpret_cast<unsigned long long*>(table + h2), (unsigned long long)-1ll,
                       (unsigned long long)c) != (unsigned long long)-1ll)
        h2 = (h2 + 1) & (uint32_t)(cap - 1);
    }
    const int nf = __popc(fm);
    if (got + nf >= S) {
      // the try that produced the S-th unique id ends the sequence
      const uint32_t last = __ballot_sync(0xffffffffu, fresh && got + rank == S - 1);
      if (lane == 0) *num_tries = t0 + __ffs(last);
      return;
    }
    got += nf;
    t0 += 32;
    __syncwarp();
  }
}

// expected_count(c) = -expm1(num_tries * log1p(-P(c)))  (unique=True), App. A14
__global__ void __launch_bounds__(256)
log_uniform_expected_kernel(const long long* __restrict__ ids, long long n, long long range_max,
                            const int32_t* __restrict__ num_tries, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double c = (double)ids[i];
  const double p = (log(c + 2.0) - log(c + 1.0)) / log((double)range_max + 1.0);
  out[i] = (float)(-expm1((double)*num_tries * log1p(-p)));
}

struct SsmParams {
  const float* x; long long x_sb;   // (B, D)
  const float* W;                   // (N, D)
  const float* bias;                // (N) or null
  const long long* labels;          // (B)
  const long long* sampled;         // (S)
  const float* true_exp;            // (B)
  const float* samp_exp;            // (S)
  long long B, N;
  int S, D, remove_hits;
  float* Ws;                        // workspace (S, D): gathered sampled rows
  float* cs;                        // workspace (S): bias[s_j] - log(samp_exp[j])
  float* loss; float* lse;          // (B)
  const float* gloss;               // (B) backward
  float* gx; long long gx_sb;       // (B, D)
  float* G;                         // (B, S+1): dL/dlogit, column 0 = true class
  int32_t* err;
};

__global__ void __launch_bounds__(256) ssm_gather(const __grid_constant__ SsmParams P) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P.S) return;
  long long id = P.sampled[warp];
  const bool ok = id >= 0 && id < P.N;
  if (!ok && lane == 0 && P.err) atomicOr(P.err, 1);
  for (int c = lane; c < P.D; c += 32) P.Ws[(long long)warp * P.D + c] = ok ? P.W[id * P.D + c] : 0.f;
  if (lane == 0) P.cs[warp] = ((ok && P.bias) ? P.bias[id] : 0.f) - logf(P.samp_exp[warp]);
}

// one warp per sample; lanes over D (CPL columns per lane)
template <int CPL, bool BWD>
__global__ void __launch_bounds__(256) ssm_kernel(const __grid_constant__ SsmParams P) {
  const int lane = threadIdx.x & 31;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= P.B) return;
  const int D = P.D, S = P.S;
  float xr[CPL];
#pragma unroll
  for (int t = 0; t < CPL; ++t) {
    const int c = lane + 32 * t;
    xr[t] = c < D ? P.x[b * P.x_sb + c] : 0.f;
  }
  const long long label = P.labels[b];
  const bool lab_ok = label >= 0 && label < P.N;
  if (!lab_ok && lane == 0 && P.err) atomicOr(P.err, 1);
  float t0 = 0.f;
#pragma unroll
  for (int t = 0; t < CPL; ++t) {
    const int c = lane + 32 * t;
    if (c < D && lab_ok) t0 = fmaf(xr[t], __ldg(P.W + label * D + c), t0);
  }
  t0 = warp_sum(t0) + ((lab_ok && P.bias) ? P.bias[label] : 0.f) - logf(P.true_exp[b]);
  if (!BWD) {
    float m = t0, l = 1.f;
    for (int j = 0; j < S; ++j) {
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < CPL; ++t) {
        const int c = lane + 32 * t;
        if (c < D) s = fmaf(xr[t], P.Ws[(long long)j * D + c], s);
      }
      s = warp_sum(s);
      if (P.remove_hits && P.sampled[j] == label) s += -FLT_MAX;
      s += P.cs[j];
      const float mn = fmaxf(m, s);
      l = l * expf(m - mn) + expf(s - mn);
      m = mn;
    }
    if (lane == 0) {
      const float lse = m + logf(l);
      P.lse[b] = lse;
      P.loss[b] = lse - t0;
    }
  } else {
    const float lse = P.lse[b], g = P.gloss[b];
    float gxr[CPL];
    const float g0 = g * (expf(t0 - lse) - 1.f);
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
      const int c = lane + 32 * t;
      gxr[t] = (c < D && lab_ok) ? g0 * __ldg(P.W + label * D + c) : 0.f;
    }
    float* Gb = P.G + b * (long long)(S + 1);
    if (lane == 0) Gb[0] = g0;
    for (int j = 0; j < S; ++j) {
      float s = 0.f;
      float wv[CPL];
#pragma unroll
      for (int t = 0; t < CPL; ++t) {
        const int c = lane + 32 * t;
        wv[t] = c < D ? P.Ws[(long long)j * D + c] : 0.f;
        s = fmaf(xr[t], wv[t], s);
      }
      s = warp_sum(s);
      if (P.remove_hits && P.sampled[j] == label) s += -FLT_MAX;
      s += P.cs[j];
      const float gj = g * expf(s - lse);
      if (lane == 0) Gb[1 + j] = gj;
#pragma unroll
      for (int t = 0; t < CPL; ++t) gxr[t] = fmaf(gj, wv[t], gxr[t]);
    }
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
      const int c = lane + 32 * t;
      if (c < D) P.gx[b * P.gx_sb + c] = gxr[t];
    }
  }
}

// ---- GEMM form (large S): the (B, S) sampled logits come from the tensor-core GEMM
// x (B,D) . Ws^T (S,D) (rtf_dense_gemm_nt, fp32-accurate) instead of every sample's warp
// re-streaming all S sampled rows from L2 (S = 1024, D = 64: 1.07 GB of L2 reads per launch,
// 0.35 ms = 2 % of the FMA peak); these kernels are the per-sample epilogues around it.
struct SsmLogitParams {
  const float* logits; long long ld;   // (B, S) raw x . w_s
  const float* x; long long x_sb;
  const float* W; const float* bias;
  const long long* labels; const long long* sampled;
  const float* true_exp; const float* cs;   // cs[j] = bias[s_j] - log(samp_exp[j]) (ssm_gather)
  long long B, N;
  int S, D, remove_hits;
  float* loss; float* lse;
  const float* gloss;
  float* g0; float* G1; long long g1_ld;    // bwd: dL/dlogit true (B) and sampled (B, S)
  float* gx; long long gx_sb;               // bwd: gx += g0 * W[label]
  int32_t* err;
};

__device__ __forceinline__ float ssm_true_logit(const SsmLogitParams& P, long long b, int lane,
                                                bool& lab_ok) {
  const long long label = P.labels[b];
  lab_ok = label >= 0 && label < P.N;
  float t0 = 0.f;
  if (lab_ok)
    for (int c = lane; c < P.D; c += 32) t0 = fmaf(P.x[b * P.x_sb + c], __ldg(P.W + label * P.D + c), t0);
  return warp_sum(t0) + ((lab_ok && P.bias) ? P.bias[label] : 0.f) - logf(P.true_exp[b]);
}

__device__ __forceinline__ float ssm_sampled_logit(const SsmLogitParams& P, long long b, int j,
                                                   long long label) {
  float sv = P.logits[b * P.ld + j];
  if (P.remove_hits && P.sampled[j] == label) sv += -FLT_MAX;
  return sv + P.cs[j];
}

__global__ void __launch_bounds__(256) ssm_logits_fwd_kernel(const __grid_constant__ SsmLogitParams P) {
  const int lane = threadIdx.x & 31;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= P.B) return;
  bool lab_ok;
  const float t0 = ssm_true_logit(P, b, lane, lab_ok);
  if (!lab_ok && lane == 0 && P.err) atomicOr(P.err, 1);
  const long long label = P.labels[b];
  float m = -INFINITY, l = 0.f;
  for (int j = lane; j < P.S; j += 32) {
    const float sv = ssm_sampled_logit(P, b, j, label);
    const float mn = fmaxf(m, sv);
    l = l * expf(m - mn) + expf(sv - mn);
    m = mn;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {   // combine the lanes' (max, sum) pairs
    const float mo = __shfl_xor_sync(0xffffffffu, m, o), lo = __shfl_xor_sync(0xffffffffu, l, o);
    const float mn = fmaxf(m, mo);
    l = (m == -INFINITY ? 0.f : l * expf(m - mn)) + (mo == -INFINITY ? 0.f : lo * expf(mo - mn));
    m = mn;
  }
  if (lane == 0) {
    const float mn = fmaxf(m, t0);
    const float tot = (m == -INFINITY ? 0.f : l * expf(m - mn)) + expf(t0 - mn);
    const float lse = mn + logf(tot);
    P.lse[b] = lse;
    P.loss[b] = lse - t0;
  }
}

__global__ void __launch_bounds__(256) ssm_logits_bwd_kernel(const __grid_constant__ SsmLogitParams P) {
  const int lane = threadIdx.x & 31;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= P.B) return;
  bool lab_ok;
  const float t0 = ssm_true_logit(P, b, lane, lab_ok);
  const long long label = P.labels[b];
  const float lse = P.lse[b], g = P.gloss[b];
  if (lane == 0) P.g0[b] = g * (expf(t0 - lse) - 1.f);
  for (int j = lane; j < P.S; j += 32)
    P.G1[b * P.g1_ld + j] = g * expf(ssm_sampled_logit(P, b, j, label) - lse);
}

// gx[b] += g0[b] * W[label[b]]  (the true class' share of d loss / d x; the sampled share is a GEMM)
__global__ void __launch_bounds__(256) ssm_true_gx_kernel(const __grid_constant__ SsmLogitParams P) {
  const int lane = threadIdx.x & 31;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= P.B) return;
  const long long label = P.labels[b];
  if (label < 0 || label >= P.N) return;
  const float g0 = P.g0[b];
  for (int c = lane; c < P.D; c += 32)
    P.gx[b * P.gx_sb + c] = fmaf(g0, __ldg(P.W + label * P.D + c), P.gx[b * P.gx_sb + c]);
}

template <bool BWD>
static int ssm_launch(const SsmParams& P, cudaStream_t st) {
  const unsigned blocks = (unsigned)((P.B * 32 + 255) / 256);
  const int cpl = (P.D + 31) / 32;
  if (cpl <= 1) ssm_kernel<1, BWD><<<blocks, 256, 0, st>>>(P);
  else if (cpl <= 2) ssm_kernel<2, BWD><<<blocks, 256, 0, st>>>(P);
  else if (cpl <= 4) ssm_kernel<4, BWD><<<blocks, 256, 0, st>>>(P);
  else if (cpl <= 8) ssm_kernel<8, BWD><<<blocks, 256, 0, st>>>(P);
  else return RTF_E_RANGE;
  RTF_CHECK_LAUNCH();
  return 0;
}

}  // namespace rtf

using namespace rtf;

extern "C" int rtf_log_uniform_workspace(int S, size_t* bytes) {
  if (!bytes || S <= 0) return RTF_E_ARG;
  size_t cap = 64;
  while (cap < (size_t)2 * S) cap <<= 1;
  *bytes = cap * 8;
  return 0;
}

static int log_uniform_sample_launch(uint64_t seed, const uint64_t* d_seed, int S, int64_t range_max,
                                     int64_t* d_sampled, int32_t* d_num_tries, void* d_ws,
                                     void* stream) {
  if (S <= 0 || range_max <= 0 || !d_sampled || !d_num_tries || !d_ws) return RTF_E_ARG;
  if ((int64_t)S > range_max) return RTF_E_RANGE;  // TF would never terminate (A14)
  int cap = 64;
  while (cap < 2 * S) cap <<= 1;
  const size_t smem = (size_t)cap * 8;
  const int use_smem = smem <= 200 * 1024;
  if (use_smem && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(log_uniform_sample_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  log_uniform_sample_kernel<<<1, 32, use_smem ? smem : 0, (cudaStream_t)stream>>>(
      seed, d_seed, S, range_max, (long long*)d_sampled, d_num_tries, (long long*)d_ws, cap,
      use_smem);
  RTF_CHECK_LAUNCH();
  return 0;
}

extern "C" int rtf_log_uniform_sample(uint64_t seed, int S, int64_t range_max, int64_t* d_sampled,
                                      int32_t* d_num_tries, void* d_ws, void* stream) {
  return log_uniform_sample_launch(seed, nullptr, S, range_max, d_sampled, d_num_tries, d_ws, stream);
}

extern "C" int rtf_log_uniform_sample_dseed(const uint64_t* d_seed, int S, int64_t range_max,
                                            int64_t* d_sampled, int32_t* d_num_tries, void* d_ws,
                                            void* stream) {
  if (!d_seed) return RTF_E_ARG;
  return log_uniform_sample_launch(0, d_seed, S, range_max, d_sampled, d_num_tries, d_ws, stream);
}

extern "C" int rtf_log_uniform_expected(const int64_t* d_ids, int64_t n, int64_t range_max,
                                        const int32_t* d_num_tries, float* d_out, void* stream) {
  if (n < 0 || range_max <= 0) return RTF_E_ARG;
  if (n == 0) return 0;
  if (!d_ids || !d_num_tries || !d_out) return RTF_E_ARG;
  log_uniform_expected_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const long long*)d_ids, n, range_max, d_num_tries, d_out);
  RTF_CHECK_LAUNCH();
  return 0;
}

static int ssm_fill(SsmParams& P, const float* x, int64_t x_sb, const float* W, const float* bias,
                    const int64_t* labels, const int64_t* sampled, const float* te, const float* se,
                    int64_t B, int64_t N, int S, int D, int remove_hits, float* ws) {
  if (B < 0 || N <= 0 || S <= 0 || D <= 0) return RTF_E_ARG;
  if (D > 256) return RTF_E_RANGE;
  if (B == 0) return 0;
  if (!x || !W || !labels || !sampled || !te || !se || !ws) return RTF_E_ARG;
  P.x = x; P.x_sb = x_sb; P.W = W; P.bias = bias; P.labels = (const long long*)labels;
  P.sampled = (const long long*)sampled; P.true_exp = te; P.samp_exp = se; P.B = B; P.N = N;
  P.S = S; P.D = D; P.remove_hits = remove_hits; P.Ws = ws; P.cs = ws + (size_t)S * D;
  return 0;
}

extern "C" int rtf_sampled_softmax_workspace(int S, int D, size_t* bytes) {
  if (!bytes || S <= 0 || D <= 0) return RTF_E_ARG;
  *bytes = ((size_t)S * D + S) * 4;
  return 0;
}

extern "C" int rtf_sampled_softmax_fwd(const float* d_x, int64_t x_sb, const float* d_W,
                                       const float* d_bias, const int64_t* d_labels,
                                       const int64_t* d_sampled, const float* d_true_exp,
                                       const float* d_samp_exp, int64_t B, int64_t N, int S, int D,
                                       int remove_hits, float* d_loss, float* d_lse, void* d_ws,
                                       int32_t* d_err, void* stream) {
  SsmParams P = {};
  int rc = ssm_fill(P, d_x, x_sb, d_W, d_bias, d_labels, d_sampled, d_true_exp, d_samp_exp, B, N, S,
                    D, remove_hits, (float*)d_ws);
  if (rc || B == 0) return rc;
  if (!d_loss || !d_lse) return RTF_E_ARG;
  P.loss = d_loss; P.lse = d_lse; P.err = d_err;
  cudaStream_t st = (cudaStream_t)stream;
  ssm_gather<<<(unsigned)((S * 32 + 255) / 256), 256, 0, st>>>(P);
  return ssm_launch<false>(P, st);
}

extern "C" int rtf_sampled_softmax_bwd(const float* d_x, int64_t x_sb, const float* d_W,
                                       const float* d_bias, const int64_t* d_labels,
                                       const int64_t* d_sampled, const float* d_true_exp,
                                       const float* d_samp_exp, int64_t B, int64_t N, int S, int D,
                                       int remove_hits, const float* d_lse, const float* d_gloss,
                                       float* d_gx, int64_t gx_sb, float* d_G, void* d_ws,
                                       void* stream) {
  SsmParams P = {};
  int rc = ssm_fill(P, d_x, x_sb, d_W, d_bias, d_labels, d_sampled, d_true_exp, d_samp_exp, B, N, S,
                    D, remove_hits, (float*)d_ws);
  if (rc || B == 0) return rc;
  if (!d_lse || !d_gloss || !d_gx || !d_G) return RTF_E_ARG;
  P.lse = const_cast<float*>(d_lse); P.gloss = d_gloss; P.gx = d_gx; P.gx_sb = gx_sb; P.G = d_G;
  cudaStream_t st = (cudaStream_t)stream;
  ssm_gather<<<(unsigned)((S * 32 + 255) / 256), 256, 0, st>>>(P);
  return ssm_launch<true>(P, st);
}

// ---- GEMM form: pieces around x . Ws^T (the caller runs the GEMMs, e.g. rtf_dense_gemm_nt/nn/tn)
extern "C" int rtf_ssm_gather(const float* d_W, const float* d_bias, const int64_t* d_sampled,
                              const float* d_samp_exp, int64_t N, int S, int D, float* d_Ws,
                              float* d_cs, int32_t* d_err, void* stream) {
  if (N <= 0 || S <= 0 || D <= 0 || !d_W || !d_sampled || !d_samp_exp || !d_Ws || !d_cs)
    return RTF_E_ARG;
  SsmParams P = {};
  P.W = d_W; P.bias = d_bias; P.sampled = (const long long*)d_sampled; P.samp_exp = d_samp_exp;
  P.N = N; P.S = S; P.D = D; P.Ws = d_Ws; P.cs = d_cs; P.err = d_err;
  ssm_gather<<<(unsigned)((S * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
  RTF_CHECK_LAUNCH();
  return 0;
}

static int ssm_logit_fill(SsmLogitParams& P, const float* logits, int64_t ld, const float* x,
                          int64_t x_sb, const float* W, const float* bias, const int64_t* labels,
                          const int64_t* sampled, const float* te, const float* cs, int64_t B,
                          int64_t N, int S, int D, int remove_hits) {
  if (B < 0 || N <= 0 || S <= 0 || D <= 0 || ld < S) return RTF_E_ARG;
  if (B == 0) return 0;
  if (!logits || !x || !W || !labels || !sampled || !te || !cs) return RTF_E_ARG;
  P.logits = logits; P.ld = ld; P.x = x; P.x_sb = x_sb; P.W = W; P.bias = bias;
  P.labels = (const long long*)labels; P.sampled = (const long long*)sampled; P.true_exp = te;
  P.cs = cs; P.B = B; P.N = N; P.S = S; P.D = D; P.remove_hits = remove_hits;
  return 0;
}

extern "C" int rtf_ssm_logits_fwd(const float* d_logits, int64_t ld, const float* d_x, int64_t x_sb,
                                  const float* d_W, const float* d_bias, const int64_t* d_labels,
                                  const int64_t* d_sampled, const float* d_true_exp,
                                  const float* d_cs, int64_t B, int64_t N, int S, int D,
                                  int remove_hits, float* d_loss, float* d_lse, int32_t* d_err,
                                  void* stream) {
  SsmLogitParams P = {};
  int rc = ssm_logit_fill(P, d_logits, ld, d_x, x_sb, d_W, d_bias, d_labels, d_sampled, d_true_exp,
                          d_cs, B, N, S, D, remove_hits);
  if (rc || B == 0) return rc;
  if (!d_loss || !d_lse) return RTF_E_ARG;
  P.loss = d_loss; P.lse = d_lse; P.err = d_err;
  ssm_logits_fwd_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
  RTF_CHECK_LAUNCH();
  return 0;
}

extern "C" int rtf_ssm_logits_bwd(const float* d_logits, int64_t ld, const float* d_x, int64_t x_sb,
                                  const float* d_W, const float* d_bias, const int64_t* d_labels,
                                  const int64_t* d_sampled, const float* d_true_exp,
                                  const float* d_cs, int64_t B, int64_t N, int S, int D,
                                  int remove_hits, const float* d_lse, const float* d_gloss,
                                  float* d_g0, float* d_G1, int64_t g1_ld, void* stream) {
  SsmLogitParams P = {};
  int rc = ssm_logit_fill(P, d_logits, ld, d_x, x_sb, d_W, d_bias, d_labels, d_sampled, d_true_exp,
                          d_cs, B, N, S, D, remove_hits);
  if (rc || B == 0) return rc;
  if (!d_lse || !d_gloss || !d_g0 || !d_G1 || g1_ld < S) return RTF_E_ARG;
  P.lse = const_cast<float*>(d_lse); P.gloss = d_gloss; P.g0 = d_g0; P.G1 = d_G1; P.g1_ld = g1_ld;
  ssm_logits_bwd_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
  RTF_CHECK_LAUNCH();
  return 0;
}

extern "C" int rtf_ssm_true_gx(const float* d_W, const int64_t* d_labels, const float* d_g0,
                               int64_t B, int64_t N, int D, float* d_gx, int64_t gx_sb,
                               void* stream) {
  if (B < 0 || N <= 0 || D <= 0) return RTF_E_ARG;
  if (B == 0) return 0;
  if (!d_W || !d_labels || !d_g0 || !d_gx) return RTF_E_ARG;
  SsmLogitParams P = {};
  P.W = d_W; P.labels = (const long long*)d_labels; P.g0 = const_cast<float*>(d_g0); P.B = B;
  P.N = N; P.D = D; P.gx = d_gx; P.gx_sb = gx_sb;
  ssm_true_gx_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
  RTF_CHECK_LAUNCH();
  return 0;
}
