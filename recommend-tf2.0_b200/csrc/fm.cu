// K3 — factorisation-machine interactions.
//
// K3a  FM layer: ctr.layers.modules.FM.call (src/ctr/layers/modules.py:57-72)
//        first  = reduce_sum(first_inputs @ w)                  (a scalar over the WHOLE batch)
//        second = 0.5 * sum_axis1((sum_axis1 x)^2 - sum_axis1 x^2)
//        out    = reshape(first + second, (-1, 1))
//      with second_inputs (B, F, D) — the 2-D tensor DeepFM passes (deep_fm/model.py:58-59)
//      is F = M, D = 1 — giving out (B*D, 1).  Flags select the paper-correct variants
//      (per-sample first order; also summing the second order over D).
// K3b  FM model in gather form: ctr.fm.model.FM.call (src/ctr/fm/model.py:34-53) builds a
//      dense one-hot (B, M) matrix and multiplies it by w (M,1) and V^T (M,k).  one_hot @ W is
//      a row gather, so each sample only needs its 13 + 26 rows of [V^T | w].
//
// All HBM-bound, a few FLOPs per byte: one lane-group per sample, warp-shuffle reductions,
// batch-wide sums (first-order scalar, weight gradients) through a deterministic two-stage
// column reduction.
#include "rtf_common.cuh"

namespace rtf {

// ---------------------------------------------------------------- deterministic column sum
// out[c] = sum_b scale(b) * x[b, c]; stage 1: chunks of COLSUM_ROWS rows, stage 2: chunks in order
constexpr int COLSUM_ROWS = 256;

__global__ void __launch_bounds__(128)
colsum_stage1(const float* __restrict__ x, long long x_sb, const float* __restrict__ rowscale,
              long long B, int Ccols, float* __restrict__ partial) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Ccols) return;
  const long long b0 = (long long)blockIdx.y * COLSUM_ROWS;
  const long long b1 = min(b0 + COLSUM_ROWS, B);
  float acc = 0.f;
  for (long long b = b0; b < b1; ++b) {
    const float v = x[b * x_sb + c];
    acc = __fadd_rn(acc, rowscale ? __fmul_rn(rowscale[b], v) : v);
  }
  partial[(long long)blockIdx.y * Ccols + c] = acc;
}
__global__ void __launch_bounds__(128)
colsum_stage2(const float* __restrict__ partial, int nchunks, int Ccols, float* __restrict__ out,
              int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Ccols) return;
  float acc = 0.f;
  for (int k = 0; k < nchunks; ++k) acc = __fadd_rn(acc, partial[(long long)k * Ccols + c]);
  out[c] = accumulate ? __fadd_rn(out[c], acc) : acc;
}

static int colsum_launch(const float* x, long long x_sb, const float* rowscale, long long B,
                         int Ccols, float* out, float* ws, cudaStream_t st) {
  const int nchunks = (int)((B + COLSUM_ROWS - 1) / COLSUM_ROWS);
  dim3 g1((Ccols + 127) / 128, nchunks);
  colsum_stage1<<<g1, 128, 0, st>>>(x, x_sb, rowscale, B, Ccols, ws);
  colsum_stage2<<<(Ccols + 127) / 128, 128, 0, st>>>(ws, nchunks, Ccols, out, 0);
  RTF_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------- K3a FM layer
struct FmLayerParams {
  const float* first;  // (B, P1)
  long long first_sb;
  const float* w;       // (P1)
  const float* second;  // (B, F, D) contiguous per sample, sample stride second_sb
  long long second_sb;
  long long B;
  int P1, F, D;
  int first_batch_scalar, sum_d;
  float* out;      // (B, D) or (B) when sum_d
  float* first_b;  // workspace (B): per-sample first-order dot
};

// one warp per sample
__global__ void __launch_bounds__(256) fm_layer_fwd(const __grid_constant__ FmLayerParams P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= P.B) return;
  float fo = 0.f;
  const float* fr = P.first + b * P.first_sb;
  for (int j = lane; j < P.P1; j += 32) fo = fmaf(fr[j], __ldg(P.w + j), fo);
  fo = warp_sum(fo);
  if (lane == 0) P.first_b[b] = fo;
  const float add = P.first_batch_scalar ? 0.f : fo;
  const float* x = P.second + b * P.second_sb;
  const int D = P.D, F = P.F;
  float tot = 0.f;  // for sum_d
  if (D < 32 && (32 % D) == 0) {
    // lanes stride the flat (F*D) row; a lane always sees the same d = lane % D
    float s = 0.f, q = 0.f;
    for (int e = lane; e < F * D; e += 32) {
      const float v = x[e];
      s += v;
      q = fmaf(v, v, q);
    }
    for (int o = 16; o >= D; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    const float sec = 0.5f * (s * s - q);
    if (P.sum_d) {
      float t = lane < D ? sec : 0.f;
      tot = warp_sum(t);
    } else if (lane < D) {
      P.out[b * D + lane] = add + sec;
    }
  } else {
    for (int d = lane; d < D; d += 32) {
      float s = 0.f, q = 0.f;
      for (int f = 0; f < F; ++f) {
        const float v = x[f * D + d];
        s += v;
        q = fmaf(v, v, q);
      }
      const float sec = 0.5f * (s * s - q);
      if (P.sum_d) tot += sec;
      else P.out[b * D + d] = add + sec;
    }
    if (P.sum_d) tot = warp_sum(tot);
  }
  if (P.sum_d && lane == 0) P.out[b] = add + tot;
}

// single CTA: total = sum_b v[b] in a fixed order; optionally add it to every out element
__global__ void __launch_bounds__(1024)
reduce_all(const float* __restrict__ v, long long n, float* __restrict__ total) {
  __shared__ float sh[1024];
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += 1024) acc = __fadd_rn(acc, v[i]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] = __fadd_rn(sh[threadIdx.x], sh[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = sh[0];
}
__global__ void __launch_bounds__(256)
add_scalar(float* __restrict__ out, long long n, const float* __restrict__ total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += *total;
}

struct FmLayerBwdParams {
  FmLayerParams f;
  const float* gout;    // (B, D) or (B)
  const float* gtotal;  // device scalar: sum of all gout (first_batch_scalar)
  float* gfirst;        // (B, P1)
  long long gfirst_sb;
  float* gsecond;  // (B, F, D)
  long long gsecond_sb;
  float* gfirst_b;  // workspace (B): dL/d first_b
};

__global__ void __launch_bounds__(256) fm_layer_bwd(const __grid_constant__ FmLayerBwdParams Q) {
  const FmLayerParams& P = Q.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= P.B) return;
  const int D = P.D, F = P.F;
  const float* g = Q.gout + (P.sum_d ? b : b * D);
  float gfb;
  if (P.first_batch_scalar) {
    gfb = *Q.gtotal;
  } else if (P.sum_d) {
    gfb = g[0];
  } else {
    float t = 0.f;
    for (int d = lane; d < D; d += 32) t += g[d];
    gfb = warp_sum(t);
  }
  if (lane == 0) Q.gfirst_b[b] = gfb;
  if (Q.gfirst) {
    float* gf = Q.gfirst + b * Q.gfirst_sb;
    for (int j = lane; j < P.P1; j += 32) gf[j] = gfb * __ldg(P.w + j);
  }
  if (!Q.gsecond) return;
  const float* x = P.second + b * P.second_sb;
  float* gx = Q.gsecond + b * Q.gsecond_sb;
  if (D < 32 && (32 % D) == 0) {
    float s = 0.f;
    for (int e = lane; e < F * D; e += 32) s += x[e];
    for (int o = 16; o >= D; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float gg = P.sum_d ? g[0] : g[lane % D];
    for (int e = lane; e < F * D; e += 32) gx[e] = gg * (s - x[e]);
  } else {
    for (int d = lane; d < D; d += 32) {
      float s = 0.f;
      for (int f = 0; f < F; ++f) s += x[f * D + d];
      const float gg = P.sum_d ? g[0] : g[d];
      for (int f = 0; f < F; ++f) gx[f * D + d] = gg * (s - x[f * D + d]);
    }
  }
}

// ---------------------------------------------------------------- K3b FM model, gather form
struct FmGatherParams {
  const float* table[RTF_MAX_FIELDS];  // (N_f, kp): columns [0,k) = V^T row, column k = w, rest 0
  long long rows[RTF_MAX_FIELDS];
  int n_fields, n_dense, k, kp;
  const float* dense_table;  // (n_dense, kp)
  const float* dense;        // (B, n_dense)
  long long dense_sb;
  const void* ids;
  long long ids_sb, ids_sf;
  const float* w0;  // device scalar
  long long B;
  float* out;  // (B) sigmoid(z)
  float* A;    // (B, kp) saved sum_i x_i * R_i[c]
  // backward
  const float* gout;  // (B)
  float* gsparse;     // (B, n_fields*kp)
  float* gdense_rows;  // (B, n_dense*kp)
  float* dz;           // (B)
  int32_t* err;
};

template <typename IdT, int G, bool BWD>
__global__ void __launch_bounds__(256) fm_gather_kernel(const __grid_constant__ FmGatherParams P) {
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int c = threadIdx.x % G;
  if (b >= P.B) return;  // whole groups exit together (blockDim % G == 0)
  const int g0 = (int)(threadIdx.x & 31) / G * G;
  const uint32_t gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << g0);
  const int kp = P.kp, k = P.k;
  const int nfeat = P.n_dense + P.n_fields;
  const bool live = c < kp;
  float A = 0.f, Qs = 0.f, dz = 0.f;
  if (BWD) {
    const float o = P.out[b];
    dz = P.gout[b] * o * (1.f - o);
    A = live ? P.A[b * kp + c] : 0.f;
    if (c == 0) P.dz[b] = dz;
  }
  constexpr int U = 4;
  for (int i0 = 0; i0 < nfeat; i0 += U) {
    const float* row[U];
    float x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u;
      row[u] = nullptr;
      x[u] = 0.f;
      if (i < P.n_dense) {
        row[u] = P.dense_table + (long long)i * kp;
        x[u] = P.dense[b * P.dense_sb + i];
      } else if (i < nfeat) {
        const int f = i - P.n_dense;
        const long long id = load_id((const IdT*)P.ids, b * P.ids_sb + (long long)f * P.ids_sf,
                                     P.rows[f], P.err);
        if (id >= 0) row[u] = P.table[f] + id * kp;
        x[u] = 1.f;
      }
    }
    float r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = (row[u] && live) ? __ldg(row[u] + c) : 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u;
      if (i >= nfeat) continue;
      if (!BWD) {
        A = fmaf(x[u], r[u], A);
        Qs = fmaf(x[u] * x[u], r[u] * r[u], Qs);
      } else if (live) {
        float g = 0.f;
        if (c < k) g = dz * (x[u] * A - x[u] * x[u] * r[u]);
        else if (c == k) g = dz * x[u];
        if (i < P.n_dense) P.gdense_rows[b * ((long long)P.n_dense * kp) + (long long)i * kp + c] = g;
        else P.gsparse[b * ((long long)P.n_fields * kp) + (long long)(i - P.n_dense) * kp + c] = g;
      }
    }
  }
  if (!BWD) {
    float t = c < k ? (A * A - Qs) : 0.f;
    float lin = c == k ? A : 0.f;
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      t += __shfl_xor_sync(gmask, t, o);
      lin += __shfl_xor_sync(gmask, lin, o);
    }
    if (live) P.A[b * kp + c] = A;
    if (c == 0) {
      const float z = *P.w0 + lin + 0.5f * t;
      P.out[b] = 1.f / (1.f + expf(-z));
    }
  }
}

template <typename IdT, bool BWD>
static int fm_gather_launch(const FmGatherParams& P, cudaStream_t st) {
  int G = 4;
  while (G < P.kp) G <<= 1;
  if (G > 32) return RTF_E_RANGE;
  const long long threads = P.B * G;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  switch (G) {
    case 4: fm_gather_kernel<IdT, 4, BWD><<<blocks, 256, 0, st>>>(P); break;
    case 8: fm_gather_kernel<IdT, 8, BWD><<<blocks, 256, 0, st>>>(P); break;
    case 16: fm_gather_kernel<IdT, 16, BWD><<<blocks, 256, 0, st>>>(P); break;
    default: fm_gather_kernel<IdT, 32, BWD><<<blocks, 256, 0, st>>>(P); break;
  }
  RTF_CHECK_LAUNCH();
  return 0;
}

}  // namespace rtf

using namespace rtf;

extern "C" int rtf_colsum_workspace(int64_t B, int cols, size_t* bytes) {
  if (!bytes || B < 0 || cols <= 0) return RTF_E_ARG;
  *bytes = (size_t)((B + COLSUM_ROWS - 1) / COLSUM_ROWS + 1) * (size_t)cols * 4;
  return 0;
}

extern "C" int rtf_colsum(const float* d_x, int64_t x_sb, const float* d_rowscale, int64_t B,
                          int cols, float* d_out, void* d_ws, void* stream) {
  if (B < 0 || cols <= 0 || !d_out) return RTF_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return (int)cudaMemsetAsync(d_out, 0, (size_t)cols * 4, st);
  if (!d_x || !d_ws) return RTF_E_ARG;
  return colsum_launch(d_x, x_sb, d_rowscale, B, cols, d_out, (float*)d_ws, st);
}

extern "C" int rtf_fm_layer_workspace(int64_t B, int P1, size_t* bytes) {
  if (!bytes || B < 0 || P1 <= 0) return RTF_E_ARG;
  size_t cs = 0;
  rtf_colsum_workspace(B, P1, &cs);
  *bytes = (size_t)(2 * B + 8) * 4 + cs;
  return 0;
}

extern "C" int rtf_fm_layer_fwd(const float* d_first, int64_t first_sb, const float* d_w, int P1,
                                const float* d_second, int64_t second_sb, int F, int D, int64_t B,
                                int first_batch_scalar, int sum_d, float* d_out, void* d_ws,
                                void* stream) {
  if (B < 0 || P1 <= 0 || F <= 0 || D <= 0) return RTF_E_ARG;
  if (B == 0) return 0;
  if (!d_first || !d_w || !d_second || !d_out || !d_ws) return RTF_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  FmLayerParams P;
  P.first = d_first; P.first_sb = first_sb; P.w = d_w; P.second = d_second;
  P.second_sb = second_sb; P.B = B; P.P1 = P1; P.F = F; P.D = D;
  P.first_batch_scalar = first_batch_scalar; P.sum_d = sum_d; P.out = d_out;
  float* ws = (float*)d_ws;
  P.first_b = ws;            // [B]
  float* total = ws + 2 * B;  // scalar
  fm_layer_fwd<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(P);
  if (first_batch_scalar) {
    reduce_all<<<1, 1024, 0, st>>>(P.first_b, B, total);
    const long long n = sum_d ? B : B * D;
    add_scalar<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_out, n, total);
  }
  RTF_CHECK_LAUNCH();
  return 0;
}

extern "C" int rtf_fm_layer_bwd(const float* d_first, int64_t first_sb, const float* d_w, int P1,
                                const float* d_second, int64_t second_sb, int F, int D, int64_t B,
                                int first_batch_scalar, int sum_d, const float* d_gout,
                                float* d_gfirst, int64_t gfirst_sb, float* d_gw, float* d_gsecond,
                                int64_t gsecond_sb, void* d_ws, void* stream) {
  if (B < 0 || P1 <= 0 || F <= 0 || D <= 0) return RTF_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    if (d_gw) return (int)cudaMemsetAsync(d_gw, 0, (size_t)P1 * 4, st);
    return 0;
  }
  if (!d_first || !d_w || !d_second || !d_gout || !d_ws) return RTF_E_ARG;
  FmLayerBwdParams Q;
  Q.f.first = d_first; Q.f.first_sb = first_sb; Q.f.w = d_w; Q.f.second = d_second;
  Q.f.second_sb = second_sb; Q.f.B = B; Q.f.P1 = P1; Q.f.F = F; Q.f.D = D;
  Q.f.first_batch_scalar = first_batch_scalar; Q.f.sum_d = sum_d; Q.f.out = nullptr;
  float* ws = (float*)d_ws;
  Q.f.first_b = nullptr;
  Q.gfirst_b = ws + B;  // [B]
  float* gtotal = ws + 2 * B + 1;
  float* cs_ws = ws + 2 * B + 8;
  Q.gout = d_gout; Q.gtotal = gtotal; Q.gfirst = d_gfirst; Q.gfirst_sb = gfirst_sb;
  Q.gsecond = d_gsecond; Q.gsecond_sb = gsecond_sb;
  if (first_batch_scalar) reduce_all<<<1, 1024, 0, st>>>(d_gout, sum_d ? B : B * D, gtotal);
  fm_layer_bwd<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(Q);
  RTF_CHECK_LAUNCH();
  if (d_gw) return colsum_launch(d_first, first_sb, Q.gfirst_b, B, P1, d_gw, cs_ws, st);
  return 0;
}

static int fm_gather_fill(FmGatherParams& P, const float* const* tables, const int64_t* rows,
                          int n_fields, int k, int kp, const float* dense_table, int n_dense,
                          const float* dense, int64_t dense_sb, const void* ids, int64_t B,
                          int64_t sb, int64_t sf) {
  if (n_fields < 0 || n_fields > RTF_MAX_FIELDS || n_dense < 0 || k <= 0 || kp <= k) return RTF_E_ARG;
  if (n_fields + n_dense == 0) return RTF_E_ARG;
  if (kp > 32) return RTF_E_RANGE;
  if (n_fields && (!tables || !rows || !ids)) return RTF_E_ARG;
  if (n_dense && (!dense_table || !dense)) return RTF_E_ARG;
  for (int f = 0; f < n_fields; ++f) {
    if (!tables[f] || rows[f] <= 0) return RTF_E_ARG;
    P.table[f] = tables[f];
    P.rows[f] = rows[f];
  }
  P.n_fields = n_fields; P.n_dense = n_dense; P.k = k; P.kp = kp; P.dense_table = dense_table;
  P.dense = dense; P.dense_sb = dense_sb; P.ids = ids; P.ids_sb = sb; P.ids_sf = sf; P.B = B;
  return 0;
}

extern "C" int rtf_fm_gather_fwd(const float* const* tables, const int64_t* rows, int n_fields,
                                 int k, int kp, const float* d_dense_table, int n_dense,
                                 const float* d_dense, int64_t dense_sb, const void* d_ids,
                                 int ids_i64, int64_t B, int64_t ids_sb, int64_t ids_sf,
                                 const float* d_w0, float* d_out, float* d_A, int32_t* d_err,
                                 void* stream) {
  if (B < 0) return RTF_E_ARG;
  if (B == 0) return 0;
  if (!d_w0 || !d_out || !d_A) return RTF_E_ARG;
  FmGatherParams P = {};
  int rc = fm_gather_fill(P, tables, rows, n_fields, k, kp, d_dense_table, n_dense, d_dense,
                          dense_sb, d_ids, B, ids_sb, ids_sf);
  if (rc) return rc;
  P.w0 = d_w0; P.out = d_out; P.A = d_A; P.err = d_err;
  cudaStream_t st = (cudaStream_t)stream;
  return ids_i64 ? fm_gather_launch<int64_t, false>(P, st) : fm_gather_launch<int32_t, false>(P, st);
}

extern "C" int rtf_fm_gather_bwd(const float* const* tables, const int64_t* rows, int n_fields,
                                 int k, int kp, const float* d_dense_table, int n_dense,
                                 const float* d_dense, int64_t dense_sb, const void* d_ids,
                                 int ids_i64, int64_t B, int64_t ids_sb, int64_t ids_sf,
                                 const float* d_out, const float* d_A, const float* d_gout,
                                 float* d_gsparse, float* d_gdense_rows, float* d_dz,
                                 void* stream) {
  if (B < 0) return RTF_E_ARG;
  if (B == 0) return 0;
  if (!d_out || !d_A || !d_gout || !d_dz) return RTF_E_ARG;
  if ((n_fields && !d_gsparse) || (n_dense && !d_gdense_rows)) return RTF_E_ARG;
  FmGatherParams P = {};
  int rc = fm_gather_fill(P, tables, rows, n_fields, k, kp, d_dense_table, n_dense, d_dense,
                          dense_sb, d_ids, B, ids_sb, ids_sf);
  if (rc) return rc;
  P.out = const_cast<float*>(d_out); P.A = const_cast<float*>(d_A); P.gout = d_gout;
  P.gsparse = d_gsparse; P.gdense_rows = d_gdense_rows; P.dz = d_dz;
  cudaStream_t st = (cudaStream_t)stream;
  return ids_i64 ? fm_gather_launch<int64_t, true>(P, st) : fm_gather_launch<int32_t, true>(P, st);
}
