// f3 — exact top-k inner-product retrieval (the serving half of the matching models).
//
// replaces: faiss.IndexFlatIP(d); index.add(item_embs); D, I = index.search(user_embs, k)
//           (src/match/fm/train.py:71-75, src/match/dssm/dssm_train.py:74-78): for every user
//           row the k items with the largest inner product, best first.
//
// Three stages, all on the caller's stream, no host sync:
//   1. scores: user (B,D) x item-tile^T (T,D) on the tcgen05 tensor cores through the library's
//      fp32-accurate GEMM (rtf_dense_gemm_nt: each fp32 operand split into 3 bf16 terms, 6
//      products, ~2e-7 relative) into an L2-sized (B, T) tile buffer;
//   2. topk_merge_tile: one warp per user row streams the tile with 128-bit loads and keeps its
//      running best C = 32 candidates (value desc, index asc): values above the row's current
//      32nd best are compacted into a shared-memory buffer by ballot and merged by a warp
//      bitonic sort — after the first tiles almost nothing passes the threshold, so the stage
//      runs at the speed of reading the tile;
//   3. topk_rescore: the 32 candidates of a row are re-scored EXACTLY (products of fp32 inputs
//      accumulated in fp64) and ranked by (score desc, index asc); the first k leave.  The
//      kernel also proves the result: if the approximate value of the weakest kept candidate is
//      not separated from the k-th exact score by more than 4x the largest |approx - exact| seen
//      in the row, bit 0 of *d_flag is set (more than 32 - k near-ties: re-run with larger k).
// So indices equal np.argsort of the fp64 scores (ties: lower index first) whenever the flag
// stays 0 — tests/test_topk_gpu.py checks this at N = 1 M.
#include <cfloat>

#include "rtf_common.cuh"

namespace rtf {

constexpr int TK_C = 32;        // candidates kept per row
constexpr int TK_BUF = 256;     // merge buffer (power of two >= TK_C + 31 + 128)
constexpr int TK_WARPS = 8;

__device__ __forceinline__ bool tk_before(float va, int ia, float vb, int ib) {
  return va > vb || (va == vb && ia < ib);
}

// sort TK_BUF (value, index) pairs in shared memory, best first; one warp
__device__ __forceinline__ void tk_bitonic(float* v, int* idx, int lane) {
  for (int k = 2; k <= TK_BUF; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < TK_BUF / 2; t += 32) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
        const int hi = lo | j;
        const bool up = (lo & k) == 0;  // ascending block = best first
        const float a = v[lo], b = v[hi];
        const int ia = idx[lo], ib = idx[hi];
        const bool swap = up ? tk_before(b, ib, a, ia) : tk_before(a, ia, b, ib);
        if (swap) {
          v[lo] = b; v[hi] = a;
          idx[lo] = ib; idx[hi] = ia;
        }
      }
      __syncwarp();
    }
  }
}

// scores (B, ld): columns [first, n) of this tile are new items with global index base + col
__global__ void __launch_bounds__(TK_WARPS * 32)
topk_merge_tile(const float* __restrict__ scores, long long ld, int first, int n, long long base,
                float* __restrict__ cval, int* __restrict__ cidx, long long B) {
  __shared__ float bv[TK_WARPS][TK_BUF];
  __shared__ int bi[TK_WARPS][TK_BUF];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * TK_WARPS + w;
  if (row >= B) return;
  float* v = bv[w];
  int* ix = bi[w];
  float cv = cval[row * TK_C + lane];
  int ci = cidx[row * TK_C + lane];
  float tau = __shfl_sync(0xffffffffu, cv, 31);
  int cnt = 0;
  const float* src = scores + row * ld;
  const int n4 = (n + 3) & ~3;  // the tile buffer is padded to a multiple of 4 columns
  for (int j0 = 0; j0 < n4; j0 += 128) {
    const int c0 = j0 + 4 * lane;
    float4 s = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    if (c0 < n4) s = ldg_nc_f4(src + c0);
    const float e[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u;
      const bool take = c >= first && c < n && e[u] > tau;
      const unsigned m = __ballot_sync(0xffffffffu, take);
      if (take) {
        const int pos = TK_C + cnt + __popc(m & ((1u << lane) - 1u));
        v[pos] = e[u];
        ix[pos] = (int)(base + c);
      }
      cnt += __popc(m);
    }
    if (cnt >= TK_BUF - TK_C - 128 || (j0 + 128 >= n4 && cnt > 0)) {
      v[lane] = cv;
      ix[lane] = ci;
      for (int t = TK_C + cnt + lane; t < TK_BUF; t += 32) {
        v[t] = -FLT_MAX;
        ix[t] = 0x7fffffff;
      }
      __syncwarp();
      tk_bitonic(v, ix, lane);
      cv = v[lane];
      ci = ix[lane];
      tau = __shfl_sync(0xffffffffu, cv, 31);
      cnt = 0;
      __syncwarp();
    }
  }
  cval[row * TK_C + lane] = cv;
  cidx[row * TK_C + lane] = ci;
}

__global__ void __launch_bounds__(256)
topk_init(float* __restrict__ cval, int* __restrict__ cidx, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    cval[i] = -FLT_MAX;
    cidx[i] = 0x7fffffff;
  }
}

// exact re-scoring of the 32 candidates of a row; one warp per row
__global__ void __launch_bounds__(TK_WARPS * 32)
topk_rescore(const float* __restrict__ users, long long u_ld, const float* __restrict__ items,
             long long i_ld, long long N, int D, const float* __restrict__ cval,
             const int* __restrict__ cidx, int k, long long* __restrict__ out_idx,
             float* __restrict__ out_score, int32_t* __restrict__ flag, long long B) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * TK_WARPS + w;
  if (row >= B) return;
  const float approx = cval[row * TK_C + lane];
  const int idx = cidx[row * TK_C + lane];
  const bool valid = idx != 0x7fffffff;
  double exact = -DBL_MAX;
  if (valid) {
    const float* u = users + row * u_ld;
    const float* it = items + (long long)idx * i_ld;
    double acc = 0.0;
    for (int d = 0; d < D; ++d) acc = fma((double)__ldg(u + d), (double)__ldg(it + d), acc);
    exact = acc;
  }
  // rank = number of candidates that come before this one (score desc, index asc)
  int rank = 0;
  for (int o = 0; o < 32; ++o) {
    const double eo = __shfl_sync(0xffffffffu, exact, o);
    const int io = __shfl_sync(0xffffffffu, idx, o);
    if (eo > exact || (eo == exact && (io < idx || (io == idx && o < lane)))) ++rank;
  }
  if (rank < k) {
    out_idx[row * k + rank] = valid ? (long long)idx : -1;
    out_score[row * k + rank] = valid ? (float)exact : -INFINITY;
  }
  // proof of exactness: the weakest kept approximate value must sit clearly below the k-th
  // exact score (anything never kept scored <= it, up to the GEMM's error)
  float dlt = valid ? fabsf(approx - (float)exact) : 0.f;
  dlt = warp_max(dlt);
  const float weakest = __shfl_sync(0xffffffffu, approx, 31);
  const bool full = __shfl_sync(0xffffffffu, (int)valid, 31) != 0;  // more than 32 items seen
  const unsigned kth_mask = __ballot_sync(0xffffffffu, rank == k - 1);
  if (kth_mask && full) {
    const double kth = __shfl_sync(0xffffffffu, exact, __ffs(kth_mask) - 1);
    if (lane == 0 && (double)weakest + 4.0 * (double)dlt + 1e-30 >= kth && N > TK_C) atomicOr(flag, 1);
  }
}

}  // namespace rtf

using namespace rtf;

extern "C" int rtf_dense_gemm_nt_workspace(int M, int N, int K, int batch, size_t* bytes);
extern "C" int rtf_dense_gemm_nt(const float* d_a, int64_t lda, int64_t stride_a, const float* d_b,
                                 int64_t ldb, int64_t stride_b, const float* d_bias, int relu,
                                 float* d_out, int64_t ldd, int64_t stride_d, int M, int N, int K,
                                 int batch, void* d_ws, size_t ws_bytes, void* stream);

static size_t tk_align(size_t x) { return (x + 255) / 256 * 256; }

static int tk_tile(int64_t B, int64_t N) {
  // item-tile width: the (B, T) score tile should stay L2-resident (~64 MB) between the GEMM
  // that writes it and the merge that reads it
  int64_t t = (64ll << 20) / (4 * (B > 0 ? B : 1));
  t = t / 128 * 128;
  if (t < 512) t = 512;
  if (t > 65536) t = 65536;
  if (t > N) t = (N + 3) / 4 * 4;
  return (int)t;
}

extern "C" int rtf_topk_ip_workspace(int64_t B, int64_t N, int D, int k, size_t* bytes) {
  if (!bytes || B < 0 || N <= 0 || D <= 0 || k <= 0) return RTF_E_ARG;
  if (B > 0x7fffffff || N > 0x7ffffffe) return RTF_E_RANGE;
  const int T = tk_tile(B, N);
  size_t g = 0;
  int rc = rtf_dense_gemm_nt_workspace((int)(B > 0 ? B : 1), T, D, 1, &g);
  if (rc) return rc;
  *bytes = tk_align((size_t)(B > 0 ? B : 1) * T * 4) + tk_align(g) +
           tk_align((size_t)(B > 0 ? B : 1) * TK_C * 4) * 2 + 256;
  return 0;
}

extern "C" int rtf_topk_ip(const float* d_users, int64_t u_ld, int64_t B, const float* d_items,
                           int64_t i_ld, int64_t N, int D, int k, int64_t* d_out_idx,
                           float* d_out_score, int32_t* d_flag, void* d_ws, size_t ws_bytes,
                           void* stream) {
  if (B < 0 || N <= 0 || D <= 0 || k <= 0) return RTF_E_ARG;
  if (k > TK_C / 2 || B > 0x7fffffff || N > 0x7ffffffe) return RTF_E_RANGE;
  if (B == 0) return 0;
  if (!d_users || !d_items || !d_out_idx || !d_out_score || !d_flag || !d_ws) return RTF_E_ARG;
  if (D % 4 || u_ld % 4 || i_ld % 4 || (uintptr_t)d_users % 16 || (uintptr_t)d_items % 16 ||
      (uintptr_t)d_ws % 256)
    return RTF_E_ALIGN;
  size_t need = 0;
  int rc = rtf_topk_ip_workspace(B, N, D, k, &need);
  if (rc) return rc;
  if (ws_bytes < need) return RTF_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int T = tk_tile(B, N);
  size_t g = 0;
  rtf_dense_gemm_nt_workspace((int)B, T, D, 1, &g);
  char* ws = (char*)d_ws;
  float* scores = (float*)ws;                  ws += tk_align((size_t)B * T * 4);
  void* gws = ws;                              ws += tk_align(g);
  float* cval = (float*)ws;                    ws += tk_align((size_t)B * TK_C * 4);
  int* cidx = (int*)ws;
  const long long nc = B * TK_C;
  topk_init<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(cval, cidx, nc);
  RTF_CHECK_LAUNCH();
  const unsigned rb = (unsigned)((B + TK_WARPS - 1) / TK_WARPS);
  // one GEMM + merge over items [start, start + n), of which the columns >= first are new
  auto run_tile = [&](int64_t start, int64_t n, int first) -> int {
    int r = rtf_dense_gemm_nt(d_users, u_ld, 0, d_items + start * i_ld, i_ld, 0, nullptr, 0, scores, T,
                              0, (int)B, (int)n, D, 1, gws, tk_align(g), stream);
    if (r) return r;
    topk_merge_tile<<<rb, TK_WARPS * 32, 0, st>>>(scores, T, first, (int)n, start, cval, cidx, B);
    RTF_CHECK_LAUNCH();
    return 0;
  };
  for (int64_t i0 = 0; i0 < N; i0 += T) {
    // the GEMM's N must be a multiple of 4: a ragged last tile is cut into its multiple-of-4 part
    // and the LAST FOUR items of the table, whose columns already covered are skipped (`first`)
    const int64_t n = N - i0 < T ? N - i0 : T;
    const int64_t n_main = n - n % 4;
    if (n_main > 0) {
      rc = run_tile(i0, n_main, 0);
      if (rc) return rc;
    }
    if (n % 4) {
      if (N < 4) return RTF_E_RANGE;
      rc = run_tile(N - 4, 4, (int)(4 - n % 4));
      if (rc) return rc;
    }
  }
  topk_rescore<<<rb, TK_WARPS * 32, 0, st>>>(d_users, u_ld, d_items, i_ld, N, D, cval, cidx, k,
                                             (long long*)d_out_idx, d_out_score, d_flag, B);
  RTF_CHECK_LAUNCH();
  return 0;
}
