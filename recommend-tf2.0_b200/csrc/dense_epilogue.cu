// Dense-layer backward epilogue (the MLP either side of the hot path).
//
// The GEMMs of the dense layers are library calls; what the framework adds around them in the
// backward is two more passes over the (B, N) gradient: the ReLU mask (threshold_backward) and
// the bias gradient (a column sum).  This kernel does both in ONE pass: g = gy * (y > 0) is
// written once and its column sums are accumulated on the fly, with the same deterministic
// two-stage order as rtf_colsum (256-row chunks, chunks combined in order).
#include "rtf_common.cuh"

namespace rtf {

constexpr int RB_ROWS = 128;       // rows per chunk up to B = 65 536 (512 chunks); larger batches take
constexpr int RB_MAX_CHUNKS = 512; // proportionally longer chunks so that stage 2 stays short

// CTA = lx column lanes (float4 each) x 128/lx row lanes; a row lane owns RB_ROWS/(128/lx)
// consecutive rows of the chunk and writes its own partial row, so narrow layers (N = 128: 32
// column lanes) still fill the CTA
__global__ void __launch_bounds__(128)
relu_bwd_colsum_stage1(const float* __restrict__ gy, const float* __restrict__ y, long long B,
                       int Ccols, int lx, int rb, float* __restrict__ g, float* __restrict__ partial) {
  const int tx = threadIdx.x % lx, ty = threadIdx.x / lx, ny = 128 / lx;
  const int c4 = (blockIdx.x * lx + tx) * 4;
  if (c4 >= Ccols) return;
  const int sub = rb / ny;
  const long long b0 = (long long)blockIdx.y * rb + (long long)ty * sub;
  const long long b1 = min(b0 + sub, B);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long b = b0; b < b1; ++b) {
    const float4 gv = *reinterpret_cast<const float4*>(gy + b * Ccols + c4);
    float4 o = gv;
    if (y) {
      const float4 yv = *reinterpret_cast<const float4*>(y + b * Ccols + c4);
      o.x = yv.x > 0.f ? gv.x : 0.f;
      o.y = yv.y > 0.f ? gv.y : 0.f;
      o.z = yv.z > 0.f ? gv.z : 0.f;
      o.w = yv.w > 0.f ? gv.w : 0.f;
      *reinterpret_cast<float4*>(g + b * Ccols + c4) = o;
    }
    acc = f4_add(acc, o);
  }
  *reinterpret_cast<float4*>(partial + ((long long)blockIdx.y * ny + ty) * Ccols + c4) = acc;
}

// 32 columns per CTA, 32 warps: warp g adds chunks g, g+32, g+64, ... (ascending) with 16 loads in
// flight, the 32 partial sums are then added in warp order — a fixed tree, so the result is
// reproducible.  (8 warps x 4 loads in flight took 11-39 us per layer on 512-2048 partial rows:
// pure load latency, 0.125 ms of the DLRM step.)
constexpr int RB2_WARPS = 32, RB2_UNROLL = 16;
__global__ void __launch_bounds__(RB2_WARPS * 32)
relu_bwd_colsum_stage2(const float* __restrict__ partial, int nchunks, int Ccols,
                       float* __restrict__ out) {
  __shared__ float part[RB2_WARPS][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (c < Ccols) {
    int k = grp;
    for (; k + (RB2_UNROLL - 1) * RB2_WARPS < nchunks; k += RB2_UNROLL * RB2_WARPS) {
      float a[RB2_UNROLL];
#pragma unroll
      for (int u = 0; u < RB2_UNROLL; ++u) a[u] = partial[(long long)(k + u * RB2_WARPS) * Ccols + c];
#pragma unroll
      for (int u = 0; u < RB2_UNROLL; ++u) acc = __fadd_rn(acc, a[u]);
    }
    for (; k < nchunks; k += RB2_WARPS) acc = __fadd_rn(acc, partial[(long long)k * Ccols + c]);
  }
  part[grp][lane] = acc;
  __syncthreads();
  if (grp == 0 && c < Ccols) {
    float t = part[0][lane];
#pragma unroll
    for (int g = 1; g < RB2_WARPS; ++g) t = __fadd_rn(t, part[g][lane]);
    out[c] = t;
  }
}

}  // namespace rtf

using namespace rtf;

static int relu_bwd_rb(long long B) {   // rows per chunk: 128 * ceil(B / 65 536)
  const long long per = (long long)RB_ROWS * RB_MAX_CHUNKS;
  return (int)(RB_ROWS * ((B + per - 1) / per > 0 ? (B + per - 1) / per : 1));
}

static int relu_bwd_lx(int cols) {   // column lanes per CTA: 32, 64 or 128
  const int v = cols / 4;
  return v <= 32 ? 32 : (v <= 64 ? 64 : 128);
}

extern "C" int rtf_relu_bwd_colsum_workspace(int64_t B, int cols, size_t* bytes) {
  if (!bytes || B < 0 || cols <= 0) return RTF_E_ARG;
  const size_t ny = 128 / relu_bwd_lx(cols);
  const int rb = relu_bwd_rb(B);
  *bytes = (size_t)((B + rb - 1) / rb + 1) * ny * (size_t)cols * 4;
  return 0;
}

// d_g = d_gy * (d_y > 0) (skipped when d_y is NULL: then only the column sums of d_gy are taken)
// d_colsum[c] = sum_b d_g[b, c].  Contiguous (B, cols) row-major, cols % 4 == 0, 16-byte aligned.
extern "C" int rtf_relu_bwd_colsum(const float* d_gy, const float* d_y, int64_t B, int cols,
                                   float* d_g, float* d_colsum, void* d_ws, void* stream) {
  if (B < 0 || cols <= 0 || !d_colsum) return RTF_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return (int)cudaMemsetAsync(d_colsum, 0, (size_t)cols * 4, st);
  if (!d_gy || !d_ws || (d_y && !d_g)) return RTF_E_ARG;
  if (cols % 4 || (uintptr_t)d_gy % 16 || (uintptr_t)d_y % 16 || (uintptr_t)d_g % 16 ||
      (uintptr_t)d_ws % 16)
    return RTF_E_ALIGN;
  const int rb = relu_bwd_rb(B);
  const int nchunks = (int)((B + rb - 1) / rb);
  const int lx = relu_bwd_lx(cols), ny = 128 / lx;
  dim3 g1((cols / 4 + lx - 1) / lx, nchunks);
  relu_bwd_colsum_stage1<<<g1, 128, 0, st>>>(d_gy, d_y, B, cols, lx, rb, d_g, (float*)d_ws);
  relu_bwd_colsum_stage2<<<(cols + 31) / 32, RB2_WARPS * 32, 0, st>>>((const float*)d_ws, nchunks * ny, cols,
                                                           d_colsum);
  RTF_CHECK_LAUNCH();
  return 0;
}
