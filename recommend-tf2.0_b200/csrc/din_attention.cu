// K5 — DIN local activation unit: ctr.layers.modules.AttentionLayer.call
// (src/ctr/layers/modules.py:149-175) with hidden_unit = 1 (the only width its reshape at :159
// admits).
//
//   info = concat([q, k, q-k, q*k], -1)            (B, L, 4d)   — q tiled over L (:150-154)
//   s    = act(info @ W + b)                       (B, L)       — Dense(1, activation) (:157)
//   s    = where(mask == 0, -2^32+1, s)            mask not a tensor => ALL positions padded
//   a    = softmax(s)  (no 1/sqrt(d) scaling)      (:169)
//   out  = a @ v                                   (B, d)
//
// The tiled q and the (B,L,4d) `info` tensor are never built:
//   info_l . W = (W1+W3).q + (W2-W3 + W4*q).k_l = c + u.k_l
// so a sample needs one pass over its k/v rows.  One 4-warp CTA per sample; its (L,d) key tile
// (and value tile when v is a different tensor) is staged in shared memory by a single TMA bulk
// copy (SASS UBLKCP) on the CTA's mbarrier; rows are split across the warps, scores use
// lane-groups over d with shuffle reductions (conflict-free shared reads), the softmax and the
// a.v product run out of shared memory.
// HBM-bound: L*d*4 bytes in per sample (x2 if v != k), d*4 out.
#include "rtf_common.cuh"

namespace rtf {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_TANH = 3 };

struct DinParams {
  const float* q; long long q_sb;     // (B, d)
  const float* k; long long k_sb;     // (B, L, d), rows contiguous
  const float* v; long long v_sb;     // (B, L, d) — may alias k
  const float* mask; long long m_sb;  // (B, L) or null (=> every score padded, as the source)
  const float* W;                     // (4d): Dense(1) kernel, rows [q | k | q-k | q*k]
  const float* bias;                  // (1)
  int act;
  long long B;
  int L, d;
  float* out; long long o_sb;         // (B, d)
  // backward
  const float* gout; long long go_sb;
  float* gq; long long gq_sb;
  float* gk; long long gk_sb;
  float* gv; long long gv_sb;
  float* gw_rows;                     // (B, 4d+1): per-sample dL/d[W | bias]
};

__device__ __forceinline__ float act_fwd(int act, float z) {
  switch (act) {
    case ACT_RELU: return fmaxf(z, 0.f);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    case ACT_TANH: return tanhf(z);
    default: return z;
  }
}
__device__ __forceinline__ float act_bwd(int act, float z, float y) {
  switch (act) {
    case ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case ACT_SIGMOID: return y * (1.f - y);
    case ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}


// dot(vec[d], tile[l][d]) for the rows l = first, first+step, ... of this warp; lane-groups of
// g lanes split d (conflict-free shared reads), shuffle reduction inside the group
__device__ __forceinline__ void rows_dot(const float* __restrict__ vec, const float* __restrict__ tile,
                                         int L, int d, int g, int lane, int warp, int nwarps,
                                         float* __restrict__ dst) {
  const int rpi = 32 / g;  // rows per warp iteration
  const int lg = lane % g, lr = lane / g;
  for (int l0 = warp * rpi; l0 < L; l0 += nwarps * rpi) {
    const int l = l0 + lr;
    float acc = 0.f;
    if (l < L)
      for (int c = lg * 4; c < d; c += g * 4) {
        const float4 a = *reinterpret_cast<const float4*>(vec + c);
        const float4 b = *reinterpret_cast<const float4*>(tile + (long long)l * d + c);
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
        acc = fmaf(a.z, b.z, acc);
        acc = fmaf(a.w, b.w, acc);
      }
    for (int o = g >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (l < L && lg == 0) dst[l] = acc;
  }
}

// One CTA (DIN_WARPS warps) per sample, persistent over samples.  The (L,d) key tile (and value
// tile when distinct) arrives by one TMA bulk copy per tile on the CTA's mbarrier; rows are split
// across warps for the score / dz / output passes, per-warp partial d-vectors are combined in
// warp order (deterministic).
template <bool BWD, int DIN_WARPS>
__global__ void __launch_bounds__(DIN_WARPS * 32) din_kernel(const __grid_constant__ DinParams P,
                                                             int kv_same) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = P.L, d = P.d;
  const int Lp = (L + 3) & ~3;
  // one tile stage per CTA: more resident CTAs per SM beat double buffering here (measured:
  // the per-sample reduction, not the tile latency, is what limits a CTA)
  const int stage_floats = (kv_same ? 1 : 2) * L * d;
  float* u = smem + stage_floats;                       // [d]
  float* qs = u + d;                                    // [d]
  float* gs = qs + d;                                   // [d]   (bwd) gout
  float* part = gs + d;                                 // [DIN_WARPS][d] per-warp partial vectors
  float* zb = part + DIN_WARPS * d;                     // [Lp]  pre-activation z
  float* ab = zb + Lp;                                  // [Lp]  attention weights
  float* db = ab + Lp;                                  // [Lp]  (bwd) da -> dz
  uint64_t* bars = reinterpret_cast<uint64_t*>(db + Lp);  // [2]
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  int g = 1;
  while (g * 4 < d && g < 32) g <<= 1;
  const unsigned tile_bytes = (unsigned)L * d * 4u;
  uint32_t parity[2] = {0, 0};
  auto issue = [&](long long bb, int st) {
    if (threadIdx.x == 0) {
      float* kd = smem;
      mbar_expect_tx(&bars[st], kv_same ? tile_bytes : 2 * tile_bytes);
      bulk_g2s(kd, P.k + bb * P.k_sb, tile_bytes, &bars[st]);
      if (!kv_same) bulk_g2s(kd + L * d, P.v + bb * P.v_sb, tile_bytes, &bars[st]);
    }
  };
  long long b = blockIdx.x;
  if (b < P.B) issue(b, 0);
  for (; b < P.B; b += gridDim.x) {
    const int st = 0;
    const float* kt = smem;
    const float* vt = kv_same ? kt : kt + L * d;
    // u = (W2 - W3) + W4*q ; c = (W1 + W3).q + bias   (every warp computes c; warp 0 stores u, q)
    float cpart = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float qv = P.q[b * P.q_sb + c];
      const float w1 = __ldg(P.W + c), w2 = __ldg(P.W + d + c), w3 = __ldg(P.W + 2 * d + c),
                  w4 = __ldg(P.W + 3 * d + c);
      if (warp == 0) {
        qs[c] = qv;
        u[c] = (w2 - w3) + w4 * qv;
        if (BWD) gs[c] = P.gout[b * P.go_sb + c];
      }
      cpart = fmaf(w1 + w3, qv, cpart);
    }
    const float cc = warp_sum(cpart) + __ldg(P.bias);
    __syncthreads();
    mbar_wait(&bars[st], parity[st]);
    parity[st] ^= 1;
    rows_dot(u, kt, L, d, g, lane, warp, DIN_WARPS, zb);
    __syncthreads();
    // activation + mask + softmax statistics: every warp scans all L (L is small), each warp
    // then writes the weights of its own strided rows
    float m = -INFINITY;
    for (int l = lane; l < L; l += 32) {
      const float s = act_fwd(P.act, zb[l] + cc);
      const bool keep = P.mask && P.mask[b * P.m_sb + l] != 0.f;
      m = fmaxf(m, keep ? s : kPadLogit);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float s = act_fwd(P.act, zb[l] + cc);
      const bool keep = P.mask && P.mask[b * P.m_sb + l] != 0.f;
      sum += expf((keep ? s : kPadLogit) - m);
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int l = warp * 32 + lane; l < L; l += DIN_WARPS * 32) {
      const float s = act_fwd(P.act, zb[l] + cc);
      const bool keep = P.mask && P.mask[b * P.m_sb + l] != 0.f;
      ab[l] = expf((keep ? s : kPadLogit) - m) * inv;
    }
    __syncthreads();
    if (!BWD) {
      // out = a @ v : rows split across warps, partial vectors combined in warp order
      for (int c = lane * 4; c < d; c += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = warp; l < L; l += DIN_WARPS) {
          const float a = ab[l];
          const float4 vv = *reinterpret_cast<const float4*>(vt + (long long)l * d + c);
          acc.x = fmaf(a, vv.x, acc.x);
          acc.y = fmaf(a, vv.y, acc.y);
          acc.z = fmaf(a, vv.z, acc.z);
          acc.w = fmaf(a, vv.w, acc.w);
        }
        *reinterpret_cast<float4*>(part + warp * d + c) = acc;
      }
      __syncthreads();
      for (int c = threadIdx.x; c < d; c += DIN_WARPS * 32) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < DIN_WARPS; ++w) acc += part[w * d + c];
        P.out[b * P.o_sb + c] = acc;
      }
    } else {
      rows_dot(gs, vt, L, d, g, lane, warp, DIN_WARPS, db);  // da_l = gout . v_l
      __syncthreads();
      float dotp = 0.f;
      for (int l = lane; l < L; l += 32) dotp = fmaf(ab[l], db[l], dotp);
      dotp = warp_sum(dotp);
      // dz for all l (every warp, for dc); each warp stores only its rows afterwards
      float dcp = 0.f;
      for (int l = lane; l < L; l += 32) {
        const bool keep = P.mask && P.mask[b * P.m_sb + l] != 0.f;
        const float z = zb[l] + cc;
        const float y = act_fwd(P.act, z);
        const float ds = keep ? ab[l] * (db[l] - dotp) : 0.f;  // a padded score is a constant
        dcp += ds * act_bwd(P.act, z, y);
      }
      const float dc = warp_sum(dcp);
      __syncthreads();  // everyone has read db (= da) before it is overwritten with dz
      for (int l = warp * 32 + lane; l < L; l += DIN_WARPS * 32) {
        const bool keep = P.mask && P.mask[b * P.m_sb + l] != 0.f;
        const float z = zb[l] + cc;
        const float y = act_fwd(P.act, z);
        const float da = db[l];
        db[l] = (keep ? ab[l] * (da - dotp) : 0.f) * act_bwd(P.act, z, y);
      }
      __syncthreads();
      float* gk = P.gk + b * P.gk_sb;
      float* gv = P.gv + b * P.gv_sb;
      for (int c = lane * 4; c < d; c += 128) {
        float4 du = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 uu = *reinterpret_cast<const float4*>(u + c);
        const float4 gg = *reinterpret_cast<const float4*>(gs + c);
        for (int l = warp; l < L; l += DIN_WARPS) {
          const float dz = db[l], a = ab[l];
          const float4 kk = *reinterpret_cast<const float4*>(kt + (long long)l * d + c);
          du.x = fmaf(dz, kk.x, du.x);
          du.y = fmaf(dz, kk.y, du.y);
          du.z = fmaf(dz, kk.z, du.z);
          du.w = fmaf(dz, kk.w, du.w);
          *reinterpret_cast<float4*>(gk + (long long)l * d + c) =
              make_float4(dz * uu.x, dz * uu.y, dz * uu.z, dz * uu.w);
          *reinterpret_cast<float4*>(gv + (long long)l * d + c) =
              make_float4(a * gg.x, a * gg.y, a * gg.z, a * gg.w);
        }
        *reinterpret_cast<float4*>(part + warp * d + c) = du;
      }
      __syncthreads();
      float* gw = P.gw_rows + b * (4LL * d + 1);
      for (int ci = threadIdx.x; ci < d; ci += DIN_WARPS * 32) {
        float duv = 0.f;
#pragma unroll
        for (int w = 0; w < DIN_WARPS; ++w) duv += part[w * d + ci];
        const float qv = qs[ci];
        const float w1 = __ldg(P.W + ci), w3 = __ldg(P.W + 2 * d + ci), w4 = __ldg(P.W + 3 * d + ci);
        P.gq[b * P.gq_sb + ci] = dc * (w1 + w3) + duv * w4;
        gw[ci] = dc * qv;
        gw[d + ci] = duv;
        gw[2 * d + ci] = dc * qv - duv;
        gw[3 * d + ci] = duv * qv;
      }
      if (threadIdx.x == 0) gw[4 * d] = dc;
    }
    __syncthreads();  // all reads of the tiles / the scratch vectors are done
    if (b + gridDim.x < P.B) issue(b + gridDim.x, 0);
  }
}

template <int NW>
static int din_launch_nw(const DinParams& P, bool bwd, int kv_same, cudaStream_t st) {
  const int L = P.L, d = P.d, Lp = (L + 3) & ~3;
  const size_t floats = (size_t)(kv_same ? 1 : 2) * L * d + 3 * (size_t)d + (size_t)NW * d +
                        3 * (size_t)Lp + 4;
  const size_t smem = floats * 4;
  if (smem > 227 * 1024) return RTF_E_RANGE;
  auto kern = bwd ? din_kernel<true, NW> : din_kernel<false, NW>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 64 / NW) per_sm = 64 / NW;
  if (per_sm > 16) per_sm = 16;
  long long blocks = P.B;
  if (blocks > (long long)kNumSMs * per_sm) blocks = (long long)kNumSMs * per_sm;
  kern<<<(unsigned)blocks, NW * 32, smem, st>>>(P, kv_same);
  RTF_CHECK_LAUNCH();
  return 0;
}

static int din_launch(const DinParams& P, bool bwd, cudaStream_t st) {
  const int kv_same = (P.k == P.v && P.k_sb == P.v_sb) ? 1 : 0;
  // small tiles: one warp per sample (more samples in flight); large tiles: 8 warps cooperate
  if ((long long)P.L * P.d <= 4096) return din_launch_nw<1>(P, bwd, kv_same, st);
  return din_launch_nw<8>(P, bwd, kv_same, st);
}

static int din_check(int64_t B, int L, int d, int act, const float* q, const float* k,
                     const float* v, int64_t k_sb, int64_t v_sb) {
  if (B < 0 || L <= 0 || d <= 0 || act < ACT_NONE || act > ACT_TANH) return RTF_E_ARG;
  if (d % 4 || d > 1024 || L > 4096) return RTF_E_RANGE;
  if (B == 0) return 0;
  if (!q || !k || !v) return RTF_E_ARG;
  if ((uintptr_t)k % 16 || (uintptr_t)v % 16 || k_sb % 4 || v_sb % 4) return RTF_E_ALIGN;
  return 0;
}

}  // namespace rtf

using namespace rtf;

extern "C" int rtf_din_attn_fwd(const float* d_q, int64_t q_sb, const float* d_k, int64_t k_sb,
                                const float* d_v, int64_t v_sb, const float* d_mask, int64_t m_sb,
                                const float* d_W, const float* d_bias, int act, int64_t B, int L,
                                int d, float* d_out, int64_t o_sb, void* stream) {
  int rc = din_check(B, L, d, act, d_q, d_k, d_v, k_sb, v_sb);
  if (rc || B == 0) return rc;
  if (!d_W || !d_bias || !d_out) return RTF_E_ARG;
  DinParams P = {};
  P.q = d_q; P.q_sb = q_sb; P.k = d_k; P.k_sb = k_sb; P.v = d_v; P.v_sb = v_sb; P.mask = d_mask;
  P.m_sb = m_sb; P.W = d_W; P.bias = d_bias; P.act = act; P.B = B; P.L = L; P.d = d;
  P.out = d_out; P.o_sb = o_sb;
  return din_launch(P, false, (cudaStream_t)stream);
}

extern "C" int rtf_din_attn_bwd(const float* d_q, int64_t q_sb, const float* d_k, int64_t k_sb,
                                const float* d_v, int64_t v_sb, const float* d_mask, int64_t m_sb,
                                const float* d_W, const float* d_bias, int act, int64_t B, int L,
                                int d, const float* d_gout, int64_t go_sb, float* d_gq,
                                int64_t gq_sb, float* d_gk, int64_t gk_sb, float* d_gv,
                                int64_t gv_sb, float* d_gw_rows, void* stream) {
  int rc = din_check(B, L, d, act, d_q, d_k, d_v, k_sb, v_sb);
  if (rc || B == 0) return rc;
  if (!d_W || !d_bias || !d_gout || !d_gq || !d_gk || !d_gv || !d_gw_rows) return RTF_E_ARG;
  if ((uintptr_t)d_gk % 16 || (uintptr_t)d_gv % 16 || gk_sb % 4 || gv_sb % 4) return RTF_E_ALIGN;
  DinParams P = {};
  P.q = d_q; P.q_sb = q_sb; P.k = d_k; P.k_sb = k_sb; P.v = d_v; P.v_sb = v_sb; P.mask = d_mask;
  P.m_sb = m_sb; P.W = d_W; P.bias = d_bias; P.act = act; P.B = B; P.L = L; P.d = d;
  P.gout = d_gout; P.go_sb = go_sb; P.gq = d_gq; P.gq_sb = gq_sb; P.gk = d_gk; P.gk_sb = gk_sb;
  P.gv = d_gv; P.gv_sb = gv_sb; P.gw_rows = d_gw_rows;
  return din_launch(P, true, (cudaStream_t)stream);
}
