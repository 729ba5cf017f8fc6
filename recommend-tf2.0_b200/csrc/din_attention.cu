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
// so a sample needs one pass over its k/v rows.  One warp per sample; its (L,d) key tile (and
// value tile when v is a different tensor) is staged in shared memory by a single TMA bulk copy
// (SASS UBLKCP) on a per-warp mbarrier; scores use lane-groups over d with shuffle reductions
// (conflict-free shared reads), the softmax and the a.v product run out of shared memory.
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

// dot(vec[d], tile[l][d]) for every l, written to dst[l]; lane-groups of g lanes split d
__device__ __forceinline__ void rows_dot(const float* __restrict__ vec, const float* __restrict__ tile,
                                         int L, int d, int g, int lane, float* __restrict__ dst) {
  const int rpi = 32 / g;  // rows per iteration
  const int lg = lane % g, lr = lane / g;
  for (int l0 = 0; l0 < L; l0 += rpi) {
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

template <bool BWD>
__global__ void __launch_bounds__(256) din_kernel(const __grid_constant__ DinParams P, int warp_floats,
                                                  int kv_same) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int L = P.L, d = P.d;
  const int Lp = (L + 3) & ~3;
  float* kt = smem + (size_t)warp * warp_floats;        // [L][d]
  float* vt = kv_same ? kt : kt + L * d;                // [L][d]
  float* u = (kv_same ? kt + L * d : vt + L * d);       // [d]   u, later reused
  float* qs = u + d;                                    // [d]   q
  float* zb = qs + d;                                   // [Lp]  pre-activation z
  float* ab = zb + Lp;                                  // [Lp]  scores -> attention weights
  float* db = ab + Lp;                                  // [Lp]  (bwd) da -> dz
  uint64_t* bar = reinterpret_cast<uint64_t*>(db + Lp);
  if (lane == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncwarp();
  int g = 1;
  while (g * 4 < d && g < 32) g <<= 1;
  const unsigned tile_bytes = (unsigned)L * d * 4u;
  const long long stride = (long long)gridDim.x * nwarps;
  long long b = (long long)blockIdx.x * nwarps + warp;
  uint32_t parity = 0;
  auto issue = [&](long long bb) {
    if (lane == 0) {
      mbar_expect_tx(bar, kv_same ? tile_bytes : 2 * tile_bytes);
      bulk_g2s(kt, P.k + bb * P.k_sb, tile_bytes, bar);
      if (!kv_same) bulk_g2s(vt, P.v + bb * P.v_sb, tile_bytes, bar);
    }
  };
  if (b < P.B) issue(b);
  for (; b < P.B; b += stride) {
    // u = (W2 - W3) + W4*q ; c = (W1 + W3).q + bias
    float cpart = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float qv = P.q[b * P.q_sb + c];
      const float w1 = __ldg(P.W + c), w2 = __ldg(P.W + d + c), w3 = __ldg(P.W + 2 * d + c),
                  w4 = __ldg(P.W + 3 * d + c);
      qs[c] = qv;
      u[c] = (w2 - w3) + w4 * qv;
      cpart = fmaf(w1 + w3, qv, cpart);
    }
    const float cc = warp_sum(cpart) + __ldg(P.bias);
    __syncwarp();
    mbar_wait(bar, parity);
    parity ^= 1;
    rows_dot(u, kt, L, d, g, lane, zb);
    __syncwarp();
    // activation, mask, softmax
    float m = -INFINITY;
    for (int l = lane; l < L; l += 32) {
      const float z = zb[l] + cc;
      zb[l] = z;
      float s = act_fwd(P.act, z);
      const bool keep = P.mask && P.mask[b * P.m_sb + l] != 0.f;
      s = keep ? s : kPadLogit;
      ab[l] = s;
      m = fmaxf(m, s);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float e = expf(ab[l] - m);
      ab[l] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int l = lane; l < L; l += 32) ab[l] *= inv;
    __syncwarp();
    if (!BWD) {
      for (int c = lane * 4; c < d; c += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < L; ++l) {
          const float a = ab[l];
          const float4 vv = *reinterpret_cast<const float4*>(vt + (long long)l * d + c);
          acc.x = fmaf(a, vv.x, acc.x);
          acc.y = fmaf(a, vv.y, acc.y);
          acc.z = fmaf(a, vv.z, acc.z);
          acc.w = fmaf(a, vv.w, acc.w);
        }
        float* o = P.out + b * P.o_sb + c;
        o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
      }
    } else {
      // da_l = gout . v_l  (gout staged over q's slot is not possible: q is needed) -> use u? no:
      // keep u; stage gout in registers per lane-group via a small shared vector reuse of db? db is
      // [Lp] floats; gout needs d floats -> stage it after db (room reserved by warp_floats).
      float* gs = reinterpret_cast<float*>(bar + 2);  // [d] gout staging (16-byte aligned)
      for (int c = lane; c < d; c += 32) gs[c] = P.gout[b * P.go_sb + c];
      __syncwarp();
      rows_dot(gs, vt, L, d, g, lane, db);
      __syncwarp();
      float dotp = 0.f;
      for (int l = lane; l < L; l += 32) dotp = fmaf(ab[l], db[l], dotp);
      dotp = warp_sum(dotp);
      float dcp = 0.f;
      for (int l = lane; l < L; l += 32) {
        const bool keep = P.mask && P.mask[b * P.m_sb + l] != 0.f;
        const float z = zb[l];
        const float y = act_fwd(P.act, z);
        const float ds = keep ? ab[l] * (db[l] - dotp) : 0.f;  // padded score is a constant
        const float dz = ds * act_bwd(P.act, z, y);
        db[l] = dz;
        dcp += dz;
      }
      const float dc = warp_sum(dcp);
      __syncwarp();
      float* gk = P.gk + b * P.gk_sb;
      float* gv = P.gv + b * P.gv_sb;
      float* gw = P.gw_rows + b * (4LL * d + 1);
      for (int c = lane * 4; c < d; c += 128) {
        float4 du = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 uu = *reinterpret_cast<const float4*>(u + c);
        const float4 gg = *reinterpret_cast<const float4*>(gs + c);
        for (int l = 0; l < L; ++l) {
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
        const float duv[4] = {du.x, du.y, du.z, du.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ci = c + e;
          const float qv = qs[ci];
          const float w1 = __ldg(P.W + ci), w3 = __ldg(P.W + 2 * d + ci), w4 = __ldg(P.W + 3 * d + ci);
          P.gq[b * P.gq_sb + ci] = dc * (w1 + w3) + duv[e] * w4;
          gw[ci] = dc * qv;
          gw[d + ci] = duv[e];
          gw[2 * d + ci] = dc * qv - duv[e];
          gw[3 * d + ci] = duv[e] * qv;
        }
      }
      if (lane == 0) gw[4 * d] = dc;
    }
    __syncwarp();
    if (b + stride < P.B) issue(b + stride);
  }
}

static int din_launch(const DinParams& P, bool bwd, cudaStream_t st) {
  const int L = P.L, d = P.d, Lp = (L + 3) & ~3;
  const int kv_same = (P.k == P.v && P.k_sb == P.v_sb) ? 1 : 0;
  // tiles + u + q + z/a/d + mbarrier slot (4 floats) + gout staging (d)
  const int warp_floats = (kv_same ? 1 : 2) * L * d + 2 * d + 3 * Lp + 4 + d;
  const size_t per_warp = (size_t)warp_floats * 4;
  int nwarps = (int)((227 * 1024) / per_warp);
  if (nwarps < 1) return RTF_E_RANGE;
  if (nwarps > 8) nwarps = 8;
  const size_t smem = per_warp * nwarps;
  auto kern = bwd ? din_kernel<true> : din_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int per_sm = (int)((227 * 1024) / smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  long long blocks = (P.B + nwarps - 1) / nwarps;
  if (blocks > (long long)kNumSMs * per_sm) blocks = (long long)kNumSMs * per_sm;
  kern<<<(unsigned)blocks, nwarps * 32, smem, st>>>(P, warp_floats, kv_same);
  RTF_CHECK_LAUNCH();
  return 0;
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
