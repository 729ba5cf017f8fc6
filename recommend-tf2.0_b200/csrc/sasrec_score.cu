// a10 — SASRec scoring + loss epilogue (src/match/sasrec/model.py:88-96), fused with the
// positive / negative item gathers (:77-79):
//   seq_info = att_outputs[:, -1]                       (B, D)      (:88)
//   pos = sum_d seq_info * pos_embed  (B, 1),  neg = sum_d seq_info * neg_embed  (B, NEG)   (:90-91)
//   loss = mean_{b,j}( -log sigmoid(pos_b) - log(1 - sigmoid(neg_bj)) ) / 2                 (:93-95)
//   logits = concat([pos, neg])       (B, 1 + NEG)                                          (:96)
// The reference materialises the two gathered tensors (B, 1+NEG, D), two products and four
// elementwise passes.  Here one warp owns a sample: its 1 + NEG rows are read straight from
// the pos / neg tables (one 16-byte load per lane per row), reduced with shuffles, and the
// sample's loss terms leave as one float (the batch mean is finished by rtf_colsum, fixed
// order).  Backward re-gathers the rows: d seq_info = sum_j ds_j row_j, and the row gradients
// ds_j * seq_info go to K2 (deterministic segment reduce + sparse optimizer).
// The literal formulas are kept (sigmoid, then log / log(1 - .)), saturation included.
#include "rtf_common.cuh"

namespace rtf {

struct SsParams {
  const float* info; long long info_sb;
  const float* tab[2]; long long rows[2];   // [0] pos table, [1] neg table
  const void* pos_ids; const void* neg_ids; long long neg_sb;
  long long B; int NEG, D;
  float* logits; float* loss_rows;          // fwd
  const float* logits_in; const float* gloss; const float* glogits;  // bwd
  float* ginfo; float* gemb;                // bwd
  int32_t* err;
};

constexpr int SS_MAXC = 4;  // float4 chunks per lane: D <= 512

template <typename IdT>
__device__ __forceinline__ const float* ss_row(const SsParams& P, long long b, int j) {
  const int t = j == 0 ? 0 : 1;
  const long long idx = j == 0 ? b : b * P.neg_sb + (j - 1);
  const long long id = load_id((const IdT*)(j == 0 ? P.pos_ids : P.neg_ids), idx, P.rows[t], P.err);
  return id < 0 ? nullptr : P.tab[t] + id * P.D;
}

template <typename IdT>
__global__ void __launch_bounds__(256)
sasrec_score_fwd_kernel(const __grid_constant__ SsParams P) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nc = P.D >> 2;
  for (long long b = warp0; b < P.B; b += nwarps) {
    float4 x[SS_MAXC];
#pragma unroll
    for (int c = 0; c < SS_MAXC; ++c) {
      const int ci = lane + 32 * c;
      x[c] = ci < nc ? *reinterpret_cast<const float4*>(P.info + b * P.info_sb + 4 * ci)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float lpos = 0.f, lneg = 0.f;
    for (int j = 0; j <= P.NEG; ++j) {
      const float* row = ss_row<IdT>(P, b, j);
      float s = 0.f;
      if (row) {
#pragma unroll
        for (int c = 0; c < SS_MAXC; ++c) {
          const int ci = lane + 32 * c;
          if (ci < nc) {
            const float4 r = ldg_nc_f4(row + 4 * ci);
            s = fmaf(x[c].x, r.x, s);
            s = fmaf(x[c].y, r.y, s);
            s = fmaf(x[c].z, r.z, s);
            s = fmaf(x[c].w, r.w, s);
          }
        }
      }
      s = warp_sum(s);
      if (lane == 0) {
        P.logits[b * (P.NEG + 1) + j] = s;
        const float sg = 1.f / (1.f + expf(-s));
        if (j == 0) lpos = -logf(sg);
        else lneg += -logf(1.f - sg);
      }
    }
    if (lane == 0) P.loss_rows[b] = fmaf((float)P.NEG, lpos, lneg);
  }
}

template <typename IdT>
__global__ void __launch_bounds__(256)
sasrec_score_bwd_kernel(const __grid_constant__ SsParams P) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nc = P.D >> 2;
  const float gl = P.gloss ? __ldg(P.gloss) / (2.f * (float)P.B * (float)P.NEG) : 0.f;
  for (long long b = warp0; b < P.B; b += nwarps) {
    float4 x[SS_MAXC], gi[SS_MAXC];
#pragma unroll
    for (int c = 0; c < SS_MAXC; ++c) {
      const int ci = lane + 32 * c;
      x[c] = ci < nc ? *reinterpret_cast<const float4*>(P.info + b * P.info_sb + 4 * ci)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
      gi[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int j = 0; j <= P.NEG; ++j) {
      const float s = __ldg(P.logits_in + b * (P.NEG + 1) + j);
      const float sg = 1.f / (1.f + expf(-s));
      float ds = j == 0 ? gl * (float)P.NEG * (sg - 1.f) : gl * sg;
      if (P.glogits) ds += __ldg(P.glogits + b * (P.NEG + 1) + j);
      const float* row = ss_row<IdT>(P, b, j);
      float* ge = P.gemb + (b * (P.NEG + 1) + j) * P.D;
#pragma unroll
      for (int c = 0; c < SS_MAXC; ++c) {
        const int ci = lane + 32 * c;
        if (ci < nc) {
          if (row) {
            const float4 r = ldg_nc_f4(row + 4 * ci);
            gi[c].x = fmaf(ds, r.x, gi[c].x);
            gi[c].y = fmaf(ds, r.y, gi[c].y);
            gi[c].z = fmaf(ds, r.z, gi[c].z);
            gi[c].w = fmaf(ds, r.w, gi[c].w);
          }
          *reinterpret_cast<float4*>(ge + 4 * ci) =
              make_float4(ds * x[c].x, ds * x[c].y, ds * x[c].z, ds * x[c].w);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < SS_MAXC; ++c) {
      const int ci = lane + 32 * c;
      if (ci < nc) *reinterpret_cast<float4*>(P.ginfo + b * P.D + 4 * ci) = gi[c];
    }
  }
}

static int ss_check(const SsParams& P) {
  if (P.B < 0 || P.NEG < 0 || P.D <= 0 || P.rows[0] <= 0 || P.rows[1] <= 0) return RTF_E_ARG;
  if (P.B > 0 && (!P.info || !P.tab[0] || !P.tab[1] || !P.pos_ids || (P.NEG > 0 && !P.neg_ids)))
    return RTF_E_ARG;
  if (P.D % 4 || P.D > 128 * SS_MAXC) return RTF_E_RANGE;
  if ((uintptr_t)P.info % 16 || P.info_sb % 4 || (uintptr_t)P.tab[0] % 16 || (uintptr_t)P.tab[1] % 16)
    return RTF_E_ALIGN;
  return 0;
}

static unsigned ss_blocks(long long B) {
  long long blocks = (B + 7) / 8;
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace rtf

using namespace rtf;

extern "C" int rtf_sasrec_score_fwd(const float* d_info, int64_t info_sb, const float* d_pos_tab,
                                    int64_t pos_rows, const float* d_neg_tab, int64_t neg_rows,
                                    const void* d_pos_ids, const void* d_neg_ids, int ids_i64,
                                    int64_t neg_sb, int64_t B, int NEG, int D, float* d_logits,
                                    float* d_loss_rows, int32_t* d_err, void* stream) {
  SsParams P = {};
  P.info = d_info; P.info_sb = info_sb; P.tab[0] = d_pos_tab; P.tab[1] = d_neg_tab;
  P.rows[0] = pos_rows; P.rows[1] = neg_rows; P.pos_ids = d_pos_ids; P.neg_ids = d_neg_ids;
  P.neg_sb = neg_sb; P.B = B; P.NEG = NEG; P.D = D; P.logits = d_logits; P.loss_rows = d_loss_rows;
  P.err = d_err;
  int rc = ss_check(P);
  if (rc) return rc;
  if (B == 0) return 0;
  if (!d_logits || !d_loss_rows) return RTF_E_ARG;
  if (ids_i64)
    sasrec_score_fwd_kernel<int64_t><<<ss_blocks(B), 256, 0, (cudaStream_t)stream>>>(P);
  else
    sasrec_score_fwd_kernel<int32_t><<<ss_blocks(B), 256, 0, (cudaStream_t)stream>>>(P);
  RTF_CHECK_LAUNCH();
  return 0;
}

extern "C" int rtf_sasrec_score_bwd(const float* d_info, int64_t info_sb, const float* d_pos_tab,
                                    int64_t pos_rows, const float* d_neg_tab, int64_t neg_rows,
                                    const void* d_pos_ids, const void* d_neg_ids, int ids_i64,
                                    int64_t neg_sb, int64_t B, int NEG, int D, const float* d_logits,
                                    const float* d_gloss, const float* d_glogits, float* d_ginfo,
                                    float* d_gemb, void* stream) {
  SsParams P = {};
  P.info = d_info; P.info_sb = info_sb; P.tab[0] = d_pos_tab; P.tab[1] = d_neg_tab;
  P.rows[0] = pos_rows; P.rows[1] = neg_rows; P.pos_ids = d_pos_ids; P.neg_ids = d_neg_ids;
  P.neg_sb = neg_sb; P.B = B; P.NEG = NEG; P.D = D; P.logits_in = d_logits; P.gloss = d_gloss;
  P.glogits = d_glogits; P.ginfo = d_ginfo; P.gemb = d_gemb;
  int rc = ss_check(P);
  if (rc) return rc;
  if (B == 0) return 0;
  if (!d_logits || !d_ginfo || !d_gemb) return RTF_E_ARG;
  if ((uintptr_t)d_ginfo % 16 || (uintptr_t)d_gemb % 16) return RTF_E_ALIGN;
  if (ids_i64)
    sasrec_score_bwd_kernel<int64_t><<<ss_blocks(B), 256, 0, (cudaStream_t)stream>>>(P);
  else
    sasrec_score_bwd_kernel<int32_t><<<ss_blocks(B), 256, 0, (cudaStream_t)stream>>>(P);
  RTF_CHECK_LAUNCH();
  return 0;
}
