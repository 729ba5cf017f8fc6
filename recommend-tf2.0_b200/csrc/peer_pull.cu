// Multi-GPU exchange, forward half, as its own copy kernel: every rank pulls the
// embedding rows of ITS samples from the holders' owner-gathered (B_global, T_g*D) buffers over
// NVLink peer memory into a local (B_local, F*D) staging buffer.  Launched on a side stream it
// overlaps the bottom MLP (tensor-core GEMMs), so the interaction kernel that follows reads
// local HBM only and the "all-to-all of pooled embeddings" leaves the critical path; the same
// buffer is what the backward re-reads (it replaces the in-kernel pull + xsave copy).
// One lane-group of D/4 lanes per (sample, field) row: a 512-byte row at D = 128 is one 16-byte
// load per lane; loads bypass L1 (ld.global.cv): the remote buffers are rewritten every step.
#include "rtf_common.cuh"

namespace rtf {

struct PullParams {
  const long long* peer_tab;   // [F][G] base of field f's column in rank g's gathered buffer
  const long long* peer_str;   // [F][G] elements between consecutive samples there
  long long rows[RTF_MAX_FIELDS];
  const void* ids; long long ids_sb, ids_sf;
  unsigned long long rw_mask;
  long long sample0, B;
  int F, D, G;
  float* out; long long out_sb;
  int32_t* err;
};

template <typename IdT>
__global__ void __launch_bounds__(256)
peer_pull_kernel(const __grid_constant__ PullParams P, int lanes) {
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / lanes;
  const int lg = (int)(threadIdx.x % lanes);
  if (gid >= P.B * P.F) return;
  const long long b = gid / P.F;
  const int f = (int)(gid - b * P.F);
  const long long id = load_id((const IdT*)P.ids, b * P.ids_sb + (long long)f * P.ids_sf, P.rows[f],
                               P.err);
  float* dst = P.out + b * P.out_sb + (long long)f * P.D;
  const float* src = nullptr;
  if (id >= 0) {
    const int g = ((P.rw_mask >> f) & 1ull) ? (int)(id % P.G) : 0;
    const int e = f * P.G + g;
    src = reinterpret_cast<const float*>(P.peer_tab[e]) + (P.sample0 + b) * P.peer_str[e];
  }
  for (int c = lg * 4; c < P.D; c += lanes * 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);   // bad id: the row reads as zeros
    if (src) v = __ldcv(reinterpret_cast<const float4*>(src + c));
    stg_cs_f4(dst + c, v);
  }
}

}  // namespace rtf

extern "C" int rtf_peer_pull_rows(const int64_t* d_peer_tab, const int64_t* d_peer_str,
                                  int64_t sample0, int G, uint64_t rw_mask, const int64_t* rows,
                                  int n_fields, int D, const void* d_ids, int ids_i64, int64_t B,
                                  int64_t ids_sb, int64_t ids_sf, float* d_out, int64_t out_sb,
                                  int32_t* d_err, void* stream) {
  using namespace rtf;
  if (!d_peer_tab || !d_peer_str || !rows || n_fields < 1 || G < 1 || B < 0 || D <= 0)
    return RTF_E_ARG;
  if (n_fields > RTF_MAX_FIELDS || D % 4) return RTF_E_RANGE;
  if (B == 0) return 0;
  if (!d_ids || !d_out) return RTF_E_ARG;
  if ((uintptr_t)d_out % 16 || out_sb % 4 || out_sb < (int64_t)n_fields * D) return RTF_E_ALIGN;
  PullParams P = {};
  for (int f = 0; f < n_fields; ++f) {
    if (rows[f] <= 0) return RTF_E_ARG;
    P.rows[f] = rows[f];
  }
  P.peer_tab = (const long long*)d_peer_tab; P.peer_str = (const long long*)d_peer_str;
  P.ids = d_ids; P.ids_sb = ids_sb; P.ids_sf = ids_sf; P.rw_mask = rw_mask; P.sample0 = sample0;
  P.B = B; P.F = n_fields; P.D = D; P.G = G; P.out = d_out; P.out_sb = out_sb; P.err = d_err;
  int lanes = 1;
  while (lanes < D / 4 && lanes < 32) lanes <<= 1;
  const long long threads = B * n_fields * lanes;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  if (ids_i64)
    peer_pull_kernel<int64_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(P, lanes);
  else
    peer_pull_kernel<int32_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(P, lanes);
  RTF_CHECK_LAUNCH();
  return 0;
}
