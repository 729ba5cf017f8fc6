// Dense optimizer steps around the hot path.
//
// (1) rtf_dense_adam — Keras-form Adam over ONE flat parameter / gradient / moment buffer
//     (the data-parallel MLP replicas): the same formulas, rounding and step-size folding as
//     K2's row update (embed_bwd.cu finish_row), i.e. what model.compile(optimizer=Adam(1e-3))
//     applies to every dense variable (src/ctr/fm/train.py:49-50):
//       m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  w -= lr_t m / (sqrt(v) + eps),
//       lr_t = lr sqrt(1-b2^t) / (1-b1^t)   (folded on the host, passed as opt.lr)
// (2) rtf_rows_apply_dense — the replicated (data-parallel) small tables of the multi-GPU path:
//     every replica applies K2's row update to the rows touched anywhere in the global batch,
//     from the all-reduced dense block [G (R,D) | touched (R)].
#include "rtf_common.cuh"

namespace rtf {

__device__ __forceinline__ void adam_elem(const rtf_opt& o, float lr, float g, float& w, float& a,
                                          float& b) {
  a = __fadd_rn(__fmul_rn(o.beta1, a), __fmul_rn(__fsub_rn(1.0f, o.beta1), g));
  b = __fadd_rn(__fmul_rn(o.beta2, b), __fmul_rn(__fsub_rn(1.0f, o.beta2), __fmul_rn(g, g)));
  w = __fsub_rn(w, __fdiv_rn(__fmul_rn(lr, a), __fadd_rn(__fsqrt_rn(b), o.eps)));
}

__global__ void __launch_bounds__(256)
dense_adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                  float* __restrict__ v, long long n4, const __grid_constant__ rtf_opt o) {
  const float lr = o.lr_dev ? __ldg(o.lr_dev) : o.lr;   // device scalar under CUDA-graph replay
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 wv = reinterpret_cast<float4*>(w)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_elem(o, lr, gv.x, wv.x, mv.x, vv.x);
    adam_elem(o, lr, gv.y, wv.y, mv.y, vv.y);
    adam_elem(o, lr, gv.z, wv.z, mv.z, vv.z);
    adam_elem(o, lr, gv.w, wv.w, mv.w, vv.w);
    reinterpret_cast<float4*>(w)[i] = wv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
}

// one lane-group of D/4 lanes (<= 32) per row; rows whose touched count is 0 do not move
__global__ void __launch_bounds__(256)
rows_apply_dense_kernel(float* __restrict__ W, float* __restrict__ s1, float* __restrict__ s2,
                        const float* __restrict__ G, const float* __restrict__ touched,
                        long long R, int D, int lanes, const __grid_constant__ rtf_opt o) {
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / lanes;
  const int lg = (int)(threadIdx.x % lanes);
  if (gid >= R || touched[gid] <= 0.f) return;
  const float lr = o.lr_dev ? __ldg(o.lr_dev) : o.lr;
  const float two_l2 = __fmul_rn(2.0f, o.l2);
  for (int c = lg * 4; c < D; c += lanes * 4) {
    const long long off = gid * D + c;
    float4 wv = *reinterpret_cast<float4*>(W + off);
    const float4 gv = *reinterpret_cast<const float4*>(G + off);
    float gq[4] = {gv.x, gv.y, gv.z, gv.w};
    float wq[4] = {wv.x, wv.y, wv.z, wv.w};
    float aq[4] = {0.f, 0.f, 0.f, 0.f}, bq[4] = {0.f, 0.f, 0.f, 0.f};
    if (o.kind >= RTF_OPT_ADAGRAD) {
      const float4 t = *reinterpret_cast<float4*>(s1 + off);
      aq[0] = t.x; aq[1] = t.y; aq[2] = t.z; aq[3] = t.w;
    }
    if (o.kind == RTF_OPT_ADAM) {
      const float4 t = *reinterpret_cast<float4*>(s2 + off);
      bq[0] = t.x; bq[1] = t.y; bq[2] = t.z; bq[3] = t.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float g = gq[e];
      if (o.l2 > 0.f) g = __fadd_rn(g, __fmul_rn(two_l2, wq[e]));
      if (o.kind == RTF_OPT_SGD) {
        wq[e] = __fsub_rn(wq[e], __fmul_rn(lr, g));
      } else if (o.kind == RTF_OPT_ADAGRAD) {
        aq[e] = __fadd_rn(aq[e], __fmul_rn(g, g));
        wq[e] = __fsub_rn(wq[e], __fdiv_rn(__fmul_rn(lr, g), __fadd_rn(__fsqrt_rn(aq[e]), o.eps)));
      } else {
        adam_elem(o, lr, g, wq[e], aq[e], bq[e]);
      }
    }
    *reinterpret_cast<float4*>(W + off) = make_float4(wq[0], wq[1], wq[2], wq[3]);
    if (o.kind >= RTF_OPT_ADAGRAD)
      *reinterpret_cast<float4*>(s1 + off) = make_float4(aq[0], aq[1], aq[2], aq[3]);
    if (o.kind == RTF_OPT_ADAM)
      *reinterpret_cast<float4*>(s2 + off) = make_float4(bq[0], bq[1], bq[2], bq[3]);
  }
}

}  // namespace rtf

extern "C" int rtf_dense_adam(float* d_w, const float* d_g, float* d_m, float* d_v, int64_t n,
                              const rtf_opt* opt, void* stream) {
  if (!opt || n < 0 || opt->kind != RTF_OPT_ADAM) return RTF_E_ARG;
  if (n == 0) return 0;
  if (!d_w || !d_g || !d_m || !d_v) return RTF_E_ARG;
  if (n % 4 || (uintptr_t)d_w % 16 || (uintptr_t)d_g % 16 || (uintptr_t)d_m % 16 ||
      (uintptr_t)d_v % 16)
    return RTF_E_ALIGN;
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > rtf::kNumSMs * 8) blocks = rtf::kNumSMs * 8;
  rtf::dense_adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_w, d_g, d_m, d_v,
                                                                            n4, *opt);
  RTF_CHECK_LAUNCH();
  return 0;
}

extern "C" int rtf_rows_apply_dense(float* d_w, float* d_s1, float* d_s2, const float* d_g,
                                    const float* d_touched, int64_t rows, int dim,
                                    const rtf_opt* opt, void* stream) {
  if (!opt || rows < 0 || dim <= 0) return RTF_E_ARG;
  if (opt->kind < RTF_OPT_SGD || opt->kind > RTF_OPT_ADAM) return RTF_E_ARG;
  if (rows == 0) return 0;
  if (!d_w || !d_g || !d_touched) return RTF_E_ARG;
  if (opt->kind >= RTF_OPT_ADAGRAD && !d_s1) return RTF_E_ARG;
  if (opt->kind == RTF_OPT_ADAM && !d_s2) return RTF_E_ARG;
  if (dim % 4 || (uintptr_t)d_w % 16 || (uintptr_t)d_g % 16 || (uintptr_t)d_s1 % 16 ||
      (uintptr_t)d_s2 % 16)
    return RTF_E_ALIGN;
  int lanes = 1;
  while (lanes < dim / 4 && lanes < 32) lanes <<= 1;
  const long long threads = rows * lanes;
  rtf::rows_apply_dense_kernel<<<(unsigned)((threads + 255) / 256), 256, 0,
                                 (cudaStream_t)stream>>>(d_w, d_s1, d_s2, d_g, d_touched, rows,
                                                         dim, lanes, *opt);
  RTF_CHECK_LAUNCH();
  return 0;
}
