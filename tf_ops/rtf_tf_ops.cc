// TensorFlow custom-op shim over librtf_b200.so: tf.load_op_library binds the
// SAME C symbols the ctypes layer binds, so the reference's model files can keep running on
// TensorFlow while the hot path runs on the sm_100a kernels.
//
// Build (only where TensorFlow headers exist — they do not in the build image):  make -C tf_ops
// The Makefile probes `import tensorflow` and skips cleanly otherwise.
//
// Ops (GPU only; TF is used for shapes, allocation and the stream, nothing else):
//   RtfEmbedFwd      K1  F x ResourceGather + ConcatV2      src/ctr/dlrm/model.py:45-46
//   RtfEmbedBwdAdam  K2  IndexedSlices -> UnsortedSegmentSum -> ResourceApplyAdam
//                                                          src/ctr/fm/train.py:49-50
//   RtfEmbedDotFwd / RtfEmbedDotBwd   K1+K4  gather + DLRM pairwise dot (src/ctr/dlrm/model.py:48,
//                                     defined from the paper cited at :7)
//   RtfBce           Keras binary_crossentropy on probabilities + d loss / d p in one pass
//                                                          src/ctr/fm/train.py:49
// The remaining entry points of include/rtf_b200.h wrap the same way.
#include <cstdint>
#include <vector>

#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#define EIGEN_USE_GPU
#include "tensorflow/core/util/gpu_device_functions.h"

#include "rtf_b200.h"

using namespace tensorflow;  // NOLINT
using shape_inference::DimensionHandle;
using shape_inference::InferenceContext;

namespace {

void* StreamOf(OpKernelContext* ctx) {
  return reinterpret_cast<void*>(ctx->eigen_gpu_device().stream());
}

struct TableArgs {
  std::vector<const float*> ptr;
  std::vector<float*> mptr;
  std::vector<int64_t> rows;
  std::vector<int32_t> dims;
  int64_t sum_dim = 0;
};

TableArgs Tables(const OpInputList& tabs) {
  TableArgs t;
  for (int f = 0; f < tabs.size(); ++f) {
    t.ptr.push_back(tabs[f].flat<float>().data());
    t.mptr.push_back(const_cast<float*>(tabs[f].flat<float>().data()));
    t.rows.push_back(tabs[f].dim_size(0));
    t.dims.push_back(static_cast<int32_t>(tabs[f].dim_size(1)));
    t.sum_dim += tabs[f].dim_size(1);
  }
  return t;
}

}  // namespace

// ----------------------------------------------------------------------------- K1
REGISTER_OP("RtfEmbedFwd")
    .Input("tables: N * float")
    .Input("ids: int32")  // (B, N)
    .Attr("N: int >= 1")
    .Output("out: float")  // (B, sum(dim))
    .SetShapeFn([](InferenceContext* c) {
      DimensionHandle total = c->MakeDim(0);
      for (int i = 0; i < c->num_inputs() - 1; ++i)
        TF_RETURN_IF_ERROR(c->Add(total, c->Dim(c->input(i), 1), &total));
      c->set_output(0, c->Matrix(c->Dim(c->input(c->num_inputs() - 1), 0), total));
      return OkStatus();
    });

class RtfEmbedFwdOp : public OpKernel {
 public:
  explicit RtfEmbedFwdOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    OpInputList tabs;
    OP_REQUIRES_OK(ctx, ctx->input_list("tables", &tabs));
    const Tensor& ids = ctx->input(tabs.size());
    const int F = tabs.size();
    const int64_t B = ids.dim_size(0);
    OP_REQUIRES(ctx, ids.dims() == 2 && ids.dim_size(1) == F,
                errors::InvalidArgument("ids must be (B, N)"));
    TableArgs t = Tables(tabs);
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {B, t.sum_dim}, &out));
    Tensor err;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_INT32, {1}, &err));
    const int rc = rtf_embed_fwd(t.ptr.data(), t.rows.data(), t.dims.data(), F,
                                 ids.flat<int32>().data(), 0, B, 1, F, 1, 0, RTF_POOL_NONE,
                                 out->flat<float>().data(), t.sum_dim, nullptr, StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("rtf_embed_fwd rc=", rc));
  }
};
REGISTER_KERNEL_BUILDER(Name("RtfEmbedFwd").Device(DEVICE_GPU), RtfEmbedFwdOp);

// ----------------------------------------------------------------------------- K2
// In-place sparse Adam on the rows the batch touches (the tables and their m / v slots are
// resource-backed buffers passed as tensors that alias the variables: use
// `tf.raw_ops.ReadVariableOp` aliases or `experimental_ref()` buffers, as any in-place
// custom optimizer op does).  lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) is folded on the host.
REGISTER_OP("RtfEmbedBwdAdam")
    .Input("tables: N * float")
    .Input("m: N * float")
    .Input("v: N * float")
    .Input("ids: int32")       // (B, N)
    .Input("grad: float")      // (B, sum(dim)) = d loss / d RtfEmbedFwd.out
    .Attr("N: int >= 1")
    .Attr("lr_t: float")
    .Attr("beta1: float = 0.9")
    .Attr("beta2: float = 0.999")
    .Attr("epsilon: float = 1e-7")
    .Attr("l2: float = 0.0")
    .Output("num_touched: int32")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->Vector(1));
      return OkStatus();
    });

class RtfEmbedBwdAdamOp : public OpKernel {
 public:
  explicit RtfEmbedBwdAdamOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("lr_t", &opt_.lr));
    OP_REQUIRES_OK(c, c->GetAttr("beta1", &opt_.beta1));
    OP_REQUIRES_OK(c, c->GetAttr("beta2", &opt_.beta2));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &opt_.eps));
    OP_REQUIRES_OK(c, c->GetAttr("l2", &opt_.l2));
    opt_.kind = RTF_OPT_ADAM;
    opt_.lr_dev = nullptr;
  }
  void Compute(OpKernelContext* ctx) override {
    OpInputList tabs, ms, vs;
    OP_REQUIRES_OK(ctx, ctx->input_list("tables", &tabs));
    OP_REQUIRES_OK(ctx, ctx->input_list("m", &ms));
    OP_REQUIRES_OK(ctx, ctx->input_list("v", &vs));
    const int F = tabs.size();
    const Tensor& ids = ctx->input(3 * F);
    const Tensor& grad = ctx->input(3 * F + 1);
    const int64_t B = ids.dim_size(0);
    TableArgs t = Tables(tabs), m = Tables(ms), v = Tables(vs);
    std::vector<int32_t> field_table(F);
    for (int f = 0; f < F; ++f) field_table[f] = f;
    int32_t dim_max = 0;
    for (int32_t d : t.dims) dim_max = d > dim_max ? d : dim_max;
    size_t ws_bytes = 0;
    OP_REQUIRES(ctx, rtf_embed_bwd_workspace(B * F, dim_max, &ws_bytes) == 0,
                errors::Internal("rtf_embed_bwd_workspace"));
    Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_INT8, {static_cast<int64_t>(ws_bytes) + 256}, &ws));
    char* wsp = reinterpret_cast<char*>(ws.flat<int8>().data());
    wsp += (256 - reinterpret_cast<uintptr_t>(wsp) % 256) % 256;
    Tensor* n = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {1}, &n));
    int row_bits = 0;
    const int rc = rtf_embed_bwd(t.mptr.data(), m.mptr.data(), v.mptr.data(), t.rows.data(),
                                 t.dims.data(), F, field_table.data(), F, ids.flat<int32>().data(),
                                 0, B, 1, F, 1, 0, RTF_POOL_NONE, grad.flat<float>().data(),
                                 t.sum_dim, &opt_, nullptr, nullptr, n->flat<int32>().data(),
                                 &row_bits, wsp, ws_bytes, StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("rtf_embed_bwd rc=", rc));
  }

 private:
  rtf_opt opt_;
};
REGISTER_KERNEL_BUILDER(Name("RtfEmbedBwdAdam").Device(DEVICE_GPU), RtfEmbedBwdAdamOp);

// ----------------------------------------------------------------------------- K1 + K4
REGISTER_OP("RtfEmbedDotFwd")
    .Input("tables: N * float")  // all (rows_t, D)
    .Input("ids: int32")         // (B, N)
    .Input("dense: float")       // (B, D) bottom-MLP output
    .Attr("N: int >= 1")
    .Attr("pad_to: int = 1")
    .Output("out: float")        // (B, D + (N+1)N/2 rounded up to pad_to)
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->Matrix(c->Dim(c->input(c->num_inputs() - 1), 0), c->UnknownDim()));
      return OkStatus();
    });

class RtfEmbedDotFwdOp : public OpKernel {
 public:
  explicit RtfEmbedDotFwdOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("pad_to", &pad_to_));
  }
  void Compute(OpKernelContext* ctx) override {
    OpInputList tabs;
    OP_REQUIRES_OK(ctx, ctx->input_list("tables", &tabs));
    const int F = tabs.size();
    const Tensor& ids = ctx->input(F);
    const Tensor& dense = ctx->input(F + 1);
    const int64_t B = ids.dim_size(0);
    const int D = static_cast<int>(dense.dim_size(1));
    TableArgs t = Tables(tabs);
    int cols = D + (F + 1) * F / 2;
    cols = (cols + pad_to_ - 1) / pad_to_ * pad_to_;
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {B, cols}, &out));
    const int rc = rtf_embed_dot_fwd(t.ptr.data(), t.rows.data(), F, D, ids.flat<int32>().data(), 0,
                                     B, F, 1, dense.flat<float>().data(), D,
                                     out->flat<float>().data(), cols, cols, nullptr, StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("rtf_embed_dot_fwd rc=", rc));
  }

 private:
  int pad_to_;
};
REGISTER_KERNEL_BUILDER(Name("RtfEmbedDotFwd").Device(DEVICE_GPU), RtfEmbedDotFwdOp);

REGISTER_OP("RtfEmbedDotBwd")
    .Input("tables: N * float")
    .Input("ids: int32")
    .Input("dense: float")
    .Input("gout: float")        // (B, cols)
    .Attr("N: int >= 1")
    .Output("gdense: float")     // (B, D)
    .Output("gemb: float")       // (B, N*D): feed to RtfEmbedBwdAdam
    .SetShapeFn([](InferenceContext* c) {
      const int n = c->num_inputs() - 3;
      c->set_output(0, c->input(n + 1));
      DimensionHandle w;
      TF_RETURN_IF_ERROR(c->Multiply(c->Dim(c->input(n + 1), 1), n, &w));
      c->set_output(1, c->Matrix(c->Dim(c->input(n), 0), w));
      return OkStatus();
    });

class RtfEmbedDotBwdOp : public OpKernel {
 public:
  explicit RtfEmbedDotBwdOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    OpInputList tabs;
    OP_REQUIRES_OK(ctx, ctx->input_list("tables", &tabs));
    const int F = tabs.size();
    const Tensor& ids = ctx->input(F);
    const Tensor& dense = ctx->input(F + 1);
    const Tensor& gout = ctx->input(F + 2);
    const int64_t B = ids.dim_size(0);
    const int D = static_cast<int>(dense.dim_size(1));
    TableArgs t = Tables(tabs);
    Tensor *gdense = nullptr, *gemb = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {B, D}, &gdense));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, {B, static_cast<int64_t>(F) * D}, &gemb));
    const int rc = rtf_embed_dot_bwd(t.ptr.data(), t.rows.data(), F, D, ids.flat<int32>().data(), 0,
                                     B, F, 1, dense.flat<float>().data(), D,
                                     gout.flat<float>().data(), gout.dim_size(1),
                                     gdense->flat<float>().data(), D, gemb->flat<float>().data(),
                                     static_cast<int64_t>(F) * D, StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("rtf_embed_dot_bwd rc=", rc));
  }
};
REGISTER_KERNEL_BUILDER(Name("RtfEmbedDotBwd").Device(DEVICE_GPU), RtfEmbedDotBwdOp);

// ----------------------------------------------------------------------------- loss
REGISTER_OP("RtfBce")
    .Input("y_true: float")   // (n) labels
    .Input("y_pred: float")   // (n) probabilities
    .Output("loss: float")    // scalar: -mean(y log(pc + 1e-7) + (1 - y) log(1 - pc + 1e-7))
    .Output("dp: float")      // (n): d loss / d y_pred (0 where the clip saturates)
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->Scalar());
      c->set_output(1, c->input(1));
      return OkStatus();
    });

class RtfBceOp : public OpKernel {
 public:
  explicit RtfBceOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& y = ctx->input(0);
    const Tensor& p = ctx->input(1);
    const int64_t n = p.NumElements();
    OP_REQUIRES(ctx, y.NumElements() == n && n > 0, errors::InvalidArgument("RtfBce: shapes"));
    Tensor *loss = nullptr, *dp = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {}, &loss));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, p.shape(), &dp));
    size_t ws_bytes = 0;
    OP_REQUIRES(ctx, rtf_bce_workspace(n, &ws_bytes) == 0, errors::Internal("rtf_bce_workspace"));
    Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, {static_cast<int64_t>(ws_bytes)}, &ws));
    const int rc = rtf_bce_fwd(y.flat<float>().data(), p.flat<float>().data(), n,
                               loss->flat<float>().data(), dp->flat<float>().data(),
                               ws.flat<uint8>().data(), StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("rtf_bce_fwd rc=", rc));
  }
};
REGISTER_KERNEL_BUILDER(Name("RtfBce").Device(DEVICE_GPU), RtfBceOp);
