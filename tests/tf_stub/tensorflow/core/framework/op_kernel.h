// TEST INFRASTRUCTURE — not TensorFlow (see op.h in this directory).
#pragma once
#include "tensorflow/core/framework/op.h"

namespace tensorflow {

enum DataType { DT_FLOAT, DT_INT32, DT_INT8, DT_UINT8, DT_INT64 };
constexpr const char* DEVICE_GPU = "GPU";

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64_t>) {}
};

template <class T>
class FlatView {
 public:
  T* data() const;
};

class Tensor {
 public:
  template <class T>
  FlatView<T> flat();
  template <class T>
  FlatView<const T> flat() const;
  int64_t dim_size(int i) const;
  int dims() const;
  int64_t NumElements() const;
  const TensorShape& shape() const;
};

class OpInputList {
 public:
  int size() const;
  const Tensor& operator[](int i) const;
};

class GpuDeviceStub {
 public:
  void* stream() const;
};

class OpKernelConstruction {
 public:
  template <class T>
  Status GetAttr(const char* name, T* value) const;
  void CtxFailure(const Status&) {}
  void CtxFailureWithWarning(const Status&) {}
};

class OpKernelContext {
 public:
  const Tensor& input(int i);
  Status input_list(const char* name, OpInputList* list);
  Status allocate_output(int i, const TensorShape& shape, Tensor** out);
  Status allocate_temp(DataType t, const TensorShape& shape, Tensor* out);
  const GpuDeviceStub& eigen_gpu_device() const;
  void CtxFailure(const Status&) {}
  void CtxFailureWithWarning(const Status&) {}
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* ctx) = 0;
};

#define OP_REQUIRES(CTX, EXP, STATUS)    \
  do {                                   \
    if (!(EXP)) {                        \
      (CTX)->CtxFailure((STATUS));       \
      return;                            \
    }                                    \
  } while (0)
#define OP_REQUIRES_OK(CTX, ...)                      \
  do {                                                \
    ::tensorflow::Status s__(__VA_ARGS__);            \
    if (!s__.ok()) {                                  \
      (CTX)->CtxFailureWithWarning(s__);              \
      return;                                         \
    }                                                 \
  } while (0)

class KernelDefBuilder {
 public:
  explicit KernelDefBuilder(const char*) {}
  KernelDefBuilder& Device(const char*) { return *this; }
  KernelDefBuilder& HostMemory(const char*) { return *this; }
};
inline KernelDefBuilder Name(const char* n) { return KernelDefBuilder(n); }

#define REGISTER_KERNEL_BUILDER(builder, ...)                                  \
  static_assert(sizeof(__VA_ARGS__) > 0, "kernel class must be complete");     \
  static ::tensorflow::KernelDefBuilder RTF_STUB_CAT(rtf_stub_kernel_, __COUNTER__) = (builder)

}  // namespace tensorflow
