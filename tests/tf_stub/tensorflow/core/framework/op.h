// TEST INFRASTRUCTURE — not TensorFlow.  A declaration-only stand-in for the handful of TensorFlow
// C++ API names tf_ops/rtf_tf_ops.cc uses, so that the shim can be TYPE-CHECKED
// (g++ -fsyntax-only) in an image without TensorFlow headers: every call into librtf_b200 is then
// checked against include/rtf_b200.h (argument count and types).  Semantics are not modelled.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <string>

namespace tensorflow {

using int32 = int32_t;
using int64 = int64_t;
using int8 = int8_t;
using uint8 = uint8_t;

class Status {
 public:
  bool ok() const { return true; }
};
inline Status OkStatus() { return Status(); }

namespace errors {
template <class... A>
Status Internal(const A&...) { return Status(); }
template <class... A>
Status InvalidArgument(const A&...) { return Status(); }
}  // namespace errors

#define TF_RETURN_IF_ERROR(expr)          \
  do {                                    \
    ::tensorflow::Status s__ = (expr);    \
    if (!s__.ok()) return s__;            \
  } while (0)

namespace shape_inference {
class DimensionHandle {};
class ShapeHandle {};
class InferenceContext {
 public:
  int num_inputs() const;
  ShapeHandle input(int i) const;
  void set_output(int i, ShapeHandle s);
  DimensionHandle Dim(ShapeHandle s, int i);
  DimensionHandle UnknownDim();
  DimensionHandle MakeDim(int64_t v);
  ShapeHandle Matrix(DimensionHandle r, DimensionHandle c);
  ShapeHandle Vector(DimensionHandle n);
  ShapeHandle Vector(int64_t n);
  ShapeHandle Scalar();
  Status Add(DimensionHandle a, DimensionHandle b, DimensionHandle* out);
  Status Multiply(DimensionHandle a, int64_t b, DimensionHandle* out);
};
}  // namespace shape_inference

namespace register_op {
class OpDefBuilderWrapper {
 public:
  explicit OpDefBuilderWrapper(const char*) {}
  OpDefBuilderWrapper& Input(const std::string&) { return *this; }
  OpDefBuilderWrapper& Output(const std::string&) { return *this; }
  OpDefBuilderWrapper& Attr(const std::string&) { return *this; }
  OpDefBuilderWrapper& SetIsStateful() { return *this; }
  OpDefBuilderWrapper& SetShapeFn(std::function<Status(shape_inference::InferenceContext*)>) {
    return *this;
  }
};
}  // namespace register_op

#define RTF_STUB_CAT2(a, b) a##b
#define RTF_STUB_CAT(a, b) RTF_STUB_CAT2(a, b)
#define REGISTER_OP(name)                                                     \
  static ::tensorflow::register_op::OpDefBuilderWrapper RTF_STUB_CAT(rtf_stub_op_, __COUNTER__) = \
      ::tensorflow::register_op::OpDefBuilderWrapper(name)

}  // namespace tensorflow
