// Minimal stand-in for the slice of the TensorFlow C++ op API that tf_ops/yolo_loss_op.cc uses
// (TEST INFRASTRUCTURE: TensorFlow is not installable in the build container).  It is functional,
// not just declarative: REGISTER_OP / REGISTER_KERNEL_BUILDER fill registries, OpKernelConstruction
// serves attributes, OpKernelContext serves caller-provided device buffers, so harness.cc can
// construct the real kernel classes and run their Compute() against libyolo_b200.so.
// Semantics follow tensorflow/core/framework/{op.h,op_kernel.h,shape_inference.h}.
#pragma once
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace tensorflow {

class Status {
 public:
    Status() : ok_(true) {}
    explicit Status(std::string msg) : ok_(false), msg_(std::move(msg)) {}
    bool ok() const { return ok_; }
    const std::string& message() const { return msg_; }

 private:
    bool ok_;
    std::string msg_;
};
inline Status OkStatus() { return Status(); }

namespace errors {
template <typename... A>
Status make(const char* kind, A... a) {
    std::ostringstream os;
    os << kind << ": ";
    (void)std::initializer_list<int>{((os << a), 0)...};
    return Status(os.str());
}
template <typename... A> Status InvalidArgument(A... a) { return make("InvalidArgument", a...); }
template <typename... A> Status Internal(A... a) { return make("Internal", a...); }
template <typename... A> Status ResourceExhausted(A... a) { return make("ResourceExhausted", a...); }
template <typename... A> Status NotFound(A... a) { return make("NotFound", a...); }
}  // namespace errors

enum DataType { DT_FLOAT = 1, DT_DOUBLE = 2, DT_UINT8 = 4 };
using uint8 = unsigned char;
constexpr const char* DEVICE_GPU = "GPU";
constexpr const char* DEVICE_CPU = "CPU";

class TensorShape {
 public:
    TensorShape() {}
    TensorShape(std::initializer_list<int64_t> d) : dims_(d) {}
    explicit TensorShape(std::vector<int64_t> d) : dims_(std::move(d)) {}
    int64_t num_elements() const {
        int64_t n = 1;
        for (auto d : dims_) n *= d;
        return n;
    }
    const std::vector<int64_t>& dims() const { return dims_; }

 private:
    std::vector<int64_t> dims_;
};

class Tensor {
 public:
    template <typename T>
    struct Flat {
        T* p;
        T* data() const { return p; }
    };
    Tensor() {}
    Tensor(void* data, TensorShape shape) : data_(data), shape_(std::move(shape)) {}
    template <typename T> Flat<T> flat() { return Flat<T>{static_cast<T*>(data_)}; }
    template <typename T> Flat<const T> flat() const { return Flat<const T>{static_cast<const T*>(data_)}; }
    int64_t NumElements() const { return shape_.num_elements(); }
    const TensorShape& shape() const { return shape_; }

 private:
    void* data_ = nullptr;
    TensorShape shape_;
};

// ---- attributes ---------------------------------------------------------------------------------
struct AttrValue {
    enum Kind { kInt, kFloat, kBool, kFloatList, kIntList } kind = kInt;
    long long i = 0;
    float f = 0.f;
    bool b = false;
    std::vector<float> fl;
    std::vector<int> il;
};
using AttrMap = std::map<std::string, AttrValue>;

class AttrReader {
 public:
    explicit AttrReader(const AttrMap* a) : attrs_(a) {}
    Status GetAttr(const std::string& n, int* v) const { return get(n, AttrValue::kInt, [&](const AttrValue& a) { *v = (int)a.i; }); }
    Status GetAttr(const std::string& n, float* v) const { return get(n, AttrValue::kFloat, [&](const AttrValue& a) { *v = a.f; }); }
    Status GetAttr(const std::string& n, bool* v) const { return get(n, AttrValue::kBool, [&](const AttrValue& a) { *v = a.b; }); }
    Status GetAttr(const std::string& n, std::vector<float>* v) const { return get(n, AttrValue::kFloatList, [&](const AttrValue& a) { *v = a.fl; }); }
    Status GetAttr(const std::string& n, std::vector<int>* v) const { return get(n, AttrValue::kIntList, [&](const AttrValue& a) { *v = a.il; }); }

 private:
    template <typename F>
    Status get(const std::string& n, AttrValue::Kind k, F f) const {
        auto it = attrs_->find(n);
        if (it == attrs_->end()) return errors::NotFound("no attr named '", n, "'");
        if (it->second.kind != k) return errors::InvalidArgument("attr '", n, "' has another type");
        f(it->second);
        return OkStatus();
    }
    const AttrMap* attrs_;
};

class OpKernelConstruction : public AttrReader {
 public:
    explicit OpKernelConstruction(const AttrMap* a) : AttrReader(a) {}
    void CtxFailure(const Status& s) {
        if (status_.ok()) status_ = s;
    }
    const Status& status() const { return status_; }

 private:
    Status status_;
};

// ---- execution context --------------------------------------------------------------------------
struct PlatformStreamHandle {
    void* stream = nullptr;
};
class StreamStub {
 public:
    explicit StreamStub(void* s) { h_.stream = s; }
    PlatformStreamHandle platform_specific_handle() const { return h_; }

 private:
    PlatformStreamHandle h_;
};
class DeviceContext {
 public:
    explicit DeviceContext(void* s) : stream_(s) {}
    StreamStub* stream() { return &stream_; }

 private:
    StreamStub stream_;
};

class OpKernelContext {
 public:
    OpKernelContext(std::vector<Tensor> inputs, std::vector<void*> output_buffers, char* temp, size_t temp_bytes,
                    void* stream)
        : inputs_(std::move(inputs)), out_buf_(std::move(output_buffers)), outputs_(out_buf_.size()),
          temp_(temp), temp_left_(temp_bytes), dc_(stream) {}
    const Tensor& input(int i) const { return inputs_.at(i); }
    Status allocate_output(int i, const TensorShape& shape, Tensor** out) {
        if (i < 0 || i >= (int)out_buf_.size() || out_buf_[i] == nullptr)
            return errors::InvalidArgument("output ", i, " has no buffer");
        outputs_[i] = Tensor(out_buf_[i], shape);
        *out = &outputs_[i];
        return OkStatus();
    }
    Status allocate_temp(DataType, const TensorShape& shape, Tensor* out) {
        const size_t need = (size_t)shape.num_elements();
        if (need > temp_left_) return errors::ResourceExhausted("temp of ", need, " bytes");
        *out = Tensor(temp_, shape);
        temp_ += need;
        temp_left_ -= need;
        return OkStatus();
    }
    DeviceContext* op_device_context() { return &dc_; }
    void CtxFailure(const Status& s) {
        if (status_.ok()) status_ = s;
    }
    const Status& status() const { return status_; }
    const Tensor& output(int i) const { return outputs_.at(i); }

 private:
    std::vector<Tensor> inputs_;
    std::vector<void*> out_buf_;
    std::vector<Tensor> outputs_;
    char* temp_;
    size_t temp_left_;
    DeviceContext dc_;
    Status status_;
};

class OpKernel {
 public:
    explicit OpKernel(OpKernelConstruction*) {}
    virtual ~OpKernel() {}
    virtual void Compute(OpKernelContext* ctx) = 0;
};

#define OP_REQUIRES(CTX, COND, STATUS)   \
    do {                                 \
        if (!(COND)) {                   \
            (CTX)->CtxFailure((STATUS)); \
            return;                      \
        }                                \
    } while (0)
#define OP_REQUIRES_OK(CTX, EXPR)                    \
    do {                                             \
        ::tensorflow::Status _s = (EXPR);            \
        if (!_s.ok()) {                              \
            (CTX)->CtxFailure(_s);                   \
            return;                                  \
        }                                            \
    } while (0)
#define TF_RETURN_IF_ERROR(EXPR)                     \
    do {                                             \
        ::tensorflow::Status _s = (EXPR);            \
        if (!_s.ok()) return _s;                     \
    } while (0)

// ---- shape inference (compiled, never run by the harness) --------------------------------------
namespace shape_inference {
struct ShapeHandle {
    int rank = -1;
    long long dim0 = -1;
};
class InferenceContext {
 public:
    explicit InferenceContext(const AttrMap* a) : attrs_(a) {}
    ShapeHandle Scalar() { return ShapeHandle{0, -1}; }
    ShapeHandle Vector(long long n) { return ShapeHandle{1, n}; }
    ShapeHandle input(int) { return ShapeHandle{}; }
    void set_output(int i, ShapeHandle s) { outputs[i] = s; }
    Status GetAttr(const std::string& n, int* v) const { return attrs_.GetAttr(n, v); }
    std::map<int, ShapeHandle> outputs;

 private:
    AttrReader attrs_;
};
}  // namespace shape_inference

// ---- registries ----------------------------------------------------------------------------------
struct OpDef {
    std::string name;
    std::vector<std::string> inputs, outputs, attrs;   // the spec strings, verbatim
    std::function<Status(shape_inference::InferenceContext*)> shape_fn;
};
struct Registry {
    std::map<std::string, OpDef> ops;
    std::map<std::string, std::function<OpKernel*(OpKernelConstruction*)>> kernels;   // GPU kernels only
    std::map<std::string, std::string> kernel_device;
    static Registry& get() {
        static Registry r;
        return r;
    }
};

class OpDefBuilder {
 public:
    explicit OpDefBuilder(const char* name) { def_.name = name; }
    OpDefBuilder& Input(const char* s) { def_.inputs.push_back(s); return *this; }
    OpDefBuilder& Output(const char* s) { def_.outputs.push_back(s); return *this; }
    OpDefBuilder& Attr(const char* s) { def_.attrs.push_back(s); return *this; }
    template <typename F>
    OpDefBuilder& SetShapeFn(F f) { def_.shape_fn = f; return *this; }
    const OpDef& def() const { return def_; }

 private:
    OpDef def_;
};
struct OpRegistrar {
    OpRegistrar(const OpDefBuilder& b) { Registry::get().ops[b.def().name] = b.def(); }
};
struct KernelName {
    std::string name, device;
    KernelName& Device(const char* d) { device = d; return *this; }
};
inline KernelName Name(const char* n) { return KernelName{n, ""}; }
struct KernelRegistrar {
    KernelRegistrar(const KernelName& n, std::function<OpKernel*(OpKernelConstruction*)> f) {
        Registry::get().kernels[n.name] = std::move(f);
        Registry::get().kernel_device[n.name] = n.device;
    }
};

#define TFSTUB_CAT2(a, b) a##b
#define TFSTUB_CAT(a, b) TFSTUB_CAT2(a, b)
#define REGISTER_OP(NAME) \
    static ::tensorflow::OpRegistrar TFSTUB_CAT(tfstub_op_, __COUNTER__) = ::tensorflow::OpDefBuilder(NAME)
#define REGISTER_KERNEL_BUILDER(KNAME, ...)                                                     \
    static ::tensorflow::KernelRegistrar TFSTUB_CAT(tfstub_kernel_, __COUNTER__)(               \
        ::tensorflow::KNAME, [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { \
            return new __VA_ARGS__(c);                                                          \
        })

}  // namespace tensorflow
