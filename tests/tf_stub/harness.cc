// Test harness around tf_ops/yolo_loss_op.cc compiled against tf_stub.h: constructs the registered
// kernel with a set of attributes and runs its Compute() on caller-provided device buffers.
// C interface for ctypes (tests/test_tf_ops.py, tests/test_gpu_tf_op.py).
#include <cstring>
#include <sstream>

#include "tf_stub.h"

using namespace tensorflow;

namespace {
std::string g_error;

// "name=i:3;name=f:0.5;name=b:1;name=lf:1,2,3;name=li:19,38"
bool parse_attrs(const char* spec, AttrMap* out) {
    std::stringstream ss(spec ? spec : "");
    std::string item;
    while (std::getline(ss, item, ';')) {
        if (item.empty()) continue;
        const size_t eq = item.find('='), colon = item.find(':', eq);
        if (eq == std::string::npos || colon == std::string::npos) return false;
        const std::string name = item.substr(0, eq), type = item.substr(eq + 1, colon - eq - 1), val = item.substr(colon + 1);
        AttrValue a;
        if (type == "i") { a.kind = AttrValue::kInt; a.i = std::stoll(val); }
        else if (type == "f") { a.kind = AttrValue::kFloat; a.f = std::stof(val); }
        else if (type == "b") { a.kind = AttrValue::kBool; a.b = val == "1"; }
        else if (type == "lf" || type == "li") {
            a.kind = type == "lf" ? AttrValue::kFloatList : AttrValue::kIntList;
            std::stringstream vs(val);
            std::string v;
            while (std::getline(vs, v, ','))
                if (!v.empty()) { if (type == "lf") a.fl.push_back(std::stof(v)); else a.il.push_back(std::stoi(v)); }
        } else return false;
        (*out)[name] = a;
    }
    return true;
}
}  // namespace

extern "C" {

const char* tfstub_last_error() { return g_error.c_str(); }

// newline-separated "op|inputs;..|outputs;..|attrs;.." of every REGISTER_OP in the library
const char* tfstub_describe_ops() {
    static std::string s;
    s.clear();
    for (auto& kv : Registry::get().ops) {
        s += kv.first + "|";
        for (auto& x : kv.second.inputs) s += x + ";";
        s += "|";
        for (auto& x : kv.second.outputs) s += x + ";";
        s += "|";
        for (auto& x : kv.second.attrs) s += x + ";";
        s += "|" + Registry::get().kernel_device[kv.first] + "\n";
    }
    return s.c_str();
}

// Construct the kernel only (attribute validation; needs no GPU).  0 = ok, 1 = construction failed.
int tfstub_construct(const char* op, const char* attr_spec) {
    g_error.clear();
    AttrMap attrs;
    if (!parse_attrs(attr_spec, &attrs)) { g_error = "bad attr spec"; return 2; }
    auto it = Registry::get().kernels.find(op);
    if (it == Registry::get().kernels.end()) { g_error = std::string("no kernel for ") + op; return 3; }
    OpKernelConstruction c(&attrs);
    std::unique_ptr<OpKernel> k(it->second(&c));
    if (!c.status().ok()) { g_error = c.status().message(); return 1; }
    return 0;
}

// Construct and Compute.  inputs[i] has input_elems[i] elements (flat); outputs[i] are caller-allocated
// device buffers large enough for the op's outputs; temp is device scratch for allocate_temp.
int tfstub_run(const char* op, const char* attr_spec, int n_in, void* const* inputs, const int64_t* input_elems,
               int n_out, void* const* outputs, void* temp, size_t temp_bytes, void* stream) {
    g_error.clear();
    AttrMap attrs;
    if (!parse_attrs(attr_spec, &attrs)) { g_error = "bad attr spec"; return 2; }
    auto it = Registry::get().kernels.find(op);
    if (it == Registry::get().kernels.end()) { g_error = std::string("no kernel for ") + op; return 3; }
    OpKernelConstruction c(&attrs);
    std::unique_ptr<OpKernel> k(it->second(&c));
    if (!c.status().ok()) { g_error = c.status().message(); return 1; }
    std::vector<Tensor> in;
    for (int i = 0; i < n_in; ++i) in.emplace_back(inputs[i], TensorShape({input_elems[i]}));
    OpKernelContext ctx(in, std::vector<void*>(outputs, outputs + n_out), static_cast<char*>(temp), temp_bytes, stream);
    k->Compute(&ctx);
    if (!ctx.status().ok()) { g_error = ctx.status().message(); return 4; }
    return 0;
}

}  // extern "C"
