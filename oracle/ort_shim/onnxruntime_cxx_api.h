// TEST INFRASTRUCTURE -- a stand-in for <onnxruntime_cxx_api.h> (ONNX Runtime 1.20 is not available offline).
//
// Purpose: let the reference's own src/tts_onnx.cpp compile UNMODIFIED (oracle/Makefile -> oracle/_ref/tts_host_ref), so
// that its host logic -- prompt assembly (src/tts_onnx.cpp:442-539), loops A and B (:782-872), the sampler filters
// (:907-950), the KV round trip (:615-732) and the vocoder hand-off (:759-776) -- runs here and pins the Python
// restatement in oracle/qwen3_tts_oracle.py (tests/test_ref_host_pin.py).
//
// Only the part of the Ort:: C++ API that src/tts_onnx.cpp uses exists. Session::Run does not execute a graph: it
// dispatches on the model file's stem to a deterministic STUB graph (stub_graphs.h) whose outputs have the contract's
// shapes and are a hash of the inputs, and appends one line per call (graph, input names/shapes/digests) to a trace.
// oracle/stub_graphs.py computes the same stubs in numpy, so two host implementations can be compared call by call.
// Nothing here is linked into or called by the product.
#pragma once
#include <cstdint>
#include <cstring>
#include <exception>
#include <memory>
#include <string>
#include <vector>

enum OrtLoggingLevel { ORT_LOGGING_LEVEL_VERBOSE, ORT_LOGGING_LEVEL_INFO, ORT_LOGGING_LEVEL_WARNING, ORT_LOGGING_LEVEL_ERROR, ORT_LOGGING_LEVEL_FATAL };
enum OrtAllocatorType { OrtInvalidAllocator = -1, OrtDeviceAllocator = 0, OrtArenaAllocator = 1 };
enum OrtMemType { OrtMemTypeCPUInput = -2, OrtMemTypeCPUOutput = -1, OrtMemTypeCPU = OrtMemTypeCPUOutput, OrtMemTypeDefault = 0 };
enum GraphOptimizationLevel { ORT_DISABLE_ALL = 0, ORT_ENABLE_BASIC = 1, ORT_ENABLE_EXTENDED = 2, ORT_ENABLE_ALL = 99 };

namespace Ort {

struct Exception : std::exception {
    explicit Exception(std::string m) : msg_(std::move(m)) {}
    const char* what() const noexcept override { return msg_.c_str(); }
    std::string msg_;
};

struct Env { Env(OrtLoggingLevel, const char*) {} };
struct MemoryInfo {
    static MemoryInfo CreateCpu(OrtAllocatorType, OrtMemType) { return MemoryInfo(); }
};
struct SessionOptions {
    void SetIntraOpNumThreads(int n) { intra_op_threads = n; }
    void SetGraphOptimizationLevel(GraphOptimizationLevel l) { opt_level = l; }
    int intra_op_threads = 0;
    GraphOptimizationLevel opt_level = ORT_DISABLE_ALL;
};
struct RunOptions { RunOptions(std::nullptr_t) {} RunOptions() {} };

struct TensorTypeAndShapeInfo {
    size_t GetElementCount() const { size_t n = 1; for (int64_t d : shape) n *= (size_t)d; return n; }
    std::vector<int64_t> GetShape() const { return shape; }
    std::vector<int64_t> shape;
};

// A tensor: either a view of caller memory (CreateTensor, zero copy like ORT) or an owned buffer (graph outputs).
struct Value {
    enum Kind { F32, I64 };
    Value() = default;
    Value(Value&&) = default;
    Value& operator=(Value&&) = default;
    Value(const Value&) = delete;
    template <typename T>
    static Value CreateTensor(const MemoryInfo&, T* data, size_t count, const int64_t* shape, size_t rank);
    template <typename T>
    T* GetTensorMutableData() { return reinterpret_cast<T*>(ptr); }
    TensorTypeAndShapeInfo GetTensorTypeAndShapeInfo() const { return TensorTypeAndShapeInfo{shape}; }

    static Value Owned(Kind k, std::vector<int64_t> shape) {
        Value v; v.kind = k; v.shape = std::move(shape);
        size_t n = 1; for (int64_t d : v.shape) n *= (size_t)d;
        v.count = n; v.own.assign(n * 8, 0); v.ptr = v.own.data();
        return v;
    }
    Kind kind = F32;
    void* ptr = nullptr;
    size_t count = 0;
    std::vector<int64_t> shape;
    std::vector<unsigned char> own;
};
template <> inline Value Value::CreateTensor<float>(const MemoryInfo&, float* d, size_t n, const int64_t* s, size_t r) {
    Value v; v.kind = F32; v.ptr = d; v.count = n; v.shape.assign(s, s + r); return v;
}
template <> inline Value Value::CreateTensor<int64_t>(const MemoryInfo&, int64_t* d, size_t n, const int64_t* s, size_t r) {
    Value v; v.kind = I64; v.ptr = d; v.count = n; v.shape.assign(s, s + r); return v;
}

struct AllocatorWithDefaultOptions {};
struct FreeDeleter { void operator()(char* p) const { delete[] p; } };
using AllocatedStringPtr = std::unique_ptr<char, FreeDeleter>;

}  // namespace Ort

// implemented in stub_graphs.h (one translation unit includes it with ORT_SHIM_IMPLEMENTATION)
namespace ort_shim {
std::vector<Ort::Value> run_stub(const std::string& stem, const char* const* in_names, const Ort::Value* in, size_t n_in,
                                 const char* const* out_names, size_t n_out);
}

namespace Ort {
struct Session {
    Session(Env&, const char* path, const SessionOptions& o) : opts(o) {
        std::string p(path);
        const size_t s = p.find_last_of('/');
        stem = (s == std::string::npos) ? p : p.substr(s + 1);
        const size_t d = stem.rfind('.');
        if (d != std::string::npos) stem.resize(d);
    }
    std::vector<Value> Run(const RunOptions&, const char* const* in_names, const Value* in, size_t n_in,
                           const char* const* out_names, size_t n_out) {
        return ort_shim::run_stub(stem, in_names, in, n_in, out_names, n_out);
    }
    static AllocatedStringPtr dup(const char* s) { char* p = new char[std::strlen(s) + 1]; std::strcpy(p, s); return AllocatedStringPtr(p); }
    AllocatedStringPtr GetInputNameAllocated(size_t, AllocatorWithDefaultOptions&) const { return dup("mel"); }
    AllocatedStringPtr GetOutputNameAllocated(size_t, AllocatorWithDefaultOptions&) const { return dup("embedding"); }
    std::string stem;
    SessionOptions opts;
};
}  // namespace Ort
