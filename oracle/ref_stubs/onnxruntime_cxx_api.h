// TEST INFRASTRUCTURE -- minimal stand-in for the ONNX Runtime C++ API surface that the reference's
// onnx_inference/src/cpp/onnx_actor.cpp touches (onnx_actor.cpp:12-16,21-35,47), so that the reference's own
// ONNXActor can be compiled here without the un-vendored onnxruntime 1.20.1 binary.  Session::Run evaluates the
// graph with the plain-C restatement (oracle_mlp.c): A7 stays "parity unpinned"; what this pins is everything the
// reference's wrapper and controller do AROUND Run.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "oracle.h"

#define ORT_API_VERSION 20
extern "C" {
typedef enum OrtLoggingLevel {
  ORT_LOGGING_LEVEL_VERBOSE = 0, ORT_LOGGING_LEVEL_INFO = 1, ORT_LOGGING_LEVEL_WARNING = 2,
  ORT_LOGGING_LEVEL_ERROR = 3, ORT_LOGGING_LEVEL_FATAL = 4
} OrtLoggingLevel;
typedef enum OrtAllocatorType { OrtInvalidAllocator = -1, OrtDeviceAllocator = 0, OrtArenaAllocator = 1 } OrtAllocatorType;
typedef enum OrtMemType { OrtMemTypeCPUInput = -2, OrtMemTypeCPUOutput = -1, OrtMemTypeDefault = 0 } OrtMemType;
}

namespace Ort {
struct Exception : std::runtime_error { using std::runtime_error::runtime_error; };
struct Env { Env(OrtLoggingLevel, const char*) {} };
struct SessionOptions { SessionOptions() = default; explicit SessionOptions(std::nullptr_t) {} };
struct RunOptions { RunOptions() = default; explicit RunOptions(std::nullptr_t) {} };
struct MemoryInfo { static MemoryInfo CreateCpu(OrtAllocatorType, OrtMemType) { return MemoryInfo{}; } };
struct AllocatorWithDefaultOptions {};
struct AllocatedStringPtr { std::string s; const char* get() const { return s.c_str(); } };
struct TensorTypeAndShapeInfo { std::vector<int64_t> shape; std::vector<int64_t> GetShape() const { return shape; } };
struct TypeInfo { std::vector<int64_t> shape; TensorTypeAndShapeInfo GetTensorTypeAndShapeInfo() const { return {shape}; } };

struct Value {
  float* data = nullptr; size_t count = 0; std::vector<int64_t> shape;
  template <class T>
  static Value CreateTensor(const MemoryInfo&, T* p, size_t n, const int64_t* shp, size_t rank) {
    Value v; v.data = p; v.count = n; v.shape.assign(shp, shp + rank); return v;
  }
};

struct Session {
  orc_model* m = nullptr;
  Session(Env&, const char* path, const SessionOptions&) {
    char err[256];
    if (orc_load(path, &m, err, sizeof err) != 0) throw Exception(err);
  }
  ~Session() { if (m) orc_free(m); }
  Session(const Session&) = delete;
  AllocatedStringPtr GetInputNameAllocated(size_t, AllocatorWithDefaultOptions&) const { return {m->input_name}; }
  AllocatedStringPtr GetOutputNameAllocated(size_t, AllocatorWithDefaultOptions&) const { return {m->output_name}; }
  TypeInfo GetInputTypeInfo(size_t) const { return {std::vector<int64_t>(m->input_shape, m->input_shape + m->input_rank)}; }
  TypeInfo GetOutputTypeInfo(size_t) const { return {std::vector<int64_t>(m->output_shape, m->output_shape + m->output_rank)}; }
  void Run(const RunOptions&, const char* const* in_names, const Value* in, size_t n_in, const char* const* out_names,
           Value* out, size_t n_out) {
    if (n_in != 1 || n_out != 1 || std::strcmp(in_names[0], m->input_name) || std::strcmp(out_names[0], m->output_name))
      throw Exception("stub ORT: unexpected feeds/fetches");
    if ((int64_t)in->count != m->dims[0] || (int64_t)out->count != m->dims[m->n_layers]) throw Exception("stub ORT: tensor size");
    orc_forward_f32(m, in->data, out->data);
  }
};
}  // namespace Ort
