// Drop-in replacement for the reference's onnx_inference/include/onnx_actor.hpp.
//
// Same public interface (reference: onnx_inference/include/onnx_actor.hpp:14-53) so that
// onnx_controller/src/controller.cpp, controller.hpp and the smoke main compile unchanged:
//   ONNXActor(model_path, std::span<float> observation, std::span<float> action, OrtLoggingLevel)
//   void act();  void print_model_info();  bool check_dims();
// The private ONNX Runtime members (onnx_actor.hpp:56-75) are replaced by a handle to the
// B200-native C ABI (include/go2policy.h): act() is one round trip to a resident sm_100a kernel
// instead of Ort::Session::Run (onnx_actor.cpp:47).  There is no CPU fallback: construction throws
// std::runtime_error when no sm_100 device is present or the graph is unsupported, where the
// reference would throw Ort::Exception (also a std::exception).
#pragma once
#include <memory>
#include <span>
#include <string>

// The reference header pulls OrtLoggingLevel from onnxruntime_cxx_api.h; provide the same C enum
// (ONNX Runtime C API numbering) when ONNX Runtime's headers are not in use.
#ifndef ORT_API_VERSION
extern "C" {
typedef enum OrtLoggingLevel {
  ORT_LOGGING_LEVEL_VERBOSE = 0,
  ORT_LOGGING_LEVEL_INFO = 1,
  ORT_LOGGING_LEVEL_WARNING = 2,
  ORT_LOGGING_LEVEL_ERROR = 3,
  ORT_LOGGING_LEVEL_FATAL = 4
} OrtLoggingLevel;
}
#endif

struct go2p_handle;

class ONNXActor
{
public:
  /// reference: onnx_actor.hpp:29-33 / onnx_actor.cpp:6-36.  The spans are captured (zero-copy
  /// binding semantics): every act() reads observation[0..in) and overwrites action[0..out).
  ONNXActor(
    const std::string & model_path,
    const std::span<float> observation,
    const std::span<float> action,
    OrtLoggingLevel log_level = ORT_LOGGING_LEVEL_WARNING);

  ~ONNXActor();
  ONNXActor(const ONNXActor &) = delete;
  ONNXActor & operator=(const ONNXActor &) = delete;

  /// reference: onnx_actor.hpp:43 / onnx_actor.cpp:38-48
  void act();
  /// reference: onnx_actor.hpp:48 / onnx_actor.cpp:60-66 (same four lines on stdout)
  void print_model_info();
  /// reference: onnx_actor.hpp:53 / onnx_actor.cpp:50-58
  bool check_dims();

  /// Extension (not in the reference): the underlying C-ABI handle, for the fused controller step
  /// (go2p_step_fused) and the batched entry points.
  go2p_handle * native_handle() const { return handle_; }

private:
  OrtLoggingLevel log_level_;
  const std::span<float> observation_;
  const std::span<float> action_;
  const std::string model_path_;
  go2p_handle * handle_ = nullptr;
  int in_dim_ = 0, out_dim_ = 0;
  std::string input_name_, output_name_;
};
