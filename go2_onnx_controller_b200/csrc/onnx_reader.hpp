// Minimal ONNX reader for MLP policies (host C++, no dependencies).
//
// Replaces the model parse the reference delegates to Ort::Session
// (reference: onnx_inference/src/cpp/onnx_actor.cpp:16) for exactly the graph
// family the Go2 policy belongs to:  Gemm(alpha=1,beta=1,transA=0) [-> Elu(alpha)] ...
// Anything else is rejected with a message; nothing is silently approximated.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace go2p {

struct MlpLayer {
  int in = 0, out = 0;
  std::vector<float> weight;   // [out][in] row-major (ONNX Gemm B with transB=1 == nn.Linear.weight)
  std::vector<float> bias;     // [out]
  bool has_elu = false;
  float elu_alpha = 1.0f;
};

struct MlpModel {
  std::vector<MlpLayer> layers;
  std::string input_name, output_name;
  std::vector<int64_t> input_shape, output_shape;   // -1 for symbolic dims
  int64_t opset = 0;
  std::string producer;

  int in_dim() const { return layers.front().in; }
  int out_dim() const { return layers.back().out; }
  int64_t n_params() const {
    int64_t n = 0;
    for (auto& l : layers) n += int64_t(l.weight.size() + l.bias.size());
    return n;
  }
};

// Throws std::runtime_error with a descriptive message on any problem.
MlpModel load_onnx_mlp(const std::string& path);
MlpModel parse_onnx_mlp(const uint8_t* data, size_t size);

}  // namespace go2p
