// placeholder until the wide-layer tcgen05 kernel lands: reports "shape not served"
#pragma once
#include <string>
#include <vector>
#include "onnx_reader.hpp"
#include "policy_dev.cuh"
namespace go2p {
struct WideModel { int n_layers = 0; };
inline int wide_prepare(const MlpModel&, std::vector<void*>&, WideModel*, std::string&) { return -1; }
inline int wide_launch(const WideModel&, const float*, const int32_t*, float*, double*, long long, bool, uint32_t,
                       const CtrlConst&, int, cudaStream_t, int*, std::string& err) {
  err = "wide tensor-core path not built";
  return 6;
}
}  // namespace go2p
