// ONNXActor over the go2policy C ABI (see include/onnx_actor.hpp).
#include "../../include/onnx_actor.hpp"

#include <iostream>
#include <stdexcept>

#include "../../include/go2policy.h"

ONNXActor::ONNXActor(
  const std::string & model_path,
  const std::span<float> observation,
  const std::span<float> action,
  OrtLoggingLevel log_level)
: log_level_(log_level)
, observation_(observation)
, action_(action)
, model_path_(model_path)
{
  go2p_config cfg;
  go2p_config_default(&cfg);
  cfg.log_level = static_cast<int>(log_level);
  if (go2p_create(model_path.c_str(), &cfg, &handle_) != GO2P_OK)
    throw std::runtime_error(std::string("ONNXActor: ") + go2p_last_error());
  go2p_model_info_t info;
  go2p_model_info(handle_, &info);
  in_dim_ = info.in_dim;
  out_dim_ = info.out_dim;
  input_name_ = info.input_name;
  output_name_ = info.output_name;
  // The reference wraps the caller's buffers with shape[1] elements regardless of the span sizes
  // (onnx_actor.cpp:31-35) and never validates them; a short span there is an out-of-bounds access
  // inside Run().  Here it is an error at construction.
  if (go2p_bind(handle_, observation_.data(), observation_.size(), action_.data(), action_.size()) != GO2P_OK) {
    const std::string msg = go2p_last_error();
    go2p_destroy(handle_);
    handle_ = nullptr;
    throw std::runtime_error("ONNXActor: " + msg);
  }
}

ONNXActor::~ONNXActor()
{
  if (handle_) go2p_destroy(handle_);
}

void ONNXActor::act()
{
  if (go2p_act(handle_) != GO2P_OK) throw std::runtime_error(std::string("ONNXActor::act: ") + go2p_last_error());
}

bool ONNXActor::check_dims()
{
  bool result = true;
  result &= observation_.size() == static_cast<size_t>(in_dim_);
  result &= action_.size() == static_cast<size_t>(out_dim_);
  return result;
}

void ONNXActor::print_model_info()
{
  std::cout << "Input dimension: " << in_dim_ << std::endl;
  std::cout << "Output dimension: " << out_dim_ << std::endl;
  std::cout << "Input name: " << input_name_ << std::endl;
  std::cout << "Output name: " << output_name_ << std::endl;
}
