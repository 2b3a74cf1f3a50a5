#pragma once
#include <cstdlib>
#include <string>
namespace ament_index_cpp {
// the "share" directory of onnx_inference is the package directory of the reference tree itself
inline std::string get_package_share_directory(const std::string& pkg) {
  const char* root = std::getenv("GO2_REF_ROOT");
  return std::string(root ? root : "/root/reference") + "/" + pkg;
}
}
