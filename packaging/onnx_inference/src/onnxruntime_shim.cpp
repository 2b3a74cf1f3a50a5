/* Empty stand-in for libonnxruntime: the reference node's CMakeLists names `onnxruntime` on its link line
 * (onnx_controller/CMakeLists.txt:49) although controller.cpp uses no Ort:: symbol; the policy runs on the
 * go2policy kernels.  One exported symbol so the library is not empty and its purpose can be read with `strings`. */
extern "C" const char* go2p_onnxruntime_shim_info(void) {
  return "onnxruntime shim of go2_onnx_controller_b200: no ONNX Runtime inside; ONNXActor runs on libgo2policy (sm_100a)";
}
