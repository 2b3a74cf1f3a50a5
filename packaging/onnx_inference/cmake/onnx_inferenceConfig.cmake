# find_package(onnx_inference) without ament: the imported targets the reference's consumers link
# (onnx_inference::onnx_actor, onnx_controller/CMakeLists.txt:45-51) plus a plain `onnxruntime` target for its bare
# link item.
include("${CMAKE_CURRENT_LIST_DIR}/export_onnx_actorExport.cmake")
if(NOT TARGET onnxruntime)
  add_library(onnxruntime INTERFACE IMPORTED)
  target_link_libraries(onnxruntime INTERFACE onnx_inference::onnxruntime)
endif()
set(onnx_inference_FOUND TRUE)
