"""B200-native (sm_100a) implementation of go2_onnx_controller's policy hot path.

CUDA kernels + C ABI live in csrc/ (built in-tree into lib/); this package is the thin Python host
mirror used by the tests and bench.py.  There is no CPU fallback anywhere in this package.
"""
from .actor import DEFAULT_MODEL, Fleet, Go2Controller, ONNXActor, PolicyBatch, default_config  # noqa: F401
from . import build, capi, shard  # noqa: F401
build_native = build.build

__all__ = ["ONNXActor", "Go2Controller", "PolicyBatch", "Fleet", "default_config", "build", "build_native", "capi", "shard", "DEFAULT_MODEL"]
