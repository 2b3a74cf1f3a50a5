"""The drop-in claim, checked where it can be without ROS 2 (SURVEY.md 8b "build-system edge"): the replacement
`onnx_inference` CMake package configures, builds (nvcc cross-compiles sm_100a) and installs; a consumer that does
`find_package(onnx_inference)` and links `onnx_inference::onnx_actor onnxruntime` like the reference's node
(onnx_controller/CMakeLists.txt:45-51) compiles the reference's OWN controller.cpp UNCHANGED against the installed
`onnx_actor.hpp` and links -- the bare `onnxruntime` item resolved by the shim library."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0, " ".join(cmd) + "\n" + r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.skipif(shutil.which("cmake") is None or shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"),
                    reason="needs cmake and nvcc")
@pytest.mark.timeout(900)
def test_package_installs_and_reference_node_links_unchanged(tmp_path):
    env_nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    build, prefix = tmp_path / "build", tmp_path / "install"
    _run(["cmake", "-S", os.path.join(ROOT, "packaging", "onnx_inference"), "-B", str(build), f"-DCMAKE_INSTALL_PREFIX={prefix}",
          f"-DCMAKE_CUDA_COMPILER={env_nvcc}", "-DCMAKE_BUILD_TYPE=Release"])
    _run(["cmake", "--build", str(build), "-j", "8"])
    _run(["cmake", "--install", str(build)])
    for rel in ("lib/libonnx_actor.so", "lib/libgo2policy.so", "lib/libonnxruntime.so", "include/onnx_actor.hpp",
                "include/go2policy.h", "share/onnx_inference/data/model.onnx", "bin/onnx_inference"):
        assert (prefix / rel).exists(), rel
    # the shim carries no ONNX Runtime: a handful of symbols, none of them Ort*
    nm = _run(["nm", "-D", "--defined-only", str(prefix / "lib" / "libonnxruntime.so")])
    assert "go2p_onnxruntime_shim_info" in nm and "Ort" not in nm
    if not os.path.exists(os.path.join(REF, "onnx_controller", "src", "controller.cpp")):
        pytest.skip("reference tree not present (GPU box): package build checked, node link check needs /root/reference")
    cbuild = tmp_path / "consumer"
    _run(["cmake", "-S", os.path.join(ROOT, "packaging", "consumer"), "-B", str(cbuild), f"-DCMAKE_PREFIX_PATH={prefix}",
          f"-DREF_ROOT={REF}", f"-DSTUBS={os.path.join(ROOT, 'oracle', 'ref_stubs')}"])
    _run(["cmake", "--build", str(cbuild), "-j", "4"])
    exe = cbuild / "controller"
    assert exe.exists()
    ldd = _run(["ldd", str(exe)])
    assert "libonnx_actor.so" in ldd and "libgo2policy.so" in ldd
    # the node's only policy symbols are ONNXActor's: no Ort:: reference survives in the reference's translation unit
    und = _run(["nm", "-C", "--undefined-only", str(exe)])
    assert "ONNXActor::act()" in und and "Ort::" not in und
