"""Minimal ONNX writer for Gemm(+Elu) policies: how the synthetic wide policy of BASELINE.json configs[4] gets on disk
 for bench.py and for users who want to try other layer shapes.  Plain protobuf wire format,
the field numbers torch.onnx.export emits for the bundled model; no onnx package needed.
The parser that reads it back is csrc/onnx_reader.cpp."""
from __future__ import annotations

import struct

import numpy as np


def _varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _field(num: int, wire: int, payload: bytes) -> bytes:
    return _varint((num << 3) | wire) + payload


def _bytes(num: int, data: bytes) -> bytes:
    return _field(num, 2, _varint(len(data)) + data)


def _int(num: int, v: int) -> bytes:
    return _field(num, 0, _varint(v))


def _float_attr(name: str, v: float) -> bytes:
    return _bytes(5, _bytes(1, name.encode()) + _field(2, 5, struct.pack("<f", v)) + _int(20, 1))


def _int_attr(name: str, v: int) -> bytes:
    return _bytes(5, _bytes(1, name.encode()) + _int(3, v) + _int(20, 2))


def _tensor(name: str, a: np.ndarray) -> bytes:
    a = np.ascontiguousarray(a, np.float32)
    dims = b"".join(_int(1, int(d)) for d in a.shape)
    return dims + _int(2, 1) + _bytes(8, name.encode()) + _bytes(9, a.tobytes())


def _io(name: str, batch, width: int) -> bytes:
    def dim(d):
        return _bytes(1, _bytes(2, d.encode()) if isinstance(d, str) else _int(1, int(d)))
    shape = _bytes(2, dim(batch) + dim(width))
    return _bytes(1, name.encode()) + _bytes(2, _bytes(1, _int(1, 1) + shape))


def write_policy(path, weights, biases, elu_alpha: float = 1.0, batch="batch") -> None:
    """weights[i]: [out, in] fp32, biases[i]: [out]; Elu after every layer but the last (the bundled policy's shape)."""
    nodes, inits, cur = b"", b"", "observation"
    for i, (w, b) in enumerate(zip(weights, biases)):
        last = i == len(weights) - 1
        out = "action" if last else f"/{2 * i}/Gemm_output_0"
        nodes += _bytes(1, _bytes(1, cur.encode()) + _bytes(1, f"{2 * i}.weight".encode()) + _bytes(1, f"{2 * i}.bias".encode())
                        + _bytes(2, out.encode()) + _bytes(3, f"/{2 * i}/Gemm".encode()) + _bytes(4, b"Gemm")
                        + _float_attr("alpha", 1.0) + _float_attr("beta", 1.0) + _int_attr("transB", 1))
        inits += _bytes(5, _tensor(f"{2 * i}.weight", w)) + _bytes(5, _tensor(f"{2 * i}.bias", b))
        cur = out
        if not last:
            act = f"/{2 * i + 1}/Elu_output_0"
            nodes += _bytes(1, _bytes(1, cur.encode()) + _bytes(2, act.encode()) + _bytes(3, f"/{2 * i + 1}/Elu".encode())
                            + _bytes(4, b"Elu") + _float_attr("alpha", float(elu_alpha)))
            cur = act
    graph = (nodes + _bytes(2, b"main_graph") + inits + _bytes(11, _io("observation", batch, weights[0].shape[1]))
             + _bytes(12, _io("action", batch, weights[-1].shape[0])))
    model = _int(1, 8) + _bytes(2, b"go2policy-b200") + _bytes(3, b"2") + _bytes(7, graph) + _bytes(8, _int(2, 17))
    with open(path, "wb") as f:
        f.write(model)


def wide_policy(seed: int = 5, dims=(245, 1024, 512, 256, 12)):
    """BASELINE.json configs[4]: 5-frame observation history (49*5 = 245 inputs), ELU MLP 1024-512-256, 12 outputs,
    PyTorch nn.Linear default init U(-1/sqrt(fan_in), 1/sqrt(fan_in)), numpy default_rng(seed)."""
    rng = np.random.default_rng(seed)
    ws, bs = [], []
    for fan_in, fan_out in zip(dims[:-1], dims[1:]):
        bound = 1.0 / np.sqrt(fan_in)
        ws.append(rng.uniform(-bound, bound, size=(fan_out, fan_in)).astype(np.float32))
        bs.append(rng.uniform(-bound, bound, size=(fan_out,)).astype(np.float32))
    return ws, bs
