"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Minimal ONNX protobuf wire-format reader *and writer* used by the oracle and by
the test-suite.  It is deliberately independent of the C++ reader that ships in
``go2_onnx_controller_b200/csrc/onnx_reader.cpp`` so that the two can be checked
against each other.  No ``onnx`` / ``onnxruntime`` package exists in this image.

Only what the Go2 policy family needs is accepted:
  Gemm(alpha=1, beta=1, transA=0, transB in {0,1}) and Elu(alpha)
chained as  Gemm -> Elu -> ... -> Gemm  (what the reference loads at
reference: onnx_inference/src/cpp/onnx_actor.cpp:16 through Ort::Session).

Field numbers follow the public onnx.proto3 schema (ModelProto.graph = 7,
GraphProto.node = 1 / initializer = 5 / input = 11 / output = 12, ...).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------
# wire-format primitives
# ----------------------------------------------------------------------------


def _varint(buf: bytes, pos: int):
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) for one message body."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            if len(val) != ln:
                raise ValueError("truncated length-delimited field")
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, val


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


# ----------------------------------------------------------------------------
# model description
# ----------------------------------------------------------------------------


@dataclass
class Node:
    op_type: str = ""
    name: str = ""
    inputs: list = field(default_factory=list)
    outputs: list = field(default_factory=list)
    attrs: dict = field(default_factory=dict)


@dataclass
class ValueInfo:
    name: str = ""
    elem_type: int = 0
    shape: list = field(default_factory=list)  # ints, or str for dim_param


@dataclass
class Layer:
    weight: np.ndarray  # [out, in] fp32, row-major (PyTorch nn.Linear layout)
    bias: np.ndarray    # [out] fp32
    elu_alpha: float | None  # None = no activation after this layer


@dataclass
class Policy:
    layers: list
    input_name: str
    output_name: str
    input_shape: list
    output_shape: list
    opset: int
    producer: str

    @property
    def in_dim(self) -> int:
        return int(self.layers[0].weight.shape[1])

    @property
    def out_dim(self) -> int:
        return int(self.layers[-1].weight.shape[0])

    @property
    def dims(self) -> list:
        return [self.in_dim] + [int(l.weight.shape[0]) for l in self.layers]

    @property
    def n_params(self) -> int:
        return sum(l.weight.size + l.bias.size for l in self.layers)


def _parse_tensor(buf: bytes):
    dims, dtype, name, raw, fdata = [], 0, "", None, []
    for fno, wt, val in _fields(buf):
        if fno == 1:  # dims: packed or un-packed varints
            if wt == 0:
                dims.append(_signed64(val))
            else:
                p = 0
                while p < len(val):
                    d, p = _varint(val, p)
                    dims.append(_signed64(d))
        elif fno == 2:
            dtype = val
        elif fno == 4:  # float_data
            if wt == 2:
                fdata.extend(struct.unpack(f"<{len(val) // 4}f", val))
            else:
                fdata.append(struct.unpack("<f", val)[0])
        elif fno == 8:
            name = val.decode()
        elif fno == 9:
            raw = bytes(val)
        elif fno in (13, 14):
            # 13 = external_data (repeated entries), 14 = data_location (enum)
            if fno == 13 or val != 0:
                raise ValueError(f"initializer '{name}': external data is not supported")
    if dtype != 1:
        raise ValueError(f"initializer '{name}': only FLOAT (1) tensors are supported, got {dtype}")
    count = int(np.prod(dims)) if dims else 1
    if raw is not None:
        arr = np.frombuffer(raw, dtype="<f4")
    else:
        arr = np.asarray(fdata, dtype=np.float32)
    if arr.size != count:
        raise ValueError(f"initializer '{name}': {arr.size} values for dims {dims}")
    return name, arr.astype(np.float32).reshape(dims)


def _parse_attr(buf: bytes):
    name, f, i, typ = "", None, None, 0
    for fno, wt, val in _fields(buf):
        if fno == 1:
            name = val.decode()
        elif fno == 2:
            f = struct.unpack("<f", val)[0]
        elif fno == 3:
            i = _signed64(val)
        elif fno == 20:
            typ = val
    if typ == 1 or (typ == 0 and f is not None):
        return name, float(f if f is not None else 0.0)
    if typ == 2 or (typ == 0 and i is not None):
        return name, int(i if i is not None else 0)
    raise ValueError(f"attribute '{name}': unsupported type {typ}")


def _parse_node(buf: bytes) -> Node:
    n = Node()
    for fno, wt, val in _fields(buf):
        if fno == 1:
            n.inputs.append(val.decode())
        elif fno == 2:
            n.outputs.append(val.decode())
        elif fno == 3:
            n.name = val.decode()
        elif fno == 4:
            n.op_type = val.decode()
        elif fno == 5:
            k, v = _parse_attr(val)
            n.attrs[k] = v
        elif fno == 7:
            if val:
                raise ValueError(f"node '{n.name}': non-default domain {val!r}")
    return n


def _parse_value_info(buf: bytes) -> ValueInfo:
    vi = ValueInfo()
    for fno, wt, val in _fields(buf):
        if fno == 1:
            vi.name = val.decode()
        elif fno == 2:  # TypeProto
            for f2, _, v2 in _fields(val):
                if f2 != 1:  # tensor_type
                    continue
                for f3, _, v3 in _fields(v2):
                    if f3 == 1:
                        vi.elem_type = v3
                    elif f3 == 2:  # TensorShapeProto
                        for f4, _, v4 in _fields(v3):
                            if f4 != 1:
                                continue
                            dim = None
                            for f5, _, v5 in _fields(v4):
                                if f5 == 1:
                                    dim = _signed64(v5)
                                elif f5 == 2:
                                    dim = v5.decode()
                            vi.shape.append(dim)
    return vi


def parse_graph(data: bytes):
    graph, opset, producer = None, 0, ""
    for fno, wt, val in _fields(data):
        if fno == 7:
            graph = val
        elif fno == 8:
            dom, ver = "", 0
            for f2, _, v2 in _fields(val):
                if f2 == 1:
                    dom = v2.decode()
                elif f2 == 2:
                    ver = v2
            if dom in ("", "ai.onnx"):
                opset = ver
        elif fno == 2:
            producer = val.decode()
    if graph is None:
        raise ValueError("no GraphProto in model")
    nodes, inits, inputs, outputs = [], {}, [], []
    for fno, wt, val in _fields(graph):
        if fno == 1:
            nodes.append(_parse_node(val))
        elif fno == 5:
            k, v = _parse_tensor(val)
            inits[k] = v
        elif fno == 11:
            inputs.append(_parse_value_info(val))
        elif fno == 12:
            outputs.append(_parse_value_info(val))
    return nodes, inits, inputs, outputs, opset, producer


def load_policy(path_or_bytes) -> Policy:
    """Parse an .onnx MLP policy into a list of (W[out,in], b, elu_alpha) layers."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        data = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as fh:
            data = fh.read()
    nodes, inits, inputs, outputs, opset, producer = parse_graph(data)
    graph_inputs = [vi for vi in inputs if vi.name not in inits]
    if len(graph_inputs) != 1 or len(outputs) != 1:
        raise ValueError("expected exactly one graph input and one graph output")
    cur = graph_inputs[0].name
    layers = []
    pre_ops = []               # observation normaliser in front of the first layer: [(op, constant)] applied in order
    bias_open = False          # last layer came from MatMul / bias-less Gemm and may still take an Add

    def add_layer(n, w, b):
        if w.ndim != 2:
            raise ValueError(f"node '{n.name}': weight must be 2-D")
        if b is None:
            b = np.zeros(w.shape[0], np.float32)
        if b.shape != (w.shape[0],):
            raise ValueError(f"node '{n.name}': bias shape {b.shape} vs weight {w.shape}")
        if layers and layers[-1].weight.shape[0] != w.shape[1]:
            raise ValueError(f"node '{n.name}': inner dimension mismatch")
        layers.append(Layer(np.ascontiguousarray(w, np.float32), np.ascontiguousarray(b, np.float32), None))

    for n in nodes:
        if n.op_type == "Gemm":
            if len(n.inputs) not in (2, 3) or n.inputs[0] != cur:
                raise ValueError(f"node '{n.name}': Gemm is not chained on '{cur}'")
            if n.attrs.get("alpha", 1.0) != 1.0 or n.attrs.get("beta", 1.0) != 1.0:
                raise ValueError(f"node '{n.name}': only alpha=beta=1 Gemm is supported")
            if n.attrs.get("transA", 0) != 0:
                raise ValueError(f"node '{n.name}': transA=1 is not supported")
            w = inits[n.inputs[1]]
            if w.ndim == 2 and n.attrs.get("transB", 0) == 0:
                w = np.ascontiguousarray(w.T)
            add_layer(n, w, inits[n.inputs[2]] if len(n.inputs) == 3 else None)
            bias_open = len(n.inputs) == 2
            cur = n.outputs[0]
        elif n.op_type == "MatMul":
            if len(n.inputs) != 2 or n.inputs[0] != cur or n.inputs[1] not in inits:
                raise ValueError(f"node '{n.name}': MatMul is not chained on '{cur}' with an initializer weight")
            w = inits[n.inputs[1]]
            add_layer(n, np.ascontiguousarray(w.T) if w.ndim == 2 else w, None)
            bias_open = True
            cur = n.outputs[0]
        elif not layers and n.op_type in ("Sub", "Add", "Mul", "Div"):
            if len(n.inputs) != 2 or cur not in n.inputs:
                raise ValueError(f"node '{n.name}': {n.op_type} is not chained on '{cur}'")
            first = n.inputs[0] == cur
            if not first and n.op_type in ("Sub", "Div"):
                raise ValueError(f"node '{n.name}': constant {n.op_type} x is not supported")
            other = n.inputs[1] if first else n.inputs[0]
            if other not in inits:
                raise ValueError(f"node '{n.name}': normaliser operand must be an initializer")
            pre_ops.append((n.op_type, np.asarray(inits[other], np.float32).reshape(-1)))
            cur = n.outputs[0]
        elif n.op_type == "Add":
            if len(n.inputs) != 2 or not bias_open or not layers or cur not in n.inputs:
                raise ValueError(f"node '{n.name}': Add is only supported as the bias of the preceding MatMul")
            other = n.inputs[1] if n.inputs[0] == cur else n.inputs[0]
            if other not in inits or inits[other].shape != (layers[-1].weight.shape[0],):
                raise ValueError(f"node '{n.name}': Add operand must be an initializer of the layer's output width")
            layers[-1].bias = np.ascontiguousarray(inits[other], np.float32)
            bias_open = False
            cur = n.outputs[0]
        elif n.op_type == "Identity":
            if len(n.inputs) != 1 or n.inputs[0] != cur:
                raise ValueError(f"node '{n.name}': Identity is not chained on '{cur}'")
            cur = n.outputs[0]
        elif n.op_type == "Elu":
            if not layers or n.inputs[0] != cur or layers[-1].elu_alpha is not None:
                raise ValueError(f"node '{n.name}': Elu must directly follow a Gemm")
            layers[-1].elu_alpha = float(n.attrs.get("alpha", 1.0))
            bias_open = False
            cur = n.outputs[0]
        elif n.op_type == "Relu":
            if not layers or n.inputs[0] != cur or layers[-1].elu_alpha is not None:
                raise ValueError(f"node '{n.name}': Relu must directly follow a Gemm")
            layers[-1].elu_alpha = 0.0      # Elu(alpha = 0) == max(x, 0) up to the sign of zero
            bias_open = False
            cur = n.outputs[0]
        else:
            raise ValueError(f"node '{n.name}': unsupported op_type '{n.op_type}'")
    if not layers:
        raise ValueError("graph has no Gemm node")
    if cur != outputs[0].name:
        raise ValueError("graph output is not produced by the Gemm/Elu chain")
    if layers[-1].elu_alpha is not None:
        pass  # allowed: activation on the output layer
    pol = Policy(layers, graph_inputs[0].name, outputs[0].name,
                 graph_inputs[0].shape, outputs[0].shape, opset, producer)
    pol.pre_ops = pre_ops
    return pol


# ----------------------------------------------------------------------------
# writer (tests + synthetic wide policy of BASELINE.json configs[4])
# ----------------------------------------------------------------------------


def _enc_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(fno: int, payload: bytes) -> bytes:
    return _enc_varint((fno << 3) | 2) + _enc_varint(len(payload)) + payload


def _vi(fno: int, v: int) -> bytes:
    return _enc_varint((fno << 3) | 0) + _enc_varint(v)


def _f32(fno: int, v: float) -> bytes:
    return _enc_varint((fno << 3) | 5) + struct.pack("<f", v)


def _tensor(name: str, arr: np.ndarray, packed_dims: bool, use_float_data: bool) -> bytes:
    arr = np.ascontiguousarray(arr, dtype="<f4")
    out = b""
    if packed_dims:
        out += _ld(1, b"".join(_enc_varint(d) for d in arr.shape))
    else:
        for d in arr.shape:
            out += _vi(1, d)
    out += _vi(2, 1)
    if use_float_data:
        out += _ld(4, arr.tobytes())
    out += _ld(8, name.encode())
    if not use_float_data:
        out += _ld(9, arr.tobytes())
    return out


def _value_info(name: str, shape) -> bytes:
    dims = b""
    for d in shape:
        if isinstance(d, str):
            dims += _ld(1, _ld(2, d.encode()))
        else:
            dims += _ld(1, _vi(1, d))
    tensor_type = _vi(1, 1) + _ld(2, dims)
    return _ld(1, name.encode()) + _ld(2, _ld(1, tensor_type))


def write_mlp_onnx(weights, biases, elu_alpha=1.0, *, batch=1, trans_b=True,
                   packed_dims=False, use_float_data=False,
                   input_name="observation", output_name="action",
                   final_activation=False, form="gemm", identity_tail=False, act_op="Elu", pre=()) -> bytes:
    """Serialise a Gemm/Elu MLP the way torch.onnx.export names things.

    weights[i] is [out, in].  ``batch`` may be an int or a symbolic dim string.  ``act_op`` "Relu" writes Relu nodes
    instead of Elu; ``pre`` = [(op, constant)] writes an observation normaliser (Sub / Add / Mul / Div with an
    initializer) in front of the first layer.
    """
    nodes, inits = b"", b""
    cur = input_name
    for q, (op, const) in enumerate(pre):
        cname, oname = f"norm.{q}", f"/norm/{op}_{q}_output_0"
        nodes += _ld(1, _ld(1, cur.encode()) + _ld(1, cname.encode()) + _ld(2, oname.encode())
                     + _ld(3, f"/norm/{op}_{q}".encode()) + _ld(4, op.encode()))
        inits += _ld(5, _tensor(cname, np.asarray(const, np.float32), packed_dims, use_float_data))
        cur = oname
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        idx = 2 * i
        wname, bname = f"{idx}.weight", f"{idx}.bias"
        last = i == n - 1
        tail_name = output_name if not identity_tail else "/tail/Identity_input"
        gemm_out = tail_name if (last and not final_activation) else f"/{idx}/Gemm_output_0"
        if form == "matmul_add" or (form == "mixed" and i % 2 == 1):
            # x @ W[in,out] followed by Add(bias): the other way exporters spell a Linear layer
            mm_out = f"/{idx}/MatMul_output_0"
            node = (_ld(1, cur.encode()) + _ld(1, wname.encode()) + _ld(2, mm_out.encode())
                    + _ld(3, f"/{idx}/MatMul".encode()) + _ld(4, b"MatMul"))
            nodes += _ld(1, node)
            node = (_ld(1, bname.encode()) + _ld(1, mm_out.encode()) + _ld(2, gemm_out.encode())
                    + _ld(3, f"/{idx}/Add".encode()) + _ld(4, b"Add"))
            nodes += _ld(1, node)
            inits += _ld(5, _tensor(wname, np.ascontiguousarray(np.asarray(w, np.float32).T), packed_dims, use_float_data))
            inits += _ld(5, _tensor(bname, np.asarray(b, np.float32), packed_dims, use_float_data))
            cur = gemm_out
            if not last or final_activation:
                elu_out = tail_name if last else f"/{idx + 1}/Elu_output_0"
                node = (_ld(1, cur.encode()) + _ld(2, elu_out.encode()) + _ld(3, f"/{idx + 1}/{act_op}".encode())
                        + _ld(4, act_op.encode())
                        + (_ld(5, _ld(1, b"alpha") + _f32(2, float(elu_alpha)) + _vi(20, 1)) if act_op == "Elu" else b""))
                nodes += _ld(1, node)
                cur = elu_out
            continue
        node = (_ld(1, cur.encode()) + _ld(1, wname.encode()) + _ld(1, bname.encode())
                + _ld(2, gemm_out.encode()) + _ld(3, f"/{idx}/Gemm".encode()) + _ld(4, b"Gemm")
                + _ld(5, _ld(1, b"alpha") + _f32(2, 1.0) + _vi(20, 1))
                + _ld(5, _ld(1, b"beta") + _f32(2, 1.0) + _vi(20, 1))
                + _ld(5, _ld(1, b"transB") + _vi(3, 1 if trans_b else 0) + _vi(20, 2)))
        nodes += _ld(1, node)
        wt = np.asarray(w, np.float32)
        inits += _ld(5, _tensor(wname, wt if trans_b else wt.T, packed_dims, use_float_data))
        inits += _ld(5, _tensor(bname, np.asarray(b, np.float32), packed_dims, use_float_data))
        cur = gemm_out
        if not last or final_activation:
            elu_out = tail_name if last else f"/{idx + 1}/Elu_output_0"
            node = (_ld(1, cur.encode()) + _ld(2, elu_out.encode()) + _ld(3, f"/{idx + 1}/{act_op}".encode())
                    + _ld(4, act_op.encode())
                    + (_ld(5, _ld(1, b"alpha") + _f32(2, float(elu_alpha)) + _vi(20, 1)) if act_op == "Elu" else b""))
            nodes += _ld(1, node)
            cur = elu_out
    if identity_tail:
        nodes += _ld(1, _ld(1, cur.encode()) + _ld(2, output_name.encode()) + _ld(3, b"/tail/Identity") + _ld(4, b"Identity"))
    in_dim = int(np.asarray(weights[0]).shape[1])
    out_dim = int(np.asarray(weights[-1]).shape[0])
    graph = (nodes + _ld(2, b"main_graph") + inits
             + _ld(11, _value_info(input_name, [batch, in_dim]))
             + _ld(12, _value_info(output_name, [batch, out_dim])))
    model = (_vi(1, 8) + _ld(2, b"go2policy-b200-writer") + _ld(3, b"1")
             + _ld(7, graph) + _ld(8, _vi(2, 17)))
    return model


def make_wide_policy(seed: int = 5, dims=(245, 1024, 512, 256, 12)):
    """BASELINE.json configs[4]: 5-frame history input (49*5 = 245), ELU MLP
    1024-512-256, PyTorch nn.Linear default init U(-1/sqrt(fan_in), 1/sqrt(fan_in))."""
    rng = np.random.default_rng(seed)
    ws, bs = [], []
    for fan_in, fan_out in zip(dims[:-1], dims[1:]):
        bound = 1.0 / np.sqrt(fan_in)
        ws.append(rng.uniform(-bound, bound, size=(fan_out, fan_in)).astype(np.float32))
        bs.append(rng.uniform(-bound, bound, size=(fan_out,)).astype(np.float32))
    return ws, bs
