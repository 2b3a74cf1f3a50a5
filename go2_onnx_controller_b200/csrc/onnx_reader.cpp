// See onnx_reader.hpp.  Protobuf wire format is decoded by hand; field numbers are
// those of the public onnx.proto3 schema (ModelProto.graph=7, GraphProto.node=1,
// .initializer=5, .input=11, .output=12, NodeProto.{input=1,output=2,name=3,
// op_type=4,attribute=5,domain=7}, AttributeProto.{name=1,f=2,i=3,type=20},
// TensorProto.{dims=1,data_type=2,float_data=4,name=8,raw_data=9,external_data=13,
// data_location=14}).
#include "onnx_reader.hpp"

#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>

namespace go2p {
namespace {

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool done() const { return p >= end; }
};

[[noreturn]] void fail(const std::string& what) { throw std::runtime_error("onnx: " + what); }

uint64_t varint(Cursor& c) {
  uint64_t v = 0;
  for (int shift = 0; shift < 64; shift += 7) {
    if (c.done()) fail("truncated varint");
    const uint8_t b = *c.p++;
    v |= uint64_t(b & 0x7F) << shift;
    if (!(b & 0x80)) return v;
  }
  fail("varint longer than 64 bits");
}

struct Field {
  int number = 0;
  int wire = 0;
  uint64_t scalar = 0;   // wire 0
  Cursor bytes{nullptr, nullptr};   // wire 1, 2, 5
  std::string str() const { return std::string(reinterpret_cast<const char*>(bytes.p), size_t(bytes.end - bytes.p)); }
};

bool next(Cursor& c, Field& f) {
  if (c.done()) return false;
  const uint64_t key = varint(c);
  f.number = int(key >> 3);
  f.wire = int(key & 7);
  size_t n = 0;
  switch (f.wire) {
    case 0: f.scalar = varint(c); return true;
    case 1: n = 8; break;
    case 5: n = 4; break;
    case 2: n = size_t(varint(c)); break;
    default: fail("unsupported protobuf wire type " + std::to_string(f.wire));
  }
  if (size_t(c.end - c.p) < n) fail("truncated field " + std::to_string(f.number));
  f.bytes = Cursor{c.p, c.p + n};
  c.p += n;
  return true;
}

struct Tensor {
  std::string name;
  std::vector<int64_t> dims;
  std::vector<float> data;
};

Tensor read_tensor(Cursor c) {
  Tensor t;
  int dtype = 0;
  Cursor raw{nullptr, nullptr};
  bool have_raw = false;
  std::vector<float> fdata;
  Field f;
  while (next(c, f)) {
    switch (f.number) {
      case 1:
        if (f.wire == 0) t.dims.push_back(int64_t(f.scalar));
        else { Cursor d = f.bytes; while (!d.done()) t.dims.push_back(int64_t(varint(d))); }
        break;
      case 2: dtype = int(f.scalar); break;
      case 4: {
        const size_t n = size_t(f.bytes.end - f.bytes.p) / 4;
        const size_t old = fdata.size();
        fdata.resize(old + n);
        std::memcpy(fdata.data() + old, f.bytes.p, n * 4);
        break;
      }
      case 8: t.name = f.str(); break;
      case 9: raw = f.bytes; have_raw = true; break;
      case 13: fail("initializer '" + t.name + "' uses external data (unsupported)");
      case 14: if (f.scalar != 0) fail("initializer '" + t.name + "' uses external data (unsupported)"); break;
      default: break;
    }
  }
  if (dtype != 1) fail("initializer '" + t.name + "' has data_type " + std::to_string(dtype) + "; only FLOAT(1) is supported");
  int64_t count = 1;
  for (int64_t d : t.dims) {
    if (d <= 0) fail("initializer '" + t.name + "' has a non-positive dim");
    count *= d;
  }
  if (have_raw) {
    if (int64_t(raw.end - raw.p) != count * 4) fail("initializer '" + t.name + "': raw_data size does not match dims");
    t.data.resize(size_t(count));
    std::memcpy(t.data.data(), raw.p, size_t(count) * 4);   // ONNX raw_data is little-endian; so is every CUDA host
  } else {
    if (int64_t(fdata.size()) != count) fail("initializer '" + t.name + "': float_data size does not match dims");
    t.data = std::move(fdata);
  }
  return t;
}

struct Node {
  std::string op, name, output;
  std::vector<std::string> inputs;
  std::map<std::string, float> fattr;
  std::map<std::string, int64_t> iattr;
};

Node read_node(Cursor c) {
  Node n;
  Field f;
  while (next(c, f)) {
    switch (f.number) {
      case 1: n.inputs.push_back(f.str()); break;
      case 2: if (n.output.empty()) n.output = f.str(); else fail("node with several outputs"); break;
      case 3: n.name = f.str(); break;
      case 4: n.op = f.str(); break;
      case 7: if (!f.str().empty() && f.str() != "ai.onnx") fail("node '" + n.name + "' is in domain '" + f.str() + "'"); break;
      case 5: {
        Cursor a = f.bytes;
        Field g;
        std::string an;
        float fv = 0.f; int64_t iv = 0; int type = 0; bool hf = false, hi = false;
        while (next(a, g)) {
          if (g.number == 1) an = g.str();
          else if (g.number == 2 && g.wire == 5) { std::memcpy(&fv, g.bytes.p, 4); hf = true; }
          else if (g.number == 3 && g.wire == 0) { iv = int64_t(g.scalar); hi = true; }
          else if (g.number == 20) type = int(g.scalar);
        }
        if (type == 1 || (type == 0 && hf)) n.fattr[an] = fv;
        else if (type == 2 || (type == 0 && hi)) n.iattr[an] = iv;
        else fail("node '" + n.name + "': attribute '" + an + "' has unsupported type " + std::to_string(type));
        break;
      }
      default: break;
    }
  }
  return n;
}

struct ValueInfo {
  std::string name;
  std::vector<int64_t> shape;
  int elem_type = 0;
};

ValueInfo read_value_info(Cursor c) {
  ValueInfo vi;
  Field f;
  while (next(c, f)) {
    if (f.number == 1) vi.name = f.str();
    if (f.number != 2) continue;
    Cursor type = f.bytes; Field a;
    while (next(type, a)) {
      if (a.number != 1) continue;                    // TypeProto.tensor_type
      Cursor tt = a.bytes; Field b;
      while (next(tt, b)) {
        if (b.number == 1) vi.elem_type = int(b.scalar);
        if (b.number != 2) continue;                  // Tensor.shape
        Cursor sh = b.bytes; Field d;
        while (next(sh, d)) {
          if (d.number != 1) continue;                // TensorShapeProto.dim
          Cursor dim = d.bytes; Field e; int64_t v = -1;
          while (next(dim, e)) if (e.number == 1 && e.wire == 0) v = int64_t(e.scalar);
          vi.shape.push_back(v);
        }
      }
    }
  }
  return vi;
}

}  // namespace

MlpModel parse_onnx_mlp(const uint8_t* data, size_t size) {
  Cursor top{data, data + size};
  Cursor graph{nullptr, nullptr};
  MlpModel m;
  Field f;
  while (next(top, f)) {
    if (f.number == 7 && f.wire == 2) graph = f.bytes;
    else if (f.number == 2 && f.wire == 2) m.producer = f.str();
    else if (f.number == 8 && f.wire == 2) {
      Cursor o = f.bytes; Field g; std::string dom; int64_t ver = 0;
      while (next(o, g)) { if (g.number == 1) dom = g.str(); else if (g.number == 2) ver = int64_t(g.scalar); }
      if (dom.empty() || dom == "ai.onnx") m.opset = ver;
    }
  }
  if (!graph.p) fail("file holds no GraphProto (not an ONNX model?)");

  std::vector<Node> nodes;
  std::map<std::string, Tensor> inits;
  std::vector<ValueInfo> inputs, outputs;
  while (next(graph, f)) {
    if (f.wire != 2) continue;
    switch (f.number) {
      case 1: nodes.push_back(read_node(f.bytes)); break;
      case 5: { Tensor t = read_tensor(f.bytes); std::string k = t.name; inits.emplace(std::move(k), std::move(t)); break; }
      case 11: inputs.push_back(read_value_info(f.bytes)); break;
      case 12: outputs.push_back(read_value_info(f.bytes)); break;
      default: break;
    }
  }
  std::vector<ValueInfo> real_inputs;
  for (auto& vi : inputs) if (!inits.count(vi.name)) real_inputs.push_back(vi);
  if (real_inputs.size() != 1 || outputs.size() != 1) fail("expected exactly one graph input and one graph output");
  if (real_inputs[0].elem_type != 1 || outputs[0].elem_type != 1) fail("graph input/output must be FLOAT tensors");
  m.input_name = real_inputs[0].name;
  m.output_name = outputs[0].name;
  m.input_shape = real_inputs[0].shape;
  m.output_shape = outputs[0].shape;

  std::string cur = m.input_name;
  bool bias_open = false;      // the last layer came from a MatMul / bias-less Gemm and may still take an Add
  // Observation normaliser in front of the first Gemm (Sub / Add / Mul / Div with an initializer: x' = (x - mean) / std
  // and friends): kept as x' = x * pre_scale + pre_shift per input column and folded into the first layer's weights
  // and bias in double precision when that layer arrives.  The reference's bundled policy has none.
  std::vector<double> pre_scale, pre_shift;
  auto pre_apply = [&](const Node& n, const Tensor& c, char op) {
    const size_t w = c.data.size();
    if (w == 0) fail("node '" + n.name + "': empty normaliser constant");
    if (pre_scale.empty()) { pre_scale.assign(w, 1.0); pre_shift.assign(w, 0.0); }
    if (w != 1 && pre_scale.size() == 1) { pre_scale.assign(w, pre_scale[0]); pre_shift.assign(w, pre_shift[0]); }
    if (w != 1 && w != pre_scale.size()) fail("node '" + n.name + "': normaliser constant width changes along the chain");
    for (size_t k = 0; k < pre_scale.size(); ++k) {
      const double v = (double)c.data[w == 1 ? 0 : k];
      switch (op) {
        case '+': pre_shift[k] += v; break;
        case '-': pre_shift[k] -= v; break;
        case '*': pre_scale[k] *= v; pre_shift[k] *= v; break;
        default:  pre_scale[k] /= v; pre_shift[k] /= v; break;
      }
    }
  };
  auto add_layer = [&](const Node& n, const Tensor& W, bool w_is_out_in, const Tensor* Bv) {
    if (W.dims.size() != 2) fail("node '" + n.name + "': weight must be 2-D");
    MlpLayer L;
    L.out = int(w_is_out_in ? W.dims[0] : W.dims[1]);
    L.in = int(w_is_out_in ? W.dims[1] : W.dims[0]);
    if (int64_t(W.data.size()) != int64_t(L.out) * L.in) fail("node '" + n.name + "': weight data does not match its dims");
    if (Bv && int64_t(Bv->data.size()) != L.out) fail("node '" + n.name + "': bias length does not match weight");
    if (!m.layers.empty() && m.layers.back().out != L.in) fail("node '" + n.name + "': inner dimension mismatch");
    L.weight.resize(size_t(L.out) * L.in);
    for (int o = 0; o < L.out; ++o)
      for (int k = 0; k < L.in; ++k)
        L.weight[size_t(o) * L.in + k] = w_is_out_in ? W.data[size_t(o) * L.in + k] : W.data[size_t(k) * L.out + o];
    if (Bv) L.bias = Bv->data; else L.bias.assign(size_t(L.out), 0.0f);
    if (m.layers.empty() && !pre_scale.empty()) {
      if (pre_scale.size() != 1 && (int)pre_scale.size() != L.in)
        fail("node '" + n.name + "': observation normaliser width does not match the first layer's input");
      for (int o = 0; o < L.out; ++o) {
        double acc = (double)L.bias[o];
        for (int k = 0; k < L.in; ++k) {
          const size_t q = pre_scale.size() == 1 ? 0 : (size_t)k;
          acc += (double)L.weight[size_t(o) * L.in + k] * pre_shift[q];
          L.weight[size_t(o) * L.in + k] = (float)((double)L.weight[size_t(o) * L.in + k] * pre_scale[q]);
        }
        L.bias[o] = (float)acc;
      }
    }
    m.layers.push_back(std::move(L));
  };
  for (const Node& n : nodes) {
    if (n.op == "Gemm") {
      if ((n.inputs.size() != 3 && n.inputs.size() != 2) || n.inputs[0] != cur)
        fail("node '" + n.name + "': Gemm is not chained on '" + cur + "'");
      auto fa = [&](const char* k, float d) { auto it = n.fattr.find(k); return it == n.fattr.end() ? d : it->second; };
      auto ia = [&](const char* k, int64_t d) { auto it = n.iattr.find(k); return it == n.iattr.end() ? d : it->second; };
      if (fa("alpha", 1.f) != 1.f || fa("beta", 1.f) != 1.f) fail("node '" + n.name + "': only alpha = beta = 1 is supported");
      if (ia("transA", 0) != 0) fail("node '" + n.name + "': transA = 1 is not supported");
      auto wi = inits.find(n.inputs[1]);
      if (wi == inits.end()) fail("node '" + n.name + "': weight must be an initializer");
      const Tensor* Bv = nullptr;
      if (n.inputs.size() == 3) {
        auto bi = inits.find(n.inputs[2]);
        if (bi == inits.end()) fail("node '" + n.name + "': bias must be an initializer");
        Bv = &bi->second;
      }
      add_layer(n, wi->second, ia("transB", 0) != 0, Bv);
      bias_open = Bv == nullptr;
      cur = n.output;
    } else if (n.op == "MatMul") {
      // x @ W with W [in, out] (what exporters emit for Linear layers on >2-D inputs); an Add may supply the bias
      if (n.inputs.size() != 2 || n.inputs[0] != cur) fail("node '" + n.name + "': MatMul is not chained on '" + cur + "'");
      auto wi = inits.find(n.inputs[1]);
      if (wi == inits.end()) fail("node '" + n.name + "': MatMul weight must be an initializer");
      add_layer(n, wi->second, false, nullptr);
      bias_open = true;
      cur = n.output;
    } else if (m.layers.empty() && (n.op == "Sub" || n.op == "Add" || n.op == "Mul" || n.op == "Div")) {
      // observation normaliser: (cur op constant); constant-first is accepted for the commutative ops only
      if (n.inputs.size() != 2) fail("node '" + n.name + "': binary op with " + std::to_string(n.inputs.size()) + " inputs");
      const bool first = n.inputs[0] == cur;
      if (!first && (n.inputs[1] != cur || n.op == "Sub" || n.op == "Div"))
        fail("node '" + n.name + "': " + n.op + " is not chained on '" + cur + "'");
      auto ci = inits.find(first ? n.inputs[1] : n.inputs[0]);
      if (ci == inits.end()) fail("node '" + n.name + "': normaliser operand must be an initializer");
      pre_apply(n, ci->second, n.op == "Sub" ? '-' : n.op == "Add" ? '+' : n.op == "Mul" ? '*' : '/');
      cur = n.output;
    } else if (n.op == "Add") {
      if (n.inputs.size() != 2 || !bias_open || m.layers.empty())
        fail("node '" + n.name + "': Add is only supported as the bias of the preceding MatMul / bias-less Gemm");
      const std::string& other = n.inputs[0] == cur ? n.inputs[1] : n.inputs[0];
      if (n.inputs[0] != cur && n.inputs[1] != cur) fail("node '" + n.name + "': Add is not chained on '" + cur + "'");
      auto bi = inits.find(other);
      if (bi == inits.end() || int64_t(bi->second.data.size()) != m.layers.back().out)
        fail("node '" + n.name + "': Add operand must be an initializer of the layer's output width");
      m.layers.back().bias = bi->second.data;
      bias_open = false;
      cur = n.output;
    } else if (n.op == "Identity") {
      if (n.inputs.size() != 1 || n.inputs[0] != cur) fail("node '" + n.name + "': Identity is not chained on '" + cur + "'");
      cur = n.output;
    } else if (n.op == "Elu") {
      if (m.layers.empty() || n.inputs.size() != 1 || n.inputs[0] != cur || m.layers.back().has_elu)
        fail("node '" + n.name + "': Elu must directly follow a Gemm");
      auto it = n.fattr.find("alpha");
      m.layers.back().has_elu = true;
      m.layers.back().elu_alpha = it == n.fattr.end() ? 1.0f : it->second;
      bias_open = false;
      cur = n.output;
    } else if (n.op == "Relu") {
      // max(x, 0) == Elu with alpha = 0: every kernel's ELU epilogue serves it (the sign of an exact zero may differ)
      if (m.layers.empty() || n.inputs.size() != 1 || n.inputs[0] != cur || m.layers.back().has_elu)
        fail("node '" + n.name + "': Relu must directly follow a Gemm");
      m.layers.back().has_elu = true;
      m.layers.back().elu_alpha = 0.0f;
      bias_open = false;
      cur = n.output;
    } else {
      fail("node '" + n.name + "': unsupported op_type '" + n.op +
           "' (supported: Gemm, MatMul, Add, Elu, Relu, Identity; Sub/Add/Mul/Div with a constant in front of the first layer)");
    }
  }
  if (m.layers.empty()) fail("graph has no Gemm node");
  if (cur != m.output_name) fail("graph output '" + m.output_name + "' is not produced by the Gemm/Elu chain");
  return m;
}

MlpModel load_onnx_mlp(const std::string& path) {
  std::ifstream in(path, std::ios::binary | std::ios::ate);
  if (!in) throw std::runtime_error("io: cannot open model file '" + path + "'");
  const std::streamsize n = in.tellg();
  in.seekg(0);
  std::vector<uint8_t> buf(size_t(n > 0 ? n : 0));
  if (n > 0 && !in.read(reinterpret_cast<char*>(buf.data()), n)) throw std::runtime_error("io: short read on '" + path + "'");
  return parse_onnx_mlp(buf.data(), buf.size());
}

}  // namespace go2p
