// sentis.cuh -- reader for the Unity Inference Engine `.sentis` model container, inside the library.
//
// The reference loads its network with ModelLoader.Load(_sentisModel) (Assets/Scripts/InferenceEngine/IEExecutor.cs:382)
// from Assets/Resources/Model/yolo11n-seg-sentis.sentis.  The package that defines the format (com.unity.ai.inference
// 2.2.1) is not vendored in the reference; the layout is restated from the file itself:
//     u32 size | FlatBuffer "Program"   { version, ExecutionPlan { values[], inputs, ..., chains[], operators[] } }
//     u32 size | FlatBuffer "Buffer"    { bytes }      (repeated: chunks of the weight blob)
// This file walks exactly what the hot path needs: the constant tensors, the DequantizeUint8 chains (u8 tensor, scale,
// zero point) feeding every Conv / ConvTranspose chain in file order -- the canonical layer order of model.cuh -- and the
// two thresholds the editor script baked into the NonMaxSuppression chain (IEModelEditorConverter.cs:76).
// Host-only code; dequantization is the graph's own (q - zp) * scale, per tensor.
#pragma once

#include <stdint.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace xrseg {

class FlatBuf {   // minimal read-only FlatBuffer accessor; every access is bounds-checked (the bytes come from a caller)
 public:
  FlatBuf(const uint8_t* b, size_t n, size_t base) : b_(b), n_(n), base_(base) {}
  size_t root() const { return base_ + u32(base_); }
  uint32_t u32(size_t p) const { chk(p, 4); uint32_t v; memcpy(&v, b_ + p, 4); return v; }
  int32_t i32(size_t p) const { chk(p, 4); int32_t v; memcpy(&v, b_ + p, 4); return v; }
  uint16_t u16(size_t p) const { chk(p, 2); uint16_t v; memcpy(&v, b_ + p, 2); return v; }
  uint8_t u8(size_t p) const { chk(p, 1); return b_[p]; }
  float f32(size_t p) const { chk(p, 4); float v; memcpy(&v, b_ + p, 4); return v; }
  // absolute offset of field `idx` of `table`, 0 when absent
  size_t field(size_t table, int idx) const {
    const size_t vt = table - static_cast<size_t>(static_cast<int64_t>(i32(table)));
    const uint16_t vt_size = u16(vt);
    const size_t slot = 4 + 2 * static_cast<size_t>(idx);
    if (slot + 2 > vt_size) return 0;
    const uint16_t off = u16(vt + slot);
    return off ? table + off : 0;
  }
  size_t indirect(size_t table, int idx) const {
    const size_t p = field(table, idx);
    return p ? p + u32(p) : 0;
  }
  std::string str(size_t table, int idx) const {
    const size_t p = indirect(table, idx);
    if (!p) return std::string();
    const uint32_t n = u32(p);
    chk(p + 4, n);
    return std::string(reinterpret_cast<const char*>(b_ + p + 4), n);
  }
  std::vector<size_t> tables(size_t table, int idx) const {
    std::vector<size_t> out;
    const size_t p = indirect(table, idx);
    if (!p) return out;
    const uint32_t n = u32(p);
    for (uint32_t i = 0; i < n; ++i) {
      const size_t e = p + 4 + 4 * static_cast<size_t>(i);
      out.push_back(e + u32(e));
    }
    return out;
  }
  std::vector<int32_t> ints(size_t table, int idx) const {
    std::vector<int32_t> out;
    const size_t p = indirect(table, idx);
    if (!p) return out;
    const uint32_t n = u32(p);
    chk(p + 4, static_cast<size_t>(n) * 4);
    out.resize(n);
    if (n) memcpy(out.data(), b_ + p + 4, static_cast<size_t>(n) * 4);
    return out;
  }
  void byte_span(size_t table, int idx, size_t* start, size_t* len) const {
    const size_t p = indirect(table, idx);
    XR_CHECK(p != 0, "sentis: weight chunk without bytes");
    *len = u32(p);
    *start = p + 4;
    chk(*start, *len);
  }

 private:
  void chk(size_t p, size_t n) const { XR_CHECK(p + n <= n_ && p + n >= p, "sentis: truncated or corrupt file (offset %zu)", p); }
  const uint8_t* b_;
  size_t n_, base_;
};

struct SentisTensor {
  int dtype = -1;                 // 0 f32, 1 i32, 2 i16, 3 u8
  std::vector<int32_t> shape;
  bool is_const = false;
  int64_t offset = 0;
  // element count; throws on a negative dimension or a product past 2^31 (the bytes come from a caller)
  size_t count() const {
    size_t n = 1;
    for (int32_t d : shape) {
      XR_CHECK(d >= 0, "sentis: negative tensor dimension %d", d);
      n *= static_cast<size_t>(d);
      XR_CHECK(n <= (static_cast<size_t>(1) << 31), "sentis: tensor too large");
    }
    return n;
  }
  // true when [offset, offset + bytes) lies inside a blob of `blob_size` bytes (offset is a sign-extended i32: reject < 0;
  // written without additions that could wrap)
  bool inside(size_t bytes, size_t blob_size) const {
    return offset >= 0 && static_cast<size_t>(offset) <= blob_size && bytes <= blob_size - static_cast<size_t>(offset);
  }
};

struct SentisConv {               // one Conv / ConvTranspose chain with its dequantized operands
  bool transposed = false;
  std::vector<int32_t> w_shape;   // [cout, cin/g, kh, kw]  (ConvTranspose: [cin, cout, kh, kw])
  std::vector<float> w, b;
};

struct SentisWeights {
  uint32_t version = 0;
  std::vector<SentisConv> convs;
  bool has_nms = false;
  float iou_threshold = 0.f, score_threshold = 0.f;
};

static inline bool looks_like_sentis(const void* data, size_t bytes) {
  if (bytes < 16) return false;
  uint32_t sz;
  memcpy(&sz, data, 4);
  return sz >= 8 && 4 + static_cast<size_t>(sz) <= bytes && memcmp(data, "XRSW", 4) != 0;
}

// Parses the container and returns every biased convolution in chain order (the DFL 1x1 conv, whose weights 0..15 are
// structural and which has no bias, is skipped exactly like tests/golden/make_golden.py does).
static inline SentisWeights sentis_load(const void* data, size_t bytes) {
  const uint8_t* raw = static_cast<const uint8_t*>(data);
  XR_CHECK(bytes >= 8, "sentis: file too small");
  uint32_t prog_size;
  memcpy(&prog_size, raw, 4);
  XR_CHECK(4 + static_cast<size_t>(prog_size) <= bytes, "sentis: program size %u exceeds the file", prog_size);
  FlatBuf fb(raw, 4 + static_cast<size_t>(prog_size), 4);
  const size_t prog = fb.root();
  SentisWeights out;
  const size_t vp = fb.field(prog, 0);
  out.version = vp ? fb.u32(vp) : 0;
  const size_t plan = fb.indirect(prog, 1);
  XR_CHECK(plan != 0, "sentis: no execution plan");

  // weight chunks: concatenated byte spans following the program
  std::vector<uint8_t> blob;
  for (size_t pos = 4 + static_cast<size_t>(prog_size); pos + 4 <= bytes;) {
    uint32_t csize;
    memcpy(&csize, raw + pos, 4);
    XR_CHECK(pos + 4 + csize <= bytes, "sentis: weight chunk exceeds the file");
    FlatBuf cfb(raw, pos + 4 + csize, pos + 4);
    size_t start, len;
    cfb.byte_span(cfb.root(), 0, &start, &len);
    blob.insert(blob.end(), raw + start, raw + start + len);
    pos += 4 + static_cast<size_t>(csize);
  }

  std::vector<std::string> operators;
  for (size_t t : fb.tables(plan, 7)) operators.push_back(fb.str(t, 0));

  // EValues: 2 int, 3 float, 6 tensor; everything else is irrelevant here
  struct Val { int kind = 0; int32_t i = 0; float f = 0.f; SentisTensor t; };
  std::vector<Val> values;
  for (size_t ev : fb.tables(plan, 1)) {
    Val v;
    const size_t tp = fb.field(ev, 0);
    const int vt = tp ? fb.u8(tp) : 0;
    const size_t body = fb.indirect(ev, 1);
    v.kind = vt;
    if (vt == 2 && body) { const size_t p = fb.field(body, 0); v.i = p ? fb.i32(p) : 0; }
    if (vt == 3 && body) { const size_t p = fb.field(body, 0); v.f = p ? fb.f32(p) : 0.f; }
    if (vt == 6 && body) {
      const size_t p0 = fb.field(body, 0), p3 = fb.field(body, 3), p4 = fb.field(body, 4);
      v.t.dtype = p0 ? fb.u8(p0) : 0;
      v.t.shape = fb.ints(body, 2);
      v.t.is_const = p3 ? fb.u32(p3) != 0 : false;
      v.t.offset = p4 ? fb.i32(p4) : 0;
    }
    values.push_back(v);
  }
  auto value = [&](int32_t id) -> const Val& {
    XR_CHECK(id >= 0 && static_cast<size_t>(id) < values.size(), "sentis: value id %d out of range", id);
    return values[id];
  };
  auto scalar_f = [&](int32_t id) { const Val& v = value(id); return v.kind == 3 ? v.f : static_cast<float>(v.i); };
  // constant f32 scalar tensor (thresholds are stored as 0-d / 1-element tensors) or plain float EValue
  auto const_f = [&](int32_t id, float* dst) {
    if (id < 0) return false;
    const Val& v = value(id);
    if (v.kind == 3) { *dst = v.f; return true; }
    if (v.kind == 6 && v.t.is_const && v.t.dtype == 0 && v.t.count() == 1 && v.t.inside(4, blob.size())) {
      memcpy(dst, blob.data() + v.t.offset, 4);
      return true;
    }
    return false;
  };

  struct Deq { std::vector<float> data; std::vector<int32_t> shape; };
  std::map<int32_t, Deq> deq;     // output value id of a DequantizeUint8 chain -> fp32 tensor
  for (size_t c : fb.tables(plan, 6)) {
    const std::vector<int32_t> ins = fb.ints(c, 0), outs = fb.ints(c, 1);
    const std::vector<size_t> instrs = fb.tables(c, 2);
    XR_CHECK(instrs.size() == 1, "sentis: one kernel call per chain expected");
    const size_t kc = fb.indirect(instrs[0], 1);
    XR_CHECK(kc != 0, "sentis: chain without kernel call");
    const size_t opp = fb.field(kc, 0);
    const int32_t op_index = opp ? fb.i32(opp) : 0;
    XR_CHECK(op_index >= 0 && static_cast<size_t>(op_index) < operators.size(), "sentis: operator index out of range");
    const std::string& op = operators[op_index];
    const std::vector<int32_t> args = fb.ints(kc, 1);
    if (op == "DequantizeUint8") {
      XR_CHECK(ins.size() >= 1 && outs.size() >= 1 && args.size() >= 2, "sentis: malformed DequantizeUint8 chain");
      const Val& q = value(ins[0]);
      XR_CHECK(q.kind == 6 && q.t.is_const && q.t.dtype == 3, "sentis: DequantizeUint8 input is not a constant u8 tensor");
      const size_t n = q.t.count();
      XR_CHECK(q.t.inside(n, blob.size()), "sentis: tensor data outside the weight blob");
      const float scale = scalar_f(args[0]);
      const float zp = scalar_f(args[1]);
      Deq d;
      d.shape = q.t.shape;
      d.data.resize(n);
      const uint8_t* src = blob.data() + q.t.offset;
      for (size_t i = 0; i < n; ++i) d.data[i] = (static_cast<float>(src[i]) - zp) * scale;
      deq[outs[0]] = std::move(d);
    } else if (op == "Conv" || op == "ConvTranspose") {
      if (ins.size() < 3 || ins[2] < 0) continue;      // the bias-free DFL conv: structural, not a layer of the pack
      auto w = deq.find(ins[1]), b = deq.find(ins[2]);
      XR_CHECK(w != deq.end() && b != deq.end(), "sentis: conv operands are not dequantized uint8 tensors");
      SentisConv cv;
      cv.transposed = op == "ConvTranspose";
      cv.w_shape = w->second.shape;
      cv.w = w->second.data;
      cv.b = b->second.data;
      out.convs.push_back(std::move(cv));
    } else if (op == "NonMaxSuppression") {
      // inputs: boxes, scores, maxOutputBoxesPerClass, iouThreshold, scoreThreshold (ONNX order)
      float iou = 0.f, sc = 0.f;
      if (ins.size() >= 5 && const_f(ins[3], &iou) && const_f(ins[4], &sc)) {
        out.has_nms = true;
        out.iou_threshold = iou;
        out.score_threshold = sc;
      }
    }
  }
  return out;
}

}  // namespace xrseg
