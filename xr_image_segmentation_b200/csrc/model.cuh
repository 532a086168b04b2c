// model.cuh -- YOLO11-seg topology builder (n / s widths, any batch) and XRSW weight-pack reader.
//
// Restates the layer graph the reference runs (the 499 chains of
// Assets/Resources/Model/yolo11n-seg-sentis.sentis, executed through IEExecutor.cs:371,397) as a list of kernel
// launches over NHWC fp16 buffers.  Concat / Split are channel-slice views of a shared buffer (no copies), residual
// Adds ride in the convolution epilogue, Swish is fused into every producing kernel.  Convolutions are emitted in
// the asset's chain order, which is the canonical layer order of the XRSW weight pack.
#pragma once

#include <string.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_umma.cuh"

namespace xrseg {

struct TV {           // tensor view inside the activation arena (element offsets, resolved at launch)
  size_t off = 0;
  int B = 0, H = 0, W = 0, C = 0, Cp = 0, pitch = 0;
};

struct LayerRec {
  std::string name;
  int cin, cout, k, stride, groups, act, transposed, h_in, w_in;
};

enum OpKind { OP_STEM, OP_CONV, OP_DW, OP_SPPF, OP_UP, OP_ATTN, OP_BNECK, OP_C3K2, OP_CHAIN };

struct Op {
  OpKind kind;
  int layer = -1;
  TV x, y, res;
  bool has_res = false;
  int k = 1, stride = 1, act = 0, transposed = 0, heads = 0;
  int in_grp = 0, in_grp_stride = 0, in_grp_off = 0;   // OP_DW channel remap (see DwParams)
  // Branch concurrency inside the captured graph: ops of branch > 0 (the head / proto chains) run on side streams.  The
  // first op of a branch segment waits for the event `wait_tag` (3/4/5 = P3/P4/P5 ready); the op producing that feature
  // map carries `signal_tag`.
  int branch = 0, wait_tag = 0, signal_tag = 0;
  // sibling fusion: a second convolution of the same input / kernel / stride / activation computed by the same launch
  // (OP_BNECK: the second 3x3 convolution of the fused Bottleneck, whose output is y)
  int layer2 = -1;
  TV y2;
  // fused trailing 1x1 convolution (conv_tma.cuh "tail"): the 1x1 layer `tail_layer` consumes this op's 64-channel output inside
  // the kernel; ytail is its destination (y stays the never-materialised intermediate)
  int tail_layer = -1, tail_act = 0;
  TV ytail;
  // OP_C3K2 (whole-block fusion): layer = X.cv1, layer2 = X.m0.cv1, layer3 = X.m0.cv2, layer4 = X.cv2
  int layer3 = -1, layer4 = -1;
  // OP_CHAIN (conv_chain.cuh): consecutive convolutions of the 20x20 / 40x40 stages run by one launch, a CTA per frame;
  // `chain` holds the OP_CONV ops in order (x / y of the chain op itself are those of its first / last convolution)
  std::vector<Op> chain;
};

// Channel triples (input, middle, output; as laid out in shared memory) the fused Bottleneck kernel is built for
// (bottleneck.cuh).  Anything else keeps its two tcgen05 launches.
// (input, split width c, bottleneck middle, output) of the C3k2 blocks the whole-block kernel is built for
static inline bool c3k2_supported(int cin, int c, int cm, int cout) { return cin == 32 && c == 16 && cm == 8 && cout == 64; }
static inline bool bneck_supported(int c1, int cm, int c2) {
  return (c1 == 16 && cm == 8 && c2 == 16) || (c1 == 32 && cm == 16 && c2 == 32);
}

struct Spec {
  int ch[5];
  int mid0, mid1;     // C3k2 output widths of layers 2 / 4 (n: 64 / 128)
  int cls_mid, box_mid, coef_mid, proto_mid, heads;
};

static inline Spec make_spec(int scale) {
  const int wnum = (scale == 's') ? 2 : 1;  // width = wnum / 4
  Spec s;
  const int base[5] = {64, 128, 256, 512, 1024};
  for (int i = 0; i < 5; ++i) s.ch[i] = base[i] * wnum / 4;
  s.mid0 = 256 * wnum / 4;
  s.mid1 = 512 * wnum / 4;
  s.cls_mid = s.ch[2] > 80 ? s.ch[2] : 80;
  s.box_mid = 64;
  if (s.ch[2] / 4 > s.box_mid) s.box_mid = s.ch[2] / 4;
  s.coef_mid = s.ch[2] / 4 > 32 ? s.ch[2] / 4 : 32;
  s.proto_mid = 256 * wnum / 4;
  s.heads = (s.ch[4] / 2) / 64;
  return s;
}

class Net {
 public:
  int B;
  Spec sp;
  std::vector<LayerRec> layers;
  std::vector<Op> ops;
  std::map<std::string, TV> named;
  size_t arena_elems = 0;
  int cur_branch = 0, pending_wait = 0;   // see Op::branch
  bool fuse_enabled = true;               // sibling fusion (fuse_siblings); off for the direct-convolution cross-check
  bool fuse_bneck = true;                 // Bottleneck fusion (fuse_bottlenecks)
  bool fuse_pe = true;                    // C2PSA positional encoding inside the attention kernel (fuse_attention_pe)
  bool fuse_tail = true;                  // proto.cv3 (1x1, 64 -> 32) inside proto.cv2's launch (fuse_tail_1x1)
  bool fuse_c3k2 = false;                 // whole-block fusion (fuse_c3k2_blocks); measured slower, opt-in (XRSEG_FUSE_C3K2=1)
  TV input;                 // [B,640,640,4] fp16
  TV box[3], cls[3], coef[3], protos;
  int fh[3], fw[3];

  Net(int scale, int batch, int in_hw = 640, bool fuse = true, bool bneck = true, bool c3k2_blocks = false)
      : B(batch), sp(make_spec(scale)), fuse_enabled(fuse), fuse_bneck(fuse && bneck), fuse_tail(fuse), fuse_c3k2(c3k2_blocks) { build(in_hw); }

  TV alloc(int H, int W, int C, int pitch_override = 0) {
    TV t;
    t.B = B; t.H = H; t.W = W; t.C = C;
    t.Cp = pitch_override ? pitch_override : round_up(C, 16);
    t.pitch = t.Cp;
    t.off = arena_elems;
    arena_elems += static_cast<size_t>(B) * H * W * t.pitch;
    arena_elems = (arena_elems + 127) / 128 * 128;
    return t;
  }
  static TV slice(const TV& t, int c0, int c) {
    XR_CHECK(c0 % 16 == 0 && c % 16 == 0 && c0 + c <= t.Cp, "bad slice %d+%d of %d", c0, c, t.Cp);
    TV s = t;
    s.off = t.off + c0;
    s.C = c;
    s.Cp = c;
    return s;
  }

  TV conv(const TV& x, int cout, int k, int s, bool act, const std::string& name, const TV* dst = nullptr,
          const TV* res = nullptr, bool transposed = false) {
    LayerRec l{name, x.C, cout, k, s, 1, act ? 1 : 0, transposed ? 1 : 0, x.H, x.W};
    layers.push_back(l);
    const int Ho = transposed ? x.H * 2 : (x.H + 2 * (k / 2) - k) / s + 1;
    const int Wo = transposed ? x.W * 2 : (x.W + 2 * (k / 2) - k) / s + 1;
    TV y = dst ? *dst : alloc(Ho, Wo, cout);
    XR_CHECK(y.H == Ho && y.W == Wo && y.C == cout, "conv %s: destination mismatch", name.c_str());
    Op o;
    o.kind = layers.size() == 1 ? OP_STEM : OP_CONV;
    o.layer = static_cast<int>(layers.size()) - 1;
    o.x = x; o.y = y; o.k = k; o.stride = s; o.act = act; o.transposed = transposed;
    if (res) { o.res = *res; o.has_res = true; }
    o.branch = cur_branch; o.wait_tag = pending_wait; pending_wait = 0;
    ops.push_back(o);
    named[name] = y;
    return y;
  }
  // grp > 0: the conv's C = n_grp * grp input channels are read from x at (c / grp) * grp_stride + grp_off + c % grp
  TV dw(const TV& x, bool act, const std::string& name, const TV* res = nullptr, int grp = 0, int grp_stride = 0,
        int grp_off = 0, int n_grp = 0) {
    const int C = grp > 0 ? grp * n_grp : x.C;
    LayerRec l{name, C, C, 3, 1, C, act ? 1 : 0, 0, x.H, x.W};
    layers.push_back(l);
    TV y = alloc(x.H, x.W, C);
    Op o;
    o.kind = OP_DW;
    o.layer = static_cast<int>(layers.size()) - 1;
    o.x = x; o.y = y; o.k = 3; o.act = act;
    o.in_grp = grp; o.in_grp_stride = grp_stride; o.in_grp_off = grp_off;
    if (res) { o.res = *res; o.has_res = true; }
    o.branch = cur_branch; o.wait_tag = pending_wait; pending_wait = 0;
    ops.push_back(o);
    named[name] = y;
    return y;
  }

  TV c3k2(const TV& x, int c_out, int c, bool c3k, const std::string& n, const TV* dst = nullptr) {
    TV cat = alloc(x.H, x.W, 3 * c);
    TV ab = slice(cat, 0, 2 * c);
    conv(x, 2 * c, 1, 1, true, n + ".cv1", &ab);
    TV b = slice(cat, c, c);
    TV m = slice(cat, 2 * c, c);
    if (!c3k) {
      TV t = conv(b, c / 2, 3, 1, true, n + ".m0.cv1");
      conv(t, c, 3, 1, true, n + ".m0.cv2", &m, &b);
    } else {
      const int c_ = c / 2;
      TV cat2 = alloc(x.H, x.W, 2 * c_);
      TV t0 = conv(b, c_, 1, 1, true, n + ".m0.cv1");
      TV h = conv(t0, c_, 3, 1, true, n + ".m0.m0.cv1");
      TV t1 = conv(h, c_, 3, 1, true, n + ".m0.m0.cv2", nullptr, &t0);
      h = conv(t1, c_, 3, 1, true, n + ".m0.m1.cv1");
      TV t2 = slice(cat2, 0, c_);
      conv(h, c_, 3, 1, true, n + ".m0.m1.cv2", &t2, &t1);
      TV u = slice(cat2, c_, c_);
      conv(b, c_, 1, 1, true, n + ".m0.cv2", &u);
      conv(cat2, c, 1, 1, true, n + ".m0.cv3", &m);
    }
    return conv(cat, c_out, 1, 1, true, n + ".cv2", dst);
  }

  void head_box_cls(const TV& p, int i, const std::string& n) {
    cur_branch = 1; pending_wait = 3 + i;            // box chain: side stream 1, after P(3+i)
    TV t = conv(p, sp.box_mid, 3, 1, true, n + ".box.0");
    t = conv(t, sp.box_mid, 3, 1, true, n + ".box.1");
    box[i] = conv(t, 64, 1, 1, false, n + ".box.2");
    cur_branch = 2; pending_wait = 3 + i;            // class chain: side stream 2
    t = dw(p, true, n + ".cls.0dw");
    t = conv(t, sp.cls_mid, 1, 1, true, n + ".cls.0pw");
    t = dw(t, true, n + ".cls.1dw");
    t = conv(t, sp.cls_mid, 1, 1, true, n + ".cls.1pw");
    cls[i] = conv(t, 80, 1, 1, false, n + ".cls.2");
    cur_branch = 0;
  }
  void head_coef(const TV& p, int i, const std::string& n) {
    cur_branch = 3; pending_wait = 3 + i;            // coefficient chain: side stream 3
    TV t = conv(p, sp.coef_mid, 3, 1, true, n + ".coef.0");
    t = conv(t, sp.coef_mid, 3, 1, true, n + ".coef.1");
    coef[i] = conv(t, 32, 1, 1, false, n + ".coef.2");
    cur_branch = 0;
  }

  // Two convolutions that read the same tensor with the same kernel / stride / activation and no residual become one
  // launch with N = Cout_a + Cout_b (the input is read once, one launch less): the two 1x1 convs at the head of every
  // C3k block (m0.cv1 | m0.cv2) and the first 3x3 of the box and coefficient branches of every scale (box.0 | coef.0).
  // The second op disappears from `ops`; its consumers wait for the fused launch through event tag 6 + pair index.
  void fuse_siblings() {
    if (!fuse_enabled) return;
    auto find = [&](const std::string& nm) {
      for (size_t i = 0; i < ops.size(); ++i)
        if (ops[i].layer >= 0 && layers[ops[i].layer].name == nm) return static_cast<int>(i);
      return -1;
    };
    std::vector<std::pair<std::string, std::string>> pairs;
    for (const char* blk : {"b6", "b8", "n22"}) pairs.push_back({std::string(blk) + ".m0.cv1", std::string(blk) + ".m0.cv2"});
    for (const char* h : {"h3", "h4", "h5"}) pairs.push_back({std::string(h) + ".box.0", std::string(h) + ".coef.0"});
    int tag = 6;
    for (const auto& pr : pairs) {
      const int ia = find(pr.first), ib = find(pr.second);
      if (ia < 0 || ib < 0) continue;
      Op& a = ops[ia];
      const Op b = ops[ib];
      const bool same = a.kind == OP_CONV && b.kind == OP_CONV && a.x.off == b.x.off && a.x.C == b.x.C && a.x.pitch == b.x.pitch &&
                        a.k == b.k && a.stride == b.stride && a.act == b.act && !a.has_res && !b.has_res && !a.transposed &&
                        !b.transposed && a.layer2 < 0 && a.y.Cp + b.y.Cp <= 256;
      if (!same) continue;
      a.layer2 = b.layer;
      a.y2 = b.y;
      if (a.branch != b.branch) {
        // the consumer chain of b runs on another stream: it must wait for the fused launch (which runs on a's stream)
        a.signal_tag = tag;
        for (size_t j = ib + 1; j < ops.size(); ++j)
          if (ops[j].branch == b.branch) {           // first later op of b's branch = b's consumer
            if (ops[j].wait_tag == 0) ops[j].wait_tag = tag;
            break;
          }
        ++tag;
      }
      ops.erase(ops.begin() + ib);
    }
  }

  // A 3x3 convolution whose ONLY consumer is a 1x1 convolution (rows are independent for a 1x1) runs that 1x1 inside its own
  // launch as a second set of MMAs on the epilogue's fp16 tile kept in tensor memory (ConvParams::tail_n, conv_tma.cuh): the
  // intermediate tensor is never written to HBM or read back.  The kernel side handles 32 / 64 channels on either side; the
  // pairs below are the ones whose TMEM budget works out on the n scale:
  //   proto.cv2 (64) -> proto.cv3 (32)   in place (512 accumulator columns are all in use)
  //   b1 (32) -> b2.cv1 (32),  b3 (64) -> b4.cv1 (64)   the stride-2 stage convs feed only the next C3k2 block's cv1;
  //                                                      256 + 256 columns
  // Anything else keeps its two launches.
  void fuse_tail_1x1() {
    static const bool env_on = [] { const char* e = getenv("XRSEG_FUSE_TAIL"); return !(e && e[0] == '0'); }();
    // the stage pairs are opt-in: measured b1 + b2.cv1 76 + 43 -> 113 us, b3 + b4.cv1 66 + 27 -> 88 us, frames/s unchanged -- these
    // layers are bound by their epilogue, and two epilogue passes in one launch cost what the second launch did
    static const bool env_stage = [] { const char* e = getenv("XRSEG_FUSE_TAIL_STAGE"); return e && e[0] == '1'; }();
    if (!fuse_tail || !env_on) return;
    for (size_t i = 0; i + 1 < ops.size(); ++i) {
      Op& a = ops[i];
      const Op b = ops[i + 1];
      if (a.kind != OP_CONV || b.kind != OP_CONV || a.layer2 >= 0 || b.layer2 >= 0 || a.tail_layer >= 0) continue;
      const std::string &na = layers[a.layer].name, &nb = layers[b.layer].name;
      const bool proto = na == "proto.cv2" && nb == "proto.cv3" && a.stride == 1 && a.y.Cp == 64 && b.y.Cp == 32;
      const bool stage = env_stage && a.stride == 2 &&
                         ((na == "b1" && nb == "b2.cv1" && a.y.Cp == 32 && b.y.Cp == 32 && a.y.W == 160) ||
                          (na == "b3" && nb == "b4.cv1" && a.y.Cp == 64 && b.y.Cp == 64 && a.y.W == 80));
      if (!proto && !stage) continue;
      const bool ok = a.k == 3 && !a.transposed && a.act && !a.has_res && b.k == 1 && b.stride == 1 && !b.transposed && !b.has_res &&
                      b.x.off == a.y.off && b.x.pitch == a.y.pitch && b.x.Cp == a.y.Cp && a.branch == b.branch && b.wait_tag == 0 &&
                      a.signal_tag == 0;
      if (!ok) continue;
      a.tail_layer = b.layer;
      a.tail_act = b.act;
      a.ytail = b.y;
      a.signal_tag = b.signal_tag;
      ops.erase(ops.begin() + i + 1);
    }
  }

  // C2PSA: x = attention(q, k, v) + pe(v).  The depthwise 3x3 `attn.pe` reads V in place from the qkv tensor and adds the
  // attention output as its residual; the attention kernel already holds the head's V map in shared memory, so the pair
  // becomes one OP_ATTN launch: layer = the pe layer (its weights), y = the pe op's destination.
  void fuse_attention_pe() {
    static const bool env_on = [] { const char* e = getenv("XRSEG_FUSE_PE"); return !(e && e[0] == '0'); }();
    if (!fuse_pe || !env_on) return;
    for (size_t i = 0; i + 1 < ops.size(); ++i) {
      Op& a = ops[i];
      const Op b = ops[i + 1];
      if (a.kind != OP_ATTN || b.kind != OP_DW || a.layer >= 0) continue;
      const bool ok = b.in_grp > 0 && b.x.off == a.x.off && b.has_res && b.res.off == a.y.off && !b.act && a.branch == b.branch &&
                      b.wait_tag == 0 && a.signal_tag == 0;
      if (!ok) continue;
      a.layer = b.layer;
      a.y = b.y;
      a.signal_tag = b.signal_tag;
      ops.erase(ops.begin() + i + 1);
    }
  }

  // The Bottleneck of a C3k2 block without C3k (X.m0.cv1 -> X.m0.cv2 (+ x)): two dependent 3x3 convolutions over 8..32
  // channels become one launch whose intermediate stays in shared memory (bottleneck.cuh).  The pair collapses into one
  // OP_BNECK op: layer = cv1, layer2 = cv2, x = the block input (also the residual), y = cv2's destination.
  void fuse_bottlenecks() {
    if (!fuse_bneck) return;
    for (size_t i = 0; i + 1 < ops.size(); ++i) {
      Op& a = ops[i];
      const Op b = ops[i + 1];
      if (a.kind != OP_CONV || b.kind != OP_CONV || a.layer2 >= 0 || b.layer2 >= 0) continue;
      const std::string &na = layers[a.layer].name, &nb = layers[b.layer].name;
      const bool names = na.size() > 7 && na.compare(na.size() - 7, 7, ".m0.cv1") == 0 && nb == na.substr(0, na.size() - 1) + "2";
      if (!names) continue;
      const bool chain = a.k == 3 && b.k == 3 && a.stride == 1 && b.stride == 1 && a.act && b.act && !a.has_res &&
                         !a.transposed && !b.transposed && b.x.off == a.y.off && a.branch == b.branch && b.wait_tag == 0 &&
                         a.signal_tag == 0 && (!b.has_res || (b.res.off == a.x.off && b.res.pitch == a.x.pitch));
      const int cm = round_up(layers[a.layer].cout, 8);
      if (!chain || !bneck_supported(a.x.Cp, cm, b.y.Cp) || (b.has_res && a.x.Cp != b.y.Cp)) continue;
      a.kind = OP_BNECK;
      a.layer2 = b.layer;
      a.y = b.y;
      a.has_res = b.has_res;
      a.res = b.res;
      a.signal_tag = b.signal_tag;
      ops.erase(ops.begin() + i + 1);
    }
  }

  // X.cv1 (1x1) -> fused Bottleneck -> X.cv2 (1x1) over the block's concat buffer become ONE launch when the channel
  // counts are small enough for everything to stay in shared memory (b2 of the n scale): the concat buffer is never
  // materialised.  Runs after fuse_bottlenecks().
  void fuse_c3k2_blocks() {
    if (!fuse_bneck || !fuse_c3k2) return;
    for (size_t i = 0; i + 2 < ops.size(); ++i) {
      Op& a = ops[i];
      const Op m = ops[i + 1], z = ops[i + 2];
      if (a.kind != OP_CONV || m.kind != OP_BNECK || z.kind != OP_CONV || a.layer2 >= 0 || z.layer2 >= 0) continue;
      const int c = m.x.Cp;
      const bool chain = a.k == 1 && z.k == 1 && a.stride == 1 && z.stride == 1 && a.act && z.act && !a.has_res && !z.has_res &&
                         !a.transposed && !z.transposed && m.has_res && a.y.Cp == 2 * c && m.x.off == a.y.off + c &&
                         m.y.off == a.y.off + 2 * c && z.x.off == a.y.off && z.x.Cp == 3 * c && z.x.pitch == a.y.pitch &&
                         a.branch == m.branch && m.branch == z.branch && m.wait_tag == 0 && z.wait_tag == 0 &&
                         a.signal_tag == 0 && m.signal_tag == 0 && layers[a.layer].cout == 2 * c && layers[z.layer].cin == 3 * c;
      if (!chain || !c3k2_supported(a.x.Cp, c, round_up(layers[m.layer].cout, 8), z.y.Cp)) continue;
      a.kind = OP_C3K2;
      a.layer2 = m.layer; a.layer3 = m.layer2; a.layer4 = z.layer;
      a.y = z.y;
      a.signal_tag = z.signal_tag;
      ops.erase(ops.begin() + i + 1, ops.begin() + i + 3);
    }
  }

  // Consecutive OP_CONV ops on small maps (input <= 40x40) of the same branch become one OP_CHAIN launch: a CTA walks the
  // whole chain for one frame (frames are independent, so no grid-wide dependency exists inside these stages).  `can_chain`
  // says whether the chain kernel has a plan for a convolution.  A chain starts at an op that waits for an event (or at any
  // eligible op) and ends at an op that signals one, so the branch / event structure of the graph is unchanged.
  void fuse_chains(const std::function<bool(const Op&)>& can_chain, int max_pixels = 1600) {
    std::vector<Op> out;
    for (size_t i = 0; i < ops.size();) {
      auto eligible = [&](const Op& o) {
        return o.kind == OP_CONV && !o.transposed && o.x.H * o.x.W <= max_pixels && can_chain(o);
      };
      if (!eligible(ops[i])) { out.push_back(ops[i++]); continue; }
      size_t j = i + 1;
      if (ops[i].signal_tag == 0)
        while (j < ops.size() && eligible(ops[j]) && ops[j].branch == ops[i].branch && ops[j].wait_tag == 0) {
          ++j;
          if (ops[j - 1].signal_tag != 0) break;     // a signalling op ends its chain
        }
      if (j - i < 2) { out.push_back(ops[i++]); continue; }
      Op c;
      c.kind = OP_CHAIN;
      c.layer = ops[i].layer;
      c.x = ops[i].x;
      c.y = ops[j - 1].y;
      c.branch = ops[i].branch;
      c.wait_tag = ops[i].wait_tag;
      c.signal_tag = ops[j - 1].signal_tag;
      c.chain.assign(ops.begin() + i, ops.begin() + j);
      out.push_back(c);
      i = j;
    }
    ops.swap(out);
  }

  void build(int hw) {
    const int c1 = sp.ch[0], c2 = sp.ch[1], c5 = sp.ch[4];
    input = alloc(hw, hw, 3, 4);
    named["input"] = input;
    const int h8 = hw / 8, h16 = hw / 16, h32 = hw / 32;
    // concat buffers of the neck, allocated first so that backbone outputs land directly in their slices
    TV cat13 = alloc(h16, h16, c5 + sp.mid1);        // [up(f10) | f6]
    TV cat16 = alloc(h8, h8, sp.mid1 + sp.mid1);     // [up(f13) | f4]
    TV cat19 = alloc(h16, h16, sp.mid0 + sp.mid1);   // [n17 | f13]
    TV cat22 = alloc(h32, h32, sp.mid1 + c5);        // [n20 | f10]

    TV t = conv(input, c1, 3, 2, true, "b0");
    t = conv(t, c2, 3, 2, true, "b1");
    t = c3k2(t, sp.mid0, sp.mid0 / 4, false, "b2");
    t = conv(t, sp.mid0, 3, 2, true, "b3");
    TV f4 = slice(cat16, sp.mid1, sp.mid1);
    c3k2(t, sp.mid1, sp.mid1 / 4, false, "b4", &f4);
    t = conv(f4, sp.mid1, 3, 2, true, "b5");
    TV f6 = slice(cat13, c5, sp.mid1);
    c3k2(t, sp.mid1, sp.mid1 / 2, true, "b6", &f6);
    t = conv(f6, c5, 3, 2, true, "b7");
    t = c3k2(t, c5, c5 / 2, true, "b8");
    {  // SPPF (chains 140-151)
      const int c_ = c5 / 2;
      TV cat = alloc(h32, h32, 4 * c_);
      TV y0 = slice(cat, 0, c_);
      conv(t, c_, 1, 1, true, "b9.cv1", &y0);
      Op o;
      o.kind = OP_SPPF;
      o.x = cat; o.y = cat;
      ops.push_back(o);
      t = conv(cat, c5, 1, 1, true, "b9.cv2");
    }
    TV f10 = slice(cat22, sp.mid1, c5);
    {  // C2PSA (chains 152-190)
      const int c = c5 / 2;
      const int kd = 32, hd = 64;
      XR_CHECK(c / sp.heads == hd, "attention head dim must be 64");
      TV cat = alloc(h32, h32, 2 * c);
      conv(t, 2 * c, 1, 1, true, "b10.cv1", &cat);
      TV b = slice(cat, c, c);
      TV qkv = conv(b, c + 2 * sp.heads * kd, 1, 1, false, "b10.attn.qkv");
      TV ao = alloc(h32, h32, c);
      Op oa;
      oa.kind = OP_ATTN; oa.x = qkv; oa.y = ao; oa.heads = sp.heads;
      ops.push_back(oa);
      // pe(v) + attention output; V is read in place from qkv (per head: kd query, kd key, hd value channels)
      TV s = dw(qkv, false, "b10.attn.pe", &ao, hd, 2 * kd + hd, 2 * kd, sp.heads);
      conv(s, c, 1, 1, false, "b10.attn.proj", &b, &b);              // b += proj(...)
      TV f = conv(b, 2 * c, 1, 1, true, "b10.ffn.0");
      conv(f, c, 1, 1, false, "b10.ffn.1", &b, &b);                  // b += ffn(b)
      conv(cat, c5, 1, 1, true, "b10.cv2", &f10);
    }
    auto upsample = [&](const TV& src, const TV& dst) {
      Op o;
      o.kind = OP_UP; o.x = src; o.y = dst;
      ops.push_back(o);
    };
    upsample(f10, slice(cat13, 0, c5));
    TV f13 = slice(cat19, sp.mid0, sp.mid1);
    c3k2(cat13, sp.mid1, sp.mid1 / 2, false, "n13", &f13);
    upsample(f13, slice(cat16, 0, sp.mid1));
    TV p3 = c3k2(cat16, sp.mid0, sp.mid0 / 2, false, "n16");
    ops.back().signal_tag = 3;
    head_box_cls(p3, 0, "h3");
    TV n17 = slice(cat19, 0, sp.mid0);
    conv(p3, sp.mid0, 3, 2, true, "n17", &n17);
    TV p4 = c3k2(cat19, sp.mid1, sp.mid1 / 2, false, "n19");
    ops.back().signal_tag = 4;
    head_box_cls(p4, 1, "h4");
    TV n20 = slice(cat22, 0, sp.mid1);
    conv(p4, sp.mid1, 3, 2, true, "n20", &n20);
    TV p5 = c3k2(cat22, c5, c5 / 2, true, "n22");
    ops.back().signal_tag = 5;
    head_box_cls(p5, 2, "h5");
    head_coef(p3, 0, "h3");
    head_coef(p4, 1, "h4");
    head_coef(p5, 2, "h5");
    cur_branch = 4; pending_wait = 3;                // prototype chain: side stream 4, after P3
    t = conv(p3, sp.proto_mid, 3, 1, true, "proto.cv1");
    t = conv(t, sp.proto_mid, 2, 2, false, "proto.up", nullptr, nullptr, true);
    t = conv(t, sp.proto_mid, 3, 1, true, "proto.cv2");
    protos = conv(t, 32, 1, 1, true, "proto.cv3");
    cur_branch = 0;
    named["p3"] = p3; named["p4"] = p4; named["p5"] = p5;
    fuse_siblings();
    fuse_bottlenecks();
    fuse_c3k2_blocks();
    fuse_attention_pe();
    fuse_tail_1x1();
    fh[0] = p3.H; fw[0] = p3.W; fh[1] = p4.H; fw[1] = p4.W; fh[2] = p5.H; fw[2] = p5.W;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// XRSW weight pack (written by xr_image_segmentation_b200/weights.py)
// ---------------------------------------------------------------------------------------------------------------
#pragma pack(push, 1)
struct XrswHeader {
  char magic[4];
  uint32_t version, n_layers, scale;
  uint64_t payload_offset;
  uint8_t reserved[8];
};
struct XrswLayer {
  char name[32];
  uint32_t cout, cin_g, k, stride, groups, act, transposed, dtype;
  float w_scale; int32_t w_zp; float b_scale; int32_t b_zp;
  uint64_t w_off, b_off;
};
#pragma pack(pop)
static_assert(sizeof(XrswHeader) == 32, "header size");
static_assert(sizeof(XrswLayer) == 96, "record size");

struct HostLayerWeights {
  std::vector<float> w, b;
};

// Dequantize exactly like the reference graph's DequantizeUint8 layers: (q - zp) * scale per tensor.
static inline void xrsw_load(const void* data, size_t bytes, const std::vector<LayerRec>& layers, int scale,
                             std::vector<HostLayerWeights>& out) {
  XR_CHECK(bytes >= sizeof(XrswHeader), "weight pack too small");
  const uint8_t* p = static_cast<const uint8_t*>(data);
  XrswHeader h;
  memcpy(&h, p, sizeof(h));
  XR_CHECK(memcmp(h.magic, "XRSW", 4) == 0 && h.version == 1, "not an XRSW v1 pack");
  XR_CHECK(h.n_layers == layers.size(), "pack has %u layers, topology has %zu", h.n_layers, layers.size());
  XR_CHECK(static_cast<int>(h.scale) == scale, "pack is for scale '%c'", static_cast<char>(h.scale));
  XR_CHECK(bytes >= sizeof(XrswHeader) + h.n_layers * sizeof(XrswLayer) && h.payload_offset <= bytes, "truncated pack");
  out.resize(h.n_layers);
  for (uint32_t i = 0; i < h.n_layers; ++i) {
    XrswLayer r;
    memcpy(&r, p + sizeof(XrswHeader) + i * sizeof(XrswLayer), sizeof(r));
    const LayerRec& l = layers[i];
    const int cin_g = l.cin / l.groups;
    XR_CHECK(static_cast<int>(r.cout) == l.cout && static_cast<int>(r.cin_g) == cin_g && static_cast<int>(r.k) == l.k &&
                 static_cast<int>(r.stride) == l.stride && static_cast<int>(r.groups) == l.groups &&
                 static_cast<int>(r.transposed) == l.transposed && static_cast<int>(r.act) == l.act,
             "layer %u (%.32s) does not match topology layer %s", i, r.name, l.name.c_str());
    const size_t nw = static_cast<size_t>(l.cout) * cin_g * l.k * l.k;
    const size_t nb = l.cout;
    const size_t esz = r.dtype == 0 ? 4 : 1;
    XR_CHECK(r.dtype == 0 || r.dtype == 3, "unsupported dtype %u", r.dtype);
    XR_CHECK(h.payload_offset + r.w_off + nw * esz <= bytes && h.payload_offset + r.b_off + nb * esz <= bytes,
             "layer %u payload out of range", i);
    const uint8_t* wp = p + h.payload_offset + r.w_off;
    const uint8_t* bp = p + h.payload_offset + r.b_off;
    out[i].w.resize(nw);
    out[i].b.resize(nb);
    if (r.dtype == 0) {
      memcpy(out[i].w.data(), wp, nw * 4);
      memcpy(out[i].b.data(), bp, nb * 4);
    } else {
      for (size_t j = 0; j < nw; ++j) out[i].w[j] = (static_cast<float>(wp[j]) - static_cast<float>(r.w_zp)) * r.w_scale;
      for (size_t j = 0; j < nb; ++j) out[i].b[j] = (static_cast<float>(bp[j]) - static_cast<float>(r.b_zp)) * r.b_scale;
    }
  }
}

}  // namespace xrseg
