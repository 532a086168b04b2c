// kernels_misc.cuh -- the memory-bound CUDA-core kernels around the tensor-core convolutions:
// frame preprocessing (↔ TextureConverter.ToTensor, IEExecutor.cs:370), the K=27 stem convolution, depthwise 3x3
// convolutions, the fused SPPF max-pool chain, nearest x2 upsampling into a concat slice, the C2PSA attention, and a
// plain direct convolution used as the on-GPU cross-check of the tcgen05 path.  All activations are NHWC fp16.
#pragma once

#include "common.cuh"

namespace xrseg {

// A view of an NHWC fp16 tensor: channel slice [coff, coff + C) of a buffer whose pixel pitch is `pitch` elements.
struct TView {
  __half* ptr = nullptr;  // already offset to the first channel of the slice
  int B = 0, H = 0, W = 0, C = 0, pitch = 0;
};

#ifdef __CUDACC__

// ------------------------------------------------------------------------------------------------
// preprocess: u8 RGB/RGBA frames (any size) -> fp16 NHWC [B,640,640,4] (4th channel 0), values 0..1
// stretch: bilinear sample at (x+0.5)*sw/640-0.5, clamp; letterbox: scale r = min(640/h, 640/w), pad 114.
// ------------------------------------------------------------------------------------------------
struct PreParams {
  const uint8_t* src;
  __half* dst;
  int B, sw, sh, stride_bytes, bpp;
  int mode;         // 0 stretch, 1 letterbox
  float scale_x, scale_y;
  int left, top, nw, nh;   // letterbox placement
};

__global__ void preprocess_kernel(const PreParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= 640) return;
  const uint8_t* img = p.src + static_cast<size_t>(b) * p.sh * p.stride_bytes;
  float rgb[3];
  bool inside = true;
  int xx = x, yy = y;
  if (p.mode == 1) {
    xx = x - p.left;
    yy = y - p.top;
    inside = (xx >= 0 && xx < p.nw && yy >= 0 && yy < p.nh);
  }
  if (inside) {
    const float fx = (static_cast<float>(xx) + 0.5f) * p.scale_x - 0.5f;
    const float fy = (static_cast<float>(yy) + 0.5f) * p.scale_y - 0.5f;
    const float x0f = floorf(fx), y0f = floorf(fy);
    const float tx = fx - x0f, ty = fy - y0f;
    int x0 = static_cast<int>(x0f), y0 = static_cast<int>(y0f);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = min(max(x0, 0), p.sw - 1); x1 = min(max(x1, 0), p.sw - 1);
    y0 = min(max(y0, 0), p.sh - 1); y1 = min(max(y1, 0), p.sh - 1);
    const uint8_t* r0 = img + static_cast<size_t>(y0) * p.stride_bytes;
    const uint8_t* r1 = img + static_cast<size_t>(y1) * p.stride_bytes;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = static_cast<float>(r0[x0 * p.bpp + c]), bq = static_cast<float>(r0[x1 * p.bpp + c]);
      const float cq = static_cast<float>(r1[x0 * p.bpp + c]), d = static_cast<float>(r1[x1 * p.bpp + c]);
      const float top = __fadd_rn(__fmul_rn(a, 1.0f - tx), __fmul_rn(bq, tx));
      const float bot = __fadd_rn(__fmul_rn(cq, 1.0f - tx), __fmul_rn(d, tx));
      const float v = __fadd_rn(__fmul_rn(top, 1.0f - ty), __fmul_rn(bot, ty));
      rgb[c] = __fdiv_rn(v, 255.0f);
    }
  } else {
    rgb[0] = rgb[1] = rgb[2] = __fdiv_rn(114.0f, 255.0f);
  }
  __half2 lo = __floats2half2_rn(rgb[0], rgb[1]);
  __half2 hi = __floats2half2_rn(rgb[2], 0.0f);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&lo);
  o.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p.dst + ((static_cast<size_t>(b) * 640 + y) * 640 + x) * 4) = o;
}

// ------------------------------------------------------------------------------------------------
// stem: 3x3 stride-2 conv on the 4-channel fp16 image (K = 27 real taps), + bias + SiLU.  One thread per output
// pixel and 16 output channels; weights fp32 in shared memory as [tap(9)][ci(4)][co].
// ------------------------------------------------------------------------------------------------
struct StemParams {
  const __half* in;    // [B,H,W,4]
  __half* out;         // [B,H/2,W/2,Cout] pitch out_pitch
  const float* w;      // [9][4][Cout]
  const float* bias;   // [Cout]
  int B, H, W, Cout, out_pitch;
};

__global__ void __launch_bounds__(256) stem_conv_kernel(const StemParams p) {
  extern __shared__ float sw[];  // 36*Cout weights + Cout bias
  for (int i = threadIdx.x; i < 36 * p.Cout; i += blockDim.x) sw[i] = p.w[i];
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) sw[36 * p.Cout + i] = p.bias[i];
  __syncthreads();
  const int Ho = p.H / 2, Wo = p.W / 2;
  const int groups = p.Cout / 16;
  const long total = static_cast<long>(p.B) * Ho * Wo * groups;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    long pix = idx / groups;
    const int ow = static_cast<int>(pix % Wo);
    const int oh = static_cast<int>((pix / Wo) % Ho);
    const int b = static_cast<int>(pix / (static_cast<long>(Wo) * Ho));
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = sw[36 * p.Cout + g * 16 + i];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * 2 - 1 + kh;
      if (ih < 0 || ih >= p.H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow * 2 - 1 + kw;
        if (iw < 0 || iw >= p.W) continue;
        const uint2 raw = *reinterpret_cast<const uint2*>(p.in + ((static_cast<size_t>(b) * p.H + ih) * p.W + iw) * 4);
        const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        const float xin[3] = {f01.x, f01.y, f23.x};
        const float* wt = sw + (kh * 3 + kw) * 4 * p.Cout + g * 16;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = fmaf(xin[ci], wt[ci * p.Cout + i], acc[i]);
      }
    }
    uint4 o0, o1;
    __half2* q0 = reinterpret_cast<__half2*>(&o0);
    __half2* q1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      q0[i] = __floats2half2_rn(silu_f(acc[2 * i]), silu_f(acc[2 * i + 1]));
      q1[i] = __floats2half2_rn(silu_f(acc[8 + 2 * i]), silu_f(acc[8 + 2 * i + 1]));
    }
    uint4* op = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(pix) * p.out_pitch + g * 16);
    op[0] = o0;
    op[1] = o1;
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 stride-1 pad-1 conv (+bias, optional SiLU, optional residual add), 8 channels per thread.
// weights fp32 [9][C], bias fp32 [C].
// ------------------------------------------------------------------------------------------------
struct DwParams {
  const __half* in; int in_pitch;
  __half* out; int out_pitch;
  const __half* res; int res_pitch;   // optional, added after activation
  const float* w; const float* bias;
  int B, H, W, C, act;
};

__global__ void __launch_bounds__(256) dwconv3x3_kernel(const DwParams p) {
  const int cg = p.C / 8;
  const long total = static_cast<long>(p.B) * p.H * p.W * cg;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % cg);
    const long pix = idx / cg;
    const int x = static_cast<int>(pix % p.W);
    const int y = static_cast<int>((pix / p.W) % p.H);
    const long b = pix / (static_cast<long>(p.W) * p.H);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = p.bias[g * 8 + i];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = y - 1 + kh;
      if (iy < 0 || iy >= p.H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ix = x - 1 + kw;
        if (ix < 0 || ix >= p.W) continue;
        const uint4 raw =
            *reinterpret_cast<const uint4*>(p.in + ((b * p.H + iy) * p.W + ix) * p.in_pitch + g * 8);
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
        const float4 w0 = *reinterpret_cast<const float4*>(p.w + (kh * 3 + kw) * p.C + g * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(p.w + (kh * 3 + kw) * p.C + g * 8 + 4);
        const float2 a = __half22float2(h[0]), bb = __half22float2(h[1]), c = __half22float2(h[2]),
                     d = __half22float2(h[3]);
        acc[0] = fmaf(a.x, w0.x, acc[0]); acc[1] = fmaf(a.y, w0.y, acc[1]);
        acc[2] = fmaf(bb.x, w0.z, acc[2]); acc[3] = fmaf(bb.y, w0.w, acc[3]);
        acc[4] = fmaf(c.x, w1.x, acc[4]); acc[5] = fmaf(c.y, w1.y, acc[5]);
        acc[6] = fmaf(d.x, w1.z, acc[6]); acc[7] = fmaf(d.y, w1.w, acc[7]);
      }
    }
    if (p.act) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = silu_f(acc[i]);
    }
    if (p.res) {
      const uint4 raw = *reinterpret_cast<const uint4*>(p.res + pix * p.res_pitch + g * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        acc[2 * i] += f.x;
        acc[2 * i + 1] += f.y;
      }
    }
    uint4 o;
    __half2* q = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
    *reinterpret_cast<uint4*>(p.out + pix * p.out_pitch + g * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// SPPF pooling chain (graph chains 144-147): y1 = maxpool5(y0), y2 = maxpool5(y1), y3 = maxpool5(y2), written into
// the three channel slices after y0 of the concat buffer.  One block per (image, 8-channel group); the HxW map
// (20x20) lives in shared memory.  Chained 5x5 pools equal 5x5, 9x9 and 13x13 windows of y0.
// ------------------------------------------------------------------------------------------------
struct SppfParams {
  __half* buf;   // concat buffer [B,H,W,4C]: slice 0 holds y0 (input), slices 1..3 are written
  int B, H, W, C, pitch;
};

__global__ void __launch_bounds__(256) sppf_pool_kernel(const SppfParams p) {
  extern __shared__ __half2 sp[];  // [H*W][4] half2 = 8 channels
  const int cg = p.C / 8;
  const int g = blockIdx.x % cg;
  const int b = blockIdx.x / cg;
  const int n = p.H * p.W;
  __half* base = p.buf + static_cast<size_t>(b) * n * p.pitch + g * 8;
  uint4* s4 = reinterpret_cast<uint4*>(sp);
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    s4[i] = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(i) * p.pitch);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int y = i / p.W, x = i - y * p.W;
    __half2 m[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) m[r][c] = __float2half2_rn(-65504.0f);
    for (int dy = -6; dy <= 6; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= p.H) continue;
      for (int dx = -6; dx <= 6; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= p.W) continue;
        const int rad = max(abs(dx), abs(dy));
        const __half2* v = sp + (yy * p.W + xx) * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          m[2][c] = __hmax2(m[2][c], v[c]);
          if (rad <= 4) m[1][c] = __hmax2(m[1][c], v[c]);
          if (rad <= 2) m[0][c] = __hmax2(m[0][c], v[c]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      uint4 o;
      __half2* q = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int c = 0; c < 4; ++c) q[c] = m[r][c];
      *reinterpret_cast<uint4*>(base + static_cast<size_t>(i) * p.pitch + (r + 1) * p.C) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// nearest x2 upsample (graph chains 191, 212: out[i] = in[i / 2]) into a channel slice of the concat buffer.
// ------------------------------------------------------------------------------------------------
struct UpParams {
  const __half* in; int in_pitch;
  __half* out; int out_pitch;
  int B, H, W, C;  // input dims
};

__global__ void __launch_bounds__(256) upsample2x_kernel(const UpParams p) {
  const int cg = p.C / 8;
  const int Ho = p.H * 2, Wo = p.W * 2;
  const long total = static_cast<long>(p.B) * Ho * Wo * cg;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % cg);
    const long pix = idx / cg;
    const int x = static_cast<int>(pix % Wo);
    const int y = static_cast<int>((pix / Wo) % Ho);
    const long b = pix / (static_cast<long>(Wo) * Ho);
    const uint4 v = *reinterpret_cast<const uint4*>(p.in + ((b * p.H + (y >> 1)) * p.W + (x >> 1)) * p.in_pitch + g * 8);
    *reinterpret_cast<uint4*>(p.out + pix * p.out_pitch + g * 8) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// C2PSA attention (graph chains 160-168): per (image, head): S = Q^T K * scale, softmax over keys, O = V A^T.
// qkv is the NHWC output of the qkv 1x1 conv: per token, head h occupies channels [h*(2kd+hd), (h+1)*(2kd+hd)):
// kd query dims, kd key dims, hd value dims.  One thread per query, K and V of the head staged in shared memory,
// online softmax in fp32.  kd = 32, hd = 64.
// ------------------------------------------------------------------------------------------------
struct AttnParams {
  const __half* qkv; int qkv_pitch;
  __half* out; int out_pitch;    // [B,N,heads*hd]
  int B, N, heads;
  float scale;
};

constexpr int ATT_KD = 32, ATT_HD = 64, ATT_THREADS = 128;

__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  __half* ks = reinterpret_cast<__half*>(att_smem);                 // [N][32]
  __half* vs = ks + static_cast<size_t>(p.N) * ATT_KD;              // [N][64]
  const int bh = blockIdx.x;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int per_head = 2 * ATT_KD + ATT_HD;
  const __half* base = p.qkv + static_cast<size_t>(b) * p.N * p.qkv_pitch + h * per_head;
  // stage K (4 x 16 B per token) and V (8 x 16 B per token)
  for (int i = threadIdx.x; i < p.N * 12; i += blockDim.x) {
    const int tok = i / 12, part = i - tok * 12;
    const uint4 v = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(tok) * p.qkv_pitch + ATT_KD + part * 8);
    if (part < 4) reinterpret_cast<uint4*>(ks)[tok * 4 + part] = v;
    else reinterpret_cast<uint4*>(vs)[tok * 8 + (part - 4)] = v;
  }
  __syncthreads();
  const int qi = blockIdx.y * blockDim.x + threadIdx.x;
  if (qi >= p.N) return;
  float q[ATT_KD];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(qi) * p.qkv_pitch);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 raw = qp[i];
      const __half2* hh = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hh[j]);
        q[i * 8 + 2 * j] = f.x * p.scale;
        q[i * 8 + 2 * j + 1] = f.y * p.scale;
      }
    }
  }
  float mx = -1e30f, sum = 0.f;
  float acc[ATT_HD];
#pragma unroll
  for (int i = 0; i < ATT_HD; ++i) acc[i] = 0.f;
  for (int j = 0; j < p.N; ++j) {
    float s = 0.f;
    const uint4* kp = reinterpret_cast<const uint4*>(ks + j * ATT_KD);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 raw = kp[i];
      const __half2* hh = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(hh[t]);
        s = fmaf(q[i * 8 + 2 * t], f.x, s);
        s = fmaf(q[i * 8 + 2 * t + 1], f.y, s);
      }
    }
    const float nm = fmaxf(mx, s);
    const float corr = __expf(mx - nm);
    const float pj = __expf(s - nm);
    sum = sum * corr + pj;
    mx = nm;
    const uint4* vp = reinterpret_cast<const uint4*>(vs + j * ATT_HD);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 raw = vp[i];
      const __half2* hh = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(hh[t]);
        acc[i * 8 + 2 * t] = fmaf(pj, f.x, acc[i * 8 + 2 * t] * corr);
        acc[i * 8 + 2 * t + 1] = fmaf(pj, f.y, acc[i * 8 + 2 * t + 1] * corr);
      }
    }
  }
  const float inv = 1.0f / sum;
  __half* op = p.out + (static_cast<size_t>(b) * p.N + qi) * p.out_pitch + h * ATT_HD;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 o;
    __half2* qo = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int t = 0; t < 4; ++t) qo[t] = __floats2half2_rn(acc[i * 8 + 2 * t] * inv, acc[i * 8 + 2 * t + 1] * inv);
    reinterpret_cast<uint4*>(op)[i] = o;
  }
}

// Copy the V part of qkv ([B,N,heads*(2kd+hd)]) into a dense [B,N,heads*hd] tensor (input of the PE depthwise conv).
struct VGatherParams {
  const __half* qkv; int qkv_pitch;
  __half* out; int out_pitch;
  long tokens; int heads;
};
__global__ void __launch_bounds__(256) gather_v_kernel(const VGatherParams p) {
  const int per_tok = p.heads * 8;  // 16-byte pieces of V per token
  const long total = p.tokens * per_tok;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int piece = static_cast<int>(idx % per_tok);
    const long tok = idx / per_tok;
    const int h = piece / 8, part = piece - h * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(p.qkv + tok * p.qkv_pitch + h * (2 * ATT_KD + ATT_HD) + 2 * ATT_KD +
                                                    part * 8);
    *reinterpret_cast<uint4*>(p.out + tok * p.out_pitch + h * ATT_HD + part * 8) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// direct convolution on CUDA cores: one thread per output element, fp32 accumulate.  Cross-check path only
// (xrseg_config.conv_impl = XRSEG_CONV_DIRECT); weights fp16 [Cout][k*k][Cin] (or [4 pos][Cout][Cin] transposed).
// ------------------------------------------------------------------------------------------------
struct DirectParams {
  const __half* in; int in_pitch;
  __half* out; int out_pitch;
  const __half* res; int res_pitch;
  const __half* w; const float* bias;
  int B, H, W, Cin, Ho, Wo, Cout, k, stride, pad, act, transposed;
};

__global__ void __launch_bounds__(256) conv_direct_kernel(const DirectParams p) {
  const long total = static_cast<long>(p.B) * p.Ho * p.Wo * p.Cout;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % p.Cout);
    const long pix = idx / p.Cout;
    const int ow = static_cast<int>(pix % p.Wo);
    const int oh = static_cast<int>((pix / p.Wo) % p.Ho);
    const long b = pix / (static_cast<long>(p.Wo) * p.Ho);
    float acc = p.bias[co];
    if (p.transposed) {
      const int ih = oh >> 1, iw = ow >> 1, pos = (oh & 1) * 2 + (ow & 1);
      const __half* xi = p.in + ((b * p.H + ih) * p.W + iw) * p.in_pitch;
      const __half* wt = p.w + (static_cast<size_t>(pos) * p.Cout + co) * p.Cin;
      for (int ci = 0; ci < p.Cin; ++ci) acc = fmaf(__half2float(xi[ci]), __half2float(wt[ci]), acc);
    } else {
      for (int kh = 0; kh < p.k; ++kh) {
        const int ih = oh * p.stride - p.pad + kh;
        if (ih < 0 || ih >= p.H) continue;
        for (int kw = 0; kw < p.k; ++kw) {
          const int iw = ow * p.stride - p.pad + kw;
          if (iw < 0 || iw >= p.W) continue;
          const __half* xi = p.in + ((b * p.H + ih) * p.W + iw) * p.in_pitch;
          const __half* wt = p.w + (static_cast<size_t>(co) * p.k * p.k + kh * p.k + kw) * p.Cin;
          for (int ci = 0; ci < p.Cin; ci += 2) {
            const float2 xv = __half22float2(*reinterpret_cast<const __half2*>(xi + ci));
            const float2 wv = __half22float2(*reinterpret_cast<const __half2*>(wt + ci));
            acc = fmaf(xv.x, wv.x, acc);
            acc = fmaf(xv.y, wv.y, acc);
          }
        }
      }
    }
    if (p.act) acc = silu_f(acc);
    if (p.res) acc += __half2float(p.res[pix * p.res_pitch + co]);
    p.out[pix * p.out_pitch + co] = __float2half_rn(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// layout converters for the debug / parity entry points
// ------------------------------------------------------------------------------------------------
__global__ void nchw_f32_to_nhwc_f16_kernel(const float* src, __half* dst, int B, int C, int H, int W, int Cpad,
                                            int pitch) {
  const long total = static_cast<long>(B) * H * W * Cpad;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % Cpad);
    const long pix = idx / Cpad;
    const int x = static_cast<int>(pix % W);
    const int y = static_cast<int>((pix / W) % H);
    const long b = pix / (static_cast<long>(W) * H);
    const float v = c < C ? src[((b * C + c) * H + y) * W + x] : 0.f;
    dst[pix * pitch + c] = __float2half_rn(v);
  }
}

__global__ void nhwc_f16_to_nchw_f32_kernel(const __half* src, float* dst, int B, int C, int H, int W, int pitch) {
  const long total = static_cast<long>(B) * C * H * W;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const int c = static_cast<int>((idx / (static_cast<long>(W) * H)) % C);
    const long b = idx / (static_cast<long>(W) * H * C);
    dst[idx] = __half2float(src[((b * H + y) * W + x) * pitch + c]);
  }
}

static inline int grid_for(long total, int block = 256, int cap = 148 * 16) {
  long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

#endif  // __CUDACC__
}  // namespace xrseg
