// Micro-benchmark: throughput of the SiLU formulations per SM (elements / clock): tanh.approx.f32, ex2 + rcp (fp32),
// tanh.approx.f16x2, and an FMA-pipe polynomial exp2 + MUFU.RCP.  16 independent values per thread, 8 / 16 / 32 warps per SM.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/probe_mufu tools/probe_mufu.cu && /tmp/probe_mufu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ float tanh_f32(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_f32(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_f32(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__device__ __forceinline__ float silu(float acc, float hb) {
  if (MODE == 0) {                       // h tanh(h) + h, h = 0.5 y
    const float h = fmaf(acc, 0.5f, hb);
    return fmaf(h, tanh_f32(h), h);
  } else if (MODE == 1) {                // y / (1 + 2^(-y log2 e))
    const float y = acc + hb;
    const float e = ex2_f32(y * -1.4426950408889634f);
    return y * rcp_f32(1.0f + e);
  } else if (MODE == 3) {                // exp2 on the FMA pipe (Cody-Waite + cubic), reciprocal on the MUFU
    const float y = acc + hb;
    float t = fmaxf(y * -1.4426950408889634f, -126.0f);
    const float fl = floorf(t);
    const float f = t - fl;
    float p = fmaf(f, 0.0555054f, 0.2402265f);
    p = fmaf(p, f, 0.6931472f);
    p = fmaf(p, f, 1.0f);
    const float e = __int_as_float(__float_as_int(p) + (static_cast<int>(fl) << 23));
    return y * rcp_f32(1.0f + e);
  }
  return acc;
}

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed + 0.01f * i + 0.001f * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 2) {                     // packed fp16: two values per MUFU op
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        __half2 h = __floats2half2_rn(0.5f * v[i] + 0.1f, 0.5f * v[i + 1] + 0.1f);
        unsigned hu = *reinterpret_cast<unsigned*>(&h), tu;
        asm("tanh.approx.f16x2 %0, %1;" : "=r"(tu) : "r"(hu));
        __half2 r = __hfma2(h, *reinterpret_cast<__half2*>(&tu), h);
        const float2 f = __half22float2(r);
        v[i] = f.x; v[i + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = silu<MODE>(v[i], 0.1f);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  float* d;
  cudaMalloc(&d, sizeof(float) * 148 * 1024 * 4);
  const char* names[4] = {"tanh.approx.f32 (1 MUFU)", "ex2 + rcp f32 (2 MUFU)", "tanh.approx.f16x2 (0.5 MUFU)", "poly exp2 + rcp (1 MUFU)"};
  for (int warps : {8, 16, 32}) {
    for (int mode = 0; mode < 4; ++mode) {
      const int iters = 4096;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<prop.multiProcessorCount, warps * 32>>>(d, iters, 0.3f);
        if (mode == 1) k<1><<<prop.multiProcessorCount, warps * 32>>>(d, iters, 0.3f);
        if (mode == 2) k<2><<<prop.multiProcessorCount, warps * 32>>>(d, iters, 0.3f);
        if (mode == 3) k<3><<<prop.multiProcessorCount, warps * 32>>>(d, iters, 0.3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double elems = static_cast<double>(iters) * 16 * warps * 32;      // per SM
      const double clk = ms * 1e-3 * prop.clockRate * 1e3;
      printf("%2d warps/SM  %-30s %6.2f SiLU / clk / SM\n", warps, names[mode], elems / clk);
    }
  }
  return 0;
}
