// Micro-benchmark: how fast can ONE SM pull tensor tiles through TMA (cp.async.bulk.tensor.2d), as a function of the box's
// inner row width (32 / 64 / 128 bytes = the 16 / 32 / 64-channel K-blocks of conv_tma.cuh), the number of SMs pulling at
// the same time (64 = one CTA per frame of a chain, 148 = every SM) and where the data lives (L2-resident vs HBM)?
// The convolution kernels move every activation byte this way, so this is their memory roofline.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/probe_tma_rate tools/probe_tma_rate.cu -lcuda && /tmp/probe_tma_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int DEPTH = 6;

// Each CTA streams `boxes` boxes of (box_rows x row_bytes) from its own slice of the tensor (rows pitch_bytes apart) through a
// DEPTH-deep ring; one thread issues, waits and re-issues.  mode 0: TMA tensor tiles; mode 1: plain bulk copies of the same size.
__global__ void pull_kernel(const __grid_constant__ CUtensorMap map, const uint8_t* base, int mode, int box_rows, int row_bytes,
                            long rows_per_cta, long pitch_bytes, int boxes, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[DEPTH];
  if (threadIdx.x != 0) return;
  for (int i = 0; i < DEPTH; ++i) mbar_init(&bar[i], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const uint32_t box_bytes = static_cast<uint32_t>(box_rows) * row_bytes;
  const uint32_t stage = (box_bytes + 1023) & ~1023u;
  const long row0 = static_cast<long>(blockIdx.x) * rows_per_cta;
  const long boxes_per_slice = rows_per_cta / box_rows;
  const long long t0 = clock64();
  for (int i = 0; i < boxes + DEPTH; ++i) {
    const int slot = i % DEPTH;
    if (i >= DEPTH) mbar_wait(&bar[slot], static_cast<uint32_t>((i / DEPTH) - 1) & 1u);
    if (i < boxes) {
      const long r = row0 + (static_cast<long>(i) % boxes_per_slice) * box_rows;
      mbar_expect(&bar[slot], box_bytes);
      if (mode == 0) tma_2d(smem_u32(smem) + slot * stage, &map, &bar[slot], 0, static_cast<int>(r));
      else bulk_1d(smem_u32(smem) + slot * stage, base + r * pitch_bytes, box_bytes, &bar[slot]);
    }
  }
  cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  PFN_encodeTiled encode = reinterpret_cast<PFN_encodeTiled>(fn);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  CK(cudaFuncSetAttribute(pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, sizeof(long long) * 256));
  printf("%-10s %-6s %-9s %-9s %-6s %-10s %10s %12s %12s\n", "source", "SMs", "row_bytes", "box_rows", "pitch", "mode", "B/clk/SM", "GB/s per SM", "TB/s chip");
  for (int big = 0; big < 2; ++big) {
    // L2-resident: 148 slices of 128 KB rows... total 32 MB; HBM: 4 GB touched once
    for (int pitch : {128, 256, 512}) {
      const long rows_per_cta = big ? (1L << 15) : 1024;          // rows of `pitch` bytes per CTA slice
      const long total_rows = rows_per_cta * sms;
      uint8_t* d = nullptr;
      CK(cudaMalloc(&d, static_cast<size_t>(total_rows) * pitch));
      CK(cudaMemset(d, 1, static_cast<size_t>(total_rows) * pitch));
      for (int row_bytes : {32, 64, 128}) {
        if (row_bytes > pitch) continue;
        for (int mode = 0; mode < 2; ++mode) {
          if (mode == 1 && row_bytes != pitch) continue;          // bulk copies: contiguous rows only
          for (int grid : {16, 64, sms}) {
            const int box_rows = 256;
            CUtensorMap m;
            cuuint64_t dims[2] = {static_cast<cuuint64_t>(row_bytes / 2), static_cast<cuuint64_t>(total_rows)};
            cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch)};
            cuuint32_t box[2] = {static_cast<cuuint32_t>(row_bytes / 2), static_cast<cuuint32_t>(box_rows)};
            cuuint32_t estr[2] = {1, 1};
            const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
            CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            const long bytes_per_cta = big ? rows_per_cta * row_bytes : 8L * 1024 * 1024;   // HBM: one pass; L2: re-read the slice
            const int boxes = static_cast<int>(bytes_per_cta / (static_cast<long>(box_rows) * row_bytes));
            for (int rep = 0; rep < 2; ++rep)
              pull_kernel<<<grid, 32, 200 * 1024>>>(m, d, mode, box_rows, row_bytes, rows_per_cta, pitch, boxes, d_cyc);
            CK(cudaDeviceSynchronize());
            long long h[256];
            CK(cudaMemcpy(h, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
            double mx = 0;
            for (int i = 0; i < grid; ++i) mx = mx > h[i] ? mx : static_cast<double>(h[i]);
            const double bpc = static_cast<double>(boxes) * box_rows * row_bytes / mx;
            const double ghz = prop.clockRate * 1e-6;
            printf("%-10s %-6d %-9d %-9d %-6d %-10s %10.1f %12.1f %12.2f\n", big ? "HBM" : "L2", grid, row_bytes, box_rows, pitch,
                   mode ? "bulk-1d" : "tma-tile", bpc, bpc * ghz, bpc * ghz * grid / 1e3);
          }
        }
      }
      CK(cudaFree(d));
    }
  }
  return 0;
}
