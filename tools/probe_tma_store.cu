// Developer probe: does a tiled TMA STORE (cp.async.bulk.tensor.4d.global.shared::cta) clip a box that starts at a negative
// coordinate / is wider than the tensor, the way loads zero-fill?   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/pts
// tools/probe_tma_store.cu -lcuda ; /tmp/pts <x0> <box_w> <W> [swizzle 0|3]
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__global__ void k(const __grid_constant__ CUtensorMap map, int x0, int y0, int rows) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __half* t = reinterpret_cast<__half*>(sm);
  for (int i = threadIdx.x; i < rows * 64; i += blockDim.x) t[i] = __float2half(1.0f + (i / 64));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(sm));
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<unsigned long long>(&map)),
                 "r"(s), "r"(0), "r"(x0), "r"(y0), "r"(0)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
int main(int argc, char** argv) {
  const int x0 = atoi(argv[1]), bw = atoi(argv[2]), W = atoi(argv[3]), sw = argc > 4 ? atoi(argv[4]) : 0;
  const int H = 6, R = 2, C = 64;
  __half* d;
  cudaMalloc(&d, sizeof(__half) * H * W * C);
  cudaMemset(d, 0, sizeof(__half) * H * W * C);
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 1};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)R, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      sw == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  k<<<1, 128, 16384>>>(m, x0, 1, R * bw);
  printf("launch: %s\n", cudaGetErrorString(cudaGetLastError()));
  cudaError_t e = cudaDeviceSynchronize();
  printf("x0 %d box_w %d W %d sw %d: %s\n", x0, bw, W, sw, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<__half> h(H * W * C);
    cudaMemcpy(h.data(), d, h.size() * 2, cudaMemcpyDeviceToHost);
    for (int y = 0; y < H; ++y) {
      for (int x = 0; x < W; ++x) printf("%4.0f", __half2float(h[(y * W + x) * C]));
      printf("\n");
    }
  }
  return 0;
}
