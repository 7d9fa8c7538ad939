// Probe (B200): does cuTensorMapEncodeTiled accept a dimension whose stride (32 B) is smaller than the extent of the dimension
// below it (64 fp16 = 128 B), i.e. OVERLAPPING rows - the horizontal im2col of a 16-channel space-to-depth image expressed
// as a tensor-map view - and does the TMA unit deliver the expected bytes?   nvcc -arch=sm_100a -o probe tma_overlap_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, __half* out, int x0, int y0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint8_t* tile = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(128 * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(tile)), "l"((uint64_t)&map), "r"((uint32_t)__cvta_generic_to_shared(&bar)),
                   "r"(0), "r"(x0), "r"(y0) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
  }
  // un-swizzle (SWIZZLE_128B: 16-byte chunk index ^= row & 7) and write rows out
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    int row = i / 64, k = i % 64;
    int chunk = (k / 8) ^ (row & 7);
    out[i] = *reinterpret_cast<__half*>(tile + row * 128 + chunk * 16 + (k % 8) * 2);
  }
}

int main() {
  const int W = 260, H = 8, C = 16;                      // rows of 260 pixels x 16 channels fp16 (32 B per pixel)
  std::vector<__half> h((size_t)H * W * C);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2half((float)(i % 2039));
  __half *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 128 * 64 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap map;
  cuuint64_t dims[3] = {64, (cuuint64_t)(W - 3), (cuuint64_t)H};           // K = 4 pixels x 16 ch; W-3 start positions
  cuuint64_t strides[2] = {32, (cuuint64_t)W * C * 2};                       // 32 B between consecutive "rows": overlapping
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t es[3] = {1, 1, 1};
  cuInit(0);
  CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: CUresult %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  for (int trial = 0; trial < 2; ++trial) {
    const int x0 = trial ? 129 : 0, y0 = trial ? 5 : 2;   // second trial: box runs past the last start position (OOB rows -> 0)
    probe<<<1, 128, 128 * 128 + 1024>>>(map, o, x0, y0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<__half> got(128 * 64);
    cudaMemcpy(got.data(), o, got.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int row = 0; row < 128; ++row)
      for (int k = 0; k < 64; ++k) {
        const int x = x0 + row;
        float want = (x < W - 3) ? __half2float(h[((size_t)y0 * W + x) * C + k]) : 0.0f;
        if (__half2float(got[row * 64 + k]) != want && bad++ < 5) printf("  mismatch row %d k %d: got %g want %g\n", row, k, __half2float(got[row * 64 + k]), want);
      }
    printf("trial %d (x0=%d y0=%d): %d mismatches of %d\n", trial, x0, y0, bad, 128 * 64);
  }
  return 0;
}
