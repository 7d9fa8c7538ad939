// Input side of the hot path: uint8 HWC images (what cv2.imread + cv2.resize produce, reference
// centernet_lightning/datasets/inference.py:28-29 and README.md:84-87 A.Resize -> A.Normalize -> ToTensorV2) are
// normalised and transposed to the fp32 NCHW tensor the model takes - on the GPU, so the host copies 1 byte per
// sample instead of 4 and spends no core time on arithmetic.
//
// albumentations.Normalize(mean, std, max_pixel_value=255) computes, in numpy:
//     img = img.astype(float32);  img -= mean * 255 (float64 array: the subtraction is evaluated in double and rounded
//     to float32);  img *= float32(1 / (std * 255))
// The kernel performs exactly those roundings (double subtract -> float -> float multiply), so the result is bit-exact.
#include <cuda_runtime.h>
#include <cstdint>
#include "cnl_common.h"

namespace cnl {

struct NormParams {
  double mean255[3];
  float inv_std255[3];
};

__device__ __forceinline__ float norm1(uint8_t v, double m, float s) {
  return __fmul_rn(__double2float_rn(__dsub_rn((double)v, m)), s);
}

// one thread = 4 consecutive pixels of one row (12 input bytes, three float4 stores); requires W % 4 == 0
__global__ void __launch_bounds__(256)
normalize_u8_vec4_kernel(const uint32_t* __restrict__ in, float* __restrict__ out, long long n_quads, int hw_quads, NormParams p) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_quads) return;
  const long long img = t / hw_quads;
  const long long q = t - img * hw_quads;
  const uint32_t a = __ldg(in + 3 * t), b = __ldg(in + 3 * t + 1), c = __ldg(in + 3 * t + 2);
  // bytes: a = r0 g0 b0 r1 | b = g1 b1 r2 g2 | c = b2 r3 g3 b3   (little endian)
  const uint8_t r[4] = {(uint8_t)a, (uint8_t)(a >> 24), (uint8_t)(b >> 16), (uint8_t)(c >> 8)};
  const uint8_t g[4] = {(uint8_t)(a >> 8), (uint8_t)b, (uint8_t)(b >> 24), (uint8_t)(c >> 16)};
  const uint8_t bl[4] = {(uint8_t)(a >> 16), (uint8_t)(b >> 8), (uint8_t)c, (uint8_t)(c >> 24)};
  const long long plane = (long long)hw_quads * 4;
  float* o = out + img * 3 * plane + q * 4;
  *reinterpret_cast<float4*>(o) = make_float4(norm1(r[0], p.mean255[0], p.inv_std255[0]), norm1(r[1], p.mean255[0], p.inv_std255[0]),
                                              norm1(r[2], p.mean255[0], p.inv_std255[0]), norm1(r[3], p.mean255[0], p.inv_std255[0]));
  *reinterpret_cast<float4*>(o + plane) = make_float4(norm1(g[0], p.mean255[1], p.inv_std255[1]), norm1(g[1], p.mean255[1], p.inv_std255[1]),
                                                      norm1(g[2], p.mean255[1], p.inv_std255[1]), norm1(g[3], p.mean255[1], p.inv_std255[1]));
  *reinterpret_cast<float4*>(o + 2 * plane) = make_float4(norm1(bl[0], p.mean255[2], p.inv_std255[2]), norm1(bl[1], p.mean255[2], p.inv_std255[2]),
                                                          norm1(bl[2], p.mean255[2], p.inv_std255[2]), norm1(bl[3], p.mean255[2], p.inv_std255[2]));
}

// any size: one thread = one pixel
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, long long n_pix, long long plane, NormParams p) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pix) return;
  const long long img = t / plane, q = t - img * plane;
  float* o = out + img * 3 * plane + q;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c * plane] = norm1(__ldg(in + 3 * t + c), p.mean255[c], p.inv_std255[c]);
}

}  // namespace cnl

using namespace cnl;

extern "C" int cnl_normalize_images_u8(const uint8_t* images_hwc, float* images_nchw, int n, int h, int w,
                                       const double* mean255, const float* inv_std255, void* stream) {
  if (!images_hwc || !images_nchw || !mean255 || !inv_std255) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_normalize_images_u8: null pointer argument");
  if (n <= 0 || h <= 0 || w <= 0) return fail(CNL_ERR_INVALID_ARGUMENT, "cnl_normalize_images_u8: bad shape (%d,%d,%d)", n, h, w);
  NormParams p;
  for (int c = 0; c < 3; ++c) { p.mean255[c] = mean255[c]; p.inv_std255[c] = inv_std255[c]; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n_pix = (long long)n * h * w;
  const bool vec = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(images_hwc) & 3) == 0) && ((reinterpret_cast<uintptr_t>(images_nchw) & 15) == 0);
  if (vec) {
    const long long n_quads = n_pix / 4;
    normalize_u8_vec4_kernel<<<(unsigned)((n_quads + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint32_t*>(images_hwc), images_nchw,
                                                                               n_quads, h * w / 4, p);
  } else {
    normalize_u8_kernel<<<(unsigned)((n_pix + 255) / 256), 256, 0, st>>>(images_hwc, images_nchw, n_pix, (long long)h * w, p);
  }
  CNL_CUDA_CHECK(cudaGetLastError());
  return CNL_OK;
}
