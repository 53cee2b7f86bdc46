// bc_detect.cu -- box rasteriser behind InformationGainObjectDetection (reference
// blockcopy/policy/information_gain.py:56-108: build_instance_mask / build_instance_mask_iou_gain paint every
// detection box into a mask with `mask[y1:y2, x1:x2] = max(mask[y1:y2, x1:x2], value)`, one torch slice
// assignment per box from a Python loop).  Here: one pass over the output, every pixel takes the maximum over
// the boxes that contain it; the max is order independent, so the result equals the reference's loop bit for bit.
#include <cuda_runtime.h>

#include "bc_common.cuh"

namespace bc {

constexpr int kBoxChunk = 512;

// out[y][x] = max(0, max{ value[i] : x1_i <= (x >> shift) < x2_i and y1_i <= (y >> shift) < y2_i })
// (shift = 1: the reference rasterises at half resolution and upsamples with nearest x2, information_gain.py:72,105)
__global__ void __launch_bounds__(256) raster_boxes_kernel(float *__restrict__ out, const int32_t *__restrict__ rects,
                                                           const float *__restrict__ values, int n, int H, int W, int shift) {
  __shared__ int4 box_s[kBoxChunk];
  __shared__ float val_s[kBoxChunk];
  pdl_trigger();
  pdl_wait();
  const int x0 = (blockIdx.x * 64 + (threadIdx.x & 15) * 4), y = blockIdx.y * 16 + (threadIdx.x >> 4);
  const int ys = y >> shift;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  // boxes that miss this CTA's 64x16 pixel window are dropped while staging
  const int wx0 = (blockIdx.x * 64) >> shift, wx1 = (blockIdx.x * 64 + 63) >> shift;
  const int wy0 = (blockIdx.y * 16) >> shift, wy1 = (blockIdx.y * 16 + 15) >> shift;
  for (int base = 0; base < n; base += kBoxChunk) {
    const int cnt = min(kBoxChunk, n - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += 256) {
      int4 b = __ldg(reinterpret_cast<const int4 *>(rects) + base + i);
      if (b.x > wx1 || b.z <= wx0 || b.y > wy1 || b.w <= wy0) b = make_int4(0, 0, 0, 0);  // empty
      box_s[i] = b;
      val_s[i] = __ldg(values + base + i);
    }
    __syncthreads();
    for (int i = 0; i < cnt; ++i) {
      const int4 b = box_s[i];
      if (ys < b.y || ys >= b.w) continue;
      const float s = val_s[i];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int xs = (x0 + k) >> shift;
        if (xs >= b.x && xs < b.z) v[k] = fmaxf(v[k], s);
      }
    }
  }
  if (y < H) {
    if (x0 + 3 < W && (W & 3) == 0) {
      *reinterpret_cast<float4 *>(out + (size_t)y * W + x0) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      for (int k = 0; k < 4; ++k)
        if (x0 + k < W) out[(size_t)y * W + x0 + k] = v[k];
    }
  }
}

int raster_boxes(float *out, const int32_t *rects, const float *values, int n, int H, int W, int shift, cudaStream_t stream) {
  BC_REQUIRE(out != nullptr && H > 0 && W > 0, BC_ERR_NULL, "bc_raster_boxes: NULL output / empty image");
  BC_REQUIRE(n == 0 || (rects != nullptr && values != nullptr), BC_ERR_NULL, "bc_raster_boxes: NULL box list");
  BC_REQUIRE(shift >= 0 && shift <= 4, BC_ERR_UNSUPPORTED, "bc_raster_boxes: shift %d (0..4)", shift);
  BC_REQUIRE((((uintptr_t)out | (uintptr_t)rects) & 15) == 0, BC_ERR_ALIGN, "bc_raster_boxes: pointers must be 16-byte aligned");
  launch_kernel(raster_boxes_kernel, dim3((unsigned)((W + 63) / 64), (unsigned)((H + 15) / 16)), dim3(256), 0, stream, 1, out,
                rects, values, n, H, W, shift);
  return check_launch("bc_raster_boxes");
}

}  // namespace bc
