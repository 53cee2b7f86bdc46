// bc_deconv.cu -- depth-to-space on packed NHWC fp16 tiles: the second half of a per-block ConvTranspose2d.
//
// The reference runs ConvTranspose2d as a pass-through op on the packed tile batch (core/tensorwrapper.py:519-520:
// every tile is an independent sample with zeros beyond its edge; Pedestron necks/csp_neck.py:37-83 upsamples the
// ResNet stages that way).  Here a transposed conv whose kernel covers each output pixel from a fixed set of input
// pixels is a plain conv with r*r times the output channels (bc_conv_igemm over the tile batch viewed as E one-block
// frames: the zero frame border IS the tile border), followed by this rearrangement:
//     out[e, r*y + a, r*x + b, c] = in[e, y, x, (a*r + b)*C + c]
// For a fixed (e, y, a) the r*C*w output elements of row r*y+a are w segments of r*C contiguous input elements.
// Pure copy: one 16-byte vector per thread and step, both sides coalesced over the r*C-element segments.
#include <cuda_fp16.h>

#include "bc_common.cuh"

namespace bc {

struct D2sParams {
  const uint4 *in;
  uint4 *out;
  FastDiv seg, per_row, per_a;  // vectors per segment (r*C/8), segments per output row (w), rows per input row (r)
  uint32_t total;               // E*h*r*w*(r*C/8)
  uint32_t in_pixel_vecs;       // r*r*C/8
  uint32_t h, w, r;
};

__global__ void __launch_bounds__(256) depth_to_space_kernel(const D2sParams p) {
  pdl_trigger();
  pdl_wait();
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += stride) {
    // i enumerates the OUTPUT in memory order: (e, y, a, x, [b, c] in vectors)
    uint32_t rest, v, x, a, ey;
    p.seg.divmod(i, rest, v);
    p.per_row.divmod(rest, rest, x);
    p.per_a.divmod(rest, ey, a);
    const size_t src = ((size_t)ey * p.w + x) * p.in_pixel_vecs + (size_t)a * p.seg.d + v;
    p.out[i] = __ldg(p.in + src);
  }
}

int depth_to_space(void *out, const void *in, int E, int C, int h, int w, int r, cudaStream_t stream) {
  BC_REQUIRE(out && in, BC_ERR_NULL, "bc_depth_to_space: NULL pointer");
  BC_REQUIRE(E > 0 && C > 0 && h > 0 && w > 0 && r >= 1 && r <= 8, BC_ERR_SHAPE, "bc_depth_to_space: E=%d C=%d %dx%d r=%d", E, C,
             h, w, r);
  BC_REQUIRE(C % 8 == 0, BC_ERR_UNSUPPORTED, "bc_depth_to_space: C=%d is not a multiple of 8", C);
  BC_REQUIRE((((uintptr_t)out | (uintptr_t)in) & 15) == 0, BC_ERR_ALIGN, "bc_depth_to_space: pointers must be 16-byte aligned");
  const long long total = (long long)E * h * w * r * r * (C / 8);
  BC_REQUIRE(total < (1ll << 31), BC_ERR_RANGE, "bc_depth_to_space: %lld vectors", total);
  D2sParams p;
  p.in = (const uint4 *)in;
  p.out = (uint4 *)out;
  p.seg = FastDiv((uint32_t)(r * C / 8));
  p.per_row = FastDiv((uint32_t)w);
  p.per_a = FastDiv((uint32_t)r);
  p.total = (uint32_t)total;
  p.in_pixel_vecs = (uint32_t)(r * r * C / 8);
  p.h = (uint32_t)h; p.w = (uint32_t)w; p.r = (uint32_t)r;
  long long grid = (total + 255) / 256;
  if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
  launch_kernel(depth_to_space_kernel, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
  return check_launch("bc_depth_to_space");
}

}  // namespace bc
