// bc_io.cu -- the two steps on either side of the block path in the reference's driver.
//   bc_frame_from_u8   : decoded uint8 frame (N,H,W,3) -> normalised (N,3,H,W) fp16/fp32 network input, i.e.
//                        ExtToTensor (x / 255) + ExtNormalize ((x - mean) / std) + the .to(half) of the driver
//                        (semantic_segmentation/lib/ext_transforms.py:317-372, test_swiftnet.py:64-65,187) in one
//                        pass; the host then uploads 3 bytes per pixel instead of 6 (fp16) or 12 (fp32).
//   bc_upsample_argmax : class map of the dense logits, = F.interpolate(out, size, mode='bilinear')
//                        followed by .max(dim=1)[1] (test_swiftnet.py:196-197): the (N,K,sH,sW) upsampled
//                        tensor (80 MB at 19x1024x2048 fp16) is never written; every thread owns one logit
//                        pixel, loads its 3x3 neighbourhood once per class and keeps the running best of its
//                        s x s output pixels in registers.
// Arithmetic follows ATen op by op: division by 255 and by std are IEEE divisions in fp32; the bilinear blend
// is  h0*(w0*a + w1*b) + h1*(w0*c + w1*d)  in fp32 (upsample_bilinear2d's accscalar_t), rounded to the logits'
// dtype before the comparison, ties -> lowest class index.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "bc_common.cuh"

namespace bc {

// ---------------------------------------------------------------------------------------------------
struct U8Params {
  const uint8_t *src;  // (N,H,W,3)
  void *out;           // (N,3,H,W)
  float mean[3], std[3];
  uint32_t groups_per_image;  // H*W / 16
  uint32_t total_groups;      // N * H*W / 16
  uint32_t hw;
};

template <typename T> __device__ __forceinline__ T cvt_out(float v);
template <> __device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }

__device__ __forceinline__ float norm_u8(uint32_t u, float mean, float std) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.f), mean), std);
}

// 16 pixels (48 bytes in, 16 values per channel plane out) per thread: 3 x 16-byte loads, 16-byte stores.  A byte
// has 256 values: every CTA first computes the 3 x 256 results (same arithmetic, so the same bits) into shared
// memory, and the per-pixel work is a table look-up instead of two IEEE divisions.
template <typename T>
__global__ void __launch_bounds__(256) frame_from_u8_kernel(const U8Params p) {
  __shared__ T lut[3][256];
  pdl_trigger();  // first: the next kernel's launch latency and prologue overlap this whole kernel
  for (int k = threadIdx.x; k < 768; k += 256) lut[k >> 8][k & 255] = cvt_out<T>(norm_u8(k & 255, p.mean[k >> 8], p.std[k >> 8]));
  __syncthreads();
  pdl_wait();
  const uint32_t gstride = gridDim.x * blockDim.x;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < p.total_groups; g += gstride) {
    const uint32_t n = g / p.groups_per_image, r = g - n * p.groups_per_image;
    const uint4 *src = reinterpret_cast<const uint4 *>(p.src + ((size_t)n * p.hw + (size_t)r * 16) * 3);
    uint32_t wd[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint4 q = __ldg(src + i);
      wd[4 * i] = q.x; wd[4 * i + 1] = q.y; wd[4 * i + 2] = q.z; wd[4 * i + 3] = q.w;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T f[16];
#pragma unroll
      for (int px = 0; px < 16; ++px) {
        const int k = px * 3 + c;  // byte index within the 48-byte group (compile-time after unrolling)
        f[px] = lut[c][(wd[k >> 2] >> ((k & 3) * 8)) & 0xffu];
      }
      T *dst = reinterpret_cast<T *>(p.out) + ((size_t)n * 3 + c) * p.hw + (size_t)r * 16;
      if constexpr (sizeof(T) == 2) {
        uint32_t h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          h[k] = (uint32_t)__half_as_ushort(f[2 * k]) | ((uint32_t)__half_as_ushort(f[2 * k + 1]) << 16);
        reinterpret_cast<uint4 *>(dst)[0] = make_uint4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<uint4 *>(dst)[1] = make_uint4(h[4], h[5], h[6], h[7]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          reinterpret_cast<uint4 *>(dst)[k] = make_uint4(__float_as_uint(f[4 * k]), __float_as_uint(f[4 * k + 1]),
                                                         __float_as_uint(f[4 * k + 2]), __float_as_uint(f[4 * k + 3]));
      }
    }
  }
}

// any size / alignment: one pixel per thread
template <typename T>
__global__ void __launch_bounds__(256) frame_from_u8_generic_kernel(const U8Params p, uint32_t total_px) {
  pdl_trigger();
  pdl_wait();
  const uint32_t gstride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total_px; i += gstride) {
    const uint32_t n = i / p.hw, r = i - n * p.hw;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      reinterpret_cast<T *>(p.out)[((size_t)n * 3 + c) * p.hw + r] =
          cvt_out<T>(norm_u8(p.src[(size_t)i * 3 + c], p.mean[c], p.std[c]));
  }
}

int frame_from_u8(void *out, const uint8_t *src, const float *mean, const float *std, int N, int H, int W, int dtype,
                  cudaStream_t stream) {
  BC_REQUIRE(out && src && mean && std, BC_ERR_NULL, "bc_frame_from_u8: NULL pointer");
  BC_REQUIRE(N > 0 && H > 0 && W > 0, BC_ERR_SHAPE, "bc_frame_from_u8: empty frame");
  BC_REQUIRE(dtype == BC_F16 || dtype == BC_F32, BC_ERR_DTYPE, "bc_frame_from_u8: dtype");
  const int64_t total_px = (int64_t)N * H * W;
  BC_REQUIRE(total_px < (1ll << 31), BC_ERR_RANGE, "bc_frame_from_u8: problem too large");
  U8Params p;
  p.src = src;
  p.out = out;
  for (int c = 0; c < 3; ++c) {
    BC_REQUIRE(std[c] != 0.f, BC_ERR_RANGE, "bc_frame_from_u8: std[%d] is zero", c);
    p.mean[c] = mean[c];
    p.std[c] = std[c];
  }
  p.hw = (uint32_t)(H * W);
  const bool vec = (p.hw % 16 == 0) && (((uintptr_t)src | (uintptr_t)out) & 15) == 0;
  if (vec) {
    p.groups_per_image = p.hw / 16;
    p.total_groups = (uint32_t)(total_px / 16);
    int64_t grid = (p.total_groups + 255) / 256;
    if (grid > (int64_t)kNumSMs * 4) grid = (int64_t)kNumSMs * 4;
    if (dtype == BC_F16)
      launch_kernel(frame_from_u8_kernel<__half>, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
    else
      launch_kernel(frame_from_u8_kernel<float>, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
  } else {
    p.groups_per_image = p.total_groups = 0;
    int64_t grid = (total_px + 255) / 256;
    if (grid > (int64_t)kNumSMs * 16) grid = (int64_t)kNumSMs * 16;
    if (dtype == BC_F16)
      launch_kernel(frame_from_u8_generic_kernel<__half>, dim3((unsigned)grid), dim3(256), 0, stream, 1, p, (uint32_t)total_px);
    else
      launch_kernel(frame_from_u8_generic_kernel<float>, dim3((unsigned)grid), dim3(256), 0, stream, 1, p, (uint32_t)total_px);
  }
  return check_launch("bc_frame_from_u8");
}

// ---------------------------------------------------------------------------------------------------
// bc_blocks_from_u8: the first gather of a frame straight from the decoded uint8 frame (input side):
// tiles[e] (3, BS, BS) <- normalise(src[n, gh*BS .. , gw*BS .. , :]) for the E executed cells -- bc_frame_from_u8 followed
// by bc_gather (reference tensorwrapper.py:335-381 on the output of ext_transforms.py:317-372) without ever writing
// the normalised full frame: 3 bytes per pixel of the executed blocks are read instead of 3 + 2 x 6.  Same table
// look-up as frame_from_u8_kernel, hence the same bits.  BS and W multiples of 16.
struct U8BlocksParams {
  const uint8_t *src;      // (N,H,W,3)
  void *tiles;             // (E,3,BS,BS)
  const int32_t *mapping;  // cell of tile e
  float mean[3], std[3];
  CellDecode cell;
  FastDiv groups_per_tile, groups_per_row;  // BS*BS/16, BS/16
  uint32_t total_groups;                    // E * BS*BS/16
  int H, W, BS;
};

template <typename T>
__global__ void __launch_bounds__(256) blocks_from_u8_kernel(const U8BlocksParams p) {
  __shared__ T lut[3][256];
  pdl_trigger();
  for (int k = threadIdx.x; k < 768; k += 256) lut[k >> 8][k & 255] = cvt_out<T>(norm_u8(k & 255, p.mean[k >> 8], p.std[k >> 8]));
  __syncthreads();
  pdl_wait();
  const uint32_t gstride = gridDim.x * blockDim.x;
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < p.total_groups; g += gstride) {
    uint32_t e, rem, y, xs, n, gh, gw;
    p.groups_per_tile.divmod(g, e, rem);
    p.groups_per_row.divmod(rem, y, xs);
    p.cell((uint32_t)__ldg(p.mapping + e), n, gh, gw);
    const size_t px = ((size_t)n * p.H + gh * p.BS + y) * p.W + gw * p.BS + xs * 16;
    const uint4 *src = reinterpret_cast<const uint4 *>(p.src + px * 3);
    uint32_t wd[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint4 q = __ldg(src + i);
      wd[4 * i] = q.x; wd[4 * i + 1] = q.y; wd[4 * i + 2] = q.z; wd[4 * i + 3] = q.w;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T f[16];
#pragma unroll
      for (int k16 = 0; k16 < 16; ++k16) {
        const int k = k16 * 3 + c;
        f[k16] = lut[c][(wd[k >> 2] >> ((k & 3) * 8)) & 0xffu];
      }
      T *dst = reinterpret_cast<T *>(p.tiles) + (((size_t)e * 3 + c) * p.BS + y) * p.BS + xs * 16;
      if constexpr (sizeof(T) == 2) {
        uint32_t h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          h[k] = (uint32_t)__half_as_ushort(f[2 * k]) | ((uint32_t)__half_as_ushort(f[2 * k + 1]) << 16);
        reinterpret_cast<uint4 *>(dst)[0] = make_uint4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<uint4 *>(dst)[1] = make_uint4(h[4], h[5], h[6], h[7]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          reinterpret_cast<uint4 *>(dst)[k] = make_uint4(__float_as_uint(f[4 * k]), __float_as_uint(f[4 * k + 1]),
                                                         __float_as_uint(f[4 * k + 2]), __float_as_uint(f[4 * k + 3]));
      }
    }
  }
}

int blocks_from_u8(void *tiles, const uint8_t *src, const float *mean, const float *std, const int32_t *mapping, int E, int N,
                   int H, int W, int BS, int dtype, cudaStream_t stream) {
  BC_REQUIRE(tiles && src && mean && std && mapping, BC_ERR_NULL, "bc_blocks_from_u8: NULL pointer");
  BC_REQUIRE(E > 0 && N > 0 && H > 0 && W > 0 && BS > 0 && H % BS == 0 && W % BS == 0, BC_ERR_SHAPE,
             "bc_blocks_from_u8: E=%d N=%d %dx%d block %d", E, N, H, W, BS);
  BC_REQUIRE(BS % 16 == 0, BC_ERR_UNSUPPORTED, "bc_blocks_from_u8: block edge %d is not a multiple of 16", BS);
  BC_REQUIRE(dtype == BC_F16 || dtype == BC_F32, BC_ERR_DTYPE, "bc_blocks_from_u8: dtype");
  BC_REQUIRE((((uintptr_t)tiles | (uintptr_t)src) & 15) == 0, BC_ERR_ALIGN, "bc_blocks_from_u8: 16-byte alignment");
  const int64_t total = (int64_t)E * BS * BS / 16;
  BC_REQUIRE(total < (1ll << 31) && (int64_t)N * H * W < (1ll << 31), BC_ERR_RANGE, "bc_blocks_from_u8: problem too large");
  U8BlocksParams p;
  p.src = src; p.tiles = tiles; p.mapping = mapping;
  for (int c = 0; c < 3; ++c) {
    BC_REQUIRE(std[c] != 0.f, BC_ERR_RANGE, "bc_blocks_from_u8: std[%d] is zero", c);
    p.mean[c] = mean[c];
    p.std[c] = std[c];
  }
  p.cell = CellDecode(H / BS, W / BS);
  p.groups_per_tile = FastDiv((uint32_t)(BS * BS / 16));
  p.groups_per_row = FastDiv((uint32_t)(BS / 16));
  p.total_groups = (uint32_t)total;
  p.H = H; p.W = W; p.BS = BS;
  int64_t grid = (total + 255) / 256;
  if (grid > (int64_t)kNumSMs * 4) grid = (int64_t)kNumSMs * 4;
  if (dtype == BC_F16)
    launch_kernel(blocks_from_u8_kernel<__half>, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
  else
    launch_kernel(blocks_from_u8_kernel<float>, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
  return check_launch("bc_blocks_from_u8");
}

// ---------------------------------------------------------------------------------------------------
struct ArgmaxParams {
  const void *logits;  // (N,K,h,w), element strides sn, sc, sh, sw
  void *labels;        // (N, h*S, w*S) uint8 or int64, contiguous
  int N, K, h, w;
  int64_t sn, sc, sh, sw;
  uint32_t total;      // N*h*w
  // block-sparse update (bc_upsample_argmax_blocks): executed cells of the (N,1,GH,GW) grid, logit block edge BSl
  const uint8_t *grid;
  int GH, GW, BSl, chunks;  // chunks = CTAs per cell = ceil((BSl+2)^2 / 128)
};

template <typename T> __device__ __forceinline__ float ld_f(const T *p);
template <> __device__ __forceinline__ float ld_f<__half>(const __half *p) { return __half2float(__ldg(p)); }
template <> __device__ __forceinline__ float ld_f<float>(const float *p) { return __ldg(p); }
template <typename T> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<__half>(float v) { return __half2float(__float2half_rn(v)); }
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }

// Integer scale S, align_corners = False.  Output row S*i + r has source coordinate i + (r + 0.5)/S - 0.5
// (ATen: scale * (dst + 0.5) - 0.5 in fp32, clamped at 0): for r < S/2 it lies in [i-1, i) -> taps (i-1, i),
// otherwise in [i, i+1) -> taps (i, min(i+1, last)).  The only exception is the clamp at i = 0, r < S/2, where
// ATen blends rows (0, 1) with weights (1, 0): the same value, bit for bit, as rows (0, 0) with weights (0, 1), so
// the tap pair is a compile-time function of r and only the weights are per thread.
// Per class: 9 loads, 3 x S horizontal blends shared by the S output rows, S x S vertical blends; for fp16 logits
// the rounded values are compared as packed half2 (2 outputs per instruction, class indices in 16-bit lanes).
template <typename T, typename L, int S>
__device__ __forceinline__ void argmax_pixel(const ArgmaxParams &p, const int n, const int i, const int j) {
  constexpr float rs = 1.f / (float)S;  // ATen: area_pixel_compute_scale = in / out (size= call), exact for S = 2^k
  constexpr int kHalf = S / 2;          // offsets r < kHalf use taps (-1, 0), the others (0, +1)
  float h0[S], h1[S], w0[S], w1[S];
#pragma unroll
  for (int r = 0; r < S; ++r) {
    float sy = rs * ((float)(S * i + r) + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    const float ly = sy - (float)(int)sy;
    float sx = rs * ((float)(S * j + r) + 0.5f) - 0.5f;
    sx = sx < 0.f ? 0.f : sx;
    const float lx = sx - (float)(int)sx;
    if (r < kHalf) {  // taps (i-1, i): at i == 0 the clamped coordinate is exactly row 0
      h0[r] = i == 0 ? 0.f : 1.f - ly; h1[r] = i == 0 ? 1.f : ly;
      w0[r] = j == 0 ? 0.f : 1.f - lx; w1[r] = j == 0 ? 1.f : lx;
    } else {
      h0[r] = 1.f - ly; h1[r] = ly;
      w0[r] = 1.f - lx; w1[r] = lx;
    }
  }
  const int ys[3] = {max(i - 1, 0), i, min(i + 1, p.h - 1)}, xs[3] = {max(j - 1, 0), j, min(j + 1, p.w - 1)};
  int off[3][3];  // element offsets within one (n, class) slice: < 2^31, checked on the host
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) off[a][b] = (int)(ys[a] * p.sh + xs[b] * p.sw);
  const T *pc = reinterpret_cast<const T *>(p.logits) + n * p.sn;

  constexpr bool kPacked = sizeof(T) == 2 && S % 2 == 0;
  constexpr int kPairs = kPacked ? S / 2 : 1;
  float best[S][S];           // scalar path
  int arg[S][S];
  __half2 best2[S][kPairs];   // packed path: outputs (a, 2q) and (a, 2q+1)
  uint32_t arg2[S][kPairs];
#pragma unroll
  for (int a = 0; a < S; ++a) {
#pragma unroll
    for (int b = 0; b < S; ++b) { best[a][b] = -INFINITY; arg[a][b] = 0; }
#pragma unroll
    for (int q = 0; q < kPairs; ++q) { best2[a][q] = __float2half2_rn(-INFINITY); arg2[a][q] = 0u; }
  }
  for (int c = 0; c < p.K; ++c, pc += p.sc) {
    float v[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) v[a][b] = ld_f<T>(pc + off[a][b]);
    float hb[3][S];  // horizontal blends of the three rows
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < S; ++b) {
        const int cb = b < kHalf ? 0 : 1;
        hb[a][b] = w0[b] * v[a][cb] + w1[b] * v[a][cb + 1];
      }
    const uint32_t c2 = (uint32_t)c * 0x10001u;
#pragma unroll
    for (int a = 0; a < S; ++a) {
      const int ra = a < kHalf ? 0 : 1;
      if constexpr (kPacked) {
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
          const float va = h0[a] * hb[ra][2 * q] + h1[a] * hb[ra + 1][2 * q];
          const float vb = h0[a] * hb[ra][2 * q + 1] + h1[a] * hb[ra + 1][2 * q + 1];
          const __half2 val = __floats2half2_rn(va, vb);
          const uint32_t m = __hgt2_mask(val, best2[a][q]);  // 0xffff per lane where val > best (false for NaN)
          best2[a][q] = __hmax2(best2[a][q], val);
          arg2[a][q] = (arg2[a][q] & ~m) | (c2 & m);
        }
      } else {
#pragma unroll
        for (int b = 0; b < S; ++b) {
          const float val = round_to<T>(h0[a] * hb[ra][b] + h1[a] * hb[ra + 1][b]);
          if (val > best[a][b]) { best[a][b] = val; arg[a][b] = c; }
        }
      }
    }
  }
  if constexpr (kPacked) {
#pragma unroll
    for (int a = 0; a < S; ++a)
#pragma unroll
      for (int q = 0; q < kPairs; ++q) { arg[a][2 * q] = (int)(arg2[a][q] & 0xffffu); arg[a][2 * q + 1] = (int)(arg2[a][q] >> 16); }
  }
  const size_t W = (size_t)p.w * S;
  L *out = reinterpret_cast<L *>(p.labels) + ((size_t)n * p.h * S + (size_t)i * S) * W + (size_t)j * S;
#pragma unroll
  for (int a = 0; a < S; ++a) {
    if constexpr (sizeof(L) == 1 && S == 4) {
      const uint32_t w = (uint32_t)arg[a][0] | ((uint32_t)arg[a][1] << 8) | ((uint32_t)arg[a][2] << 16) | ((uint32_t)arg[a][3] << 24);
      *reinterpret_cast<uint32_t *>(out + (size_t)a * W) = w;
    } else {
#pragma unroll
      for (int b = 0; b < S; ++b) out[(size_t)a * W + b] = (L)arg[a][b];
    }
  }
}

template <typename T, typename L, int S>
__global__ void __launch_bounds__(128, 7) upsample_argmax_kernel(const ArgmaxParams p) {
  pdl_trigger();
  pdl_wait();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.total) return;
  const int j = (int)(t % (uint32_t)p.w), i = (int)((t / (uint32_t)p.w) % (uint32_t)p.h);
  const int n = (int)(t / ((uint32_t)p.w * p.h));
  argmax_pixel<T, L, S>(p, n, i, j);
}

// Block-sparse update of a label map that already holds the previous frame's labels: only the logit pixels of the
// EXECUTED cells, plus a ring of one logit pixel around each (bilinear taps reach one pixel into the neighbours), are
// recomputed from the dense logits; every other label is unchanged because every logit it depends on is.
template <typename T, typename L, int S>
__global__ void __launch_bounds__(128, 7) upsample_argmax_blocks_kernel(const ArgmaxParams p) {
  pdl_trigger();
  pdl_wait();
  const int cell = (int)(blockIdx.x / (uint32_t)p.chunks), chunk = (int)(blockIdx.x - (uint32_t)cell * p.chunks);
  if (!__ldg(p.grid + cell)) return;
  const int edge = p.BSl + 2, local = chunk * 128 + (int)threadIdx.x;
  if (local >= edge * edge) return;
  const int ly = local / edge, lx = local - ly * edge;
  const int per = p.GH * p.GW, n = cell / per, rem = cell - n * per, gh = rem / p.GW, gw = rem - gh * p.GW;
  const int i = gh * p.BSl + ly - 1, j = gw * p.BSl + lx - 1;
  if (i < 0 || j < 0 || i >= p.h || j >= p.w) return;
  argmax_pixel<T, L, S>(p, n, i, j);
}

template <typename T, typename L>
static void launch_argmax(const ArgmaxParams &p, int scale, cudaStream_t stream) {
  if (p.grid != nullptr) {
    const dim3 grid((unsigned)(p.N * p.GH * p.GW * p.chunks)), block(128);
    if (scale == 1) launch_kernel(upsample_argmax_blocks_kernel<T, L, 1>, grid, block, 0, stream, 1, p);
    else if (scale == 2) launch_kernel(upsample_argmax_blocks_kernel<T, L, 2>, grid, block, 0, stream, 1, p);
    else launch_kernel(upsample_argmax_blocks_kernel<T, L, 4>, grid, block, 0, stream, 1, p);
    return;
  }
  const dim3 grid((p.total + 127) / 128), block(128);
  if (scale == 1) launch_kernel(upsample_argmax_kernel<T, L, 1>, grid, block, 0, stream, 1, p);
  else if (scale == 2) launch_kernel(upsample_argmax_kernel<T, L, 2>, grid, block, 0, stream, 1, p);
  else launch_kernel(upsample_argmax_kernel<T, L, 4>, grid, block, 0, stream, 1, p);
}

int upsample_argmax(void *labels, const void *logits, int N, int K, int h, int w, const int64_t *strides, int scale,
                    int dtype, int label_bytes, cudaStream_t stream, const uint8_t *grid, int GH, int GW) {
  BC_REQUIRE(labels && logits && strides, BC_ERR_NULL, "bc_upsample_argmax: NULL pointer");
  BC_REQUIRE(N > 0 && K > 0 && h > 0 && w > 0, BC_ERR_SHAPE, "bc_upsample_argmax: empty problem");
  BC_REQUIRE(scale == 1 || scale == 2 || scale == 4, BC_ERR_UNSUPPORTED, "bc_upsample_argmax: scale %d (1, 2 or 4)", scale);
  BC_REQUIRE(dtype == BC_F16 || dtype == BC_F32, BC_ERR_DTYPE, "bc_upsample_argmax: dtype");
  BC_REQUIRE(label_bytes == 1 || label_bytes == 8, BC_ERR_DTYPE, "bc_upsample_argmax: labels are uint8 or int64");
  BC_REQUIRE(label_bytes == 8 || K <= 256, BC_ERR_RANGE, "bc_upsample_argmax: %d classes do not fit uint8 labels", K);
  BC_REQUIRE(((uintptr_t)labels & 7) == 0, BC_ERR_ALIGN, "bc_upsample_argmax: labels must be 8-byte aligned");
  const int64_t total = (int64_t)N * h * w;
  BC_REQUIRE(total * scale * scale < (1ll << 31), BC_ERR_RANGE, "bc_upsample_argmax: problem too large");
  BC_REQUIRE((h - 1) * strides[2] + (w - 1) * strides[3] < (1ll << 31) && strides[2] >= 0 && strides[3] >= 0, BC_ERR_RANGE,
             "bc_upsample_argmax: slice strides out of range");
  ArgmaxParams p;
  p.logits = logits; p.labels = labels;
  p.N = N; p.K = K; p.h = h; p.w = w;
  p.sn = strides[0]; p.sc = strides[1]; p.sh = strides[2]; p.sw = strides[3];
  p.total = (uint32_t)total;
  p.grid = grid;
  p.GH = GH; p.GW = GW; p.BSl = 0; p.chunks = 0;
  if (grid != nullptr) {
    BC_REQUIRE(GH > 0 && GW > 0 && h % GH == 0 && w % GW == 0 && h / GH == w / GW, BC_ERR_SHAPE,
               "bc_upsample_argmax_blocks: %dx%d logits on a %dx%d grid", h, w, GH, GW);
    p.BSl = h / GH;
    p.chunks = ((p.BSl + 2) * (p.BSl + 2) + 127) / 128;
  }
  if (dtype == BC_F16) {
    if (label_bytes == 1) launch_argmax<__half, uint8_t>(p, scale, stream);
    else launch_argmax<__half, long long>(p, scale, stream);
  } else {
    if (label_bytes == 1) launch_argmax<float, uint8_t>(p, scale, stream);
    else launch_argmax<float, long long>(p, scale, stream);
  }
  return check_launch(grid ? "bc_upsample_argmax_blocks" : "bc_upsample_argmax");
}

}  // namespace bc
