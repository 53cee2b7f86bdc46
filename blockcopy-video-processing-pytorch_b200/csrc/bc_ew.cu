// bc_ew.cu -- fused elementwise stage between two convolutions of the wrapped CNN, on packed NHWC
// fp16 tiles:   y = relu?( affine?( up2x?(a) + residual? ) )   written to the packed tile batch and,
// in the same pass, scattered into the next padded op's persistent plane.
//
// One launch replaces what the reference issues as separate torch kernels on the tile batch
// (core/tensorwrapper.py:519-520 pass-through ops and :577-598 bilinear): per-block bilinear x2
// (taps clamped at the BLOCK edge), `x += skip`, eval-mode BatchNorm (3 kernels in ATen), ReLU, plus
// this repo's scatter into the plane.  Every intermediate is rounded to fp16 exactly where the
// unfused op sequence rounds, so results match the op-by-op path.
#include <cuda_fp16.h>

#include "bc_common.cuh"

namespace bc {

struct EwParams {
  const __half *a;         // (E, BSa, BSa, C), BSa = BS/2 if up2x else BS
  const __half *residual;  // (E, BS, BS, C) or nullptr
  const float *mean, *invstd, *weight, *shift;  // [C] fp32 or nullptr (weight/shift may be null individually)
  __half *out;             // (E, BS, BS, C) or nullptr
  __half *plane;           // (N, H, W, C) or nullptr
  const int32_t *mapping;  // cell of tile b (needed when plane != nullptr)
  CellDecode cell;
  FastDiv chunks_per_pixel, bs_div, px_per_tile;
  int E, C, BS, H, W, up2x, relu, has_affine;
  uint32_t total;          // E * BS * BS * C / 8
};

__device__ __forceinline__ void load8(const __half *p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p));
  const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = __half22float2(h[t]);
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
}

__device__ __forceinline__ float round_h(float x) { return __half2float(__float2half_rn(x)); }

// Everything one unit (8 channels of one output pixel) needs from memory, loaded before any of it is used:
// the kernel is a pure stream, so what counts is bytes in flight per thread (two units = up to 160 B).
struct EwUnit {
  uint4 a00, a01, a10, a11, res;  // a00 only unless up2x
  float ly1, lx1;
  uint32_t pix, b, y, x;
  int c0;
};

__device__ __forceinline__ void ew_load(const EwParams &p, uint32_t i, EwUnit &u) {
  uint32_t ch, rem;
  p.chunks_per_pixel.divmod(i, u.pix, ch);
  p.px_per_tile.divmod(u.pix, u.b, rem);
  p.bs_div.divmod(rem, u.y, u.x);
  u.c0 = (int)ch * 8;
  if (p.up2x) {
    // PyTorch upsample_bilinear2d, align_corners = False, scale 0.5: src = 0.5 * (dst + 0.5) - 0.5, clamped at 0
    const int hs = p.BS >> 1;
    const float sy = fmaxf(0.5f * ((float)u.y + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.5f * ((float)u.x + 0.5f) - 0.5f, 0.f);
    const int y1 = (int)sy, x1 = (int)sx;
    const int yp = y1 < hs - 1 ? 1 : 0, xp = x1 < hs - 1 ? 1 : 0;
    u.ly1 = sy - (float)y1;
    u.lx1 = sx - (float)x1;
    const __half *base = p.a + (((size_t)u.b * hs + y1) * hs + x1) * p.C + u.c0;
    u.a00 = __ldg(reinterpret_cast<const uint4 *>(base));
    u.a01 = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)xp * p.C));
    u.a10 = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)yp * hs * p.C));
    u.a11 = __ldg(reinterpret_cast<const uint4 *>(base + ((size_t)yp * hs + xp) * p.C));
  } else {
    u.a00 = __ldg(reinterpret_cast<const uint4 *>(p.a + (size_t)u.pix * p.C + u.c0));
  }
  if (p.residual) u.res = __ldg(reinterpret_cast<const uint4 *>(p.residual + (size_t)u.pix * p.C + u.c0));
}

__device__ __forceinline__ void unpack8(const uint4 &u, float (&v)[8]) {
  const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = __half22float2(h[t]);
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
}

// per-channel batch-norm parameters of 8 consecutive channels
struct EwAffine {
  float mean[8], istd[8], w[8], sh[8];
};
__device__ __forceinline__ void ew_load_affine(const EwParams &p, int c0, EwAffine &f) {
  *reinterpret_cast<float4 *>(f.mean) = __ldg(reinterpret_cast<const float4 *>(p.mean + c0));
  *reinterpret_cast<float4 *>(f.mean + 4) = __ldg(reinterpret_cast<const float4 *>(p.mean + c0 + 4));
  *reinterpret_cast<float4 *>(f.istd) = __ldg(reinterpret_cast<const float4 *>(p.invstd + c0));
  *reinterpret_cast<float4 *>(f.istd + 4) = __ldg(reinterpret_cast<const float4 *>(p.invstd + c0 + 4));
  if (p.weight) {
    *reinterpret_cast<float4 *>(f.w) = __ldg(reinterpret_cast<const float4 *>(p.weight + c0));
    *reinterpret_cast<float4 *>(f.w + 4) = __ldg(reinterpret_cast<const float4 *>(p.weight + c0 + 4));
  } else {
#pragma unroll
    for (int t = 0; t < 8; ++t) f.w[t] = 1.f;
  }
  if (p.shift) {
    *reinterpret_cast<float4 *>(f.sh) = __ldg(reinterpret_cast<const float4 *>(p.shift + c0));
    *reinterpret_cast<float4 *>(f.sh + 4) = __ldg(reinterpret_cast<const float4 *>(p.shift + c0 + 4));
  } else {
#pragma unroll
    for (int t = 0; t < 8; ++t) f.sh[t] = 0.f;
  }
}

__device__ __forceinline__ void ew_finish(const EwParams &p, const EwUnit &u, const EwAffine &f) {
  float v[8];
  if (p.up2x) {
    float v00[8], v01[8], v10[8], v11[8];
    unpack8(u.a00, v00);
    unpack8(u.a01, v01);
    unpack8(u.a10, v10);
    unpack8(u.a11, v11);
    const float ly0 = 1.f - u.ly1, lx0 = 1.f - u.lx1;
#pragma unroll
    for (int t = 0; t < 8; ++t)
      v[t] = round_h(ly0 * (lx0 * v00[t] + u.lx1 * v01[t]) + u.ly1 * (lx0 * v10[t] + u.lx1 * v11[t]));
  } else {
    unpack8(u.a00, v);
  }
  if (p.residual) {
    float r[8];
    unpack8(u.res, r);
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = round_h(v[t] + r[t]);
  }
  if (p.has_affine) {
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = round_h(f.w[t] * (v[t] - f.mean[t]) * f.istd[t] + f.sh[t]);  // ATen's eval-BN expression
  }
  if (p.relu) {
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = fmaxf(v[t], 0.f);
  }
  uint4 o;
  __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
  for (int t = 0; t < 4; ++t) oh[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
  if (p.out) *reinterpret_cast<uint4 *>(p.out + (size_t)u.pix * p.C + u.c0) = o;
  if (p.plane) {
    uint32_t n, gh, gw;
    p.cell((uint32_t)__ldg(p.mapping + u.b), n, gh, gw);
    const size_t off = (((size_t)n * p.H + gh * p.BS + u.y) * p.W + gw * p.BS + u.x) * p.C + u.c0;
    *reinterpret_cast<uint4 *>(p.plane + off) = o;
  }
}

// FIXED_CH: the grid stride is a multiple of the chunks per pixel, so a thread always works on the same 8
// channels and loads their batch-norm parameters once.
template <bool FIXED_CH>
__global__ void __launch_bounds__(256) ew_fused_kernel(const EwParams p) {
  pdl_trigger();
  pdl_wait();
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  EwAffine f;
  if (FIXED_CH && p.has_affine && i < p.total) {
    uint32_t pix, ch;
    p.chunks_per_pixel.divmod(i, pix, ch);
    ew_load_affine(p, (int)ch * 8, f);
  }
  for (; i + stride < p.total; i += 2 * stride) {  // two units in flight
    EwUnit u0, u1;
    ew_load(p, i, u0);
    ew_load(p, i + stride, u1);
    if (!FIXED_CH && p.has_affine) ew_load_affine(p, u0.c0, f);
    ew_finish(p, u0, f);
    if (!FIXED_CH && p.has_affine) ew_load_affine(p, u1.c0, f);
    ew_finish(p, u1, f);
  }
  if (i < p.total) {
    EwUnit u0;
    ew_load(p, i, u0);
    if (!FIXED_CH && p.has_affine) ew_load_affine(p, u0.c0, f);
    ew_finish(p, u0, f);
  }
}

int ew_fused(void *out, void *plane, const void *a, const void *residual, const float *mean, const float *invstd,
             const float *weight, const float *shift, const int32_t *mapping, int E, int C, int BS, int N, int H,
             int W, int up2x, int relu, cudaStream_t stream) {
  BC_REQUIRE(a && (out || plane), BC_ERR_NULL, "bc_ew_fused: NULL pointer");
  BC_REQUIRE(E > 0 && C > 0 && BS > 0, BC_ERR_SHAPE, "bc_ew_fused: empty problem");
  BC_REQUIRE(C % 8 == 0, BC_ERR_UNSUPPORTED, "bc_ew_fused: C=%d is not a multiple of 8", C);
  BC_REQUIRE(!up2x || BS % 2 == 0, BC_ERR_SHAPE, "bc_ew_fused: up2x needs an even output block edge");
  BC_REQUIRE((mean == nullptr) == (invstd == nullptr), BC_ERR_NULL, "bc_ew_fused: mean and invstd come together");
  BC_REQUIRE((((uintptr_t)out | (uintptr_t)plane | (uintptr_t)a | (uintptr_t)residual) & 15) == 0, BC_ERR_ALIGN,
             "bc_ew_fused: pointers must be 16-byte aligned");
  if (plane) {
    BC_REQUIRE(mapping != nullptr, BC_ERR_NULL, "bc_ew_fused: plane output needs mapping_exec");
    BC_REQUIRE(N > 0 && H > 0 && W > 0 && H % BS == 0 && W % BS == 0, BC_ERR_SHAPE,
               "bc_ew_fused: plane %dx%d / block %d", H, W, BS);
  }
  EwParams p;
  p.a = (const __half *)a; p.residual = (const __half *)residual;
  p.mean = mean; p.invstd = invstd; p.weight = weight; p.shift = shift;
  p.out = (__half *)out; p.plane = (__half *)plane; p.mapping = mapping;
  p.cell = plane ? CellDecode(H / BS, W / BS) : CellDecode(1, 1);
  p.chunks_per_pixel = FastDiv((uint32_t)(C / 8));
  p.bs_div = FastDiv((uint32_t)BS);
  p.px_per_tile = FastDiv((uint32_t)(BS * BS));
  p.E = E; p.C = C; p.BS = BS; p.H = H; p.W = W; p.up2x = up2x; p.relu = relu; p.has_affine = mean != nullptr;
  const int64_t total = (int64_t)E * BS * BS * (C / 8);
  BC_REQUIRE(total < (1ll << 31), BC_ERR_RANGE, "bc_ew_fused: problem too large");
  p.total = (uint32_t)total;
  // two units per thread and iteration; at most 8 CTAs per SM (whole waves)
  int64_t grid = (total + 511) / 512;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (grid > cap) grid = cap;
  const bool fixed_ch = (grid * 256) % (C / 8) == 0;
  if (fixed_ch)
    launch_kernel(ew_fused_kernel<true>, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
  else
    launch_kernel(ew_fused_kernel<false>, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
  return check_launch("bc_ew_fused");
}

}  // namespace bc

// ---------------------------------------------------------------------------------------------------
// max-pool on the executed blocks, reading the op's persistent plane (halo = neighbouring cells,
// ZEROS outside the frame: the reference pads the tile with zeros -- not -inf -- and pools with
// padding 0, utils/blockpad.py:114-120 + core/tensorwrapper.py:565-571).  Replaces transfer + repad +
// at::max_pool2d + this repo's scatter of the result.
// ---------------------------------------------------------------------------------------------------
namespace bc {

struct PoolParams {
  const __half *plane;  // (N, H, W, C)
  __half *out;          // (E, BSo, BSo, C)
  __half *plane_out;    // (N, GH*BSo, GW*BSo, C) or nullptr
  const int32_t *mapping;
  CellDecode cell;
  FastDiv chunks_per_pixel, bs_div, px_per_tile;
  int C, H, W, BS_in, BSo, k, stride, pad;
  uint32_t total;
};

// K > 0: window size known at compile time: all K*K taps are loaded before the first max (a runtime-k loop keeps
// one load in flight at a time: nine serialised L2 round trips per output).  K == 0: any window, runtime loop.
template <int K>
__global__ void __launch_bounds__(256) maxpool_halo_kernel(const PoolParams p) {
  pdl_trigger();
  pdl_wait();
  const uint32_t gstride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += gstride) {
    uint32_t pix, ch, b, rem, oy, ox, n, gh, gw;
    p.chunks_per_pixel.divmod(i, pix, ch);
    p.px_per_tile.divmod(pix, b, rem);
    p.bs_div.divmod(rem, oy, ox);
    p.cell((uint32_t)__ldg(p.mapping + b), n, gh, gw);
    const int c0 = (int)ch * 8;
    const int y0 = (int)(gh * p.BS_in + oy * p.stride) - p.pad, x0 = (int)(gw * p.BS_in + ox * p.stride) - p.pad;
    const __half *base = p.plane + (size_t)n * p.H * p.W * p.C + c0;
    __half2 m[4];
    if (K > 0) {
      uint4 u[K > 0 ? K * K : 1];
#pragma unroll
      for (int dy = 0; dy < K; ++dy)
#pragma unroll
        for (int dx = 0; dx < K; ++dx) {
          const int yy = y0 + dy, xx = x0 + dx;
          u[dy * K + dx] = make_uint4(0, 0, 0, 0);  // zero padding outside the frame
          if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
            u[dy * K + dx] = __ldg(reinterpret_cast<const uint4 *>(base + ((size_t)yy * p.W + xx) * p.C));
        }
#pragma unroll
      for (int t = 0; t < 4; ++t) m[t] = reinterpret_cast<const __half2 *>(&u[0])[t];
#pragma unroll
      for (int j = 1; j < K * K; ++j)
#pragma unroll
        for (int t = 0; t < 4; ++t) m[t] = __hmax2_nan(m[t], reinterpret_cast<const __half2 *>(&u[j])[t]);
    } else {
      bool first = true;
      for (int dy = 0; dy < p.k; ++dy)
        for (int dx = 0; dx < p.k; ++dx) {
          const int yy = y0 + dy, xx = x0 + dx;
          uint4 u = make_uint4(0, 0, 0, 0);
          if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
            u = __ldg(reinterpret_cast<const uint4 *>(base + ((size_t)yy * p.W + xx) * p.C));
          const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
          for (int t = 0; t < 4; ++t) m[t] = first ? h[t] : __hmax2_nan(m[t], h[t]);
          first = false;
        }
    }
    const uint4 o = *reinterpret_cast<uint4 *>(m);
    *reinterpret_cast<uint4 *>(p.out + (size_t)pix * p.C + c0) = o;
    if (p.plane_out) {
      const int Wo = p.W / p.BS_in * p.BSo, Ho = p.H / p.BS_in * p.BSo;
      const size_t off = (((size_t)n * Ho + gh * p.BSo + oy) * Wo + gw * p.BSo + ox) * p.C + c0;
      *reinterpret_cast<uint4 *>(p.plane_out + off) = o;
    }
  }
}

int maxpool_halo(void *out, void *plane_out, const void *plane, const int32_t *mapping, int E, int N, int C, int H,
                 int W, int BS_in, int k, int stride, int pad, cudaStream_t stream) {
  BC_REQUIRE(out && plane && mapping, BC_ERR_NULL, "bc_maxpool_halo: NULL pointer");
  BC_REQUIRE(E > 0 && N > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0, BC_ERR_SHAPE, "bc_maxpool_halo: bad sizes");
  BC_REQUIRE(C % 8 == 0, BC_ERR_UNSUPPORTED, "bc_maxpool_halo: C=%d is not a multiple of 8", C);
  BC_REQUIRE(H % BS_in == 0 && W % BS_in == 0, BC_ERR_SHAPE, "bc_maxpool_halo: plane %dx%d / block %d", H, W, BS_in);
  BC_REQUIRE(BS_in % stride == 0 && (BS_in + 2 * pad - k) / stride + 1 == BS_in / stride, BC_ERR_UNSUPPORTED,
             "bc_maxpool_halo: kernel %d stride %d pad %d does not map a %d-px block onto a %d-px block", k, stride, pad,
             BS_in, BS_in / stride);
  BC_REQUIRE((((uintptr_t)out | (uintptr_t)plane | (uintptr_t)plane_out) & 15) == 0, BC_ERR_ALIGN,
             "bc_maxpool_halo: pointers must be 16-byte aligned");
  PoolParams p;
  p.plane = (const __half *)plane; p.out = (__half *)out; p.plane_out = (__half *)plane_out; p.mapping = mapping;
  p.cell = CellDecode(H / BS_in, W / BS_in);
  p.C = C; p.H = H; p.W = W; p.BS_in = BS_in; p.BSo = BS_in / stride; p.k = k; p.stride = stride; p.pad = pad;
  p.chunks_per_pixel = FastDiv((uint32_t)(C / 8));
  p.bs_div = FastDiv((uint32_t)p.BSo);
  p.px_per_tile = FastDiv((uint32_t)(p.BSo * p.BSo));
  const int64_t total = (int64_t)E * p.BSo * p.BSo * (C / 8);
  BC_REQUIRE(total < (1ll << 31), BC_ERR_RANGE, "bc_maxpool_halo: problem too large");
  p.total = (uint32_t)total;
  int64_t grid = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (grid > cap) grid = cap;
  if (k == 3)
    launch_kernel(maxpool_halo_kernel<3>, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
  else if (k == 2)
    launch_kernel(maxpool_halo_kernel<2>, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
  else
    launch_kernel(maxpool_halo_kernel<0>, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
  return check_launch("bc_maxpool_halo");
}

}  // namespace bc
