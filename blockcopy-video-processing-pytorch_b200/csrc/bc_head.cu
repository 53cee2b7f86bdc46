// bc_head.cu -- the network's output head on executed blocks, fused with the final combine:
//     y = conv1x1( relu?( batch_norm?( x ) ) ) + bias      (few output channels: class logits)
//     dense_out = dense_prev with the executed cells replaced by y
// One streaming kernel instead of what the reference path issues at the end of every frame
// (SwiftNet `logits = _BNReluConv(128, 19, k=1)`, swiftnet/util.py + semseg.py, reached through
// core/tensorwrapper.py:519-520; then core/blockcopy.py:79-86 out.combine() = clone + combine_kernel):
// BN + ReLU on the tile batch, a 19-channel cuDNN conv, a bias-add kernel, the copy of the previous output and
// the scatter.  Rounding points are those of the op-by-op sequence: BN output -> fp16, conv accumulated in
// fp32 -> fp16, + bias -> fp16.
#include <cuda_fp16.h>

#include "bc_common.cuh"

namespace bc {

constexpr int kHeadMaxCout = 32;
constexpr int kHeadThreads = 256;
constexpr int kHeadWStride = kHeadMaxCout + 4;  // floats per weight row: consecutive rows start 4 banks apart

// Shared-memory row of input channel c.  The 8 lanes that share a pixel read, in the same instruction, the
// channels 64k + 8*sub + t (sub = 0..7): those get CONSECUTIVE rows, so their 16-byte reads of the weight rows
// (stride 36 floats) and of the batch-norm records (16 bytes each) fall into distinct banks.
__device__ __forceinline__ int head_row(int c) { return (((c >> 6) << 3) + (c & 7)) * 8 + ((c >> 3) & 7); }

struct HeadParams {
  const __half *x;         // (E, BS, BS, Cin) NHWC tiles
  const __half *weight;    // [Cout][Cin]
  const __half *bias;      // [Cout] or nullptr
  const float *mean, *invstd, *bn_w, *bn_b;  // [Cin] or nullptr (bn_w / bn_b individually optional)
  __half *tiles_out;       // (E, Cout, BS, BS) in tiles_layout, or nullptr
  __half *dense_out;       // (N, Cout, H, W) in dense_layout, or nullptr
  const __half *dense_prev;  // same shape / layout as dense_out, or nullptr
  const int32_t *grid_idx;   // (N, GH, GW): >= 0 = packed index of an executed cell
  const int32_t *mapping;    // cell of packed tile b
  CellDecode cell;
  FastDiv bs_div, px_per_tile, w_div, hw_div;
  int E, Cin, Cout, BS, H, W, GW, cells_per_image, relu_in, has_bn, tiles_nhwc, dense_nhwc;
  int staged;  // tiles NHWC (or none), dense none or (NCHW and 16 | BS): outputs leave through a staging tile
  int rows;  // shared-memory rows: Cin rounded up to a multiple of 64 (head_row permutes within 64)
  uint32_t exec_px, total_px;  // E*BS*BS, N*H*W
};

__device__ __forceinline__ float head_round(float v) { return __half2float(__float2half_rn(v)); }

__device__ __forceinline__ size_t dense_off(const HeadParams &p, uint32_t n, int c, uint32_t y, uint32_t x) {
  return p.dense_nhwc ? (((size_t)n * p.H + y) * p.W + x) * p.Cout + c : (((size_t)n * p.Cout + c) * p.H + y) * p.W + x;
}

// cells that were not executed keep the previous frame's values (non-in-place combine only)
__device__ __forceinline__ void head_copy_rest(const HeadParams &p) {
  if (!(p.dense_out && p.dense_prev)) return;
  const uint32_t stride = gridDim.x * kHeadThreads;
  for (uint32_t i = blockIdx.x * kHeadThreads + threadIdx.x; i < p.total_px; i += stride) {
    uint32_t n, rem, Y, X;
    p.hw_div.divmod(i, n, rem);
    p.w_div.divmod(rem, Y, X);
    const uint32_t cell = n * p.cells_per_image + p.bs_div.div(Y) * p.GW + p.bs_div.div(X);
    if (__ldg(p.grid_idx + cell) >= 0) continue;
    for (int o = 0; o < p.Cout; ++o) {
      const size_t off = dense_off(p, n, o, Y, X);
      p.dense_out[off] = p.dense_prev[off];
    }
  }
}

// LANES threads share one executed pixel (each takes every LANES-th group of 8 channels, so a warp reads whole
// 128-byte runs), CO4 = ceil(Cout / 4) groups of 4 accumulators per thread; the partial sums meet in a butterfly.
template <int LANES, int CO4>
__global__ void __launch_bounds__(kHeadThreads) head_1x1_kernel(const HeadParams p) {
  extern __shared__ float head_smem[];
  float *w_s = head_smem;                                  // [Cin][kHeadMaxCout] fp32, zero padded
  float4 *bn_s = reinterpret_cast<float4 *>(w_s + (size_t)p.rows * kHeadWStride);  // [rows] (mean, invstd, w, b)
  pdl_trigger();
  pdl_wait();
  // weights [Cout][Cin] fp16 -> [Cin][32] fp32 (zero padded): coalesced 16-byte reads of 8 input channels
  for (int i = threadIdx.x; i < kHeadMaxCout * (p.Cin >> 3); i += kHeadThreads) {
    const int o = i / (p.Cin >> 3), c8 = (i - o * (p.Cin >> 3)) * 8;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (o < p.Cout) u = __ldg(reinterpret_cast<const uint4 *>(p.weight + (size_t)o * p.Cin + c8));
    const __half *h = reinterpret_cast<const __half *>(&u);
#pragma unroll
    for (int t = 0; t < 8; ++t) w_s[(size_t)head_row(c8 + t) * kHeadWStride + o] = __half2float(h[t]);
  }
  if (p.has_bn)
    for (int c = threadIdx.x; c < p.Cin; c += kHeadThreads)
      bn_s[head_row(c)] = make_float4(__ldg(p.mean + c), __ldg(p.invstd + c), p.bn_w ? __ldg(p.bn_w + c) : 1.f,
                            p.bn_b ? __ldg(p.bn_b + c) : 0.f);
  __syncthreads();

  const uint32_t stride = gridDim.x * (kHeadThreads / LANES);
  const int sub = threadIdx.x % LANES;
  // ---- part 1: executed pixels (the loop bound is uniform over the LANES threads of a pixel)
  for (uint32_t i0 = blockIdx.x * (kHeadThreads / LANES) + (threadIdx.x & ~31u) / LANES; i0 < p.exec_px; i0 += stride) {
    // i0 is warp-uniform (the butterfly below needs the whole warp); surplus pixel slots of the last warp
    // recompute the last pixel and skip the stores
    const uint32_t i_raw = i0 + (threadIdx.x & 31u) / LANES;
    const bool valid = i_raw < p.exec_px;
    const uint32_t i = valid ? i_raw : p.exec_px - 1;
    uint32_t b, rem, y, x;
    p.px_per_tile.divmod(i, b, rem);
    p.bs_div.divmod(rem, y, x);
    float acc[CO4 * 4];
#pragma unroll
    for (int o = 0; o < CO4 * 4; ++o) acc[o] = 0.f;
    const uint4 *src = reinterpret_cast<const uint4 *>(p.x + (size_t)i * p.Cin);
#pragma unroll 2
    for (int c8 = sub * 8; c8 < p.Cin; c8 += LANES * 8) {
      const uint4 u = __ldg(src + (c8 >> 3));
      const __half2 *h = reinterpret_cast<const __half2 *>(&u);
      float a[8];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(h[t]);
        a[2 * t] = f.x;
        a[2 * t + 1] = f.y;
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        float v = a[t];
        if (p.has_bn) {
          const float4 q = bn_s[head_row(c8 + t)];
          v = head_round(q.z * (v - q.x) * q.y + q.w);  // ATen's eval-BN expression, fp16 result
        }
        if (p.relu_in) v = fmaxf(v, 0.f);
        const float4 *wr = reinterpret_cast<const float4 *>(w_s + (size_t)head_row(c8 + t) * kHeadWStride);
#pragma unroll
        for (int o4 = 0; o4 < CO4; ++o4) {
          const float4 w4 = wr[o4];
          acc[4 * o4] = fmaf(v, w4.x, acc[4 * o4]);
          acc[4 * o4 + 1] = fmaf(v, w4.y, acc[4 * o4 + 1]);
          acc[4 * o4 + 2] = fmaf(v, w4.z, acc[4 * o4 + 2]);
          acc[4 * o4 + 3] = fmaf(v, w4.w, acc[4 * o4 + 3]);
        }
      }
    }
#pragma unroll
    for (int d = 1; d < LANES; d <<= 1)
#pragma unroll
      for (int o = 0; o < CO4 * 4; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], d);
    uint32_t n = 0, gh = 0, gw = 0;
    if (p.dense_out) p.cell((uint32_t)__ldg(p.mapping + b), n, gh, gw);
    const uint32_t Y = gh * p.BS + y, X = gw * p.BS + x;
#pragma unroll
    for (int o = 0; o < CO4 * 4; ++o) {
      if (o >= p.Cout || (o % LANES) != sub || !valid) continue;  // the LANES threads share the stores
      float v = head_round(acc[o]);
      if (p.bias) v = head_round(v + __half2float(__ldg(p.bias + o)));
      const __half hv = __float2half_rn(v);
      if (p.tiles_out)
        p.tiles_out[p.tiles_nhwc ? (size_t)i * p.Cout + o : (((size_t)b * p.Cout + o) * p.BS + y) * p.BS + x] = hv;
      if (p.dense_out) p.dense_out[dense_off(p, n, o, Y, X)] = hv;
    }
  }
  head_copy_rest(p);
}

// ---------------------------------------------------------------------------------------------------
// Warp-level tensor-core form (mma.sync m16n8k16, fp16 x fp16 -> fp32): a warp owns 16 consecutive executed
// pixels.  A fragments come straight from global memory as 16-byte loads (thread (g, t) of the warp: pixel rows
// g and g + 8, channels 32j + 8t .. + 7 of every 32-channel chunk j), so batch-norm + ReLU are applied in
// registers before the MMA -- which is why this op is not a bc_conv_igemm launch (its operands go TMA -> shared
// memory -> tcgen05 untouched).  The K slots of the two MMAs of a chunk are a fixed permutation of the chunk's
// channels (slot 2t+{0,1} <-> channel 8t+{0,1} / 8t+{4,5}; slot 2t+8+{0,1} <-> 8t+{2,3} / 8t+{6,7}); the weight
// fragments use the same permutation.  The op is a stream (10.5 MB in, 3 MB out); the SIMT form above needed
// 23 M warp instructions for it, this one ~1.5 M.
constexpr int kHeadWPad = 32;  // halfs: weight rows start 64 bytes apart modulo 128 -> conflict-free LDS.128

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 8 fp16 channels -> batch-norm (fp16 result) -> ReLU, in place
__device__ __forceinline__ void head_prep8(const HeadParams &p, uint4 &u, const float4 *bn8 /* stride 4 */) {
  __half2 *h = reinterpret_cast<__half2 *>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float2 f = __half22float2(h[t]);
    if (p.has_bn) {
      const float4 q0 = bn8[(2 * t) * 4], q1 = bn8[(2 * t + 1) * 4];
      f.x = q0.z * (f.x - q0.x) * q0.y + q0.w;  // ATen's eval-BN expression; rounded to fp16 by the pack below
      f.y = q1.z * (f.y - q1.x) * q1.y + q1.w;
    }
    h[t] = __floats2half2_rn(f.x, f.y);
    if (p.relu_in) h[t] = __hmax2(h[t], __float2half2_rn(0.f));
  }
}

template <int NT>  // number of 8-channel output tiles: ceil(Cout / 8)
__global__ void __launch_bounds__(kHeadThreads) head_mma_kernel(const HeadParams p) {
  extern __shared__ float head_smem[];
  __half *w_s = reinterpret_cast<__half *>(head_smem);  // [NT*8][Cin + kHeadWPad] fp16, zero rows beyond Cout
  const int wrow = p.Cin + kHeadWPad;
  float4 *bn_s = reinterpret_cast<float4 *>(w_s + (size_t)NT * 8 * wrow);  // permuted, see head_prep8
  __half *stage_s = reinterpret_cast<__half *>(bn_s + p.Cin);                 // [warps][16][NT * 8]
  pdl_trigger();
  pdl_wait();
  for (int i = threadIdx.x; i < NT * 8 * (p.Cin >> 3); i += kHeadThreads) {
    const int o = i / (p.Cin >> 3), c8 = (i - o * (p.Cin >> 3)) * 8;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (o < p.Cout) u = __ldg(reinterpret_cast<const uint4 *>(p.weight + (size_t)o * p.Cin + c8));
    *reinterpret_cast<uint4 *>(w_s + (size_t)o * wrow + c8) = u;
  }
  if (p.has_bn)
    for (int c = threadIdx.x; c < p.Cin; c += kHeadThreads)  // channel 32j + 8t + u lives at 32j + 4u + t
      bn_s[(c & ~31) + ((c & 7) << 2) + ((c >> 3) & 3)] =
          make_float4(__ldg(p.mean + c), __ldg(p.invstd + c), p.bn_w ? __ldg(p.bn_w + c) : 1.f, p.bn_b ? __ldg(p.bn_b + c) : 0.f);
  __syncthreads();

  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const uint32_t groups = p.exec_px >> 4;
  const uint32_t wstride = gridDim.x * (kHeadThreads / 32);
  for (uint32_t grp = blockIdx.x * (kHeadThreads / 32) + (threadIdx.x >> 5); grp < groups; grp += wstride) {
    const uint32_t i_lo = (grp << 4) + g, i_hi = i_lo + 8;  // this thread's two pixel rows of the 16
    const uint4 *x_lo = reinterpret_cast<const uint4 *>(p.x + (size_t)i_lo * p.Cin) + t;
    const uint4 *x_hi = reinterpret_cast<const uint4 *>(p.x + (size_t)i_hi * p.Cin) + t;
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[n][k] = 0.f;
#pragma unroll 4  // all eight 16-byte loads of a 128-channel pixel pair in flight before the first use
    for (int c0 = 0; c0 < p.Cin; c0 += 32) {
      uint4 lo = __ldg(x_lo + (c0 >> 3)), hi = __ldg(x_hi + (c0 >> 3));
      const float4 *bn8 = bn_s + c0 + t;
      head_prep8(p, lo, bn8);
      head_prep8(p, hi, bn8);
      const uint32_t a0[4] = {lo.x, hi.x, lo.y, hi.y}, a1[4] = {lo.z, hi.z, lo.w, hi.w};
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const uint4 w = *reinterpret_cast<const uint4 *>(w_s + (size_t)(n * 8 + g) * wrow + c0 + 8 * t);
        mma_16816(acc[n], a0, w.x, w.y);
        mma_16816(acc[n], a1, w.z, w.w);
      }
    }
    // accumulator (row g / g+8, output channels n*8 + 2t, +1) -> fp16 -> + bias -> fp16
    if (p.staged) {
      // tiles NHWC + dense NCHW + 16 | BS: the 16 pixels are consecutive in a block row.  Through a per-warp
      // staging tile [16][Cout] so that the tile batch gets one contiguous run of 16*Cout halfs and every
      // channel plane of the dense output 32 contiguous bytes (scattered 2-byte stores cost ~10 us per frame)
      __half *st = stage_s + (size_t)(threadIdx.x >> 5) * (16 * NT * 8);
#pragma unroll
      for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int o = n * 8 + 2 * t + k;
            if (o >= p.Cout) continue;
            float v = head_round(acc[n][2 * half + k]);
            if (p.bias) v = head_round(v + __half2float(__ldg(p.bias + o)));
            st[(g + 8 * half) * p.Cout + o] = __float2half_rn(v);
          }
      __syncwarp();
      const uint32_t i0 = grp << 4;
      if (p.tiles_out) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(st);
        uint32_t *dst = reinterpret_cast<uint32_t *>(p.tiles_out + (size_t)i0 * p.Cout);
        for (int w = lane; w < 8 * p.Cout; w += 32) dst[w] = src[w];
      }
      if (p.dense_out) {
        uint32_t b, rem, y, x0, nn, gh, gw;
        p.px_per_tile.divmod(i0, b, rem);
        p.bs_div.divmod(rem, y, x0);
        p.cell((uint32_t)__ldg(p.mapping + b), nn, gh, gw);
        const uint32_t Y = gh * p.BS + y, X0 = gw * p.BS + x0;
        for (int task = lane; task < 4 * p.Cout; task += 32) {  // (channel, group of 4 pixels) -> one 8-byte store
          const int o = task >> 2, q = task & 3;
          __align__(8) __half v4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v4[u] = st[(4 * q + u) * p.Cout + o];
          *reinterpret_cast<uint2 *>(p.dense_out + dense_off(p, nn, o, Y, X0 + 4 * q)) = *reinterpret_cast<const uint2 *>(v4);
        }
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t i = half ? i_hi : i_lo;
        uint32_t b, rem, y, x;
        p.px_per_tile.divmod(i, b, rem);
        p.bs_div.divmod(rem, y, x);
        uint32_t nn = 0, gh = 0, gw = 0;
        if (p.dense_out) p.cell((uint32_t)__ldg(p.mapping + b), nn, gh, gw);
        const uint32_t Y = gh * p.BS + y, X = gw * p.BS + x;
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int o = n * 8 + 2 * t + k;
            if (o >= p.Cout) continue;
            float v = head_round(acc[n][2 * half + k]);
            if (p.bias) v = head_round(v + __half2float(__ldg(p.bias + o)));
            const __half hv = __float2half_rn(v);
            if (p.tiles_out)
              p.tiles_out[p.tiles_nhwc ? (size_t)i * p.Cout + o : (((size_t)b * p.Cout + o) * p.BS + y) * p.BS + x] = hv;
            if (p.dense_out) p.dense_out[dense_off(p, nn, o, Y, X)] = hv;
          }
      }
    }
  }
  head_copy_rest(p);
}

template <int NT>
static int launch_head_mma(const HeadParams &p, cudaStream_t stream) {
  const size_t smem = (size_t)NT * 8 * (p.Cin + kHeadWPad) * sizeof(__half) + (size_t)p.Cin * sizeof(float4) +
                      (size_t)(kHeadThreads / 32) * 16 * NT * 8 * sizeof(__half);
  static cudaError_t attr = cudaFuncSetAttribute(head_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  BC_REQUIRE(attr == cudaSuccess, (int)attr, "cudaFuncSetAttribute(head_mma_kernel): %s", cudaGetErrorString(attr));
  int64_t grid = ((int64_t)(p.exec_px >> 4) * 32 + kHeadThreads - 1) / kHeadThreads;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (grid > cap) grid = cap;
  launch_kernel(head_mma_kernel<NT>, dim3((unsigned)grid), dim3(kHeadThreads), smem, stream, 1, p);
  return check_launch("bc_head_1x1");
}

template <int LANES, int CO4>
static int launch_head(const HeadParams &p, size_t smem, cudaStream_t stream) {
  static cudaError_t attr = cudaFuncSetAttribute(head_1x1_kernel<LANES, CO4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  BC_REQUIRE(attr == cudaSuccess, (int)attr, "cudaFuncSetAttribute(head_1x1_kernel): %s", cudaGetErrorString(attr));
  int64_t grid = ((int64_t)p.exec_px * LANES + kHeadThreads - 1) / kHeadThreads;
  const int64_t cap = (int64_t)kNumSMs * 4;  // every CTA converts the weights once: few, long-lived CTAs
  if (grid > cap) grid = cap;
  launch_kernel(head_1x1_kernel<LANES, CO4>, dim3((unsigned)grid), dim3(kHeadThreads), smem, stream, 1, p);
  return check_launch("bc_head_1x1");
}

int head_1x1(void *tiles_out, void *dense_out, const void *dense_prev, const void *tiles_in, const void *weight,
             const void *bias, const float *bn_mean, const float *bn_invstd, const float *bn_weight,
             const float *bn_shift, int relu_in, const int32_t *grid_idx, const int32_t *mapping, int E, int N, int GH,
             int GW, int BS, int Cin, int Cout, int tiles_layout, int dense_layout, cudaStream_t stream) {
  BC_REQUIRE(tiles_in && weight && (tiles_out || dense_out), BC_ERR_NULL, "bc_head_1x1: NULL pointer");
  BC_REQUIRE(E > 0 && N > 0 && GH > 0 && GW > 0 && BS > 0, BC_ERR_SHAPE, "bc_head_1x1: empty problem");
  BC_REQUIRE(Cin % 8 == 0 && Cin <= 1024, BC_ERR_UNSUPPORTED, "bc_head_1x1: Cin=%d (multiple of 8, <= 1024)", Cin);
  const bool mma_ok = Cin % 32 == 0 && ((int64_t)E * BS * BS) % 16 == 0;
  // (A 128-output-channel instantiation for the BN -> ReLU -> 1x1 skip bottlenecks was tried and measured slower
  //  than bc_ew_fused + bc_conv_igemm: 15.7 / 10.6 / 20.3 us against ~12 us per pair; every CTA re-stages 16-64 KB
  //  of weights for 128 pixels.  Those convs stay on the tcgen05 kernel.)
  BC_REQUIRE(Cout >= 1 && Cout <= kHeadMaxCout, BC_ERR_UNSUPPORTED, "bc_head_1x1: Cout=%d (1..%d)", Cout, kHeadMaxCout);
  BC_REQUIRE((bn_mean == nullptr) == (bn_invstd == nullptr), BC_ERR_NULL, "bc_head_1x1: mean and invstd come together");
  BC_REQUIRE((((uintptr_t)tiles_in | (uintptr_t)weight) & 15) == 0, BC_ERR_ALIGN, "bc_head_1x1: tiles_in / weight must be 16-byte aligned");
  BC_REQUIRE((unsigned)tiles_layout <= 1u && (unsigned)dense_layout <= 1u, BC_ERR_DTYPE, "bc_head_1x1: layout enum");
  if (dense_out) BC_REQUIRE(grid_idx && mapping, BC_ERR_NULL, "bc_head_1x1: dense output needs grid_idx and mapping_exec");
  HeadParams p;
  p.x = (const __half *)tiles_in; p.weight = (const __half *)weight; p.bias = (const __half *)bias;
  p.mean = bn_mean; p.invstd = bn_invstd; p.bn_w = bn_weight; p.bn_b = bn_shift;
  p.tiles_out = (__half *)tiles_out; p.dense_out = (__half *)dense_out; p.dense_prev = (const __half *)dense_prev;
  p.grid_idx = grid_idx; p.mapping = mapping;
  p.cell = CellDecode(GH, GW);
  p.bs_div = FastDiv((uint32_t)BS);
  p.px_per_tile = FastDiv((uint32_t)(BS * BS));
  p.H = GH * BS; p.W = GW * BS; p.GW = GW; p.cells_per_image = GH * GW;
  p.w_div = FastDiv((uint32_t)p.W);
  p.hw_div = FastDiv((uint32_t)(p.H * p.W));
  p.E = E; p.Cin = Cin; p.Cout = Cout; p.BS = BS; p.relu_in = relu_in; p.has_bn = bn_mean != nullptr;
  p.tiles_nhwc = tiles_layout == BC_NHWC; p.dense_nhwc = dense_layout == BC_NHWC;
  const int64_t exec_px = (int64_t)E * BS * BS, total_px = (int64_t)N * p.H * p.W;
  BC_REQUIRE(exec_px < (1ll << 31) && total_px < (1ll << 31), BC_ERR_RANGE, "bc_head_1x1: problem too large");
  p.exec_px = (uint32_t)exec_px; p.total_px = (uint32_t)total_px;
  p.rows = (Cin + 63) / 64 * 64;
  p.staged = (!tiles_out || p.tiles_nhwc) && (!dense_out || (!p.dense_nhwc && BS % 16 == 0)) &&
             (((uintptr_t)tiles_out | (uintptr_t)dense_out) & 7) == 0;
  const size_t smem = (size_t)p.rows * kHeadWStride * sizeof(float) + (size_t)p.rows * sizeof(float4);
  if (mma_ok) {  // tensor-core form
    const int nt = (Cout + 7) / 8;
    if (nt <= 1) return launch_head_mma<1>(p, stream);
    if (nt <= 2) return launch_head_mma<2>(p, stream);
    if (nt <= 3) return launch_head_mma<3>(p, stream);
    return launch_head_mma<4>(p, stream);
  }
  return Cout <= 20 ? launch_head<1, 5>(p, smem, stream) : launch_head<1, 8>(p, smem, stream);
}

}  // namespace bc
