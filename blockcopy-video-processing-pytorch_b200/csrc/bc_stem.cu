// bc_stem.cu -- the ResNet stem (conv 7x7, stride 2, padding 3, 3 input channels) on executed blocks,
// on the tensor cores.
//
// A 7x7/s2 conv on 3 channels is a 4x4/s1 conv on the space-to-depth(2) image with 12 channels (kernel
// zero-extended to 8x8): tap (kh', kw') of output (oy, ox) reads s2d pixel (oy+kh'-2, ox+kw'-2).  With
// the 12 channels padded to 16 one s2d pixel is 32 bytes = one K=16 slab of tcgen05.mma.kind::f16, so
// the implicit GEMM is M = 128 output pixels, N = 64, K = 16 taps x 16.
//   stem_pack_kernel : executed NCHW input tiles (E,3,BS,BS) -> persistent s2d plane (N,H/2,W/2+4,16) NHWC
//                      (this plane is the op's temporal state: skipped cells keep older frames' pixels);
//                      every row carries BC_STEM_XPAD = 2 zero pixels on either side (zeroed once by the owner)
//   conv_stem_kernel : the 4 taps of one kernel row are 4 CONSECUTIVE pixels = 128 contiguous bytes, so the
//                      im2col row of output pixel ox is the 128-byte window starting at padded pixel ox.  The
//                      A tensor map describes exactly that: inner dim 64 elements, next dim = pixels with a
//                      stride of ONE pixel (32 B, windows overlap).  One 16 KB SWIZZLE_128B box per kernel
//                      row (128-byte TMA rows instead of 4x as many 32-byte ones) + one 8 KB weight box,
//                      4 MMAs of K = 16; vertical halo = neighbouring cells, OOB rows = zeros, horizontal
//                      frame border = the physical zero columns.
//                      epilogue = bias + ReLU -> NHWC tiles (+ scatter into the next op's plane)
// Replaces transfer + repad + cuDNN's 3-channel fprop + bias + ReLU (+ NCHW<->NHWC conversions) of the
// reference path (core/tensorwrapper.py:529-575 for backbone.conv1).
#include "bc_conv.cuh"
#include "bc_tma.cuh"

namespace bc {

// ------------------------------------------------------------------ pack: NCHW tiles -> s2d plane
constexpr int kStemXPad = BC_STEM_XPAD;  // zero pixels on either side of every s2d row

struct PackParams {
  const __half *tiles;  // (E, 3, BS, BS) NCHW
  __half *plane;        // (N, H/2, W/2 + 2*kStemXPad, 16) NHWC
  const int32_t *mapping;
  CellDecode cell;
  FastDiv half_bs, px_per_tile;
  int BS, Hh, Wh;       // BS of the input tiles; s2d plane extent
  uint32_t total;       // E * (BS/2)^2
};

__global__ void __launch_bounds__(256) stem_pack_kernel(const PackParams p) {
  pdl_trigger();
  pdl_wait();
  const uint32_t gstride = gridDim.x * blockDim.x;
  const int hb = p.BS >> 1;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += gstride) {
    uint32_t b, rem, Y, X, n, gh, gw;
    p.px_per_tile.divmod(i, b, rem);
    p.half_bs.divmod(rem, Y, X);
    p.cell((uint32_t)__ldg(p.mapping + b), n, gh, gw);
    __align__(16) __half v[16];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const __half2 pr = __ldg(reinterpret_cast<const __half2 *>(
            p.tiles + (((size_t)b * 3 + c) * p.BS + 2 * Y + dy) * p.BS + 2 * X));
        v[(dy * 2 + 0) * 3 + c] = __low2half(pr);
        v[(dy * 2 + 1) * 3 + c] = __high2half(pr);
      }
#pragma unroll
    for (int t = 12; t < 16; ++t) v[t] = __float2half(0.f);
    __half *dst = p.plane + (((size_t)n * p.Hh + gh * hb + Y) * (p.Wh + 2 * kStemXPad) + kStemXPad + gw * hb + X) * 16;
    reinterpret_cast<uint4 *>(dst)[0] = reinterpret_cast<const uint4 *>(v)[0];
    reinterpret_cast<uint4 *>(dst)[1] = reinterpret_cast<const uint4 *>(v)[1];
  }
}

int stem_pack(void *plane, const void *tiles, const int32_t *mapping, int E, int N, int H, int W, int BS,
              cudaStream_t stream) {
  BC_REQUIRE(plane && tiles && mapping, BC_ERR_NULL, "bc_stem_pack: NULL pointer");
  BC_REQUIRE(E > 0 && BS % 2 == 0 && H % BS == 0 && W % BS == 0, BC_ERR_SHAPE, "bc_stem_pack: %dx%d / block %d", H, W, BS);
  BC_REQUIRE((((uintptr_t)plane & 15) | ((uintptr_t)tiles & 3)) == 0, BC_ERR_ALIGN, "bc_stem_pack: alignment");
  PackParams p;
  p.tiles = (const __half *)tiles; p.plane = (__half *)plane; p.mapping = mapping;
  p.cell = CellDecode(H / BS, W / BS);
  p.half_bs = FastDiv((uint32_t)(BS / 2));
  p.px_per_tile = FastDiv((uint32_t)((BS / 2) * (BS / 2)));
  p.BS = BS; p.Hh = H / 2; p.Wh = W / 2;
  const int64_t total = (int64_t)E * (BS / 2) * (BS / 2);
  BC_REQUIRE(total < (1ll << 31), BC_ERR_RANGE, "bc_stem_pack: problem too large");
  p.total = (uint32_t)total;
  int64_t grid = (total + 255) / 256;
  if (grid > (int64_t)kNumSMs * 8) grid = (int64_t)kNumSMs * 8;
  launch_kernel(stem_pack_kernel, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
  return check_launch("bc_stem_pack");
}

// ------------------------------------------------------------------ conv on the s2d plane
// One CTA = kStemTiles vertically adjacent 128-pixel tiles of one block and 64 output channels:
//   * ONE A box of (kStemTiles*rows + 3) pixel rows x BS_out windows (the tiles and the 4 kernel rows share
//     pixel rows: tap kh of tile t is the SAME shared memory shifted by (t*rows + kh) * BS_out windows),
//   * the whole 64 x 256 weight slab once (4 boxes, one per kernel row),
//   * kStemTiles accumulators of 64 TMEM columns, 16 MMAs (K = 16) each,
//   * epilogue through shared memory so that global stores are whole 128-byte pixel rows.
// L1<->L2 traffic per output pixel falls from ~1.3 KB (per-tile weight reload, 4x4 window reuse from L2,
// half-filled store sectors) to ~0.6 KB, which is what bounds this kernel (profiles/r01c_stem.md).
constexpr int kStemN = 64;
constexpr int kStemTiles = 2;
constexpr uint32_t kStemBBox = kStemN * 128;          // one kernel row of the weights: 64 x (4 taps x 16 ch) fp16
constexpr uint32_t kStemBBytes = 4 * kStemBBox;       // 32 KB
constexpr int kStemRowB = kStemN * 2 + 16;            // staged output row (+16 B: bank spread)

__host__ __device__ constexpr uint32_t stem_a_bytes(int BS_out) {
  return (uint32_t)(kStemTiles * kTileM + 3 * BS_out) * 128u;
}

__global__ void __launch_bounds__(kConvThreads)
conv_stem_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap b_map,
                 const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar;
  __shared__ __align__(8) uint64_t acc_bar[kStemTiles];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float bias_s[kStemN];
  __shared__ long long row_off_s[kTileM], row_pl_s[kTileM];  // tile 0's pixel offsets (see bc_conv.cu)
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x, n0 = blockIdx.y * kStemN;
  const int groups_per_block = p.tiles_per_block / kStemTiles;
  const int b0 = group / groups_per_block;
  const int r0 = (group - b0 * groups_per_block) * kStemTiles * p.rows_per_tile;
  const uint32_t a_bytes = stem_a_bytes(p.BS_out);
  if (threadIdx.x == 0) { trace_wall(p, 8); trace_mark(p, 0); }

  if (warp == 0 && lane == 0) {
    prefetch_map(&a_map);
    prefetch_map(&b_map);
    mbar_init(&full_bar, 1);
    for (int t = 0; t < kStemTiles; ++t) mbar_init(&acc_bar[t], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, kStemTiles * kStemN);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_trigger();
  pdl_wait();  // everything above overlapped the previous kernel's tail
  if (threadIdx.x == 0) trace_mark(p, 1);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t n, gh, gw;
      p.cell((uint32_t)__ldg(p.mapping + b0), n, gh, gw);
      // x in padded pixels (window x = taps of output column x); rows r0-2 .. r0 + tiles*rows: OOB rows = zeros
      mbar_expect_tx(&full_bar, a_bytes + kStemBBytes);
      tma_load_4d(smem, &a_map, &full_bar, 0, (int)gw * p.BS_out, (int)gh * p.BS_out + r0 - 2, (int)n);
      for (int kh = 0; kh < 4; ++kh) tma_load_2d(smem + a_bytes + kh * kStemBBox, &b_map, &full_bar, kh * 64, n0);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kTileM, kStemN);
      mbar_wait(&full_bar, 0);
      trace_mark(p, 2);
      tc_fence_after_sync();
      const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + a_bytes;
      for (int t = 0; t < kStemTiles; ++t) {
        for (int kh = 0; kh < 4; ++kh) {
          const uint32_t a_tap = a_addr + (uint32_t)((t * p.rows_per_tile + kh) * p.BS_out) * 128u;
#pragma unroll
          for (int kw = 0; kw < 4; ++kw)
            umma_f16_ss(tmem_base + (uint32_t)(t * kStemN), umma_desc_sw128(a_tap + kw * 32),
                        umma_desc_sw128(b_addr + kh * kStemBBox + kw * 32), idesc, (uint32_t)((kh | kw) != 0));
        }
        umma_commit(&acc_bar[t]);
      }
      trace_mark(p, 3);
    }
  } else {
    const int q = warp & 3;
    if (threadIdx.x - 64 < kStemN) bias_s[threadIdx.x - 64] = p.bias ? __half2float(p.bias[n0 + threadIdx.x - 64]) : 0.f;
    {
      const int m = q * 32 + lane;
      int blk, y, x;
      pixel_of_row(p, m, r0, blk, y, x);
      row_off_s[m] = (long long)((((size_t)b0 * p.BS_out + y) * p.BS_out + x) * p.Cout + n0);
      row_pl_s[m] = p.plane_out ? (long long)(plane_row(p, b0, y, x) - p.plane_out) + n0 : -1ll;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // the 4 epilogue warps only
    // staging rows live in the A region: free once the LAST accumulator is complete (all MMAs have read smem)
    uint8_t *stage = smem + (size_t)q * 32 * kStemRowB;
    constexpr int kTPR = kStemN / 8, kRPI = 32 / kTPR;  // 8 lanes cover one pixel's 128 bytes, 4 pixels per access
    const int c8 = (lane % kTPR) * 8;
    mbar_wait(&acc_bar[kStemTiles - 1], 0);
    if (threadIdx.x == 64) trace_mark(p, 4);
    tc_fence_after_sync();
#pragma unroll 1
    for (int t = 0; t < kStemTiles; ++t) {
      // phase A: this thread's accumulator row -> + bias -> ReLU -> fp16 -> staging row `lane`
#pragma unroll 1
      for (int c0 = 0; c0 < kStemN; c0 += 32) {
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * kStemN + c0), acc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            v[u] = __uint_as_float(acc[j + u]) + bias_s[c0 + j + u];
            if (p.relu) v[u] = fmaxf(v[u], 0.f);
          }
          uint4 o;
          __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
          for (int u = 0; u < 4; ++u) oh[u] = __floats2half2_rn(v[2 * u], v[2 * u + 1]);
          *reinterpret_cast<uint4 *>(stage + (size_t)lane * kStemRowB + (c0 + j) * 2) = o;
        }
      }
      __syncwarp();
      // phase B: whole 128-byte pixel rows to the tile batch and to the next op's plane
      // tile t lies t * rows_per_tile pixel rows below tile 0
      const long long dt_out = (long long)t * kTileM * p.Cout;
      const long long dt_pl = (long long)t * p.rows_per_tile * p.out_W * p.Cout;
#pragma unroll
      for (int i0 = 0; i0 < 32; i0 += kRPI) {
        const int r = i0 + lane / kTPR;
        const long long ro = row_off_s[q * 32 + r], rp = row_pl_s[q * 32 + r];
        const uint4 o = *reinterpret_cast<const uint4 *>(stage + (size_t)r * kStemRowB + c8 * 2);
        if (p.out) *reinterpret_cast<uint4 *>(p.out + ro + dt_out + c8) = o;
        if (rp >= 0) *reinterpret_cast<uint4 *>(p.plane_out + rp + dt_pl + c8) = o;
      }
      __syncwarp();
    }
    if (threadIdx.x == 64) trace_mark(p, 5);
    tc_fence_before_sync();
  }
  __syncthreads();
  if (threadIdx.x == 0) { trace_mark(p, 6); trace_wall(p, 10); }
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kStemTiles * kStemN);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder();  // bc_tma.cu

// s2d_plane (N, Hs, Ws + 2*kStemXPad, 16) fp16 NHWC, pad columns zero; weight fp16 [Cout][4][4][16] (see blockcopy/_C.py: pack_stem_weight)
int conv_stem(void *out, const void *s2d_plane, const void *weight, const void *bias, const int32_t *mapping, int E,
              int N, int Hs, int Ws, int BS_out, int Cout, int relu, void *plane_out, cudaStream_t stream) {
  BC_REQUIRE((out || plane_out) && s2d_plane && weight && mapping, BC_ERR_NULL, "bc_conv_stem: NULL pointer");
  BC_REQUIRE(E > 0 && N > 0, BC_ERR_SHAPE, "bc_conv_stem: empty problem");
  BC_REQUIRE(Cout % kStemN == 0, BC_ERR_UNSUPPORTED, "bc_conv_stem: Cout=%d is not a multiple of 64", Cout);
  const int px = BS_out * BS_out;
  BC_REQUIRE(BS_out >= 16 && BS_out <= 128 && (BS_out & (BS_out - 1)) == 0 && px % kTileM == 0, BC_ERR_UNSUPPORTED,
             "bc_conv_stem: output block edge %d (power of two, 16..128)", BS_out);
  BC_REQUIRE(Hs % BS_out == 0 && Ws % BS_out == 0, BC_ERR_SHAPE, "bc_conv_stem: plane %dx%d / block %d", Hs, Ws, BS_out);
  BC_REQUIRE((((uintptr_t)out | (uintptr_t)s2d_plane | (uintptr_t)weight | (uintptr_t)bias | (uintptr_t)plane_out) & 15) == 0,
             BC_ERR_ALIGN, "bc_conv_stem: pointers must be 16-byte aligned");
  EncodeTiledFn enc = tensor_map_encoder();
  BC_REQUIRE(enc != nullptr, BC_ERR_NO_DEVICE, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");

  ConvParams p = {};
  p.mapping = mapping;
  p.cell = CellDecode(Hs / BS_out, Ws / BS_out);
  p.bias = (const __half *)bias;
  p.out = (__half *)out;
  p.E = E; p.BS_out = BS_out; p.BS_in = BS_out; p.stride = 1; p.pad = 2; p.ksize = 4; p.Cout = Cout;
  p.blocks_per_tile = 1;
  p.tiles_per_block = px / kTileM;
  p.rows_per_tile = kTileM / BS_out;
  p.relu = relu;
  p.plane_out = (__half *)plane_out;
  p.out_mapping = mapping;
  p.out_cell = p.cell;
  p.out_H = Hs; p.out_W = Ws;
  p.splits = 1;
  p.trace = debug_trace_buffer();

  CUtensorMap a_map, b_map;
  {
    // dim 1 walks WINDOWS of 4 pixels with a stride of one pixel: window x = padded pixels x .. x+3
    const cuuint64_t Wp = (cuuint64_t)Ws + 2 * kStemXPad;
    cuuint64_t gdim[4] = {64, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)N};
    cuuint64_t gstr[3] = {32, Wp * 32, (cuuint64_t)Hs * Wp * 32};
    cuuint32_t box[4] = {64, (cuuint32_t)BS_out, (cuuint32_t)(kStemTiles * p.rows_per_tile + 3), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(&a_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(s2d_plane), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BC_REQUIRE(r == CUDA_SUCCESS, BC_ERR_UNSUPPORTED, "bc_conv_stem: tensor map (plane) failed: CUresult %d", (int)r);
  }
  {
    cuuint64_t gdim[2] = {256, (cuuint64_t)Cout};
    cuuint64_t gstr[1] = {512};
    cuuint32_t box[2] = {64, (cuuint32_t)kStemN};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&b_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(weight), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BC_REQUIRE(r == CUDA_SUCCESS, BC_ERR_UNSUPPORTED, "bc_conv_stem: tensor map (weights) failed: CUresult %d", (int)r);
  }
  const size_t smem = (size_t)stem_a_bytes(BS_out) + kStemBBytes + 1024;
  static cudaError_t attr = cudaFuncSetAttribute(conv_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)(stem_a_bytes(128) + kStemBBytes + 1024));
  BC_REQUIRE(attr == cudaSuccess, (int)attr, "cudaFuncSetAttribute(conv_stem_kernel): %s", cudaGetErrorString(attr));
  launch_kernel(conv_stem_kernel, dim3((unsigned)(E * p.tiles_per_block / kStemTiles), (unsigned)(Cout / kStemN)),
                dim3(kConvThreads), smem, stream, 1, a_map, b_map, p);
  return check_launch("bc_conv_stem");
}

}  // namespace bc
