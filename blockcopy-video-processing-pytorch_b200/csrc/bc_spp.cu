// bc_spp.cu -- the dense middle of a pyramid-pooling module (SwiftNet's SpatialPyramidPooling,
// semantic_segmentation/lib/models/swiftnet/util.py:85-138), which the reference runs as ~25 tiny dense
// torch launches inside @blockcopy_noblocks.  Three kernels replace everything between its first and its last
// 1x1 convolution (which run on the tensor cores through bc_conv_igemm):
//   spp_pool   : all pyramid levels' average pools of x0 in one launch (fp32 sums)
//   spp_levels : per level BatchNorm + ReLU + 1x1 conv (bt -> level_size channels) on the pooled maps
//   spp_prep   : concat[x0, bilinear(level_i)] + BatchNorm + ReLU, channels padded to a multiple of 64,
//                i.e. the A operand of the last 1x1 conv
// Dense NHWC fp16 (N = 1 .. few), fp32 arithmetic, fp16 rounding where the op-by-op sequence rounds.
#include <cuda_fp16.h>

#include "bc_common.cuh"

namespace bc {

constexpr int kMaxLevels = 4;

struct SppGeo {
  int N, C, H, W;              // x0: (N, H, W, C)
  int L;                       // levels
  int gh[kMaxLevels], gw[kMaxLevels];
  int cell_off[kMaxLevels + 1];  // prefix sums of N*gh*gw
};

__device__ __forceinline__ float rh_(float x) { return __half2float(__float2half_rn(x)); }

// ---- pool: one CTA per (level, n, cell); 256 threads = (C/8 chunks) x (pixel lanes) ------------------------------
__global__ void __launch_bounds__(256) spp_pool_kernel(const __half *__restrict__ x0, __half *__restrict__ pooled,
                                                        const SppGeo g) {
  __shared__ float red[256 * 8];
  pdl_trigger();
  pdl_wait();
  int lvl = 0;
  while (lvl + 1 < g.L && (int)blockIdx.x >= g.cell_off[lvl + 1]) ++lvl;
  const int cell = (int)blockIdx.x - g.cell_off[lvl];
  const int gw = g.gw[lvl], gh = g.gh[lvl];
  const int n = cell / (gh * gw), cy = (cell / gw) % gh, cx = cell % gw;
  const int wh = g.H / gh, ww = g.W / gw;        // window
  const int chunks = g.C / 8;                    // <= 256
  const int lanes = 256 / chunks;                // pixel lanes per chunk (>= 1)
  const int chunk = threadIdx.x % chunks, lane = threadIdx.x / chunks;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes) {
    for (int p = lane; p < wh * ww; p += lanes) {
      const int y = cy * wh + p / ww, x = cx * ww + p % ww;
      const uint4 u = __ldg(reinterpret_cast<const uint4 *>(x0 + (((size_t)n * g.H + y) * g.W + x) * g.C + chunk * 8));
      const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(h[t]);
        acc[2 * t] += f.x;
        acc[2 * t + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) red[threadIdx.x * 8 + t] = acc[t];
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] += red[(l * chunks + chunk) * 8 + t];
    const float inv = 1.f / (float)(wh * ww);
    __align__(16) __half o[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) o[t] = __float2half_rn(acc[t] * inv);
    *reinterpret_cast<uint4 *>(pooled + (size_t)blockIdx.x * g.C + chunk * 8) = *reinterpret_cast<const uint4 *>(o);
  }
}

// ---- levels: one CTA per pooled pixel: BN + ReLU into smem, then one thread per output channel --------------------
struct LevelParams {
  const __half *pooled;   // [cells][C]
  __half *out;            // [cells][Lc]
  const float *bn;        // [L][4][C]: mean, invstd, weight, shift
  const __half *w;        // [L][Lc][C]
  SppGeo g;
  int Lc;
};

__global__ void __launch_bounds__(128) spp_levels_kernel(const LevelParams p) {
  extern __shared__ float act[];  // C
  pdl_trigger();
  pdl_wait();
  int lvl = 0;
  while (lvl + 1 < p.g.L && (int)blockIdx.x >= p.g.cell_off[lvl + 1]) ++lvl;
  const int C = p.g.C;
  const float *bn = p.bn + (size_t)lvl * 4 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = __half2float(p.pooled[(size_t)blockIdx.x * C + c]);
    const float y = rh_(bn[2 * C + c] * (v - bn[c]) * bn[C + c] + bn[3 * C + c]);
    act[c] = fmaxf(y, 0.f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < p.Lc; o += blockDim.x) {
    const __half *w = p.w + ((size_t)lvl * p.Lc + o) * C;
    float acc = 0.f;
    for (int c = 0; c < C; c += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4 *>(w + c));
      const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(h[t]);
        acc += act[c + 2 * t] * f.x + act[c + 2 * t + 1] * f.y;
      }
    }
    p.out[(size_t)blockIdx.x * p.Lc + o] = __float2half_rn(acc);
  }
}

// ---- prep: y = relu(bn(cat[x0, up(level_0), ...])) with channels padded to Cp ------------------------------------
struct PrepParams {
  const __half *x0;       // (N,H,W,C)
  const __half *lev;      // [cells][Lc]
  __half *y;              // (N,H,W,Cp)
  const float *bn;        // [4][Cp] (padding channels: weight = shift = 0)
  SppGeo g;
  int Lc, Cp;
  uint32_t total;         // N*H*W*Cp/8
};

__global__ void __launch_bounds__(256) spp_prep_kernel(const PrepParams p) {
  pdl_trigger();
  pdl_wait();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  const int chunks = p.Cp / 8;
  const int ch = (int)(i % (uint32_t)chunks);
  const uint32_t pix = i / (uint32_t)chunks;
  const int x = (int)(pix % (uint32_t)p.g.W), y = (int)((pix / (uint32_t)p.g.W) % (uint32_t)p.g.H);
  const int n = (int)(pix / ((uint32_t)p.g.W * p.g.H));
  const int C = p.g.C, Ctot = C + p.g.L * p.Lc;
  __align__(16) __half o[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int c = ch * 8 + t;
    float v = 0.f;
    if (c < C) {
      v = __half2float(p.x0[(size_t)pix * C + c]);
    } else if (c < Ctot) {
      const int lvl = (c - C) / p.Lc, j = (c - C) - lvl * p.Lc;
      const int gh = p.g.gh[lvl], gw = p.g.gw[lvl];
      // upsample_bilinear2d, align_corners = False: src = (dst + 0.5) * in/out - 0.5, clamped at 0
      const float sy = fmaxf(((float)y + 0.5f) * ((float)gh / (float)p.g.H) - 0.5f, 0.f);
      const float sx = fmaxf(((float)x + 0.5f) * ((float)gw / (float)p.g.W) - 0.5f, 0.f);
      const int y1 = (int)sy, x1 = (int)sx;
      const int yp = y1 < gh - 1 ? 1 : 0, xp = x1 < gw - 1 ? 1 : 0;
      const float ly1 = sy - (float)y1, ly0 = 1.f - ly1, lx1 = sx - (float)x1, lx0 = 1.f - lx1;
      const __half *b = p.lev + ((size_t)p.g.cell_off[lvl] + (size_t)n * gh * gw) * p.Lc + j;
      const float v00 = __half2float(b[((size_t)y1 * gw + x1) * p.Lc]), v01 = __half2float(b[((size_t)y1 * gw + x1 + xp) * p.Lc]);
      const float v10 = __half2float(b[((size_t)(y1 + yp) * gw + x1) * p.Lc]);
      const float v11 = __half2float(b[((size_t)(y1 + yp) * gw + x1 + xp) * p.Lc]);
      v = rh_(ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11));
    }
    float r = 0.f;
    if (c < Ctot) r = fmaxf(rh_(p.bn[2 * p.Cp + c] * (v - p.bn[c]) * p.bn[p.Cp + c] + p.bn[3 * p.Cp + c]), 0.f);
    o[t] = __float2half_rn(r);
  }
  *reinterpret_cast<uint4 *>(p.y + (size_t)pix * p.Cp + ch * 8) = *reinterpret_cast<const uint4 *>(o);
}

static int make_geo(SppGeo &g, int N, int C, int H, int W, int L, const int *gh, const int *gw) {
  BC_REQUIRE(N > 0 && C > 0 && C % 8 == 0 && C <= 2048 && H > 0 && W > 0, BC_ERR_SHAPE, "bc_spp: bad plane %dx%dx%dx%d", N, C, H, W);
  BC_REQUIRE(L > 0 && L <= kMaxLevels, BC_ERR_UNSUPPORTED, "bc_spp: %d levels (1..%d)", L, kMaxLevels);
  g.N = N; g.C = C; g.H = H; g.W = W; g.L = L;
  g.cell_off[0] = 0;
  for (int i = 0; i < L; ++i) {
    BC_REQUIRE(gh[i] > 0 && gw[i] > 0 && H % gh[i] == 0 && W % gw[i] == 0, BC_ERR_UNSUPPORTED,
               "bc_spp: grid %dx%d does not divide the %dx%d plane", gh[i], gw[i], H, W);
    g.gh[i] = gh[i]; g.gw[i] = gw[i];
    g.cell_off[i + 1] = g.cell_off[i] + N * gh[i] * gw[i];
  }
  return BC_OK;
}

int spp_pool(void *pooled, const void *x0, int N, int C, int H, int W, int L, const int *gh, const int *gw, cudaStream_t s) {
  BC_REQUIRE(pooled && x0 && gh && gw, BC_ERR_NULL, "bc_spp_pool: NULL pointer");
  SppGeo g;
  int rc = make_geo(g, N, C, H, W, L, gh, gw);
  if (rc != BC_OK) return rc;
  BC_REQUIRE(C / 8 <= 256, BC_ERR_UNSUPPORTED, "bc_spp_pool: C=%d too large", C);
  launch_kernel(spp_pool_kernel, dim3((unsigned)g.cell_off[L]), dim3(256), 0, s, 1, (const __half *)x0, (__half *)pooled, g);
  return check_launch("bc_spp_pool");
}

int spp_levels(void *out, const void *pooled, const float *bn, const void *w, int N, int C, int H, int W, int L,
               const int *gh, const int *gw, int Lc, cudaStream_t s) {
  BC_REQUIRE(out && pooled && bn && w, BC_ERR_NULL, "bc_spp_levels: NULL pointer");
  LevelParams p;
  int rc = make_geo(p.g, N, C, H, W, L, gh, gw);
  if (rc != BC_OK) return rc;
  BC_REQUIRE(Lc > 0, BC_ERR_SHAPE, "bc_spp_levels: level size %d", Lc);
  p.pooled = (const __half *)pooled; p.out = (__half *)out; p.bn = bn; p.w = (const __half *)w; p.Lc = Lc;
  launch_kernel(spp_levels_kernel, dim3((unsigned)p.g.cell_off[L]), dim3(128), (size_t)C * sizeof(float), s, 1, p);
  return check_launch("bc_spp_levels");
}

int spp_prep(void *y, const void *x0, const void *lev, const float *bn, int N, int C, int H, int W, int L, const int *gh,
             const int *gw, int Lc, int Cp, cudaStream_t s) {
  BC_REQUIRE(y && x0 && lev && bn, BC_ERR_NULL, "bc_spp_prep: NULL pointer");
  PrepParams p;
  int rc = make_geo(p.g, N, C, H, W, L, gh, gw);
  if (rc != BC_OK) return rc;
  BC_REQUIRE(Cp % 8 == 0 && Cp >= C + L * Lc, BC_ERR_SHAPE, "bc_spp_prep: padded channel count %d < %d", Cp, C + L * Lc);
  p.x0 = (const __half *)x0; p.lev = (const __half *)lev; p.y = (__half *)y; p.bn = bn; p.Lc = Lc; p.Cp = Cp;
  const int64_t total = (int64_t)N * H * W * (Cp / 8);
  BC_REQUIRE(total < (1ll << 31), BC_ERR_RANGE, "bc_spp_prep: problem too large");
  p.total = (uint32_t)total;
  launch_kernel(spp_prep_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, 1, p);
  return check_launch("bc_spp_prep");
}

}  // namespace bc
