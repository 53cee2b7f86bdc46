// bc_train.cu -- backward pass of the policy CNN on sm_100a (backward half).
//
// The reference trains its policy net online: every block_train_interval frames PolicyTrainRL.optim
// (policy/policy.py:319-370) calls loss.backward() through PolicyNet (policy/net.py:78-125, policy/resnet.py:60-115:
// ResNet-8 x2 + three stride-2 convs, BatchNorm in TRAIN mode) -- ~90 cuDNN / ATen launches in fp32.  Here every unit
// conv -> BN(batch statistics) -> [+ shortcut] -> [ReLU] is differentiated by four kernels over the fp16 NHWC planes the
// fused forward (policy/fused_net.py) keeps:
//   bc_bwd_mask_add      g = (out > 0 ? dOut : 0) [+ other]                  ReLU backward / gradient joins
//   bc_bn_bwd_reduce     sum_p g, sum_p g * xhat  per channel                (xhat = (z - mean) * invstd), reproducible
//   bc_bn_bwd_apply      dz = gamma * invstd * (g - sum_g / P - xhat * sum_gx / P), optionally also written with a
//                        zero between neighbours (the input of a stride-2 conv's data gradient)
//   bc_conv_wgrad        dW[co,ci,kh,kw] = sum_p dz[p,co] * x[s*p + (kh,kw) - pad, ci]: warp-level tensor-core MMA
//                        (mma.sync m16n8k16, fp32 accumulate) with BOTH operands pixel-major in shared memory
//                        (ldmatrix.trans), per-CTA partial sums added in CTA order, scaled by 1/loss-scale and written in
//                        the parameter's own fp32 layout -- together with the BatchNorm's d(gamma), d(beta)
// The data gradient of a conv is a conv: bc_conv_igemm (tcgen05) with the flipped / transposed weights bc_pack_params
// writes from the live parameters (a stride-2 conv's over the zero-interleaved dz).
#include <cuda_fp16.h>

#include "bc_common.cuh"
#include "bc_ptx.cuh"

namespace bc {

// =====================================================================================================
// bc_bwd_mask_add
// =====================================================================================================
struct MaskAddParams {
  const uint4 *grad, *out, *add;
  uint4 *dst;
  uint32_t total;  // 16-byte vectors
};

__device__ __forceinline__ uint4 mask_by_positive(uint4 g, const uint4 o) {
  const __half2 zero = __float2half2_rn(0.f);
  uint32_t *gw = reinterpret_cast<uint32_t *>(&g);
  const __half2 *oh = reinterpret_cast<const __half2 *>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) gw[k] &= __hgt2_mask(oh[k], zero);  // 0xffff per half where out > 0
  return g;
}

__global__ void __launch_bounds__(256) bwd_mask_add_kernel(const MaskAddParams p) {
  pdl_trigger();
  pdl_wait();
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += stride) {
    uint4 g = __ldg(p.grad + i);
    if (p.out) g = mask_by_positive(g, __ldg(p.out + i));
    if (p.add) {
      const uint4 a = __ldg(p.add + i);
      __half2 *gh = reinterpret_cast<__half2 *>(&g);
      const __half2 *ah = reinterpret_cast<const __half2 *>(&a);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 x = __half22float2(gh[k]), y = __half22float2(ah[k]);
        gh[k] = __floats2half2_rn(x.x + y.x, x.y + y.y);
      }
    }
    p.dst[i] = g;
  }
}

int bwd_mask_add(void *dst, const void *grad, const void *out, const void *add, long long n, cudaStream_t stream) {
  BC_REQUIRE(dst && grad, BC_ERR_NULL, "bc_bwd_mask_add: NULL pointer");
  BC_REQUIRE(n > 0 && n % 8 == 0 && n / 8 < (1ll << 31), BC_ERR_SHAPE, "bc_bwd_mask_add: %lld elements (multiple of 8)", n);
  BC_REQUIRE((((uintptr_t)dst | (uintptr_t)grad | (uintptr_t)out | (uintptr_t)add) & 15) == 0, BC_ERR_ALIGN,
             "bc_bwd_mask_add: pointers must be 16-byte aligned");
  MaskAddParams p;
  p.grad = (const uint4 *)grad; p.out = (const uint4 *)out; p.add = (const uint4 *)add; p.dst = (uint4 *)dst;
  p.total = (uint32_t)(n / 8);
  long long grid = (p.total + 255) / 256;
  if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
  launch_kernel(bwd_mask_add_kernel, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
  return check_launch("bc_bwd_mask_add");
}

// =====================================================================================================
// bc_bn_bwd_reduce: sums[0][c] = sum_p g[p,c], sums[1][c] = sum_p g[p,c] * xhat[p,c]
// Same scheme as bn_stats_kernel (bc_policy.cu): fp32 per thread, fixed-order shared-memory reduction per CTA, the last
// CTA (atomic ticket) adds the per-CTA partials in CTA order in double precision.
// =====================================================================================================
constexpr int kBwdThreads = 512;

struct BnBwdReduceParams {
  const __half *g, *out, *z;   // (P, C) fp16; out may be NULL (no ReLU mask)
  const float *mean, *invstd;  // [C]
  float *sums;                 // [2][C]
  float *partial;              // [gridDim.x][2][C]
  unsigned int *ticket;
  uint32_t P;
  int C;
  uint32_t zero;
};

__global__ void __launch_bounds__(kBwdThreads) bn_bwd_reduce_kernel(const BnBwdReduceParams p) {
  __shared__ float red[kBwdThreads][17];
  __shared__ double comb[kBwdThreads];
  __shared__ bool last;
  pdl_trigger();
  pdl_wait();
  const int lanes = p.C >> 3;
  const int sub = threadIdx.x % lanes, row = threadIdx.x / lanes, rows = kBwdThreads / lanes;
  float mean[8], istd[8], s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mean[k] = __ldg(p.mean + sub * 8 + k);
    istd[k] = __ldg(p.invstd + sub * 8 + k);
    s[k] = q[k] = 0.f;
  }
  const uint32_t step = gridDim.x * rows;
  const size_t c0 = (size_t)sub * 8;
  for (uint32_t px = blockIdx.x * rows + row; px < p.P; px += 4 * step) {
    uint4 ug[4], uz[4], uo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t pj = px + j * step;
      const size_t off = (size_t)(pj < p.P ? pj : p.P - 1) * p.C + c0;
      ug[j] = __ldg(reinterpret_cast<const uint4 *>(p.g + off));
      uz[j] = __ldg(reinterpret_cast<const uint4 *>(p.z + off));
      uo[j] = p.out ? __ldg(reinterpret_cast<const uint4 *>(p.out + off)) : make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    }
    uint32_t fold = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) fold ^= ug[j].x ^ uz[j].x ^ uo[j].x;
    fold &= p.zero;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ug[j].x ^= fold; uz[j].x ^= fold; uo[j].x ^= fold;
      if (px + j * step >= p.P) ug[j] = make_uint4(0, 0, 0, 0);
      const uint4 gm = mask_by_positive(ug[j], uo[j]);
      const __half2 *gh = reinterpret_cast<const __half2 *>(&gm);
      const __half2 *zh = reinterpret_cast<const __half2 *>(&uz[j]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 g = __half22float2(gh[k]), z = __half22float2(zh[k]);
        s[2 * k] += g.x; s[2 * k + 1] += g.y;
        q[2 * k] = fmaf(g.x, (z.x - mean[2 * k]) * istd[2 * k], q[2 * k]);
        q[2 * k + 1] = fmaf(g.y, (z.y - mean[2 * k + 1]) * istd[2 * k + 1], q[2 * k + 1]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { red[threadIdx.x][k] = s[k]; red[threadIdx.x][8 + k] = q[k]; }
  __syncthreads();
  const int items = 2 * p.C;  // (sum_g | sum_gx, channel); kBwdThreads % items == 0
  const int slices = kBwdThreads / items, item = threadIdx.x % items, slice = threadIdx.x / items;
  {
    const int which = item / p.C, c = item - which * p.C;
    float acc = 0.f;
    for (int r = slice; r < rows; r += slices) acc += red[r * lanes + (c >> 3)][which * 8 + (c & 7)];
    comb[threadIdx.x] = (double)acc;
    __syncthreads();
    if (threadIdx.x < items) {
      double t = 0.0;
      for (int zz = 0; zz < slices; ++zz) t += comb[zz * items + threadIdx.x];
      p.partial[(size_t)blockIdx.x * items + threadIdx.x] = (float)t;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  {
    double t = 0.0;
    for (unsigned b0 = (unsigned)slice; b0 < gridDim.x; b0 += 8 * slices) {  // eight partial loads in flight
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const unsigned b = b0 + j * slices;
        v[j] = b < gridDim.x ? __ldcg(p.partial + (size_t)b * items + item) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) t += (double)v[j];
    }
    comb[threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.x < items) {
      double sum = 0.0;
      for (int zz = 0; zz < slices; ++zz) sum += comb[zz * items + threadIdx.x];
      p.sums[threadIdx.x] = (float)sum;
    }
  }
  if (threadIdx.x == 0) *p.ticket = 0u;
}

int bn_bwd_reduce(float *sums, const void *g, const void *out, const void *z, const float *mean, const float *invstd,
                  long long P, int C, void *workspace, long long workspace_bytes, cudaStream_t stream) {
  BC_REQUIRE(sums && g && z && mean && invstd && workspace, BC_ERR_NULL, "bc_bn_bwd_reduce: NULL pointer");
  BC_REQUIRE(P > 0 && P < (1ll << 31), BC_ERR_SHAPE, "bc_bn_bwd_reduce: %lld pixels", P);
  BC_REQUIRE(C >= 8 && C <= 128 && C % 8 == 0 && kBwdThreads % (2 * C) == 0, BC_ERR_UNSUPPORTED,
             "bc_bn_bwd_reduce: C=%d (8, 16, 32, 64 or 128 channels)", C);
  BC_REQUIRE((((uintptr_t)g | (uintptr_t)out | (uintptr_t)z | (uintptr_t)workspace) & 15) == 0, BC_ERR_ALIGN,
             "bc_bn_bwd_reduce: 16-byte alignment");
  const int rows = kBwdThreads / (C / 8);
  long long grid = (P + 4 * rows - 1) / (4 * rows);
  if (grid > kNumSMs) grid = kNumSMs;
  const long long need = 16 + grid * 2 * C * (long long)sizeof(float);
  BC_REQUIRE(workspace_bytes >= need, BC_ERR_RANGE, "bc_bn_bwd_reduce: workspace of %lld bytes, %lld needed", workspace_bytes, need);
  BnBwdReduceParams p;
  p.g = (const __half *)g; p.out = (const __half *)out; p.z = (const __half *)z;
  p.mean = mean; p.invstd = invstd; p.sums = sums;
  p.ticket = (unsigned int *)workspace;
  p.partial = (float *)((char *)workspace + 16);
  p.P = (uint32_t)P; p.C = C; p.zero = 0u;
  launch_kernel(bn_bwd_reduce_kernel, dim3((unsigned)grid), dim3(kBwdThreads), 0, stream, 1, p);
  return check_launch("bc_bn_bwd_reduce");
}

// =====================================================================================================
// bc_bn_bwd_apply: dz = gamma * invstd * (g - sum_g / P - xhat * sum_gx / P),  g = out > 0 ? dOut : 0
// dz_up (optional): the same values at (n, 2y, 2x) of a (N, 2H, 2W, C) plane whose other positions stay zero.
// =====================================================================================================
struct BnBwdApplyParams {
  const __half *g, *out, *z;
  const float *mean, *invstd, *gamma, *sums;
  __half *dz, *dz_up;
  FastDiv chunks_per_pixel, per_image, per_row;  // C/8, H*W, W
  uint32_t total;                                // P * C / 8
  int C, H, W;
  float inv_count;
};

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnBwdApplyParams p) {
  pdl_trigger();
  pdl_wait();
  const uint32_t stride = gridDim.x * blockDim.x;  // a multiple of C/8: a thread keeps its 8 channels
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  uint32_t pix, ch;
  p.chunks_per_pixel.divmod(i, pix, ch);
  const int c0 = (int)ch * 8;
  float mean[8], istd[8], w[8], b[8], c[8];  // dz = w * g - b - xhat * c
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mean[k] = __ldg(p.mean + c0 + k);
    istd[k] = __ldg(p.invstd + c0 + k);
    w[k] = (p.gamma ? __ldg(p.gamma + c0 + k) : 1.f) * istd[k];
    b[k] = w[k] * __ldg(p.sums + c0 + k) * p.inv_count;
    c[k] = w[k] * __ldg(p.sums + p.C + c0 + k) * p.inv_count;
  }
  for (; i < p.total; i += stride) {
    pix = p.chunks_per_pixel.div(i);
    const size_t off = (size_t)pix * p.C + c0;
    uint4 ug = __ldg(reinterpret_cast<const uint4 *>(p.g + off));
    const uint4 uz = __ldg(reinterpret_cast<const uint4 *>(p.z + off));
    if (p.out) ug = mask_by_positive(ug, __ldg(reinterpret_cast<const uint4 *>(p.out + off)));
    const __half2 *gh = reinterpret_cast<const __half2 *>(&ug);
    const __half2 *zh = reinterpret_cast<const __half2 *>(&uz);
    uint4 o;
    __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 g = __half22float2(gh[k]), z = __half22float2(zh[k]);
      const float x0 = (z.x - mean[2 * k]) * istd[2 * k], x1 = (z.y - mean[2 * k + 1]) * istd[2 * k + 1];
      oh[k] = __floats2half2_rn(w[2 * k] * g.x - b[2 * k] - x0 * c[2 * k], w[2 * k + 1] * g.y - b[2 * k + 1] - x1 * c[2 * k + 1]);
    }
    if (p.dz) *reinterpret_cast<uint4 *>(p.dz + off) = o;
    if (p.dz_up) {
      uint32_t n, rem, y, x;
      p.per_image.divmod(pix, n, rem);
      p.per_row.divmod(rem, y, x);
      const size_t up = (((size_t)n * 2 * p.H + 2 * y) * 2 * p.W + 2 * x) * p.C + c0;
      *reinterpret_cast<uint4 *>(p.dz_up + up) = o;
    }
  }
}

int bn_bwd_apply(void *dz, void *dz_up, const void *g, const void *out, const void *z, const float *mean, const float *invstd,
                 const float *gamma, const float *sums, int N, int H, int W, int C, cudaStream_t stream) {
  BC_REQUIRE((dz || dz_up) && g && z && mean && invstd && sums, BC_ERR_NULL, "bc_bn_bwd_apply: NULL pointer");
  BC_REQUIRE(N > 0 && H > 0 && W > 0 && C >= 8 && C % 8 == 0 && 256 % (C / 8) == 0, BC_ERR_SHAPE,
             "bc_bn_bwd_apply: N=%d %dx%d C=%d (C/8 must divide 256)", N, H, W, C);
  const long long total = (long long)N * H * W * (C / 8);
  BC_REQUIRE(total < (1ll << 31), BC_ERR_RANGE, "bc_bn_bwd_apply: problem too large");
  BC_REQUIRE((((uintptr_t)dz | (uintptr_t)dz_up | (uintptr_t)g | (uintptr_t)out | (uintptr_t)z) & 15) == 0, BC_ERR_ALIGN,
             "bc_bn_bwd_apply: 16-byte alignment");
  BnBwdApplyParams p;
  p.g = (const __half *)g; p.out = (const __half *)out; p.z = (const __half *)z;
  p.mean = mean; p.invstd = invstd; p.gamma = gamma; p.sums = sums;
  p.dz = (__half *)dz; p.dz_up = (__half *)dz_up;
  p.chunks_per_pixel = FastDiv((uint32_t)(C / 8));
  p.per_image = FastDiv((uint32_t)(H * W));
  p.per_row = FastDiv((uint32_t)W);
  p.total = (uint32_t)total; p.C = C; p.H = H; p.W = W;
  p.inv_count = 1.f / (float)((long long)N * H * W);
  long long grid = (total + 255) / 256;
  if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
  launch_kernel(bn_bwd_apply_kernel, dim3((unsigned)grid), dim3(256), 0, stream, 1, p);
  return check_launch("bc_bn_bwd_apply");
}

// =====================================================================================================
// bc_conv_wgrad
// =====================================================================================================
// One CTA tile = 64 output pixels (TR rows x TC columns, TC a power of two <= 64) of one image: the dz tile
// [64][CO] and the input window [(TR-1)*s + k][(TC-1)*s + k][CI] are staged in shared memory (cp.async, zero fill
// outside the image; row pitch + 16 bytes so that the eight 16-byte rows of an ldmatrix land in distinct banks), and
// for each of the CTA's taps the warps accumulate D[co][ci] += sum_p dz[p][co] * x[win(p, tap)][ci] with
// mma.sync.m16n8k16 (A = dz^T and B = x both come out of ldmatrix.trans, because both are stored pixel-major).
// Warp w owns output rows 16*(w % WM) .. +15 and CI / WN columns, for all the CTA's taps, over ALL of the CTA's tiles;
// at the end the accumulators go to partial[blockIdx.x][tap][co][ci] and wgrad_finish_kernel adds the partials in
// CTA order.
constexpr int kWgThreads = 256;

struct WgradParams {
  const __half *dz;  // (N, Ho, Wo, CO)
  const __half *x;   // (N, H, W, CI)
  float *partial;    // [gridDim.x][k*k][CO][CI]
  int N, H, W, Ho, Wo, ksize, stride, pad;
  int TR, TC, tc_shift, WR, WC;  // tile rows / columns (TR * TC = 64), window rows / columns
  int tiles_y, tiles_x, ntiles, stages;
  int co_pitch, ci_pitch;  // channels per pixel of the dz / x planes in memory (>= the CO / CI the kernel works on)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// KSPLIT = 2 (the 32-channel layers: too few output elements for eight warps): warps 4..7 take the second half of the
// tile's 64 pixels and write a partial of their own.
template <int CO, int CI, int TAPS, int KSPLIT = 1>
__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_kernel(const WgradParams p) {
  constexpr int WM = CO / 16 < 8 ? CO / 16 : 8;  // warps along the output channels
  constexpr int WN = 8 / (WM * KSPLIT);           // warps along the input channels
  constexpr int NEXT = CI / WN;                   // input channels per warp
  constexpr int NT = NEXT / 8;                    // n8 tiles per warp and tap
  constexpr int DZ_PITCH = CO * 2 + 16, X_PITCH = CI * 2 + 16;
  static_assert(CO / 16 == WM, "one m16 block per warp row");
  static_assert(NT % 2 == 0, "ldmatrix.x4 covers two n8 tiles");
  extern __shared__ __align__(128) unsigned char smem[];
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp % WM, wn = (warp / WM) % WN, kg = warp / (WM * WN);
  const int stage_bytes = 64 * DZ_PITCH + p.WR * p.WC * X_PITCH;
  const uint32_t smem_base = smem_u32(smem);
  const int tap0 = blockIdx.y * TAPS;

  float acc[TAPS][NT][4];
#pragma unroll
  for (int t = 0; t < TAPS; ++t)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[t][n][e] = 0.f;

  // per-lane ldmatrix row addresses (relative to the stage base)
  const int q = lane >> 3, j = lane & 7;
  uint32_t a_off[4], b_off[4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int ka = ks * 16 + j + 8 * (q >> 1);  // A: matrices 2, 3 hold pixels 8..15
    a_off[ks] = (uint32_t)(ka * DZ_PITCH + (wm * 16 + 8 * (q & 1)) * 2);
    const int kb = ks * 16 + j + 8 * (q & 1);   // B: matrices 1, 3 hold pixels 8..15
    const int r = kb >> p.tc_shift, c = kb & (p.TC - 1);
    b_off[ks] = (uint32_t)(64 * DZ_PITCH + (r * p.stride * p.WC + c * p.stride) * X_PITCH + (wn * NEXT + 8 * (q >> 1)) * 2);
  }

  uint32_t tap_off[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
    const int tap = tap0 + t, kh = tap / p.ksize, kw = tap - kh * p.ksize;
    tap_off[t] = (uint32_t)((kh * p.WC + kw) * X_PITCH);
  }

  auto issue = [&](int tile, int buf) {
    const int n = tile / (p.tiles_y * p.tiles_x), rem = tile - n * p.tiles_y * p.tiles_x;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    const uint32_t base = smem_base + buf * stage_bytes;
    constexpr int DZ_CH = CO / 8, X_CH = CI / 8;
    for (int i = threadIdx.x; i < 64 * DZ_CH; i += kWgThreads) {
      const int px = i / DZ_CH, ch = i - px * DZ_CH;
      const int y = ty * p.TR + (px >> p.tc_shift), x = tx * p.TC + (px & (p.TC - 1));
      const bool ok = y < p.Ho && x < p.Wo;
      const __half *src = p.dz + (ok ? (((size_t)n * p.Ho + y) * p.Wo + x) * p.co_pitch + ch * 8 : 0);
      cp_async16(base + px * DZ_PITCH + ch * 16, src, ok);
    }
    const int oy = ty * p.TR * p.stride - p.pad, ox = tx * p.TC * p.stride - p.pad;
    const int wpx = p.WR * p.WC;
    for (int i = threadIdx.x; i < wpx * X_CH; i += kWgThreads) {
      const int px = i / X_CH, ch = i - px * X_CH;
      const int wr = px / p.WC, wc = px - wr * p.WC;
      const int y = oy + wr, x = ox + wc;
      const bool ok = y >= 0 && y < p.H && x >= 0 && x < p.W;
      const __half *src = p.x + (ok ? (((size_t)n * p.H + y) * p.W + x) * p.ci_pitch + ch * 8 : 0);
      cp_async16(base + 64 * DZ_PITCH + px * X_PITCH + ch * 16, src, ok);
    }
  };

  int buf = 0;
  int tile = blockIdx.x;
  if (tile < p.ntiles) issue(tile, 0);
  cp_async_commit();
  for (; tile < p.ntiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (p.stages == 2) {
      if (next < p.ntiles) issue(next, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t base = smem_base + buf * stage_bytes;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (KSPLIT == 2 && (ks >> 1) != kg) continue;
      uint32_t a[4];
      ldmatrix_x4_trans(a, base + a_off[ks]);
#pragma unroll
      for (int t = 0; t < TAPS; ++t) {
#pragma unroll
        for (int n2 = 0; n2 < NT / 2; ++n2) {
          uint32_t b[4];
          ldmatrix_x4_trans(b, base + b_off[ks] + tap_off[t] + n2 * 32);
          mma_16816(acc[t][2 * n2], a, b[0], b[1]);
          mma_16816(acc[t][2 * n2 + 1], a, b[2], b[3]);
        }
      }
    }
    __syncthreads();
    if (p.stages == 2) {
      buf ^= 1;
    } else {
      if (next < p.ntiles) issue(next, 0);
      cp_async_commit();
    }
  }
  cp_async_wait<0>();

  const int g = lane >> 2, t4 = lane & 3;
  const int taps_total = p.ksize * p.ksize;
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
    float *dst = p.partial + (((size_t)(blockIdx.x * KSPLIT + kg) * taps_total + tap0 + t) * CO + wm * 16 + g) * CI + wn * NEXT + 2 * t4;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      *reinterpret_cast<float2 *>(dst + n * 8) = make_float2(acc[t][n][0], acc[t][n][1]);
      *reinterpret_cast<float2 *>(dst + 8 * CI + n * 8) = make_float2(acc[t][n][2], acc[t][n][3]);
    }
  }
}

// grad[co, ci, kh, kw] = inv_scale * sum_b partial[b][tap][co][ci] for the real (unpadded) channels, in the fp32
// parameter's own strides; the BatchNorm's d(gamma) = inv_scale * sums[1][c], d(beta) = inv_scale * sums[0][c] ride
// along (last CTA row).
struct WgradFinishParams {
  const float *partial;
  float *grad;
  long long gs[4];  // element strides of grad: (co, ci, kh, kw)
  const float *inv_scale;
  const float *bn_sums;
  float *dgamma, *dbeta;
  int nparts, ksize, CO, CI, Cout, Cin, bn_C, bn_Cp;
};

// Block = 32 outputs x 8 slices: slice s adds partials s, s+8, ... (all loads independent and in flight), the slices are
// then added in order through shared memory: the same sum order every run.
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const WgradFinishParams p) {
  __shared__ float red[8][33];
  pdl_trigger();
  pdl_wait();
  const float inv = p.inv_scale ? __ldg(p.inv_scale) : 1.f;
  const int taps = p.ksize * p.ksize;
  const int total = taps * p.Cout * p.Cin;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  if (blockIdx.x * 32 >= total) {  // trailing block(s): the BatchNorm gradients
    const int c = (blockIdx.x * 32 - (total + 31) / 32 * 32) + lane;
    if (slice == 0 && p.bn_sums && c < p.bn_C) {
      if (p.dbeta) p.dbeta[c] = __ldg(p.bn_sums + c) * inv;
      if (p.dgamma) p.dgamma[c] = __ldg(p.bn_sums + p.bn_Cp + c) * inv;
    }
    return;
  }
  float s = 0.f;
  int ci = 0, co = 0, tap = 0;
  if (i < total) {
    ci = i % p.Cin; co = (i / p.Cin) % p.Cout; tap = i / (p.Cin * p.Cout);
    const size_t per = (size_t)taps * p.CO * p.CI;
    const float *src = p.partial + ((size_t)tap * p.CO + co) * p.CI + ci;
    for (int b0 = slice; b0 < p.nparts; b0 += 32) {  // four independent loads in flight; few partials = few rounds
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int b = b0 + 8 * k;
        v[k] = b < p.nparts ? __ldcg(src + (size_t)b * per) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) s += v[k];
    }
  }
  red[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && i < total) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
    p.grad[co * p.gs[0] + ci * p.gs[1] + kh * p.gs[2] + kw * p.gs[3]] = t * inv;
  }
}

template <int CO, int CI, int TAPS, int KSPLIT = 1>
static int launch_wgrad(const WgradParams &p, int grid_x, int smem, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel<CO, CI, TAPS, KSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail((int)e, "bc_conv_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  launch_kernel(conv_wgrad_kernel<CO, CI, TAPS, KSPLIT>, dim3((unsigned)grid_x, (unsigned)(p.ksize * p.ksize / TAPS)), dim3(kWgThreads),
                (size_t)smem, stream, 1, p);
  return check_launch("bc_conv_wgrad");
}

int conv_wgrad(float *grad_w, const long long *grad_strides, const void *dz, const void *x, int N, int H, int W, int Cin_p,
               int Cout_p, int Cin, int Cout, int ksize, int stride, const float *inv_scale, float *dgamma, float *dbeta,
               const float *bn_sums, void *workspace, long long workspace_bytes, cudaStream_t stream) {
  BC_REQUIRE(grad_w && grad_strides && dz && x && workspace, BC_ERR_NULL, "bc_conv_wgrad: NULL pointer");
  BC_REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2), BC_ERR_UNSUPPORTED, "bc_conv_wgrad: k=%d stride=%d", ksize, stride);
  BC_REQUIRE((Cout_p == 64 || Cout_p == 128) && (Cin_p == 64 || Cin_p == 128) && Cin_p <= Cout_p, BC_ERR_UNSUPPORTED,
             "bc_conv_wgrad: padded channels %d -> %d (64 or 128, Cin <= Cout)", Cin_p, Cout_p);
  BC_REQUIRE(Cin > 0 && Cin <= Cin_p && Cout > 0 && Cout <= Cout_p, BC_ERR_SHAPE, "bc_conv_wgrad: real channels %d -> %d", Cin, Cout);
  BC_REQUIRE(N > 0 && H > 0 && W > 0 && H % stride == 0 && W % stride == 0, BC_ERR_SHAPE, "bc_conv_wgrad: N=%d %dx%d", N, H, W);
  BC_REQUIRE((((uintptr_t)dz | (uintptr_t)x | (uintptr_t)workspace) & 15) == 0, BC_ERR_ALIGN, "bc_conv_wgrad: 16-byte alignment");
  WgradParams p;
  p.dz = (const __half *)dz; p.x = (const __half *)x; p.partial = (float *)workspace;
  p.N = N; p.H = H; p.W = W; p.ksize = ksize; p.stride = stride; p.pad = ksize / 2;
  p.Ho = H / stride; p.Wo = W / stride;  // pad = k/2, even sizes: (H + 2*pad - k) / s + 1 == H / s
  int tc = 64;
  while (tc > 8 && tc / 2 >= p.Wo) tc /= 2;
  p.TC = tc; p.TR = 64 / tc;
  p.tc_shift = 0;
  while ((1 << p.tc_shift) < tc) ++p.tc_shift;
  p.WR = (p.TR - 1) * stride + ksize; p.WC = (p.TC - 1) * stride + ksize;
  p.tiles_y = (p.Ho + p.TR - 1) / p.TR; p.tiles_x = (p.Wo + p.TC - 1) / p.TC;
  p.ntiles = N * p.tiles_y * p.tiles_x;
  // channel counts the kernel works on: the planes' padded counts, or 32 x 32 when the real counts fit (the policy net's
  // 256x512-pixel layers: a quarter of the padded flops, half of the operand bytes)
  const bool small = ksize == 3 && Cout <= 32 && Cin <= 32;
  const int co = small ? 32 : Cout_p, ci = small ? 32 : Cin_p, ksplit = small ? 2 : 1;
  p.co_pitch = Cout_p; p.ci_pitch = Cin_p;
  const int stage = 64 * (co * 2 + 16) + p.WR * p.WC * (ci * 2 + 16);
  p.stages = 2 * stage <= 200 * 1024 ? 2 : 1;
  BC_REQUIRE(stage <= 200 * 1024, BC_ERR_UNSUPPORTED, "bc_conv_wgrad: tile of %d bytes", stage);
  const int taps = ksize * ksize;
  const int taps_per_cta = (co <= 64) ? taps : (ci == 64 ? (ksize == 3 ? 3 : 1) : 1);
  int grid_x = kNumSMs / (taps / taps_per_cta);  // ~one CTA per SM over both grid dimensions: few partials to add
  if (grid_x > p.ntiles) grid_x = p.ntiles;
  const long long need = (long long)grid_x * ksplit * taps * co * ci * (long long)sizeof(float);
  BC_REQUIRE(workspace_bytes >= need, BC_ERR_RANGE, "bc_conv_wgrad: workspace of %lld bytes, %lld needed", workspace_bytes, need);
  const int smem = p.stages * stage;
  int rc;
  if (small) rc = launch_wgrad<32, 32, 9, 2>(p, grid_x, smem, stream);
  else if (co == 64 && ci == 64) rc = ksize == 3 ? launch_wgrad<64, 64, 9>(p, grid_x, smem, stream) : launch_wgrad<64, 64, 1>(p, grid_x, smem, stream);
  else if (co == 128 && ci == 64) rc = ksize == 3 ? launch_wgrad<128, 64, 3>(p, grid_x, smem, stream) : launch_wgrad<128, 64, 1>(p, grid_x, smem, stream);
  else rc = launch_wgrad<128, 128, 1>(p, grid_x, smem, stream);
  if (rc) return rc;
  WgradFinishParams f;
  f.partial = p.partial; f.grad = grad_w;
  for (int i = 0; i < 4; ++i) f.gs[i] = grad_strides[i];
  f.inv_scale = inv_scale; f.bn_sums = bn_sums; f.dgamma = dgamma; f.dbeta = dbeta;
  f.nparts = grid_x * ksplit; f.ksize = ksize; f.CO = co; f.CI = ci; f.Cout = Cout; f.Cin = Cin;
  f.bn_C = bn_sums ? Cout : 0; f.bn_Cp = Cout_p;
  const int blocks = (taps * Cout * Cin + 31) / 32 + (f.bn_C + 31) / 32;
  launch_kernel(wgrad_finish_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, 1, f);
  return check_launch("bc_conv_wgrad (finish)");
}

}  // namespace bc
