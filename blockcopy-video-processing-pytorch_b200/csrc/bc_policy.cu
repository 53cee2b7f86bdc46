// bc_policy.cu -- the two feature builders around the policy CNN.
//   bc_policy_features : PolicyNet's input tensor (reference policy/net.py:84-113): nearest-resized frame |
//                        frame_state | previous output - 0.5 | previous grid - 0.5, fp32 NCHW, one pass
//                        instead of 4 interpolates + casts + subtractions + cat
//   bc_info_gain       : InformationGainSemSeg.forward (policy/information_gain.py:32-41): bilinear 1/4 of the
//                        current and previous logits, log_softmax over classes, KL(prev || cur) per class,
//                        mean over classes -> (N,1,h/4,w/4), one pass instead of 6 kernels
#include <cuda_fp16.h>

#include "bc_common.cuh"

namespace bc {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }

struct FeatParams {
  const void *frame, *state, *repr;  // (N,3,H,W) x2 NCHW; (N,K,h,w) with explicit strides
  const uint8_t *grid;               // (N,1,GH,GW) bool
  float *out;                        // (N, 7+K, Ho, Wo) fp32 NCHW
  int N, K, H, W, h, w, GH, GW, Ho, Wo;
  int64_t repr_sn, repr_sc, repr_sh, repr_sw;  // element strides of repr
  float sy_frame, sx_frame, sy_repr, sx_repr, sy_grid, sx_grid;  // nearest scales (in / out)
  uint32_t total;
};

__device__ __forceinline__ int nearest_src(float scale, int dst, int in_size) {
  return min((int)floorf((float)dst * scale), in_size - 1);  // ATen nearest_neighbor_compute_source_index
}

template <typename T>
__device__ __forceinline__ float feat_value(const FeatParams &p, int n, int c, int y, int x) {
  if (c < 6) {
    const T *src = reinterpret_cast<const T *>(c < 3 ? p.frame : p.state);
    const int cc = c < 3 ? c : c - 3;
    const int yy = nearest_src(p.sy_frame, y, p.H), xx = nearest_src(p.sx_frame, x, p.W);
    return to_f<T>(src[(((size_t)n * 3 + cc) * p.H + yy) * p.W + xx]);
  }
  if (c < 6 + p.K) {
    const int yy = nearest_src(p.sy_repr, y, p.h), xx = nearest_src(p.sx_repr, x, p.w);
    return to_f<T>(reinterpret_cast<const T *>(p.repr)[n * p.repr_sn + (c - 6) * p.repr_sc + yy * p.repr_sh + xx * p.repr_sw]) - 0.5f;
  }
  const int yy = nearest_src(p.sy_grid, y, p.GH), xx = nearest_src(p.sx_grid, x, p.GW);
  return (p.grid[((size_t)n * p.GH + yy) * p.GW + xx] ? 1.f : 0.f) - 0.5f;
}

template <typename T>
__global__ void __launch_bounds__(256) policy_features_kernel(const FeatParams p) {
  pdl_trigger();
  pdl_wait();
  const int C = 7 + p.K;
  const uint32_t gstride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += gstride) {
    const int x = (int)(i % (uint32_t)p.Wo);
    const int y = (int)((i / (uint32_t)p.Wo) % (uint32_t)p.Ho);
    const int c = (int)((i / ((uint32_t)p.Wo * p.Ho)) % (uint32_t)C);
    const int n = (int)(i / ((uint32_t)p.Wo * p.Ho * C));
    p.out[i] = feat_value<T>(p, n, c, y, x);
  }
}

// The same features as fp16 NHWC with the channel count padded to Cp (the input plane of the fused policy trunk,
// policy/fused_net.py): one thread per (pixel, 8-channel chunk), 16-byte stores; chunks beyond the last real channel
// are not touched (the caller zeroes the plane once).  Values = the fp32 features rounded to fp16.
template <typename T>
__global__ void __launch_bounds__(256) policy_features_nhwc16_kernel(const FeatParams p, __half *out16, int Cp, int chunks,
                                                                     uint32_t total) {
  pdl_trigger();
  pdl_wait();
  const int C = 7 + p.K;
  const uint32_t gstride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {
    const int chunk = (int)(i % (uint32_t)chunks);
    const uint32_t pix = i / (uint32_t)chunks;
    const int x = (int)(pix % (uint32_t)p.Wo), y = (int)((pix / (uint32_t)p.Wo) % (uint32_t)p.Ho);
    const int n = (int)(pix / ((uint32_t)p.Wo * p.Ho));
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c0 = chunk * 8 + 2 * k;
      const float a = c0 < C ? feat_value<T>(p, n, c0, y, x) : 0.f, b = c0 + 1 < C ? feat_value<T>(p, n, c0 + 1, y, x) : 0.f;
      const __half2 h = __floats2half2_rn(a, b);
      w[k] = *reinterpret_cast<const uint32_t *>(&h);
    }
    *reinterpret_cast<uint4 *>(out16 + (size_t)pix * Cp + chunk * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

static int fill_feat_params(FeatParams &p, const void *frame, const void *state, const void *repr, const uint8_t *grid, int N,
                            int K, int H, int W, int h, int w, int GH, int GW, int Ho, int Wo, const int64_t *repr_strides,
                            float sy_frame, float sx_frame, int dtype, const char *who) {
  BC_REQUIRE(frame && state && repr && grid && repr_strides, BC_ERR_NULL, "%s: NULL pointer", who);
  BC_REQUIRE(N > 0 && K > 0 && Ho > 0 && Wo > 0, BC_ERR_SHAPE, "%s: empty problem", who);
  BC_REQUIRE(dtype == BC_F16 || dtype == BC_F32, BC_ERR_DTYPE, "%s: dtype", who);
  p.frame = frame; p.state = state; p.repr = repr; p.grid = grid; p.out = nullptr;
  p.N = N; p.K = K; p.H = H; p.W = W; p.h = h; p.w = w; p.GH = GH; p.GW = GW; p.Ho = Ho; p.Wo = Wo;
  p.repr_sn = repr_strides[0]; p.repr_sc = repr_strides[1]; p.repr_sh = repr_strides[2]; p.repr_sw = repr_strides[3];
  p.sy_frame = sy_frame; p.sx_frame = sx_frame;        // = 1 / scale_factor, as ATen uses for scale_factor= calls
  p.sy_repr = (float)h / Ho; p.sx_repr = (float)w / Wo;  // size= calls: in / out
  p.sy_grid = (float)GH / Ho; p.sx_grid = (float)GW / Wo;
  p.total = 0;
  return BC_OK;
}

int policy_features(float *out, const void *frame, const void *state, const void *repr, const uint8_t *grid, int N,
                    int K, int H, int W, int h, int w, int GH, int GW, int Ho, int Wo, const int64_t *repr_strides,
                    float sy_frame, float sx_frame, int dtype, cudaStream_t stream) {
  BC_REQUIRE(out, BC_ERR_NULL, "bc_policy_features: NULL pointer");
  FeatParams p;
  const int rc = fill_feat_params(p, frame, state, repr, grid, N, K, H, W, h, w, GH, GW, Ho, Wo, repr_strides, sy_frame,
                                  sx_frame, dtype, "bc_policy_features");
  if (rc != BC_OK) return rc;
  p.out = out;
  const int64_t total = (int64_t)N * (7 + K) * Ho * Wo;
  BC_REQUIRE(total < (1ll << 31), BC_ERR_RANGE, "bc_policy_features: problem too large");
  p.total = (uint32_t)total;
  int64_t gridsz = (total + 255) / 256;
  if (gridsz > (int64_t)kNumSMs * 16) gridsz = (int64_t)kNumSMs * 16;
  if (dtype == BC_F16)
    launch_kernel(policy_features_kernel<__half>, dim3((unsigned)gridsz), dim3(256), 0, stream, 1, p);
  else
    launch_kernel(policy_features_kernel<float>, dim3((unsigned)gridsz), dim3(256), 0, stream, 1, p);
  return check_launch("bc_policy_features");
}

int policy_features_nhwc16(void *out, int Cp, const void *frame, const void *state, const void *repr, const uint8_t *grid,
                           int N, int K, int H, int W, int h, int w, int GH, int GW, int Ho, int Wo,
                           const int64_t *repr_strides, float sy_frame, float sx_frame, int dtype, cudaStream_t stream) {
  BC_REQUIRE(out && ((uintptr_t)out & 15) == 0, BC_ERR_ALIGN, "bc_policy_features_nhwc16: out must be 16-byte aligned");
  BC_REQUIRE(Cp % 8 == 0 && Cp >= 7 + K, BC_ERR_SHAPE, "bc_policy_features_nhwc16: padded channel count %d", Cp);
  FeatParams p;
  const int rc = fill_feat_params(p, frame, state, repr, grid, N, K, H, W, h, w, GH, GW, Ho, Wo, repr_strides, sy_frame,
                                  sx_frame, dtype, "bc_policy_features_nhwc16");
  if (rc != BC_OK) return rc;
  const int chunks = (7 + K + 7) / 8;
  const int64_t total = (int64_t)N * Ho * Wo * chunks;
  BC_REQUIRE((int64_t)N * Ho * Wo * Cp < (1ll << 31), BC_ERR_RANGE, "bc_policy_features_nhwc16: problem too large");
  int64_t gridsz = (total + 255) / 256;
  if (gridsz > (int64_t)kNumSMs * 16) gridsz = (int64_t)kNumSMs * 16;
  if (dtype == BC_F16)
    launch_kernel(policy_features_nhwc16_kernel<__half>, dim3((unsigned)gridsz), dim3(256), 0, stream, 1, p, (__half *)out, Cp,
                  chunks, (uint32_t)total);
  else
    launch_kernel(policy_features_nhwc16_kernel<float>, dim3((unsigned)gridsz), dim3(256), 0, stream, 1, p, (__half *)out, Cp,
                  chunks, (uint32_t)total);
  return check_launch("bc_policy_features_nhwc16");
}

// ---------------------------------------------------------------------------------------------------
struct IgParams {
  const __half *cur, *prev;  // (N,K,h,w) with explicit strides (same for both)
  __half *out;               // (N,1,ho,wo) contiguous
  int N, K, h, w, ho, wo;
  int64_t sn, sc, sh, sw;
  uint32_t total;
};

constexpr int kMaxClasses = 64;

__device__ __forceinline__ float rh(float x) { return __half2float(__float2half_rn(x)); }

// bilinear, align_corners = False, scale 4: src = 4 * (dst + 0.5) - 0.5 = 4 dst + 1.5 -> taps 4dst+1, 4dst+2, weights 1/2
__device__ __forceinline__ float quarter_tap(const __half *t, const IgParams &p, int n, int c, int y, int x) {
  const int y1 = 4 * y + 1, x1 = 4 * x + 1;
  const int yp = y1 < p.h - 1 ? 1 : 0, xp = x1 < p.w - 1 ? 1 : 0;
  const __half *b = t + n * p.sn + c * p.sc;
  const float v00 = __half2float(b[y1 * p.sh + x1 * p.sw]), v01 = __half2float(b[y1 * p.sh + (x1 + xp) * p.sw]);
  const float v10 = __half2float(b[(y1 + yp) * p.sh + x1 * p.sw]), v11 = __half2float(b[(y1 + yp) * p.sh + (x1 + xp) * p.sw]);
  return rh(0.5f * (0.5f * v00 + 0.5f * v01) + 0.5f * (0.5f * v10 + 0.5f * v11));
}

__global__ void __launch_bounds__(128) info_gain_kernel(const IgParams p) {
  pdl_trigger();
  pdl_wait();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.total) return;
  const int x = (int)(i % (uint32_t)p.wo), y = (int)((i / (uint32_t)p.wo) % (uint32_t)p.ho);
  const int n = (int)(i / ((uint32_t)p.wo * p.ho));
  float a[kMaxClasses], b[kMaxClasses];
  float ma = -INFINITY, mb = -INFINITY;
  for (int c = 0; c < p.K; ++c) {
    a[c] = quarter_tap(p.cur, p, n, c, y, x);
    b[c] = quarter_tap(p.prev, p, n, c, y, x);
    ma = fmaxf(ma, a[c]);
    mb = fmaxf(mb, b[c]);
  }
  float sa = 0.f, sb = 0.f;
  for (int c = 0; c < p.K; ++c) {
    sa += expf(a[c] - ma);
    sb += expf(b[c] - mb);
  }
  const float la = logf(sa), lb = logf(sb);
  float acc = 0.f;
  for (int c = 0; c < p.K; ++c) {
    const float lpa = rh(a[c] - ma - la), lpb = rh(b[c] - mb - lb);  // log_softmax outputs are fp16 tensors
    acc += rh(expf(lpb) * (lpb - lpa));                                 // kl_div(log_target=True), pointwise, fp16
  }
  p.out[i] = __float2half_rn(acc / (float)p.K);
}

int info_gain(void *out, const void *cur, const void *prev, int N, int K, int h, int w, const int64_t *strides,
              cudaStream_t stream) {
  BC_REQUIRE(out && cur && prev && strides, BC_ERR_NULL, "bc_info_gain: NULL pointer");
  BC_REQUIRE(N > 0 && K > 0 && K <= kMaxClasses, BC_ERR_UNSUPPORTED, "bc_info_gain: K=%d (1..%d classes)", K, kMaxClasses);
  BC_REQUIRE(h % 4 == 0 && w % 4 == 0 && h >= 4 && w >= 4, BC_ERR_SHAPE, "bc_info_gain: %dx%d is not a multiple of 4", h, w);
  IgParams p;
  p.cur = (const __half *)cur; p.prev = (const __half *)prev; p.out = (__half *)out;
  p.N = N; p.K = K; p.h = h; p.w = w; p.ho = h / 4; p.wo = w / 4;
  p.sn = strides[0]; p.sc = strides[1]; p.sh = strides[2]; p.sw = strides[3];
  const int64_t total = (int64_t)N * p.ho * p.wo;
  BC_REQUIRE(total < (1ll << 31), BC_ERR_RANGE, "bc_info_gain: problem too large");
  p.total = (uint32_t)total;
  launch_kernel(info_gain_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, stream, 1, p);
  return check_launch("bc_info_gain");
}

// ---------------------------------------------------------------------------------------------------
// Batch statistics of a train-mode BatchNorm2d (the policy net runs in train mode, policy/policy.py:240-250 and
// policy/net.py:115-125): per channel mean and 1/sqrt(biased variance + eps) over all P = N*H*W pixels of a dense
// NHWC fp16 tensor.  One pass: every CTA sums its pixel slice (fp32 per thread, 8 channels x 16-byte loads),
// reduces over its threads in shared memory (fixed order) and writes one partial per channel; the CTA that arrives last
// (atomic ticket) adds the partials IN CTA ORDER in double precision -- run-to-run reproducible -- and resets the
// ticket for the next launch.
struct StatsParams {
  const __half *x;
  float *mean, *invstd;
  float *partial;        // [gridDim.x][2][C]
  unsigned int *ticket;  // zero before the first launch; left at zero
  uint32_t P;
  int C;
  float eps;
  uint32_t zero;         // always 0 (see the note on load batching below)
  // APPLY form (bc_bn_norm): statistics, a grid-wide barrier, then out = relu?(weight * (x - mean) * invstd + shift)
  const float *weight, *shift;  // [C] or nullptr
  __half *out;
  int relu;
  unsigned int *done;    // second counter of the barrier (CTAs that have left it); zero between launches
};

constexpr int kStatsThreads = 512;

// Left alone, ptxas sinks every load of an unrolled batch next to its use (SASS: math interleaved after every LDG,
// ~2 loads in flight per thread) to save registers, which bounded the first versions of this kernel at ~1 TB/s.
// A data dependency keeps a batch together: every value is XOR-ed with (fold of ALL values of the batch) & zero,
// `zero` being a kernel parameter the compiler cannot fold -- no use can be scheduled before the last load.
// APPLY = true (bc_bn_norm): the normalisation follows in the SAME launch.  All CTAs (<= one per SM: co-resident) meet
// at a grid-wide barrier once their partials are written, every CTA then adds the partials itself (same order, same bits
// as the last-CTA reduction of the plain form) and normalises the pixels it has just read: the policy trunk's
// conv -> batch norm -> ReLU units are two dependent launches instead of three (~7 us each, mostly launch + latency).
template <bool APPLY>
__global__ void __launch_bounds__(kStatsThreads) bn_stats_kernel(const StatsParams p) {
  __shared__ float red[kStatsThreads][17];
  __shared__ double comb[kStatsThreads];
  __shared__ float s_mean[128], s_istd[128];
  __shared__ bool last;
  pdl_trigger();
  pdl_wait();
  const int lanes = p.C >> 3;                 // threads per pixel (8 channels each); C <= 128 -> <= 16
  const int sub = threadIdx.x % lanes, row = threadIdx.x / lanes, rows = kStatsThreads / lanes;
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
  const uint32_t step = gridDim.x * rows;
  const __half *base = p.x + sub * 8;
  uint32_t px = blockIdx.x * rows + row;
  // sixteen independent 16-byte loads in flight per thread, the tail included: the largest plane of the policy trunk
  // (131072 pixels x 64 channels on 148 CTAs) is ONE round of loads
  for (; px < p.P; px += 16 * step) {
    uint4 u[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {  // unconditional loads from a clamped pixel: nothing keeps them from being batched
      const uint32_t pj = px + j * step;
      u[j] = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)(pj < p.P ? pj : p.P - 1) * p.C));
    }
    uint32_t fold = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) fold ^= u[j].x;
    fold &= p.zero;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      u[j].x ^= fold; u[j].y ^= fold; u[j].z ^= fold; u[j].w ^= fold;
      if (px + j * step >= p.P) u[j] = make_uint4(0, 0, 0, 0);
      const __half2 *h = reinterpret_cast<const __half2 *>(&u[j]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h[k]);
        s[2 * k] += f.x; s[2 * k + 1] += f.y;
        q[2 * k] = fmaf(f.x, f.x, q[2 * k]); q[2 * k + 1] = fmaf(f.y, f.y, q[2 * k + 1]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { red[threadIdx.x][k] = s[k]; red[threadIdx.x][8 + k] = q[k]; }
  __syncthreads();
  const int items = 2 * p.C;  // (sum | sumsq, channel)
  {
    // item -> `slices` threads, each adds every slices-th pixel row of the CTA in order; then the slices in order
    const int slices = kStatsThreads / items, item = threadIdx.x % items, slice = threadIdx.x / items;
    const int which = item / p.C, c = item - which * p.C;
    float acc = 0.f;
    for (int r = slice; r < rows; r += slices) acc += red[r * lanes + (c >> 3)][which * 8 + (c & 7)];
    comb[threadIdx.x] = (double)acc;
    __syncthreads();
    if (threadIdx.x < items) {
      double t = 0.0;
      for (int z = 0; z < slices; ++z) t += comb[z * items + threadIdx.x];
      p.partial[(size_t)blockIdx.x * items + threadIdx.x] = (float)t;
    }
  }
  __threadfence();
  __syncthreads();
  if (APPLY) {
    if (threadIdx.x == 0) {
      atomicAdd(p.ticket, 1u);
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.ticket) : "memory");
      } while (seen < gridDim.x);
    }
    __syncthreads();
  } else {
    if (threadIdx.x == 0) last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
  }
  __threadfence();
  {
    // the last CTA (APPLY: every CTA): item -> `slices` threads, each adds every slices-th CTA partial in order
    const int slices = kStatsThreads / items, item = threadIdx.x % items, slice = threadIdx.x / items;
    double t = 0.0;
    // all of this thread's partials in flight at once (<= 148 CTAs / slices, padded to batches of 40)
    for (unsigned b0 = (unsigned)slice; b0 < gridDim.x; b0 += 40 * slices) {
      float v[40];
#pragma unroll
      for (int j = 0; j < 40; ++j) {
        const unsigned b = b0 + j * slices;
        v[j] = __ldcg(p.partial + (size_t)(b < gridDim.x ? b : gridDim.x - 1) * items + item);
      }
      uint32_t fold = 0;
#pragma unroll
      for (int j = 0; j < 40; ++j) fold ^= __float_as_uint(v[j]);
      fold &= p.zero;
#pragma unroll
      for (int j = 0; j < 40; ++j) v[j] = __uint_as_float(__float_as_uint(v[j]) ^ fold);
#pragma unroll
      for (int j = 0; j < 40; ++j) t += (b0 + j * slices < gridDim.x) ? (double)v[j] : 0.0;
    }
    comb[threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.x < p.C) {
      double sum = 0.0, sq = 0.0;
      for (int z = 0; z < slices; ++z) { sum += comb[z * items + threadIdx.x]; sq += comb[z * items + p.C + threadIdx.x]; }
      const double m = sum / (double)p.P;
      double var = sq / (double)p.P - m * m;
      var = var < 0.0 ? 0.0 : var;
      const float mf = (float)m, isf = (float)(1.0 / sqrt(var + (double)p.eps));
      if (!APPLY || blockIdx.x == 0) {
        p.mean[threadIdx.x] = mf;
        p.invstd[threadIdx.x] = isf;
      }
      if (APPLY) { s_mean[threadIdx.x] = mf; s_istd[threadIdx.x] = isf; }
    }
  }
  if (!APPLY) {
    if (threadIdx.x == 0) *p.ticket = 0u;
    return;
  }
  __syncthreads();
  // ---- normalise the pixels this CTA read for the statistics (ATen's eval-BN expression, as bc_ew_fused)
  {
    float mean[8], istd[8], w[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mean[k] = s_mean[sub * 8 + k];
      istd[k] = s_istd[sub * 8 + k];
      w[k] = p.weight ? __ldg(p.weight + sub * 8 + k) : 1.f;
      sh[k] = p.shift ? __ldg(p.shift + sub * 8 + k) : 0.f;
    }
    __half *obase = p.out + sub * 8;
    for (px = blockIdx.x * rows + row; px < p.P; px += 8 * step) {
      uint4 u[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t pj = px + j * step;
        u[j] = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)(pj < p.P ? pj : p.P - 1) * p.C));
      }
      uint32_t fold = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) fold ^= u[j].x;
      fold &= p.zero;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (px + j * step >= p.P) continue;
        u[j].x ^= fold;
        const __half2 *h = reinterpret_cast<const __half2 *>(&u[j]);
        uint4 o;
        __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(h[k]);
          float a = __half2float(__float2half_rn(w[2 * k] * (f.x - mean[2 * k]) * istd[2 * k] + sh[2 * k]));
          float b = __half2float(__float2half_rn(w[2 * k + 1] * (f.y - mean[2 * k + 1]) * istd[2 * k + 1] + sh[2 * k + 1]));
          if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
          oh[k] = __floats2half2_rn(a, b);
        }
        *reinterpret_cast<uint4 *>(obase + (size_t)(px + j * step) * p.C) = o;
      }
    }
  }
  // leave the barrier: the last CTA out resets both counters for the next launch
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(p.done, 1u) == gridDim.x - 1) {
      *p.ticket = 0u;
      *p.done = 0u;
    }
  }
}

int bn_stats(float *mean, float *invstd, const void *x, long long P, int C, float eps, void *workspace,
             long long workspace_bytes, cudaStream_t stream) {
  BC_REQUIRE(mean && invstd && x && workspace, BC_ERR_NULL, "bc_bn_stats: NULL pointer");
  BC_REQUIRE(P > 0 && P < (1ll << 31), BC_ERR_SHAPE, "bc_bn_stats: %lld pixels", P);
  BC_REQUIRE(C >= 8 && C <= 128 && C % 8 == 0 && kStatsThreads % (2 * C) == 0, BC_ERR_UNSUPPORTED,
             "bc_bn_stats: C=%d (8, 16, 32, 64 or 128 channels)", C);
  BC_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)workspace & 15) == 0, BC_ERR_ALIGN, "bc_bn_stats: 16-byte alignment");
  const int rows = kStatsThreads / (C / 8);
  // one CTA per SM at most, and up to sixteen pixel rows per thread and round: few partials for the last CTA to add
  long long grid = (P + 16 * rows - 1) / (16 * rows);
  if (grid > kNumSMs) grid = kNumSMs;
  const long long need = 16 + grid * 2 * C * (long long)sizeof(float);
  BC_REQUIRE(workspace_bytes >= need, BC_ERR_RANGE, "bc_bn_stats: workspace of %lld bytes, %lld needed", workspace_bytes, need);
  StatsParams p;
  p.x = (const __half *)x; p.mean = mean; p.invstd = invstd;
  p.ticket = (unsigned int *)workspace;
  p.partial = (float *)((char *)workspace + 16);
  p.P = (uint32_t)P; p.C = C; p.eps = eps; p.zero = 0u;
  p.weight = p.shift = nullptr; p.out = nullptr; p.relu = 0; p.done = nullptr;
  launch_kernel(bn_stats_kernel<false>, dim3((unsigned)grid), dim3(kStatsThreads), 0, stream, 1, p);
  return check_launch("bc_bn_stats");
}

int bn_norm(void *out, float *mean, float *invstd, const void *x, const float *weight, const float *shift, long long P, int C,
            float eps, int relu, void *workspace, long long workspace_bytes, cudaStream_t stream) {
  BC_REQUIRE(out && mean && invstd && x && workspace, BC_ERR_NULL, "bc_bn_norm: NULL pointer");
  BC_REQUIRE(P > 0 && P < (1ll << 31), BC_ERR_SHAPE, "bc_bn_norm: %lld pixels", P);
  BC_REQUIRE(C >= 8 && C <= 128 && C % 8 == 0 && kStatsThreads % (2 * C) == 0, BC_ERR_UNSUPPORTED,
             "bc_bn_norm: C=%d (8, 16, 32, 64 or 128 channels)", C);
  BC_REQUIRE((((uintptr_t)x | (uintptr_t)out | (uintptr_t)workspace) & 15) == 0, BC_ERR_ALIGN, "bc_bn_norm: 16-byte alignment");
  const int rows = kStatsThreads / (C / 8);
  long long grid = (P + 16 * rows - 1) / (16 * rows);
  if (grid > kNumSMs) grid = kNumSMs;  // one CTA per SM at most: the grid-wide barrier needs every CTA resident
  const long long need = 16 + grid * 2 * C * (long long)sizeof(float);
  BC_REQUIRE(workspace_bytes >= need, BC_ERR_RANGE, "bc_bn_norm: workspace of %lld bytes, %lld needed", workspace_bytes, need);
  StatsParams p;
  p.x = (const __half *)x; p.mean = mean; p.invstd = invstd;
  p.ticket = (unsigned int *)workspace;
  p.done = (unsigned int *)workspace + 1;
  p.partial = (float *)((char *)workspace + 16);
  p.P = (uint32_t)P; p.C = C; p.eps = eps; p.zero = 0u;
  p.weight = weight; p.shift = shift; p.out = (__half *)out; p.relu = relu;
  launch_kernel(bn_stats_kernel<true>, dim3((unsigned)grid), dim3(kStatsThreads), 0, stream, 1, p);
  return check_launch("bc_bn_norm");
}

// ---------------------------------------------------------------------------------------------------
// Parameter re-packing for the fused policy trunk: the live fp32 parameters change with every online optimiser
// step, the kernels want fp16 channels_last conv weights with channel counts padded to 64 (and fp32 affine
// vectors padded likewise).  One launch converts them all from a device-side table (built once: the addresses of
// parameters updated in place never change); padded positions are left as initialised by the caller.
//   entry e = 12 int64: src ptr | dst ptr | first flat element | Cout | Cin | k | padded Cin | src element strides
//   (co, ci, kh, kw) | dst is fp16 (1) or fp32 (0).   dst layout: [Cout][k][k][padded Cin].
__global__ void __launch_bounds__(256) pack_params_kernel(const long long *table, int n, long long total) {
  pdl_trigger();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int e = 0, hi = n - 1;
    while (e < hi) {  // last entry whose first element <= i
      const int mid = (e + hi + 1) >> 1;
      if (table[mid * 12 + 2] <= i) e = mid; else hi = mid - 1;
    }
    const long long *t = table + e * 12;
    const long long j = i - t[2], cin = t[4], k = t[5];
    const long long ci = j % cin, kw = (j / cin) % k, kh = (j / (cin * k)) % k, co = j / (cin * k * k);
    const float v = reinterpret_cast<const float *>(t[0])[co * t[7] + ci * t[8] + kh * t[9] + kw * t[10]];
    const long long d = ((co * k + kh) * k + kw) * t[6] + ci;
    if (t[11]) reinterpret_cast<__half *>(t[1])[d] = __float2half_rn(v);
    else reinterpret_cast<float *>(t[1])[d] = v;
  }
}

int pack_params(const long long *table, int n, long long total, cudaStream_t stream) {
  BC_REQUIRE(table && n > 0 && total > 0, BC_ERR_NULL, "bc_pack_params: empty table");
  long long grid = (total + 255) / 256;
  if (grid > (long long)kNumSMs * 8) grid = (long long)kNumSMs * 8;
  launch_kernel(pack_params_kernel, dim3((unsigned)grid), dim3(256), 0, stream, 1, table, n, total);
  return check_launch("bc_pack_params");
}

// ---------------------------------------------------------------------------------------------------
// RMSprop step of the online policy update (policy/policy.py:56-59 builds torch.optim.RMSprop; the reference steps
// it every block_train_interval frames, policy.py:361-362) for ALL parameter tensors in one launch:
//   g  = grad + weight_decay * p
//   sq = alpha * sq + (1 - alpha) * g * g;   avg = sqrt(sq) + eps
//   momentum > 0:  buf = momentum * buf + g / avg;  p -= lr * buf      else:  p -= lr * g / avg
// (torch's _single_tensor_rmsprop, centered = False).  table: n entries x 6 int64 = { param, grad, square_avg,
// momentum buffer (0 if unused), first flat element (cumulative), numel }; all tensors fp32 and dense.
constexpr int kRmsMaxEntries = 80;
struct RmsTable {
  long long e[kRmsMaxEntries][6];
};

__global__ void __launch_bounds__(256) rmsprop_kernel(const __grid_constant__ RmsTable table, int n, long long total, float lr,
                                                      float alpha, float eps, float wd, float momentum) {
  pdl_trigger();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {  // last entry whose first element <= i
      const int mid = (lo + hi + 1) >> 1;
      if (table.e[mid][4] <= i) lo = mid; else hi = mid - 1;
    }
    const long long *t = table.e[lo];
    const long long j = i - t[4];
    float *p = reinterpret_cast<float *>(t[0]) + j, *sq = reinterpret_cast<float *>(t[2]) + j;
    float g = reinterpret_cast<const float *>(t[1])[j];
    if (wd != 0.f) g = __fadd_rn(g, __fmul_rn(wd, *p));
    const float s = __fadd_rn(__fmul_rn(alpha, *sq), __fmul_rn(__fmul_rn(1.f - alpha, g), g));
    *sq = s;
    const float avg = __fadd_rn(__fsqrt_rn(s), eps);
    if (momentum > 0.f) {
      float *buf = reinterpret_cast<float *>(t[3]) + j;
      const float b = __fadd_rn(__fmul_rn(momentum, *buf), __fdiv_rn(g, avg));
      *buf = b;
      *p = __fsub_rn(*p, __fmul_rn(lr, b));
    } else {
      *p = __fsub_rn(*p, __fmul_rn(lr, __fdiv_rn(g, avg)));
    }
  }
}

// table: HOST array; it travels as a kernel parameter (<= 80 entries per launch, more are chunked), so a step needs
// no host-to-device copy and no synchronisation
int rmsprop_step(const long long *table, int n, long long total, float lr, float alpha, float eps, float wd, float momentum,
                 cudaStream_t stream) {
  BC_REQUIRE(table && n > 0 && total > 0, BC_ERR_NULL, "bc_rmsprop_step: empty table");
  for (int e0 = 0; e0 < n; e0 += kRmsMaxEntries) {
    const int m = n - e0 < kRmsMaxEntries ? n - e0 : kRmsMaxEntries;
    RmsTable t;
    const long long base = table[(size_t)e0 * 6 + 4];
    long long count = 0;
    for (int i = 0; i < m; ++i) {
      for (int k = 0; k < 6; ++k) t.e[i][k] = table[(size_t)(e0 + i) * 6 + k];
      t.e[i][4] -= base;
      BC_REQUIRE(t.e[i][4] == count && t.e[i][5] > 0, BC_ERR_RANGE, "bc_rmsprop_step: entry %d: offsets must be cumulative", e0 + i);
      count += t.e[i][5];
    }
    long long grid = (count + 255) / 256;
    if (grid > (long long)kNumSMs * 8) grid = (long long)kNumSMs * 8;
    launch_kernel(rmsprop_kernel, dim3((unsigned)grid), dim3(256), 0, stream, 1, t, m, count, lr, alpha, eps, wd, momentum);
  }
  (void)total;
  return check_launch("bc_rmsprop_step");
}

// ---------------------------------------------------------------------------------------------------
// Convolution with very few output channels on a dense NHWC fp16 tensor, fp32 weights and result: the policy
// net's last layer (128 -> 1, 3x3, stride 2: one logit per block, policy/net.py:46-50).  One CTA per output
// pixel, one thread per input channel, k*k taps each, block reduction per output channel.
struct FewOutParams {
  const __half *x;   // (N, H, W, Cx): the first C of Cx channels are used
  const float *w;    // (Cout, C, k, k) with element strides ws[4]
  const float *bias; // [Cout] or nullptr
  float *out;        // (N, Cout, Ho, Wo) contiguous
  int N, H, W, C, Cx, Cout, k, stride, pad, Ho, Wo;
  long long ws[4];
};

__global__ void __launch_bounds__(128) conv_fewout_kernel(const FewOutParams p) {
  __shared__ float red[4];
  pdl_trigger();
  pdl_wait();
  const int ox = blockIdx.x % p.Wo, oy = (blockIdx.x / p.Wo) % p.Ho, n = blockIdx.x / (p.Wo * p.Ho);
  for (int co = 0; co < p.Cout; ++co) {
    float acc = 0.f;
    for (int c = threadIdx.x; c < p.C; c += 128)
      for (int kh = 0; kh < p.k; ++kh) {
        const int y = oy * p.stride + kh - p.pad;
        if (y < 0 || y >= p.H) continue;
        for (int kw = 0; kw < p.k; ++kw) {
          const int x = ox * p.stride + kw - p.pad;
          if (x < 0 || x >= p.W) continue;
          acc = fmaf(__half2float(p.x[(((size_t)n * p.H + y) * p.W + x) * p.Cx + c]),
                     __ldg(p.w + co * p.ws[0] + c * p.ws[1] + kh * p.ws[2] + kw * p.ws[3]), acc);
        }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0)
      p.out[(((size_t)n * p.Cout + co) * p.Ho + oy) * p.Wo + ox] = red[0] + red[1] + red[2] + red[3] + (p.bias ? p.bias[co] : 0.f);
    __syncthreads();
  }
}

int conv_fewout(float *out, const void *x, const float *w, const float *bias, int N, int H, int W, int C, int Cx, int Cout,
                int k, int stride, int pad, const int64_t *w_strides, cudaStream_t stream) {
  BC_REQUIRE(out && x && w && w_strides, BC_ERR_NULL, "bc_conv_fewout: NULL pointer");
  BC_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && Cx >= C && Cout > 0 && Cout <= 16 && k > 0 && stride > 0 && pad >= 0,
             BC_ERR_SHAPE, "bc_conv_fewout: bad sizes (Cout <= 16)");
  FewOutParams p;
  p.x = (const __half *)x; p.w = w; p.bias = bias; p.out = out;
  p.N = N; p.H = H; p.W = W; p.C = C; p.Cx = Cx; p.Cout = Cout; p.k = k; p.stride = stride; p.pad = pad;
  p.Ho = (H + 2 * pad - k) / stride + 1;
  p.Wo = (W + 2 * pad - k) / stride + 1;
  BC_REQUIRE(p.Ho > 0 && p.Wo > 0 && (long long)N * p.Ho * p.Wo < (1ll << 31), BC_ERR_SHAPE, "bc_conv_fewout: output size");
  for (int i = 0; i < 4; ++i) p.ws[i] = w_strides[i];
  launch_kernel(conv_fewout_kernel, dim3((unsigned)(N * p.Ho * p.Wo)), dim3(128), 0, stream, 1, p);
  return check_launch("bc_conv_fewout");
}

}  // namespace bc

// =====================================================================================================
// bc_sample_grid -- Bernoulli draw of the execution grid + rounding of the executed-block count UP to a multiple,
// on the device (reference policy/policy.py:124-144 does the rounding on the host: D2H of the grid, Python
// random.sample over the skipped cells, index_put).  One CTA; the grid has 128 .. a few thousand cells.
//   exec0[g]  = uniforms[g] < probs[g]                      (what torch.bernoulli(probs) computes from its draw)
//   E0        = sum(exec0);  target = E0 == 0 ? 0 : multiple * (1 + (E0 - 1) / multiple)      (policy.py:139-140)
//   the (target - E0) skipped cells with the smallest (uniforms[G + g], g) are switched on: a uniformly random
//   subset of the skipped cells, like random.sample (policy.py:141-142), drawn from the caller's uniforms.
// =====================================================================================================
namespace bc {

constexpr int kSampleMaxCells = 8192;

__global__ void __launch_bounds__(1024) sample_grid_kernel(uint8_t *__restrict__ grid, int32_t *__restrict__ counts,
                                                           const float *__restrict__ probs, const float *__restrict__ uni,
                                                           int G, int multiple, int at_least_one) {
  __shared__ float key_s[kSampleMaxCells];   // tie-break key of a skipped cell, +inf for executed cells
  __shared__ int warp_cnt[32];
  __shared__ int total_s;
  pdl_trigger();
  pdl_wait();
  const int t = threadIdx.x;
  int mine = 0;
  for (int g = t; g < G; g += 1024) {
    const bool e = __ldg(uni + g) < __ldg(probs + g);
    key_s[g] = e ? __int_as_float(0x7f800000) : __ldg(uni + G + g);
    mine += e ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((t & 31) == 0) warp_cnt[t >> 5] = mine;
  __syncthreads();
  if (t < 32) {
    int v = warp_cnt[t];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (t == 0) {
      if (v == 0 && at_least_one && G > 0) {  // policy.py:262-263: grid[0,0,0,0] = 1
        key_s[0] = __int_as_float(0x7f800000);
        v = 1;
      }
      total_s = v;
    }
  }
  __syncthreads();
  const int E0 = total_s;
  int target = E0;
  if (multiple > 0) target = E0 == 0 ? 0 : multiple * (1 + (E0 - 1) / multiple);
  int need = target - E0;
  if (need > G - E0) need = G - E0;
  const float inf = __int_as_float(0x7f800000);
  for (int g = t; g < G; g += 1024) {
    const float k = key_s[g];
    bool e = k == inf;
    if (!e && need > 0) {
      int rank = 0;
      for (int j = 0; j < G; ++j) {  // broadcast reads: every thread of a warp reads the same shared word
        const float kj = key_s[j];
        rank += (kj < k || (kj == k && j < g)) ? 1 : 0;
      }
      e = rank < need;
    }
    grid[g] = e ? 1 : 0;
  }
  if (t == 0) {
    counts[0] = E0 + need;
    counts[1] = E0;
  }
}

int sample_grid(uint8_t *grid, int32_t *counts, const float *probs, const float *uniforms, int G, int multiple,
                int at_least_one, cudaStream_t stream) {
  BC_REQUIRE(grid && counts && probs && uniforms, BC_ERR_NULL, "bc_sample_grid: NULL pointer");
  BC_REQUIRE(G > 0 && G <= kSampleMaxCells, BC_ERR_UNSUPPORTED, "bc_sample_grid: %d cells (1..%d)", G, kSampleMaxCells);
  BC_REQUIRE(multiple >= 0, BC_ERR_RANGE, "bc_sample_grid: multiple %d", multiple);
  launch_kernel(sample_grid_kernel, dim3(1), dim3(1024), 0, stream, 1, grid, counts, probs, uniforms, G, multiple,
                at_least_one);
  return check_launch("bc_sample_grid");
}

}  // namespace bc

// =====================================================================================================
// bc_bn_update_running -- the train-mode side effect of BatchNorm2d (running_mean / running_var /
// num_batches_tracked, torch.nn.functional.batch_norm with training=True, as the reference's policy net runs every
// frame: policy/net.py:115-125, policy/resnet.py:60-115) for ALL batch norms of the fused trunk in one launch.
// Row l of the device table (8 x int64): { batch mean fp32*, batch invstd fp32*, running_mean fp32*,
// running_var fp32*, num_batches_tracked int64*, C, count = N*H*W, momentum as float bits (< 0: cumulative
// average, momentum=None) | eps as float bits << 32 }.  var_biased = 1/invstd^2 - eps; running_var takes the
// UNBIASED variance (count / (count - 1)), like torch.
// =====================================================================================================
namespace bc {

__global__ void __launch_bounds__(128) bn_update_running_kernel(const long long *__restrict__ table, int n) {
  pdl_trigger();
  pdl_wait();
  const int l = blockIdx.x;
  if (l >= n) return;
  const long long *row = table + (size_t)l * 8;
  const float *mean = reinterpret_cast<const float *>(row[0]);
  const float *invstd = reinterpret_cast<const float *>(row[1]);
  float *rmean = reinterpret_cast<float *>(row[2]);
  float *rvar = reinterpret_cast<float *>(row[3]);
  long long *nbt = reinterpret_cast<long long *>(row[4]);
  const int C = (int)row[5];
  const double count = (double)row[6];
  const float momentum = __int_as_float((int)(row[7] & 0xffffffffll));
  const float eps = __int_as_float((int)(row[7] >> 32));
  long long tracked = nbt ? *nbt + 1 : 1;
  const float f = momentum < 0.f ? (float)(1.0 / (double)tracked) : momentum;
  const float unbias = count > 1.0 ? (float)(count / (count - 1.0)) : 1.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float m = mean[c], is = invstd[c];
    const float var_b = 1.f / (is * is) - eps;
    if (rmean) rmean[c] = (1.f - f) * rmean[c] + f * m;
    if (rvar) rvar[c] = (1.f - f) * rvar[c] + f * (var_b * unbias);
  }
  __syncthreads();
  if (threadIdx.x == 0 && nbt) *nbt = tracked;
}

int bn_update_running(const long long *table, int n, cudaStream_t stream) {
  BC_REQUIRE(table != nullptr && n > 0, BC_ERR_NULL, "bc_bn_update_running: empty table");
  launch_kernel(bn_update_running_kernel, dim3((unsigned)n), dim3(128), 0, stream, 1, table, n);
  return check_launch("bc_bn_update_running");
}

}  // namespace bc

// =====================================================================================================
// bc_gn_stats -- GroupNorm statistics over ALL executed blocks (reference core/tensorwrapper.py:600-633,
// `_func_batched`: the (E, C, h, w) tile batch is folded into one (1, C, E*h*w, 1) sample, so a group's mean / variance
// run over the group's channels of every executed block).  x: packed NHWC fp16 tiles = (P pixels, C channels);
// output: per CHANNEL mean[c] / invstd[c] of the channel's group, which bc_ew_fused applies like a batch norm
// (y = weight * (x - mean) * invstd + bias).  One pass; per-CTA partials are added in CTA order by the last CTA
// (atomic ticket) in double precision: run-to-run reproducible.  C % 8 == 0, (C / groups) % 8 == 0, C <= 2048.
// =====================================================================================================
namespace bc {

constexpr int kGnThreads = 256;

struct GnParams {
  const __half *x;
  float *mean, *invstd;  // [C]
  double *partial;       // [gridDim.x][groups][2]
  unsigned int *ticket;
  uint32_t P;
  int C, groups;
  float eps;
  uint32_t zero;  // always 0 (load batching, see bn_stats_kernel)
};

// Loads are issued in batches of eight per thread and kept together by the data dependency bn_stats_kernel uses (see the
// note there); the last CTA adds the per-CTA partials of a (group, statistic) pair with one warp: lane l takes CTAs
// l, l+32, ... in order, then a fixed shuffle tree -- all loads of the final reduction are in flight at once.
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const GnParams p) {
  __shared__ __align__(8) float red[kGnThreads][2];
  __shared__ double fin[2 * 256];
  __shared__ bool last;
  pdl_trigger();
  pdl_wait();
  const int lanes = p.C >> 3;  // threads per pixel (8 channels each), <= 256
  const int sub = threadIdx.x % lanes, row = threadIdx.x / lanes, rows = kGnThreads / lanes;
  float s = 0.f, q = 0.f;
  if (row < rows) {
    const __half *base = p.x + sub * 8;
    const uint32_t step = gridDim.x * rows;
    for (uint32_t px = blockIdx.x * rows + row; px < p.P; px += 8 * step) {
      uint4 u[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t pj = px + j * step;
        u[j] = __ldg(reinterpret_cast<const uint4 *>(base + (size_t)(pj < p.P ? pj : p.P - 1) * p.C));
      }
      uint32_t fold = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) fold ^= u[j].x;
      fold &= p.zero;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        u[j].x ^= fold; u[j].y ^= fold; u[j].z ^= fold; u[j].w ^= fold;
        if (px + j * step >= p.P) u[j] = make_uint4(0, 0, 0, 0);
        const __half2 *h = reinterpret_cast<const __half2 *>(&u[j]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(h[k]);
          s += f.x + f.y;
          q = fmaf(f.x, f.x, fmaf(f.y, f.y, q));
        }
      }
    }
  }
  red[threadIdx.x][0] = s;
  red[threadIdx.x][1] = q;
  __syncthreads();
  const int lpg = (p.C / p.groups) >> 3;  // lanes (8-channel chunks) per group
  if (threadIdx.x < p.groups) {           // group g: its lanes of every pixel row of the CTA, in a fixed order
    double ts = 0.0, tq = 0.0;
    for (int r = 0; r < rows; ++r)
      for (int l = 0; l < lpg; ++l) {
        const int t = r * lanes + threadIdx.x * lpg + l;
        ts += (double)red[t][0];
        tq += (double)red[t][1];
      }
    p.partial[((size_t)blockIdx.x * p.groups + threadIdx.x) * 2] = ts;
    p.partial[((size_t)blockIdx.x * p.groups + threadIdx.x) * 2 + 1] = tq;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  // (group, statistic) pair -> `slices` threads, each adds every slices-th CTA partial (eight loads in flight), then the
  // slices are added in order: all 256 threads work, the sum order is fixed
  {
    const int pairs = 2 * p.groups;  // <= 512
    double *comb = reinterpret_cast<double *>(red);  // 256 x 2 floats = 256 doubles, free by now
    for (int base = 0; base < pairs; base += kGnThreads) {
      const int np = pairs - base < kGnThreads ? pairs - base : kGnThreads;  // pairs handled in this round
      int slices = 1;
      while (slices * 2 * np <= kGnThreads) slices *= 2;
      const int pr = threadIdx.x % np, slice = threadIdx.x / np;
      double t = 0.0;
      if (slice < slices) {
        for (unsigned b0 = (unsigned)slice; b0 < gridDim.x; b0 += 8u * slices) {
          double v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const unsigned b = b0 + (unsigned)(j * slices);
            v[j] = b < gridDim.x ? __ldcg(p.partial + (size_t)b * pairs + base + pr) : 0.0;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) t += v[j];
        }
      }
      __syncthreads();
      comb[threadIdx.x] = t;
      __syncthreads();
      if (threadIdx.x < np) {
        double sum = 0.0;
        for (int z = 0; z < slices; ++z) sum += comb[z * np + threadIdx.x];
        fin[base + threadIdx.x] = sum;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < p.groups) {
    const double n = (double)p.P * (double)(p.C / p.groups);
    const double m = fin[2 * threadIdx.x] / n;
    double var = fin[2 * threadIdx.x + 1] / n - m * m;
    var = var < 0.0 ? 0.0 : var;
    fin[2 * threadIdx.x] = m;
    fin[2 * threadIdx.x + 1] = 1.0 / sqrt(var + (double)p.eps);
  }
  __syncthreads();
  const int cpg = p.C / p.groups;
  for (int c = threadIdx.x; c < p.C; c += kGnThreads) {
    p.mean[c] = (float)fin[2 * (c / cpg)];
    p.invstd[c] = (float)fin[2 * (c / cpg) + 1];
  }
  if (threadIdx.x == 0) *p.ticket = 0u;
}

int gn_stats(float *mean, float *invstd, const void *x, long long P, int C, int groups, float eps, void *workspace,
             long long workspace_bytes, cudaStream_t stream) {
  BC_REQUIRE(mean && invstd && x && workspace, BC_ERR_NULL, "bc_gn_stats: NULL pointer");
  BC_REQUIRE(P > 0 && P < (1ll << 31), BC_ERR_SHAPE, "bc_gn_stats: %lld pixels", P);
  BC_REQUIRE(C >= 8 && C <= 2048 && C % 8 == 0 && groups >= 1 && groups <= 256 && C % groups == 0 && (C / groups) % 8 == 0,
             BC_ERR_UNSUPPORTED, "bc_gn_stats: C=%d in %d groups (C and C/groups multiples of 8, C <= 2048, <= 256 groups)",
             C, groups);
  BC_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)workspace & 15) == 0, BC_ERR_ALIGN, "bc_gn_stats: 16-byte alignment");
  const int rows = kGnThreads / (C / 8);
  long long grid = (P + 8 * rows - 1) / (8 * rows);
  if (grid > 2 * kNumSMs) grid = 2 * kNumSMs;  // two CTAs per SM: 16 loads in flight per thread pair
  const long long need = 16 + grid * groups * 2 * (long long)sizeof(double);
  BC_REQUIRE(workspace_bytes >= need, BC_ERR_RANGE, "bc_gn_stats: workspace of %lld bytes, %lld needed", workspace_bytes, need);
  GnParams p;
  p.x = (const __half *)x; p.mean = mean; p.invstd = invstd;
  p.ticket = (unsigned int *)workspace;
  p.partial = (double *)((char *)workspace + 16);
  p.P = (uint32_t)P; p.C = C; p.groups = groups; p.eps = eps; p.zero = 0u;
  launch_kernel(gn_stats_kernel, dim3((unsigned)grid), dim3(kGnThreads), 0, stream, 1, p);
  return check_launch("bc_gn_stats");
}

}  // namespace bc
