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
    float v;
    if (c < 6) {
      const T *src = reinterpret_cast<const T *>(c < 3 ? p.frame : p.state);
      const int cc = c < 3 ? c : c - 3;
      const int yy = nearest_src(p.sy_frame, y, p.H), xx = nearest_src(p.sx_frame, x, p.W);
      v = to_f<T>(src[(((size_t)n * 3 + cc) * p.H + yy) * p.W + xx]);
    } else if (c < 6 + p.K) {
      const int yy = nearest_src(p.sy_repr, y, p.h), xx = nearest_src(p.sx_repr, x, p.w);
      v = to_f<T>(reinterpret_cast<const T *>(p.repr)[n * p.repr_sn + (c - 6) * p.repr_sc + yy * p.repr_sh + xx * p.repr_sw]) - 0.5f;
    } else {
      const int yy = nearest_src(p.sy_grid, y, p.GH), xx = nearest_src(p.sx_grid, x, p.GW);
      v = (p.grid[((size_t)n * p.GH + yy) * p.GW + xx] ? 1.f : 0.f) - 0.5f;
    }
    p.out[i] = v;
  }
}

int policy_features(float *out, const void *frame, const void *state, const void *repr, const uint8_t *grid, int N,
                    int K, int H, int W, int h, int w, int GH, int GW, int Ho, int Wo, const int64_t *repr_strides,
                    float sy_frame, float sx_frame, int dtype, cudaStream_t stream) {
  BC_REQUIRE(out && frame && state && repr && grid && repr_strides, BC_ERR_NULL, "bc_policy_features: NULL pointer");
  BC_REQUIRE(N > 0 && K > 0 && Ho > 0 && Wo > 0, BC_ERR_SHAPE, "bc_policy_features: empty problem");
  BC_REQUIRE(dtype == BC_F16 || dtype == BC_F32, BC_ERR_DTYPE, "bc_policy_features: dtype");
  FeatParams p;
  p.frame = frame; p.state = state; p.repr = repr; p.grid = grid; p.out = out;
  p.N = N; p.K = K; p.H = H; p.W = W; p.h = h; p.w = w; p.GH = GH; p.GW = GW; p.Ho = Ho; p.Wo = Wo;
  p.repr_sn = repr_strides[0]; p.repr_sc = repr_strides[1]; p.repr_sh = repr_strides[2]; p.repr_sw = repr_strides[3];
  p.sy_frame = sy_frame; p.sx_frame = sx_frame;        // = 1 / scale_factor, as ATen uses for scale_factor= calls
  p.sy_repr = (float)h / Ho; p.sx_repr = (float)w / Wo;  // size= calls: in / out
  p.sy_grid = (float)GH / Ho; p.sx_grid = (float)GW / Wo;
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

}  // namespace bc
