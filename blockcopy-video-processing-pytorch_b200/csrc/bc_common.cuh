// bc_common.cuh -- shared host/device helpers of libblockcopy_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <tuple>
#include <type_traits>
#include <utility>

#include "blockcopy_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libblockcopy_sm100 is written for sm_100a only"
#endif

namespace bc {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of it

// ------------------------------------------------------------------ errors (thread-local text)
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);
int check_launch(const char *what);  // cudaPeekAtLastError -> positive code + message

#define BC_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) return ::bc::fail((code), __VA_ARGS__); \
  } while (0)

inline int elem_size(int dtype) { return dtype == BC_F16 ? 2 : dtype == BC_F32 ? 4 : 0; }

// ------------------------------------------------------------------ exact division by a runtime constant
// q = n / d for 0 <= n < 2^31, 1 <= d < 2^31, as one mul.hi + shift (Granlund-Montgomery
// round-up variant).  All the block kernels decode a flat index into (tile, row, pixel,
// cell) coordinates; with runtime shapes that would otherwise be ~20-instruction divides.
struct FastDiv {
  uint32_t d, mul, shr;
  FastDiv() : d(1), mul(0), shr(0) {}
  explicit FastDiv(uint32_t div) : d(div), mul(0), shr(0) {
    if (div > 1) {
      uint32_t lg = 0;
      while ((1ull << lg) < div) ++lg;  // ceil(log2(d))
      const uint32_t p = 31 + lg;
      mul = (uint32_t)(((1ull << p) + div - 1) / div);
      shr = p - 32;
    }
  }
  __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
    return d == 1 ? n : (__umulhi(n, mul) >> shr);
#else
    return d == 1 ? n : (uint32_t)(((uint64_t)n * mul) >> 32) >> shr;
#endif
  }
  __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t &q, uint32_t &r) const {
    q = div(n);
    r = n - q * d;
  }
};

// flat cell id -> (n, gh, gw)
struct CellDecode {
  FastDiv per_image, per_row;  // GH*GW, GW
  CellDecode() {}
  CellDecode(int GH, int GW) : per_image((uint32_t)(GH * GW)), per_row((uint32_t)GW) {}
  __device__ __forceinline__ void operator()(uint32_t g, uint32_t &n, uint32_t &gh, uint32_t &gw) const {
    uint32_t r;
    per_image.divmod(g, n, r);
    per_row.divmod(r, gh, gw);
  }
};

// ------------------------------------------------------------------ vector access of V bytes
template <int V> struct Vec;
template <> struct Vec<16> { using T = uint4; };
template <> struct Vec<8> { using T = uint2; };
template <> struct Vec<4> { using T = uint32_t; };
template <> struct Vec<2> { using T = uint16_t; };

template <int V> __device__ __forceinline__ typename Vec<V>::T zero_vec();
template <> __device__ __forceinline__ uint4 zero_vec<16>() { return make_uint4(0, 0, 0, 0); }
template <> __device__ __forceinline__ uint2 zero_vec<8>() { return make_uint2(0, 0); }
template <> __device__ __forceinline__ uint32_t zero_vec<4>() { return 0u; }
template <> __device__ __forceinline__ uint16_t zero_vec<2>() { return (uint16_t)0; }

// streaming read through the read-only path
template <int V> __device__ __forceinline__ typename Vec<V>::T ld_stream(const char *p) {
  return __ldg(reinterpret_cast<const typename Vec<V>::T *>(p));
}
template <int V> __device__ __forceinline__ void st_vec(char *p, const typename Vec<V>::T &v) {
  *reinterpret_cast<typename Vec<V>::T *>(p) = v;
}

// ------------------------------------------------------------------ programmatic dependent launch
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization: the
// next kernel's CTAs may be scheduled (and run their prologue: barrier init, TMEM alloc, descriptor
// prefetch) while this one drains.  Contract inside every kernel: pdl_trigger() first, then pdl_wait()
// BEFORE the first global-memory access that could depend on (or be read by) an earlier kernel.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();  // BC_PDL=0 disables the launch attribute (bc_api.cu)

// ------------------------------------------------------------------ CUDA-graph node recording / patching (bc_api.cu)
// The Python host replays one captured graph per frame; what differs between frames is a handful of POINTERS (the
// caller's frame, its grid, the output buffer).  Instead of launching those kernels eagerly in front of the graph
// (a launch gap each), the host captures them too and re-points their nodes before every replay:
//   mode 1 (record): after a launch on a capturing stream, remember the graph node the launch created;
//   mode 2 (patch, one shot): the next launch does not run -- its grid / block / arguments are written into `node` of
//   the instantiated graph `exec` (cudaGraphExecKernelNodeSetParams).
struct GraphHook {
  int mode = 0;
  cudaGraphExec_t exec = nullptr;
  cudaGraphNode_t node = nullptr;
  const void *func = nullptr;  // record: kernel of the recorded node; patch: kernel the node was captured with (checked)
  int recorded = 0;
  cudaError_t patch_error = cudaSuccess;  // result of the last node update, reported by check_launch()
};
GraphHook &graph_hook();
cudaError_t graph_record_after_launch(GraphHook &h, const void *func, cudaStream_t stream);

template <typename Tuple, size_t... I>
inline void fill_arg_pointers(Tuple &t, void **argv, std::index_sequence<I...>) {
  ((argv[I] = (void *)&std::get<I>(t)), ...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                         dim3 cluster, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster.x * cluster.y * cluster.z > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster.x;
    at[n].val.clusterDim.y = cluster.y;
    at[n].val.clusterDim.z = cluster.z;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  GraphHook &h = graph_hook();
  if (h.mode == 2) {  // re-point an existing graph node instead of launching
    h.mode = 0;
    if (h.func != (const void *)kernel) {  // another kernel variant than the captured one
      h.patch_error = cudaErrorInvalidDeviceFunction;
      return h.patch_error;
    }
    std::tuple<std::remove_cv_t<std::remove_reference_t<KArgs>>...> copies{static_cast<KArgs>(args)...};
    void *argv[sizeof...(KArgs) > 0 ? sizeof...(KArgs) : 1];
    fill_arg_pointers(copies, argv, std::index_sequence_for<KArgs...>{});
    cudaKernelNodeParams np = {};
    np.func = (void *)kernel;
    np.gridDim = grid;
    np.blockDim = block;
    np.sharedMemBytes = (unsigned int)smem;
    np.kernelParams = argv;
    np.extra = nullptr;
    h.patch_error = cudaGraphExecKernelNodeSetParams(h.exec, h.node, &np);
    return h.patch_error;
  }
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (h.mode == 1 && e == cudaSuccess) return graph_record_after_launch(h, (const void *)kernel, stream);
  return e;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 unsigned cluster_z, Args &&...args) {
  return launch_kernel_cluster(kernel, grid, block, smem, stream, dim3(1, 1, cluster_z ? cluster_z : 1),
                               static_cast<Args &&>(args)...);
}

inline int gcd_pow2_bytes(uint64_t a) {  // largest power of two <= 16 dividing a (a > 0)
  int v = 16;
  while (v > 1 && (a % (uint64_t)v)) v >>= 1;
  return v;
}

}  // namespace bc
