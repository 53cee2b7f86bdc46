// bc_api.cu -- extern "C" entry points of libblockcopy_sm100.so (include/blockcopy_b200.h):
// argument validation, SIMT/TMA path selection, error text.  No torch headers anywhere.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "bc_move.cuh"
#include "bc_tma.cuh"

namespace bc {

static thread_local char g_err[512] = "no error";
static std::atomic<int> g_tma_enabled{1};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char *what) {
  GraphHook &h = graph_hook();
  if (h.patch_error != cudaSuccess) {  // the launch was a graph-node update (bc_graph_patch_next) and it failed
    const cudaError_t pe = h.patch_error;
    h.patch_error = cudaSuccess;
    set_error("%s: graph node update failed: %s (%s)", what, cudaGetErrorName(pe), cudaGetErrorString(pe));
    return (int)pe;
  }
  const cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess) return BC_OK;
  cudaGetLastError();  // clear the sticky launch-configuration error
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return (int)e;
}

int launch_compact_mask(const uint8_t *grid, int G, int32_t *grid_idx, int32_t *mapping_exec, int32_t *counts,
                        const int32_t *prev_grid_idx, int32_t *transfer_idx, cudaStream_t s);

int conv_igemm(void *out, const void *plane, const void *weight, const void *bias, const void *residual,
               const int32_t *mapping, int E, int N, int Cin, int H, int W, int BS_in, int Cout, int ksize, int stride,
               int pad, int relu, void *plane_out, const int32_t *out_mapping, int out_N, int out_GH, int out_GW,
               int allow_split_k, void *workspace, long long workspace_bytes, cudaStream_t stream);
int ew_fused(void *out, void *plane, const void *a, const void *residual, const float *mean, const float *invstd,
             const float *weight, const float *shift, const int32_t *mapping, int E, int C, int BS, int N, int H,
             int W, int up2x, int relu, cudaStream_t stream);

int maxpool_halo(void *out, void *plane_out, const void *plane, const int32_t *mapping, int E, int N, int C, int H,
                 int W, int BS_in, int k, int stride, int pad, cudaStream_t stream);

int stem_pack(void *plane, const void *tiles, const int32_t *mapping, int E, int N, int H, int W, int BS,
              cudaStream_t stream);
int conv_stem(void *out, const void *s2d_plane, const void *weight, const void *bias, const int32_t *mapping, int E,
              int N, int Hs, int Ws, int BS_out, int Cout, int relu, void *plane_out, cudaStream_t stream);

int policy_features(float *out, const void *frame, const void *state, const void *repr, const uint8_t *grid, int N,
                    int K, int H, int W, int h, int w, int GH, int GW, int Ho, int Wo, const int64_t *repr_strides,
                    float sy_frame, float sx_frame, int dtype, cudaStream_t stream);
int policy_features_nhwc16(void *out, int Cp, const void *frame, const void *state, const void *repr, const uint8_t *grid,
                           int N, int K, int H, int W, int h, int w, int GH, int GW, int Ho, int Wo,
                           const int64_t *repr_strides, float sy_frame, float sx_frame, int dtype, cudaStream_t stream);
int info_gain(void *out, const void *cur, const void *prev, int N, int K, int h, int w, const int64_t *strides,
              cudaStream_t stream);
int sample_grid(uint8_t *grid, int32_t *counts, const float *probs, const float *uniforms, int G, int multiple,
                int at_least_one, cudaStream_t stream);
int bn_update_running(const long long *table, int n, cudaStream_t stream);
int gn_stats(float *mean, float *invstd, const void *x, long long P, int C, int groups, float eps, void *workspace,
             long long workspace_bytes, cudaStream_t stream);
int depth_to_space(void *out, const void *in, int E, int C, int h, int w, int r, cudaStream_t stream);
int bwd_mask_add(void *dst, const void *grad, const void *out, const void *add, long long n, cudaStream_t stream);
int bn_bwd_reduce(float *sums, const void *g, const void *out, const void *z, const float *mean, const float *invstd,
                  long long P, int C, void *workspace, long long workspace_bytes, cudaStream_t stream);
int bn_bwd_apply(void *dz, void *dz_up, const void *g, const void *out, const void *z, const float *mean, const float *invstd,
                 const float *gamma, const float *sums, int N, int H, int W, int C, cudaStream_t stream);
int conv_wgrad(float *grad_w, const long long *grad_strides, const void *dz, const void *x, int N, int H, int W, int Cin_p,
               int Cout_p, int Cin, int Cout, int ksize, int stride, const float *inv_scale, float *dgamma, float *dbeta,
               const float *bn_sums, void *workspace, long long workspace_bytes, cudaStream_t stream);
int blocks_from_u8(void *tiles, const uint8_t *src, const float *mean, const float *std, const int32_t *mapping, int E, int N,
                   int H, int W, int BS, int dtype, cudaStream_t stream);
int bn_norm(void *out, float *mean, float *invstd, const void *x, const float *weight, const float *shift, long long P, int C,
            float eps, int relu, void *workspace, long long workspace_bytes, cudaStream_t stream);
int raster_boxes(float *out, const int32_t *rects, const float *values, int n, int H, int W, int shift, cudaStream_t stream);

int bn_stats(float *mean, float *invstd, const void *x, long long P, int C, float eps, void *workspace,
             long long workspace_bytes, cudaStream_t stream);
int pack_params(const long long *table, int n, long long total, cudaStream_t stream);
int rmsprop_step(const long long *table, int n, long long total, float lr, float alpha, float eps, float wd, float momentum,
                 cudaStream_t stream);
int conv_fewout(float *out, const void *x, const float *w, const float *bias, int N, int H, int W, int C, int Cx, int Cout,
                int k, int stride, int pad, const int64_t *w_strides, cudaStream_t stream);
int frame_from_u8(void *out, const uint8_t *src, const float *mean, const float *std, int N, int H, int W, int dtype,
                  cudaStream_t stream);
int upsample_argmax(void *labels, const void *logits, int N, int K, int h, int w, const int64_t *strides, int scale,
                    int dtype, int label_bytes, cudaStream_t stream, const uint8_t *grid = nullptr, int GH = 0, int GW = 0);

int spp_pool(void *pooled, const void *x0, int N, int C, int H, int W, int L, const int *gh, const int *gw, cudaStream_t s);
int spp_levels(void *out, const void *pooled, const float *bn, const void *w, int N, int C, int H, int W, int L,
               const int *gh, const int *gw, int Lc, cudaStream_t s);
int spp_prep(void *y, const void *x0, const void *lev, const float *bn, int N, int C, int H, int W, int L, const int *gh,
             const int *gw, int Lc, int Cp, cudaStream_t s);

int head_1x1(void *tiles_out, void *dense_out, const void *dense_prev, const void *tiles_in, const void *weight,
             const void *bias, const float *bn_mean, const float *bn_invstd, const float *bn_weight,
             const float *bn_shift, int relu_in, const int32_t *grid_idx, const int32_t *mapping, int E, int N, int GH,
             int GW, int BS, int Cin, int Cout, int tiles_layout, int dense_layout, cudaStream_t stream);

GraphHook &graph_hook() {
  static thread_local GraphHook h;
  return h;
}

cudaError_t graph_record_after_launch(GraphHook &h, const void *func, cudaStream_t stream) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  unsigned long long id = 0;
  cudaGraph_t g = nullptr;
  const cudaGraphNode_t *deps = nullptr;
  const cudaGraphEdgeData *edges = nullptr;
  size_t n = 0;
  const cudaError_t e = cudaStreamGetCaptureInfo_v3(stream, &st, &id, &g, &deps, &edges, &n);
  if (e != cudaSuccess) return e;
  if (st == cudaStreamCaptureStatusActive && n >= 1) {  // the stream's only dependency now is the node just created
    h.node = deps[n - 1];
    h.func = func;
    ++h.recorded;
  }
  return cudaSuccess;
}

bool pdl_enabled() {
  // programmatic dependent launch: the next kernel's prologue (barrier init, TMEM alloc, descriptor prefetch)
  // overlaps this one's tail; ~1.3 % of a SwiftNet frame inside a CUDA graph.  BC_PDL=0 switches it off.
  static const bool on = !(getenv("BC_PDL") && getenv("BC_PDL")[0] == '0');
  return on;
}

static inline bool tma_on() { return g_tma_enabled.load(std::memory_order_relaxed) != 0; }

static std::atomic<unsigned long long *> g_trace{nullptr};
unsigned long long *debug_trace_buffer() { return g_trace.load(); }
}  // namespace bc

using namespace bc;

extern "C" {

BC_API int bc_version(void) { return BC_ABI_VERSION; }

BC_API const char *bc_last_error_string(void) { return g_err; }

BC_API const char *bc_build_info(void) {
  return "libblockcopy_sm100 abi 1 | sm_100a | built " __DATE__ " " __TIME__;
}

BC_API int bc_set_tma_enabled(int enabled) {
  g_tma_enabled.store(enabled ? 1 : 0);
  return BC_OK;
}

BC_API int bc_debug_trace(void *device_buffer) {
  bc::g_trace.store((unsigned long long *)device_buffer);
  return BC_OK;
}

BC_API int bc_compact_mask(const uint8_t *grid, int G, int32_t *grid_idx, int32_t *mapping_exec, int32_t *counts,
                    const int32_t *prev_grid_idx, int32_t *transfer_idx, bc_stream_t stream) {
  BC_REQUIRE(grid && grid_idx && mapping_exec && counts, BC_ERR_NULL, "bc_compact_mask: NULL pointer");
  BC_REQUIRE(G > 0, BC_ERR_SHAPE, "bc_compact_mask: G=%d", G);
  BC_REQUIRE((transfer_idx == nullptr) || (prev_grid_idx != nullptr), BC_ERR_NULL,
             "bc_compact_mask: transfer_idx requested without prev_grid_idx");
  BC_REQUIRE(prev_grid_idx != grid_idx, BC_ERR_UNSUPPORTED, "bc_compact_mask: prev_grid_idx must not alias grid_idx");
  return launch_compact_mask(grid, G, grid_idx, mapping_exec, counts, prev_grid_idx, transfer_idx,
                             (cudaStream_t)stream);
}

BC_API int bc_gather(void *blocks, const void *image, const int32_t *mapping_exec, int E, int N, int C, int H, int W, int BS,
              bc_dtype_t dtype, bc_layout_t layout, bc_stream_t stream) {
  BC_REQUIRE(E >= 0, BC_ERR_SHAPE, "bc_gather: E=%d", E);
  if (E == 0) return BC_OK;  // reference: no launch when B == 0 (block_funcs.py:33)
  BC_REQUIRE(blocks && image && mapping_exec, BC_ERR_NULL, "bc_gather: NULL pointer");
  const int es = elem_size(dtype);
  MoveGeo g;
  const void *ptrs[2] = {blocks, image};
  int rc = make_geo(g, E, N, C, H, W, BS, BS, 0, es, layout, false, ptrs, 2);
  if (rc != BC_OK) return rc;
  if (tma_on() && tma_move_eligible(blocks, image, E, C, W, BS, BS, es, layout))
    return launch_tma_move(blocks, const_cast<void *>(image), mapping_exec, E, N, C, H, W, BS, 0, BS, es, layout,
                           false, (cudaStream_t)stream);
  return launch_gather_simt(blocks, image, mapping_exec, g, false, (cudaStream_t)stream);
}

BC_API int bc_scatter(const void *blocks, void *image, const int32_t *mapping_exec, int E, int N, int C, int H, int W,
               int BS, bc_dtype_t dtype, bc_layout_t layout, bc_stream_t stream) {
  BC_REQUIRE(E >= 0, BC_ERR_SHAPE, "bc_scatter: E=%d", E);
  if (E == 0) return BC_OK;
  BC_REQUIRE(blocks && image && mapping_exec, BC_ERR_NULL, "bc_scatter: NULL pointer");
  const int es = elem_size(dtype);
  MoveGeo g;
  const void *ptrs[2] = {blocks, image};
  int rc = make_geo(g, E, N, C, H, W, BS, BS, 0, es, layout, false, ptrs, 2);
  if (rc != BC_OK) return rc;
  if (tma_on() && tma_move_eligible(blocks, image, E, C, W, BS, BS, es, layout))
    return launch_tma_move(const_cast<void *>(blocks), image, mapping_exec, E, N, C, H, W, BS, 0, BS, es, layout, true,
                           (cudaStream_t)stream);
  return launch_scatter_simt(blocks, image, mapping_exec, g, (cudaStream_t)stream);
}

BC_API int bc_copy_blocks(void *out, const void *prev, const void *blocks, const int32_t *grid_idx, int N, int C, int H,
                   int W, int BS, bc_dtype_t dtype, bc_layout_t layout, bc_stream_t stream) {
  BC_REQUIRE(out && prev && grid_idx, BC_ERR_NULL, "bc_copy_blocks: NULL pointer");
  BC_REQUIRE(out != prev, BC_ERR_UNSUPPORTED, "bc_copy_blocks: out must not alias prev (use bc_scatter in place)");
  BC_REQUIRE(BS > 0 && H > 0 && W > 0 && H % BS == 0 && W % BS == 0, BC_ERR_SHAPE,
             "bc_copy_blocks: plane %dx%d / block %d", H, W, BS);
  const int es = elem_size(dtype);
  const int G = N * (H / BS) * (W / BS);
  MoveGeo g;
  const void *ptrs[3] = {out, prev, blocks ? blocks : out};
  int rc = make_geo(g, G, N, C, H, W, BS, BS, 0, es, layout, false, ptrs, 3);
  if (rc != BC_OK) return rc;
  return launch_copy_blocks_simt(out, prev, blocks, grid_idx, g, (cudaStream_t)stream);
}

BC_API int bc_transfer(void *out, const void *prev_exec, const void *prev_transfer, const int32_t *transfer_idx, int T, int G,
                int C, int BS, int padding, bc_dtype_t dtype, bc_layout_t layout, bc_stream_t stream) {
  BC_REQUIRE(T >= 0, BC_ERR_SHAPE, "bc_transfer: T=%d", T);
  if (T == 0) return BC_OK;
  BC_REQUIRE(out && prev_exec && transfer_idx, BC_ERR_NULL, "bc_transfer: NULL pointer");
  BC_REQUIRE(G > 0, BC_ERR_SHAPE, "bc_transfer: G=%d", G);
  const int es = elem_size(dtype);
  MoveGeo g;
  // prev_transfer may legitimately be an empty tensor on the second frame (all blocks executed on the first)
  const void *ptrs[3] = {out, prev_exec, prev_transfer ? prev_transfer : prev_exec};
  // tile-to-tile: the "plane" extent is irrelevant; describe a 1 x 1 cell grid of edge BS
  int rc = make_geo(g, T, 1, C, BS, BS, BS, BS, padding, es, layout, true, ptrs, 3);
  if (rc != BC_OK) return rc;
  return launch_transfer_simt(out, prev_exec, prev_transfer, transfer_idx, g, G, (cudaStream_t)stream);
}

BC_API int bc_gather_halo_tiles(void *out, const void *exec, const void *transfer, const int32_t *grid_idx,
                         const int32_t *mapping_exec, int E, int N, int C, int GH, int GW, int BS, int pad,
                         bc_dtype_t dtype, bc_layout_t layout, bc_stream_t stream) {
  BC_REQUIRE(E >= 0, BC_ERR_SHAPE, "bc_gather_halo_tiles: E=%d", E);
  if (E == 0) return BC_OK;
  BC_REQUIRE(out && exec && grid_idx && mapping_exec, BC_ERR_NULL, "bc_gather_halo_tiles: NULL pointer");
  BC_REQUIRE(pad > 0, BC_ERR_SHAPE, "bc_gather_halo_tiles: pad must be > 0 (blockpad.py:32), got %d", pad);
  BC_REQUIRE(GH > 0 && GW > 0 && N > 0, BC_ERR_SHAPE, "bc_gather_halo_tiles: grid %dx%dx%d", N, GH, GW);
  const int es = elem_size(dtype);
  MoveGeo g, src;
  const void *ptrs[3] = {out, exec, transfer ? transfer : exec};
  int rc = make_geo(g, E, N, C, GH * BS, GW * BS, BS, BS + 2 * pad, pad, es, layout, true, ptrs, 3);
  if (rc != BC_OK) return rc;
  rc = make_geo(src, E, N, C, GH * BS, GW * BS, BS, BS, 0, es, layout, true, ptrs, 3);
  if (rc != BC_OK) return rc;
  return launch_halo_tiles_simt(out, exec, transfer, grid_idx, mapping_exec, g, src, N * GH * GW,
                                (cudaStream_t)stream);
}

BC_API int bc_gather_halo(void *out, const void *plane, const int32_t *mapping_exec, int E, int N, int C, int H, int W,
                   int BS, int pad, bc_dtype_t dtype, bc_layout_t layout, bc_stream_t stream) {
  BC_REQUIRE(E >= 0, BC_ERR_SHAPE, "bc_gather_halo: E=%d", E);
  if (E == 0) return BC_OK;
  BC_REQUIRE(out && plane && mapping_exec, BC_ERR_NULL, "bc_gather_halo: NULL pointer");
  BC_REQUIRE(pad >= 0, BC_ERR_SHAPE, "bc_gather_halo: pad=%d", pad);
  const int es = elem_size(dtype);
  MoveGeo g;
  const void *ptrs[2] = {out, plane};
  int rc = make_geo(g, E, N, C, H, W, BS, BS + 2 * pad, pad, es, layout, true, ptrs, 2);
  if (rc != BC_OK) return rc;
  if (tma_on() && tma_move_eligible(out, plane, E, C, W, BS, BS + 2 * pad, es, layout))
    return launch_tma_move(out, const_cast<void *>(plane), mapping_exec, E, N, C, H, W, BS, pad, BS + 2 * pad, es,
                           layout, false, (cudaStream_t)stream);
  if (layout == BC_NCHW && pad > 0 && gather_halo_nchw_eligible(out, plane, BS, pad, W, es))
    return launch_gather_halo_nchw(out, plane, mapping_exec, g, E, es, (cudaStream_t)stream);
  return launch_gather_simt(out, plane, mapping_exec, g, true, (cudaStream_t)stream);
}

BC_API int bc_conv_igemm(void *out, const void *plane, const void *weight, const void *bias, const void *residual,
                         const int32_t *mapping_exec, int E, int N, int Cin, int H, int W, int BS_in, int Cout,
                         int ksize, int stride, int pad, int relu, void *plane_out, const int32_t *out_mapping,
                         int out_N, int out_GH, int out_GW, int allow_split_k, void *workspace,
                         long long workspace_bytes, bc_stream_t stream) {
  BC_REQUIRE(E >= 0, BC_ERR_SHAPE, "bc_conv_igemm: E=%d", E);
  if (E == 0) return BC_OK;
  return conv_igemm(out, plane, weight, bias, residual, mapping_exec, E, N, Cin, H, W, BS_in, Cout, ksize, stride,
                    pad, relu, plane_out, out_mapping, out_N, out_GH, out_GW, allow_split_k, workspace, workspace_bytes,
                    (cudaStream_t)stream);
}

BC_API int bc_head_1x1(void *tiles_out, void *dense_out, const void *dense_prev, const void *tiles_in,
                       const void *weight, const void *bias, const float *bn_mean, const float *bn_invstd,
                       const float *bn_weight, const float *bn_shift, int relu_in, const int32_t *grid_idx,
                       const int32_t *mapping_exec, int E, int N, int GH, int GW, int BS, int Cin, int Cout,
                       bc_layout_t tiles_layout, bc_layout_t dense_layout, bc_stream_t stream) {
  BC_REQUIRE(E >= 0, BC_ERR_SHAPE, "bc_head_1x1: E=%d", E);
  if (E == 0) return BC_OK;
  return head_1x1(tiles_out, dense_out, dense_prev, tiles_in, weight, bias, bn_mean, bn_invstd, bn_weight, bn_shift,
                  relu_in, grid_idx, mapping_exec, E, N, GH, GW, BS, Cin, Cout, (int)tiles_layout, (int)dense_layout,
                  (cudaStream_t)stream);
}

BC_API int bc_ew_fused(void *out, void *plane_out, const void *a, const void *residual, const float *bn_mean,
                       const float *bn_invstd, const float *bn_weight, const float *bn_shift,
                       const int32_t *mapping_exec, int E, int C, int BS, int N, int H, int W, int up2x, int relu,
                       bc_stream_t stream) {
  BC_REQUIRE(E >= 0, BC_ERR_SHAPE, "bc_ew_fused: E=%d", E);
  if (E == 0) return BC_OK;
  return ew_fused(out, plane_out, a, residual, bn_mean, bn_invstd, bn_weight, bn_shift, mapping_exec, E, C, BS, N, H,
                  W, up2x, relu, (cudaStream_t)stream);
}

BC_API int bc_maxpool_halo(void *out, void *plane_out, const void *plane, const int32_t *mapping_exec, int E, int N,
                           int C, int H, int W, int BS_in, int ksize, int stride, int pad, bc_stream_t stream) {
  BC_REQUIRE(E >= 0, BC_ERR_SHAPE, "bc_maxpool_halo: E=%d", E);
  if (E == 0) return BC_OK;
  return maxpool_halo(out, plane_out, plane, mapping_exec, E, N, C, H, W, BS_in, ksize, stride, pad,
                      (cudaStream_t)stream);
}

BC_API int bc_stem_pack(void *s2d_plane, const void *tiles, const int32_t *mapping_exec, int E, int N, int H, int W,
                        int BS, bc_stream_t stream) {
  BC_REQUIRE(E >= 0, BC_ERR_SHAPE, "bc_stem_pack: E=%d", E);
  if (E == 0) return BC_OK;
  return stem_pack(s2d_plane, tiles, mapping_exec, E, N, H, W, BS, (cudaStream_t)stream);
}

BC_API int bc_conv_stem(void *out, const void *s2d_plane, const void *weight, const void *bias,
                        const int32_t *mapping_exec, int E, int N, int Hs, int Ws, int BS_out, int Cout, int relu,
                        void *plane_out, bc_stream_t stream) {
  BC_REQUIRE(E >= 0, BC_ERR_SHAPE, "bc_conv_stem: E=%d", E);
  if (E == 0) return BC_OK;
  return conv_stem(out, s2d_plane, weight, bias, mapping_exec, E, N, Hs, Ws, BS_out, Cout, relu, plane_out,
                   (cudaStream_t)stream);
}

BC_API int bc_policy_features(float *out, const void *frame, const void *frame_state, const void *output_repr,
                              const uint8_t *grid, int N, int K, int H, int W, int h, int w, int GH, int GW, int Ho,
                              int Wo, const int64_t *repr_strides, float inv_scale_y, float inv_scale_x,
                              bc_dtype_t dtype, bc_stream_t stream) {
  return policy_features(out, frame, frame_state, output_repr, grid, N, K, H, W, h, w, GH, GW, Ho, Wo, repr_strides,
                         inv_scale_y, inv_scale_x, dtype, (cudaStream_t)stream);
}

BC_API int bc_bn_stats(float *mean, float *invstd, const void *x, long long P, int C, float eps, void *workspace,
                       long long workspace_bytes, bc_stream_t stream) {
  return bn_stats(mean, invstd, x, P, C, eps, workspace, workspace_bytes, (cudaStream_t)stream);
}

BC_API int bc_pack_params(const long long *table, int n, long long total, bc_stream_t stream) {
  return pack_params(table, n, total, (cudaStream_t)stream);
}

BC_API int bc_rmsprop_step(const long long *table, int n, long long total, float lr, float alpha, float eps,
                           float weight_decay, float momentum, bc_stream_t stream) {
  return rmsprop_step(table, n, total, lr, alpha, eps, weight_decay, momentum, (cudaStream_t)stream);
}

BC_API int bc_conv_fewout(float *out, const void *x, const float *w, const float *bias, int N, int H, int W, int C, int Cx,
                          int Cout, int k, int stride, int pad, const int64_t *w_strides, bc_stream_t stream) {
  return conv_fewout(out, x, w, bias, N, H, W, C, Cx, Cout, k, stride, pad, w_strides, (cudaStream_t)stream);
}

BC_API int bc_frame_from_u8(void *out, const uint8_t *src, const float *mean, const float *std, int N, int H, int W,
                            bc_dtype_t dtype, bc_stream_t stream) {
  return frame_from_u8(out, src, mean, std, N, H, W, (int)dtype, (cudaStream_t)stream);
}

BC_API int bc_upsample_argmax(void *labels, const void *logits, int N, int K, int h, int w, const int64_t *strides,
                              int scale, bc_dtype_t dtype, int label_bytes, bc_stream_t stream) {
  return upsample_argmax(labels, logits, N, K, h, w, strides, scale, (int)dtype, label_bytes, (cudaStream_t)stream);
}

BC_API int bc_upsample_argmax_blocks(void *labels, const void *logits, const uint8_t *grid, int N, int K, int h, int w,
                                     const int64_t *strides, int scale, bc_dtype_t dtype, int label_bytes, int GH, int GW,
                                     bc_stream_t stream) {
  BC_REQUIRE(grid != nullptr, BC_ERR_NULL, "bc_upsample_argmax_blocks: NULL grid");
  return upsample_argmax(labels, logits, N, K, h, w, strides, scale, (int)dtype, label_bytes, (cudaStream_t)stream, grid, GH,
                         GW);
}

BC_API int bc_policy_features_nhwc16(void *out, int Cp, const void *frame, const void *frame_state,
                                     const void *output_repr, const uint8_t *grid, int N, int K, int H, int W, int h,
                                     int w, int GH, int GW, int Ho, int Wo, const int64_t *repr_strides,
                                     float inv_scale_y, float inv_scale_x, bc_dtype_t dtype, bc_stream_t stream) {
  return policy_features_nhwc16(out, Cp, frame, frame_state, output_repr, grid, N, K, H, W, h, w, GH, GW, Ho, Wo,
                                repr_strides, inv_scale_y, inv_scale_x, (int)dtype, (cudaStream_t)stream);
}

BC_API int bc_info_gain(void *out, const void *outputs, const void *outputs_prev, int N, int K, int h, int w,
                        const int64_t *strides, bc_stream_t stream) {
  return info_gain(out, outputs, outputs_prev, N, K, h, w, strides, (cudaStream_t)stream);
}

BC_API int bc_sample_grid(uint8_t *grid, int32_t *counts, const float *probs, const float *uniforms, int G, int multiple,
                          int at_least_one, bc_stream_t stream) {
  return sample_grid(grid, counts, probs, uniforms, G, multiple, at_least_one, (cudaStream_t)stream);
}

BC_API int bc_bn_update_running(const long long *table, int n, bc_stream_t stream) {
  return bn_update_running(table, n, (cudaStream_t)stream);
}

BC_API int bc_gn_stats(float *mean, float *invstd, const void *x, long long P, int C, int groups, float eps, void *workspace,
                       long long workspace_bytes, bc_stream_t stream) {
  return gn_stats(mean, invstd, x, P, C, groups, eps, workspace, workspace_bytes, (cudaStream_t)stream);
}

BC_API int bc_depth_to_space(void *out, const void *in, int E, int C, int h, int w, int r, bc_stream_t stream) {
  return depth_to_space(out, in, E, C, h, w, r, (cudaStream_t)stream);
}

BC_API int bc_bwd_mask_add(void *dst, const void *grad, const void *out, const void *add, long long n, bc_stream_t stream) {
  return bwd_mask_add(dst, grad, out, add, n, (cudaStream_t)stream);
}

BC_API int bc_bn_bwd_reduce(float *sums, const void *g, const void *out, const void *z, const float *mean, const float *invstd,
                            long long P, int C, void *workspace, long long workspace_bytes, bc_stream_t stream) {
  return bn_bwd_reduce(sums, g, out, z, mean, invstd, P, C, workspace, workspace_bytes, (cudaStream_t)stream);
}

BC_API int bc_bn_bwd_apply(void *dz, void *dz_up, const void *g, const void *out, const void *z, const float *mean,
                           const float *invstd, const float *gamma, const float *sums, int N, int H, int W, int C,
                           bc_stream_t stream) {
  return bn_bwd_apply(dz, dz_up, g, out, z, mean, invstd, gamma, sums, N, H, W, C, (cudaStream_t)stream);
}

BC_API int bc_conv_wgrad(float *grad_w, const long long *grad_strides, const void *dz, const void *x, int N, int H, int W,
                         int Cin_p, int Cout_p, int Cin, int Cout, int ksize, int stride, const float *inv_scale, float *dgamma,
                         float *dbeta, const float *bn_sums, void *workspace, long long workspace_bytes, bc_stream_t stream) {
  return conv_wgrad(grad_w, grad_strides, dz, x, N, H, W, Cin_p, Cout_p, Cin, Cout, ksize, stride, inv_scale, dgamma, dbeta,
                    bn_sums, workspace, workspace_bytes, (cudaStream_t)stream);
}

BC_API int bc_blocks_from_u8(void *tiles, const uint8_t *src, const float *mean, const float *std, const int32_t *mapping_exec,
                             int E, int N, int H, int W, int BS, bc_dtype_t dtype, bc_stream_t stream) {
  return blocks_from_u8(tiles, src, mean, std, mapping_exec, E, N, H, W, BS, (int)dtype, (cudaStream_t)stream);
}

BC_API int bc_bn_norm(void *out, float *mean, float *invstd, const void *x, const float *weight, const float *shift, long long P,
                      int C, float eps, int relu, void *workspace, long long workspace_bytes, bc_stream_t stream) {
  return bn_norm(out, mean, invstd, x, weight, shift, P, C, eps, relu, workspace, workspace_bytes, (cudaStream_t)stream);
}

BC_API int bc_graph_record(int on) {
  GraphHook &h = graph_hook();
  h.mode = on ? 1 : 0;
  h.node = nullptr;
  h.func = nullptr;
  h.recorded = 0;
  h.patch_error = cudaSuccess;
  return BC_OK;
}

BC_API int bc_graph_last_node(void **node, const void **func) {
  GraphHook &h = graph_hook();
  BC_REQUIRE(node && func, BC_ERR_NULL, "bc_graph_last_node: NULL pointer");
  BC_REQUIRE(h.node != nullptr, BC_ERR_RANGE, "bc_graph_last_node: no launch was recorded (stream not capturing?)");
  *node = (void *)h.node;
  *func = h.func;
  return BC_OK;
}

BC_API int bc_graph_patch_next(void *exec, void *node, const void *func) {
  BC_REQUIRE(exec && node && func, BC_ERR_NULL, "bc_graph_patch_next: NULL handle");
  GraphHook &h = graph_hook();
  h.mode = 2;
  h.exec = (cudaGraphExec_t)exec;
  h.node = (cudaGraphNode_t)node;
  h.func = func;
  return BC_OK;
}

BC_API int bc_graph_memcpy(void *dst, const void *src, long long bytes, bc_stream_t stream, void **node) {
  BC_REQUIRE(dst && src && bytes > 0, BC_ERR_NULL, "bc_graph_memcpy: NULL pointer / empty copy");
  const cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail((int)e, "bc_graph_memcpy: %s", cudaGetErrorString(e));
  if (node) {
    GraphHook h;
    const cudaError_t r = graph_record_after_launch(h, nullptr, (cudaStream_t)stream);
    if (r != cudaSuccess) return fail((int)r, "bc_graph_memcpy: %s", cudaGetErrorString(r));
    *node = (void *)h.node;  // NULL when the stream is not capturing
  }
  return BC_OK;
}

BC_API int bc_graph_patch_memcpy(void *exec, void *node, void *dst, const void *src, long long bytes) {
  BC_REQUIRE(exec && node && dst && src && bytes > 0, BC_ERR_NULL, "bc_graph_patch_memcpy: NULL handle / pointer");
  const cudaError_t e = cudaGraphExecMemcpyNodeSetParams1D((cudaGraphExec_t)exec, (cudaGraphNode_t)node, dst, src, (size_t)bytes,
                                                           cudaMemcpyDeviceToDevice);
  if (e != cudaSuccess) return fail((int)e, "bc_graph_patch_memcpy: %s", cudaGetErrorString(e));
  return BC_OK;
}

BC_API int bc_raster_boxes(float *out, const int32_t *rects, const float *values, int n, int H, int W, int shift,
                           bc_stream_t stream) {
  return raster_boxes(out, rects, values, n, H, W, shift, (cudaStream_t)stream);
}

BC_API int bc_spp_pool(void *pooled, const void *x0, int N, int C, int H, int W, int L, const int32_t *grid_h,
                       const int32_t *grid_w, bc_stream_t stream) {
  return spp_pool(pooled, x0, N, C, H, W, L, grid_h, grid_w, (cudaStream_t)stream);
}

BC_API int bc_spp_levels(void *out, const void *pooled, const float *bn, const void *weights, int N, int C, int H, int W,
                         int L, const int32_t *grid_h, const int32_t *grid_w, int level_channels, bc_stream_t stream) {
  return spp_levels(out, pooled, bn, weights, N, C, H, W, L, grid_h, grid_w, level_channels, (cudaStream_t)stream);
}

BC_API int bc_spp_prep(void *y, const void *x0, const void *levels, const float *bn, int N, int C, int H, int W, int L,
                       const int32_t *grid_h, const int32_t *grid_w, int level_channels, int padded_channels,
                       bc_stream_t stream) {
  return spp_prep(y, x0, levels, bn, N, C, H, W, L, grid_h, grid_w, level_channels, padded_channels,
                  (cudaStream_t)stream);
}

}  // extern "C"
