/*
 * blockcopy_b200.h -- C ABI of libblockcopy_sm100.so
 *
 * B200 (sm_100a) implementation of BlockCopy's block-sparse execution path.
 * This header is the drop-in boundary: every entry point replaces one place
 * where the reference (thomasverelst/blockcopy-video-processing-pytorch) hands
 * raw device pointers to a JIT-compiled CuPy kernel.  The reference-side stubs
 * a maintainer would write against it are shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C, PODs only: device pointers as void*, sizes as int, the CUDA stream
 *    as an opaque handle (pass torch.cuda.current_stream().cuda_stream, exactly
 *    what the reference passes at blockcopy/utils/block_funcs.py:48).
 *  - every function returns 0 on success, a NEGATIVE bc_status for an argument
 *    error detected on the host (nothing launched), a POSITIVE cudaError_t when
 *    the CUDA runtime refused the launch.  Nothing throws, nothing synchronises,
 *    nothing allocates device memory.  bc_last_error_string() describes the last
 *    failure of the calling thread.
 *  - the caller owns all buffers.  Outputs are written in place; inputs are
 *    never modified.  Tensors must be dense in the stated layout.
 *  - grid conventions are the reference's (core/tensorwrapper.py:108-128):
 *    cells are numbered row-major over (n, gh, gw); mapping_exec[b] is the cell
 *    of packed tile b; grid_idx[g] >= 0 is the packed-tile row of an executed
 *    cell, grid_idx[g] < 0 encodes row grid_idx[g] + N*GH*GW of the "transfer"
 *    tensor of a skipped cell.
 */
#ifndef BLOCKCOPY_B200_H_
#define BLOCKCOPY_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BC_ABI_VERSION 1

#if defined(__GNUC__)
#define BC_API __attribute__((visibility("default")))
#else
#define BC_API
#endif

typedef void *bc_stream_t; /* cudaStream_t */

typedef enum {
  BC_F16 = 0, /* 2-byte elements (fp16; bf16 moves through the same path) */
  BC_F32 = 1  /* 4-byte elements */
} bc_dtype_t;

typedef enum {
  BC_NCHW = 0, /* the reference's layout: planes (N,C,H,W), tiles (E,C,BS,BS)      */
  BC_NHWC = 1  /* B200 product layout: planes (N,H,W,C), tiles (E,BS,BS,C)         */
} bc_layout_t;

typedef enum {
  BC_OK = 0,
  BC_ERR_NULL = -1,      /* required pointer is NULL                               */
  BC_ERR_SHAPE = -2,     /* H or W not a multiple of BS, non-positive size, ...    */
  BC_ERR_DTYPE = -3,     /* dtype / layout enum out of range                       */
  BC_ERR_ALIGN = -4,     /* pointer not aligned to the element size                */
  BC_ERR_RANGE = -5,     /* problem too large for 32-bit chunk indexing            */
  BC_ERR_UNSUPPORTED = -6,
  BC_ERR_NO_DEVICE = -7  /* no sm_100 device / driver entry point missing          */
} bc_status_t;

/* ---- library ---------------------------------------------------------------- */
BC_API int bc_version(void);                     /* BC_ABI_VERSION the library was built with */
BC_API const char *bc_last_error_string(void);   /* thread-local, never NULL                  */
BC_API const char *bc_build_info(void);          /* "sm_100a nvcc 12.9 ..."                   */

/* ---- index bookkeeping -------------------------------------------------------
 * Replaces BlockFeatures._process_grid + get_grid_mappings
 * (core/tensorwrapper.py:150-178, :108-128), which the reference runs on the HOST
 * (D2H of the grid, CPU TorchScript, three H2D copies).  One single-CTA kernel.
 *   grid           uint8/bool  [G]   in   (G = N*GH*GW)
 *   grid_idx       int32       [G]   out
 *   mapping_exec   int32       [G]   out  (first E entries valid)
 *   counts         int32       [2]   out  {E, G-E}
 *   prev_grid_idx  int32       [G]   in   nullable
 *   transfer_idx   int32       [G]   out  nullable (first G-E entries valid)
 */
BC_API int bc_compact_mask(const uint8_t *grid, int G, int32_t *grid_idx, int32_t *mapping_exec,
                    int32_t *counts, const int32_t *prev_grid_idx, int32_t *transfer_idx,
                    bc_stream_t stream);

/* ---- active-block gather (no halo) -------------------------------------------
 * Replaces SplitFunction / split_kernel (utils/block_funcs.py:10-83).
 *   blocks[b, :, h, w] = image[n, :, gh*BS+h, gw*BS+w],  (n,gh,gw) = cell mapping_exec[b]
 */
BC_API int bc_gather(void *blocks, const void *image, const int32_t *mapping_exec, int E, int N, int C,
              int H, int W, int BS, bc_dtype_t dtype, bc_layout_t layout, bc_stream_t stream);

/* ---- scatter into the previous frame's plane, in place ------------------------
 * Replaces CombineFunction / combine_kernel (utils/block_funcs.py:85-158).
 */
BC_API int bc_scatter(const void *blocks, void *image, const int32_t *mapping_exec, int E, int N, int C,
               int H, int W, int BS, bc_dtype_t dtype, bc_layout_t layout, bc_stream_t stream);

/* ---- copy unchanged blocks from the previous frame, fused with the scatter ----
 * Replaces the NON-in-place combine (clone + combine_kernel,
 * core/tensorwrapper.py:421-434): for every cell g of the plane
 *   out[cell g] = grid_idx[g] >= 0 ? blocks[grid_idx[g]] : prev[cell g]
 * in one pass; out must not alias prev.
 */
BC_API int bc_copy_blocks(void *out, const void *prev, const void *blocks, const int32_t *grid_idx, int N,
                   int C, int H, int W, int BS, bc_dtype_t dtype, bc_layout_t layout,
                   bc_stream_t stream);

/* ---- ring transfer from the previous frame's tiles ----------------------------
 * Replaces TransferFunction / transfer_kernel (utils/block_funcs.py:161-237).
 * Writes only the border ring of width `padding` of each of the T transferred
 * tiles (all pixels if padding < 0); interiors of `out` are left untouched.
 *   G = N*GH*GW of the grid the indices refer to.
 */
BC_API int bc_transfer(void *out, const void *prev_exec, const void *prev_transfer,
                const int32_t *transfer_idx, int T, int G, int C, int BS, int padding,
                bc_dtype_t dtype, bc_layout_t layout, bc_stream_t stream);

/* ---- gather with halo, reference tile protocol --------------------------------
 * Replaces BlockPadFunction / repad_kernel (utils/blockpad.py:14-156):
 * (E,C,BS,BS) -> (E,C,BS+2p,BS+2p); halo from the neighbouring cell's tile in
 * `exec` (grid_idx >= 0) or `transfer` (grid_idx < 0), zeros outside the frame.
 */
BC_API int bc_gather_halo_tiles(void *out, const void *exec, const void *transfer, const int32_t *grid_idx,
                         const int32_t *mapping_exec, int E, int N, int C, int GH, int GW, int BS,
                         int pad, bc_dtype_t dtype, bc_layout_t layout, bc_stream_t stream);

/* ---- gather with halo from a dense persistent plane (B200 design) -------------
 * Same result as transfer + repad when the plane holds, for every cell, the most
 * recently executed tile (proven bit-identical in tests):
 *   out[b] = zero_pad(plane, p)[n, :, gh*BS : gh*BS+BS+2p, gw*BS : gw*BS+BS+2p]
 * NHWC + 16-byte-aligned channel runs take the TMA path (cp.async.bulk.tensor
 * box loads with hardware zero fill -> shared memory -> bulk store).
 */
BC_API int bc_gather_halo(void *out, const void *plane, const int32_t *mapping_exec, int E, int N, int C,
                   int H, int W, int BS, int pad, bc_dtype_t dtype, bc_layout_t layout,
                   bc_stream_t stream);

/* ---- implicit-GEMM convolution on the executed blocks (tcgen05 / TMEM / TMA) -----
 * Replaces, for one conv of the wrapped CNN, transfer_kernel + repad_kernel + the cuDNN call on the
 * padded tile batch (core/tensorwrapper.py:529-575; utils/blockpad.py; utils/block_funcs.py:161-237):
 * the operand load takes the block indices directly and reads the op's persistent NHWC fp16 plane
 * (N,H,W,Cin) -- halos are the neighbouring cells of the plane, zeros outside the frame --
 * so padded tiles never exist in memory.  Epilogue: + bias, + residual, ReLU, fp16 store into the
 * packed NHWC tile batch out (E, BS_in/stride, BS_in/stride, Cout).
 *   weight   fp16 [Cout][k][k][Cin]  (= a channels_last (Cout,Cin,k,k) tensor)
 *   bias     fp16 [Cout] or NULL;  residual fp16, layout of out, or NULL
 *   mapping_exec NULL: the "plane" is itself a packed tile batch (N = E images of BS_in x BS_in;
 *                      used for the 1x1 convs, which need no halo)
 * Optional fused scatter (north-star item (b)/(c)): when plane_out != NULL the epilogue ALSO writes
 * every output pixel into the NEXT padded op's persistent plane (out_N, out_GH*BS_out, out_GW*BS_out,
 * Cout) at the block's position (out_mapping = mapping_exec of the output grid), so no separate
 * scatter kernel runs and cells that were not executed keep the previous frame's values.  With plane_out given,
 * out may be NULL: the tile batch is then not written at all (the consumer reads the plane).
 * allow_split_k != 0: layers whose output tiles cannot fill 148 SMs (4..8-px blocks) are split along K
 * over a thread-block cluster (1,1,S<=8) and the partial accumulators are summed in rank order, so
 * results are run-to-run reproducible.  workspace (optional, 16-byte aligned device memory private to the
 * stream, workspace_bytes long): the partials travel through it (L2); 4 * 128 * Cout * ceil(pixels/128) * S
 * bytes are needed -- 32 MiB covers every supported shape up to 64 Ki output pixels.  NULL or too small:
 * they stay in shared memory and are reduced through distributed shared memory (slower, same result).
 * Supported: k = 1 (pad 0) or k = 3 with pad = dilation = p in 1..4 (the size-preserving dilated conv: taps at
 * -p, 0, +p; p = 1 is the ordinary 3x3; p > 1 needs stride 1), stride in {1,2}, Cin % 64 == 0, Cout % 64 == 0,
 * output block edge a power of two in [2,128]; anything else returns BC_ERR_UNSUPPORTED.
 */
BC_API int bc_conv_igemm(void *out, const void *plane, const void *weight, const void *bias,
                         const void *residual, const int32_t *mapping_exec, int E, int N, int Cin, int H,
                         int W, int BS_in, int Cout, int ksize, int stride, int pad, int relu,
                         void *plane_out, const int32_t *out_mapping, int out_N, int out_GH, int out_GW,
                         int allow_split_k, void *workspace, long long workspace_bytes, bc_stream_t stream);

/* ---- fused elementwise stage between two convs, on packed NHWC fp16 tiles ----------------------
 * Replaces the pass-through torch ops the reference issues on the tile batch between padded ops
 * (core/tensorwrapper.py:519-520, :577-598): per-block bilinear x2 (taps clamped at the block edge),
 * `x += skip`, eval-mode batch_norm, ReLU -- and this design's scatter into the next plane:
 *   y = relu?( bn?( up2x?(a) + residual? ) );   out <- y (if out);  plane_out[cells] <- y (if plane_out)
 * a: (E, BS/2, BS/2, C) if up2x else (E, BS, BS, C); residual/out: (E, BS, BS, C); plane_out:
 * (N, H, W, C) with H, W multiples of BS; bn_*: fp32 [C] (bn_weight / bn_shift may be NULL), NULL
 * bn_mean = no batch norm.  Intermediates are rounded to fp16 where the unfused sequence rounds.
 * out may alias a (when !up2x) or residual.  C % 8 == 0.
 */
BC_API int bc_ew_fused(void *out, void *plane_out, const void *a, const void *residual, const float *bn_mean,
                       const float *bn_invstd, const float *bn_weight, const float *bn_shift,
                       const int32_t *mapping_exec, int E, int C, int BS, int N, int H, int W, int up2x, int relu,
                       bc_stream_t stream);

/* ---- max-pool on the executed blocks, straight from the op's persistent plane -------------------
 * Replaces transfer + repad + at::max_pool2d(padding=0) of a padded pooling op
 * (core/tensorwrapper.py:529-575 with func = max_pool2d): out[b] = maxpool_k,stride(zero_pad(plane, pad))
 * on block b; halo = neighbouring cells, ZEROS (not -inf) outside the frame, as in the reference.
 * NHWC fp16, C % 8 == 0; (BS_in + 2*pad - k)/stride + 1 must equal BS_in/stride.  plane_out (optional):
 * the next padded op's plane (N, H/stride, W/stride, C), written in the same pass.
 */
BC_API int bc_maxpool_halo(void *out, void *plane_out, const void *plane, const int32_t *mapping_exec, int E, int N,
                           int C, int H, int W, int BS_in, int ksize, int stride, int pad, bc_stream_t stream);

/* ---- ResNet stem (conv 7x7, stride 2, padding 3, 3 input channels) on executed blocks -------------
 * Replaces transfer + repad + cuDNN conv + bias + ReLU for backbone.conv1 (core/tensorwrapper.py:529-575).
 * The op's temporal state is a space-to-depth(2) plane s2d (N, H/2, W/2 + 2*BC_STEM_XPAD, 16) fp16 NHWC:
 *   s2d[n, Y, BC_STEM_XPAD + X, (dy*2+dx)*3 + c] = frame[n, c, 2Y+dy, 2X+dx],  channels 12..15 = 0,
 *   and the BC_STEM_XPAD pixels on either side of every row are ZERO (the owner zero-fills the plane once;
 *   neither entry point writes them).  They are the conv's horizontal frame-border padding.
 * bc_stem_pack writes the executed cells of it from the packed NCHW input tiles (E,3,BS,BS);
 * bc_conv_stem runs the conv as a 4x4 stride-1 implicit GEMM on that plane (tcgen05, K = 16 taps x 16):
 *   weight fp16 [Cout][4][4][16], w'[o,kh',kw',(dy*2+dx)*3+c] = w[o,c,2kh'+dy-1,2kw'+dx-1] (0 outside 0..6)
 *   out    fp16 (E, BS_out, BS_out, Cout) NHWC, BS_out = BS/2;  epilogue: + bias, ReLU
 *   plane_out optional: the next padded op's plane (N, H/2, W/2, Cout) NHWC, written in the same pass.
 *   (out may be NULL when plane_out is given: tiles are not written.)
 */
#define BC_STEM_XPAD 2
BC_API int bc_stem_pack(void *s2d_plane, const void *tiles, const int32_t *mapping_exec, int E, int N, int H, int W,
                        int BS, bc_stream_t stream);
BC_API int bc_conv_stem(void *out, const void *s2d_plane, const void *weight, const void *bias,
                        const int32_t *mapping_exec, int E, int N, int Hs, int Ws, int BS_out, int Cout, int relu,
                        void *plane_out, bc_stream_t stream);

/* ---- policy network input (replaces PolicyNet.forward's feature building, policy/net.py:84-113) ---
 * out (N, 3+3+K+1, Ho, Wo) fp32 NCHW = [ nearest(frame) | nearest(frame_state) | nearest(output_repr) - 0.5 |
 * nearest(grid) - 0.5 ].  frame / frame_state: (N,3,H,W) NCHW, dtype F16 or F32; output_repr: (N,K,h,w) of the
 * same dtype with ELEMENT strides repr_strides[4] (NCHW or channels_last); grid: (N,1,GH,GW) bool.
 * inv_scale_*: 1/scale_factor of the frame resize (ATen's nearest index = floor(dst * inv_scale)).
 */
BC_API int bc_policy_features(float *out, const void *frame, const void *frame_state, const void *output_repr,
                              const uint8_t *grid, int N, int K, int H, int W, int h, int w, int GH, int GW, int Ho,
                              int Wo, const int64_t *repr_strides, float inv_scale_y, float inv_scale_x,
                              bc_dtype_t dtype, bc_stream_t stream);

/* The same features as the input plane of the fused policy trunk (blockcopy/policy/fused_net.py): fp16 NHWC
 * (N, Ho, Wo, Cp) with the channel count padded to Cp (multiple of 8, >= 7+K); values = the fp32 features rounded to
 * fp16; channels >= 8*ceil((7+K)/8) are not written (zero the plane once).
 */
BC_API int bc_policy_features_nhwc16(void *out, int Cp, const void *frame, const void *frame_state,
                                     const void *output_repr, const uint8_t *grid, int N, int K, int H, int W, int h,
                                     int w, int GH, int GW, int Ho, int Wo, const int64_t *repr_strides,
                                     float inv_scale_y, float inv_scale_x, bc_dtype_t dtype, bc_stream_t stream);

/* ---- output head fused with the final combine ------------------------------------------------------
 * Replaces, at the end of every frame, eval batch_norm + ReLU on the tile batch, the few-channel 1x1 conv
 * (class logits; cuDNN in the reference path, core/tensorwrapper.py:519-520), its bias add, and
 * out.combine() = clone of the previous output + combine_kernel (core/blockcopy.py:79-86,
 * core/tensorwrapper.py:421-434):
 *   y = conv1x1(relu?(bn?(tiles_in))) + bias;  tiles_out <- y;  dense_out <- dense_prev with executed cells = y
 * tiles_in (E,BS,BS,Cin) NHWC fp16, Cin % 8 == 0; weight fp16 [Cout][Cin], Cout <= 32; bn_*: fp32 [Cin] or NULL;
 * tiles_out (E,Cout,BS,BS) / dense_out, dense_prev (N,Cout,GH*BS,GW*BS) in the given layouts, each optional
 * (dense_prev NULL: only executed cells are written).  Rounding as op by op: BN -> fp16, conv (fp32 sum) ->
 * fp16, + bias -> fp16.
 */
BC_API int bc_head_1x1(void *tiles_out, void *dense_out, const void *dense_prev, const void *tiles_in,
                       const void *weight, const void *bias, const float *bn_mean, const float *bn_invstd,
                       const float *bn_weight, const float *bn_shift, int relu_in, const int32_t *grid_idx,
                       const int32_t *mapping_exec, int E, int N, int GH, int GW, int BS, int Cin, int Cout,
                       bc_layout_t tiles_layout, bc_layout_t dense_layout, bc_stream_t stream);

/* ---- information gain of semantic segmentation (policy/information_gain.py:32-41) ------------------
 * out (N,1,h/4,w/4) fp16 = mean_c[ p_prev * (log p_prev - log p_cur) ] of the bilinearly 1/4-resized
 * logits (align_corners False); outputs / outputs_prev: fp16 (N,K,h,w) with element strides[4];
 * h, w multiples of 4; K <= 64.  Intermediates are rounded to fp16 where the op-by-op sequence rounds.
 */
BC_API int bc_info_gain(void *out, const void *outputs, const void *outputs_prev, int N, int K, int h, int w,
                        const int64_t *strides, bc_stream_t stream);

/* ---- Bernoulli draw of the execution grid + count quantisation (policy/policy.py:124-144, :255-266) ----------
 * grid[g] = uniforms[g] < probs[g] (the comparison torch.bernoulli makes with its own draw); if nothing executes
 * and at_least_one, cell 0 does (policy.py:262-263); then the executed count E0 is rounded UP to
 * target = multiple * (1 + (E0 - 1) / multiple) (0 stays 0; multiple 0 = no rounding) by switching on the
 * (target - E0) skipped cells with the smallest (uniforms[G + g], g): a uniformly random subset, the device-side
 * counterpart of the reference's host round trip + random.sample.  probs fp32 [G], uniforms fp32 [2G] in [0,1),
 * grid uint8 [G], counts int32 [2] = {executed after rounding, executed before}.  G <= 8192.  Deterministic in its
 * inputs (blockcopy/policy/policy.py::sample_grid_host is the host restatement the tests compare with).
 */
BC_API int bc_sample_grid(uint8_t *grid, int32_t *counts, const float *probs, const float *uniforms, int G, int multiple,
                          int at_least_one, bc_stream_t stream);

/* ---- GroupNorm statistics over all executed blocks (core/tensorwrapper.py:600-633, `_func_batched`) -------------
 * The reference folds the (E,C,h,w) tile batch into ONE sample before F.group_norm, so a group's statistics run over
 * that group's channels of every executed block.  x: packed NHWC fp16 tiles = P = E*h*w pixels x C channels;
 * mean[c], invstd[c] = 1/sqrt(biased var + eps) of channel c's GROUP (fp32 [C]), ready for bc_ew_fused, which then
 * computes weight * (x - mean) * invstd + bias.  C and C/groups multiples of 8, C <= 2048, groups <= 256.
 * workspace: caller-owned, >= BC_GN_STATS_WORKSPACE bytes, 16-byte aligned, first 4 bytes zero before the first call
 * (left at zero), private to the stream.  Reproducible run to run.
 */
#define BC_GN_STATS_WORKSPACE (16 + 2 * 148 * 256 * 2 * 8)
BC_API int bc_gn_stats(float *mean, float *invstd, const void *x, long long P, int C, int groups, float eps, void *workspace,
                       long long workspace_bytes, bc_stream_t stream);

/* ---- depth-to-space on packed NHWC fp16 tiles: second half of a per-block ConvTranspose2d ---------------------
 * The reference runs ConvTranspose2d as a pass-through op on the tile batch (core/tensorwrapper.py:519-520; Pedestron
 * necks/csp_neck.py:37-83): every tile is its own sample, zeros beyond the TILE edge.  A transposed conv with stride r
 * is then a plain conv with r*r times the output channels -- bc_conv_igemm over the tile batch viewed as E one-block
 * frames (mapping 0..E-1), so the frame border of the kernel is the tile border -- followed by
 *     out[e, r*y + a, r*x + b, c] = in[e, y, x, (a*r + b)*C + c]
 * in: (E, h, w, r*r*C), out: (E, r*h, r*w, C), fp16, C % 8 == 0, 1 <= r <= 8.
 */
BC_API int bc_depth_to_space(void *out, const void *in, int E, int C, int h, int w, int r, bc_stream_t stream);

/* ---- train-mode batch norm in ONE launch: bc_bn_stats + the normalisation (policy/net.py:115-125, policy/resnet.py) ----
 * out (P, C) fp16 = relu?(weight * (x - mean) * invstd + shift) with the batch statistics of x itself; mean / invstd are
 * also written out (running-statistics update, backward pass).  The kernel's CTAs (at most one per SM) meet at a grid-wide
 * barrier between the two phases.  Arguments as bc_bn_stats; workspace: same size, first 8 bytes zero before the first call
 * (left at zero), private to the stream.  Same bits as bc_bn_stats followed by bc_ew_fused; measured SLOWER than those two
 * launches on the policy trunk (330 vs 308 us per forward), so the Python host uses it only with BLOCKCOPY_BN_NORM=1.
 */
BC_API int bc_bn_norm(void *out, float *mean, float *invstd, const void *x, const float *weight, const float *shift, long long P,
                      int C, float eps, int relu, void *workspace, long long workspace_bytes, bc_stream_t stream);

/* ---- running statistics of train-mode batch norms (policy/net.py:115-125 runs the policy net in train mode) ----
 * For every row l < n of the DEVICE table (8 x int64 per row: batch mean fp32*, batch invstd fp32* -- the outputs of
 * bc_bn_stats --, running_mean fp32* | 0, running_var fp32* | 0, num_batches_tracked int64* | 0, C, count = N*H*W,
 * momentum float bits (negative: cumulative average) | eps float bits << 32): running = (1 - f) * running + f * batch
 * statistic with the UNBIASED batch variance, num_batches_tracked += 1 -- what F.batch_norm(training=True) does.
 */
BC_API int bc_bn_update_running(const long long *table, int n, bc_stream_t stream);

/* ---- backward pass of the policy CNN (policy/policy.py:319-370 `loss.backward()` through policy/net.py:78-125 and
 * policy/resnet.py:60-115; the reference runs it as ~90 fp32 cuDNN / ATen launches every block_train_interval frames) ----
 * All planes are dense NHWC fp16 with channel counts padded to 64 / 128 (padded channels carry zeros), statistics and
 * parameter gradients are fp32.  One unit = conv -> BatchNorm(batch statistics) -> [+ shortcut] -> [ReLU]:
 *   z = conv(x);  out = relu?(gamma * (z - mean) * invstd + beta [+ shortcut])
 * and, given dOut (scaled by a power-of-two loss scale so that fp16 holds it),
 *   g     = out > 0 ? dOut : 0                     (when the unit, or the block it ends, has a ReLU)
 *   sums  = [ sum_p g, sum_p g * xhat ]            bc_bn_bwd_reduce   (xhat = (z - mean) * invstd; reproducible)
 *   dz    = gamma * invstd * (g - sums[0]/P - xhat * sums[1]/P)        bc_bn_bwd_apply
 *   dW    = wgrad(dz, x) / scale, dgamma = sums[1] / scale, dbeta = sums[0] / scale      bc_conv_wgrad
 *   dx    = conv(dz, flipped / transposed W): bc_conv_igemm; for a stride-2 conv over dz_up, the (N,2H,2W,C) plane
 *           bc_bn_bwd_apply writes dz into at the even positions (all other positions stay zero: zeroed once by the owner)
 */
/* dst = (out ? (out > 0 ? grad : 0) : grad) + (add ? add : 0); n fp16 elements (multiple of 8), dst may alias grad / add. */
BC_API int bc_bwd_mask_add(void *dst, const void *grad, const void *out, const void *add, long long n, bc_stream_t stream);
/* sums fp32 [2][C]; g, out (NULL: no mask), z: (P, C) fp16; C in {8,16,32,64,128}; workspace >= BC_BN_STATS_WORKSPACE bytes,
 * first 4 bytes zero before the first call (left at zero), private to the stream. */
BC_API int bc_bn_bwd_reduce(float *sums, const void *g, const void *out, const void *z, const float *mean, const float *invstd,
                            long long P, int C, void *workspace, long long workspace_bytes, bc_stream_t stream);
/* dz (N,H,W,C) and / or dz_up (N,2H,2W,C), either may be NULL (not both); gamma NULL = 1. */
BC_API int bc_bn_bwd_apply(void *dz, void *dz_up, const void *g, const void *out, const void *z, const float *mean,
                           const float *invstd, const float *gamma, const float *sums, int N, int H, int W, int C,
                           bc_stream_t stream);
/* grad_w: the fp32 parameter gradient (Cout, Cin, k, k) with element strides grad_strides[4]; dz (N, H/s, W/s, Cout_p),
 * x (N, H, W, Cin_p); k in {1,3} with pad k/2, stride s in {1,2}; Cin_p, Cout_p in {64,128}, Cin_p <= Cout_p.
 * inv_scale: device scalar (NULL = 1) multiplied into every output.  bn_sums (NULL: none): the unit's bc_bn_bwd_reduce
 * output, from which dgamma / dbeta (fp32 [Cout], either may be NULL) are written in the same launch.
 * workspace: >= BC_WGRAD_WORKSPACE bytes, private to the stream.  Reproducible run to run. */
#define BC_WGRAD_WORKSPACE (148ll * 9 * 64 * 64 * 4)
BC_API int bc_conv_wgrad(float *grad_w, const long long *grad_strides, const void *dz, const void *x, int N, int H, int W,
                         int Cin_p, int Cout_p, int Cin, int Cout, int ksize, int stride, const float *inv_scale, float *dgamma,
                         float *dbeta, const float *bn_sums, void *workspace, long long workspace_bytes, bc_stream_t stream);

/* ---- CUDA-graph node recording / patching (used by the Python host's block_cuda_graphs mode) -------------------
 * The host replays one captured graph per frame.  The kernels whose ARGUMENTS differ between frames (index compaction of
 * the caller's grid, the gather from the caller's frame, the copy into the output buffer) are captured as well and
 * re-pointed before every replay instead of being launched eagerly in front of the graph:
 *   bc_graph_record(1) ... launch on a capturing stream ... bc_graph_last_node(&node, &func) ... bc_graph_record(0)
 *   per replay: bc_graph_patch_next(exec, node, func); <the same entry point with this frame's pointers>  -- that call
 *   does not launch: it writes its grid / block / arguments into `node` of the instantiated graph `exec`
 *   (cudaGraphExecKernelNodeSetParams) and fails if it would have chosen another kernel variant than the captured one.
 * bc_graph_memcpy: device-to-device copy on `stream` (captured as a memcpy node; *node = its handle while capturing, else
 * NULL); bc_graph_patch_memcpy re-points such a node.  exec / node are cudaGraphExec_t / cudaGraphNode_t as void*.
 * State is per host thread.
 */
BC_API int bc_graph_record(int on);
BC_API int bc_graph_last_node(void **node, const void **func);
BC_API int bc_graph_patch_next(void *exec, void *node, const void *func);
BC_API int bc_graph_memcpy(void *dst, const void *src, long long bytes, bc_stream_t stream, void **node);
BC_API int bc_graph_patch_memcpy(void *exec, void *node, void *dst, const void *src, long long bytes);

/* ---- box rasteriser of the object-detection information gain (policy/information_gain.py:56-108) ----
 * Replaces the reference's per-box torch slice assignments `mask[y1:y2, x1:x2] = max(mask[...], value)`
 * (build_instance_mask :56-66, build_instance_mask_iou_gain :68-108): out (H,W) fp32 <- for every pixel the
 * maximum of 0 and the values of the boxes containing (x >> shift, y >> shift); rects = device int32 [n][4]
 * (x1, y1, x2, y2), half-open like the Python slices; values = device fp32 [n].  shift = 1 reproduces the
 * reference's half-resolution raster + nearest x2 upsampling (:72,105).  n = 0 writes zeros.
 */
BC_API int bc_raster_boxes(float *out, const int32_t *rects, const float *values, int n, int H, int W, int shift,
                           bc_stream_t stream);

/* ---- batch statistics of a train-mode BatchNorm2d (the policy net, policy/net.py:115-125 in train mode) -----
 * mean[c], invstd[c] = 1/sqrt(biased var + eps) over the P = N*H*W pixels of a dense NHWC fp16 tensor x (P, C),
 * C in {8,16,32,64,128}.  workspace: caller-owned, >= BC_BN_STATS_WORKSPACE bytes, 16-byte aligned, its first 4
 * bytes ZERO before the first call (the kernel leaves them zero); private to the stream.  Reproducible run to run.
 * The pair feeds bc_ew_fused's (mean, invstd, weight, shift) directly: no host round trip.
 */
#define BC_BN_STATS_WORKSPACE (16 + 2 * 148 * 2 * 128 * 4)
BC_API int bc_bn_stats(float *mean, float *invstd, const void *x, long long P, int C, float eps, void *workspace,
                       long long workspace_bytes, bc_stream_t stream);

/* ---- parameter re-packing for the fused policy trunk (policy/fused_net.py): fp32 parameters -> fp16 channels_last
 * conv weights / fp32 vectors with padded channel counts, all tensors in ONE launch.  table: DEVICE array of n
 * entries x 12 int64 = { src ptr, dst ptr, first flat element, Cout, Cin, k, padded Cin, src element strides
 * (co, ci, kh, kw), dst is fp16 }; entries ordered by first flat element; total = elements of all sources;
 * dst layout [Cout][k][k][padded Cin] (padding left untouched).
 */
BC_API int bc_pack_params(const long long *table, int n, long long total, bc_stream_t stream);

/* ---- RMSprop step of the online policy update (policy/policy.py:56-59, :361-362) for all parameter tensors in one
 * launch; torch.optim.RMSprop semantics (centered = False):  g = grad + weight_decay * p;  sq = alpha * sq +
 * (1 - alpha) * g * g;  avg = sqrt(sq) + eps;  momentum > 0: buf = momentum * buf + g / avg, p -= lr * buf;
 * else p -= lr * g / avg.  table: HOST array of n entries x 6 int64 = { param, grad, square_avg, momentum buffer
 * (0 if momentum == 0), first flat element (cumulative numel), numel }; all tensors fp32 and dense.  The table
 * travels as a kernel parameter: no host-to-device copy, no synchronisation.
 */
BC_API int bc_rmsprop_step(const long long *table, int n, long long total, float lr, float alpha, float eps,
                           float weight_decay, float momentum, bc_stream_t stream);

/* ---- convolution with <= 16 output channels on a dense NHWC fp16 tensor (the policy net's 128 -> 1 logit layer,
 * policy/net.py:46-50): out (N,Cout,Ho,Wo) fp32 contiguous = conv(x[..., :C], w) + bias; x (N,H,W,Cx) fp16 of which
 * the first C channels are used; w fp32 (Cout,C,k,k) with ELEMENT strides w_strides[4]; zero padding `pad`.
 */
BC_API int bc_conv_fewout(float *out, const void *x, const float *w, const float *bias, int N, int H, int W, int C, int Cx,
                          int Cout, int k, int stride, int pad, const int64_t *w_strides, bc_stream_t stream);

/* ---- the steps on either side of the path in the reference's driver -----------------
 * bc_frame_from_u8: decoded frame -> network input.  Replaces ExtToTensor + ExtNormalize + .to(device, half)
 *   (semantic_segmentation/lib/ext_transforms.py:317-372, test_swiftnet.py:64-65,187):
 *   out (N,3,H,W) NCHW of dtype F16|F32 = ((src / 255) - mean[c]) / std[c], src (N,H,W,3) uint8 on the device,
 *   mean / std: HOST arrays of 3 floats.  fp32 IEEE arithmetic in ATen's op order, one rounding to F16 at the end:
 *   bit-identical to the torch sequence.
 * bc_upsample_argmax: logits -> label map.  Replaces F.interpolate(out, size=(scale*h, scale*w), mode='bilinear')
 *   + out.max(dim=1)[1] (test_swiftnet.py:196-197) without materialising the upsampled logits:
 *   labels (N, scale*h, scale*w) contiguous, uint8 (label_bytes 1, K <= 256) or int64 (label_bytes 8);
 *   logits (N,K,h,w) F16|F32 with ELEMENT strides[4]; scale in {1,2,4}; align_corners False; the blended value is
 *   rounded to the logits' dtype before comparing (as the upsampled tensor would be); ties -> lowest class.
 */
BC_API int bc_frame_from_u8(void *out, const uint8_t *src, const float *mean, const float *std, int N, int H, int W,
                            bc_dtype_t dtype, bc_stream_t stream);
/* bc_blocks_from_u8: the frame's first gather fused with the normalisation (input side) --
 *   tiles (E,3,BS,BS) NCHW of dtype F16|F32 <- normalised pixels of the E executed cells (mapping_exec[e] = flat cell
 *   index over (n, gh, gw)) of src (N,H,W,3) uint8: what bc_frame_from_u8 + bc_gather (reference
 *   core/tensorwrapper.py:335-381 `_split` after ext_transforms.py:317-372) produce, same bits, without writing the
 *   normalised full frame.  BS a multiple of 16, H and W multiples of BS; mean / std: HOST arrays of 3 floats.
 */
BC_API int bc_blocks_from_u8(void *tiles, const uint8_t *src, const float *mean, const float *std, const int32_t *mapping_exec,
                             int E, int N, int H, int W, int BS, bc_dtype_t dtype, bc_stream_t stream);
BC_API int bc_upsample_argmax(void *labels, const void *logits, int N, int K, int h, int w, const int64_t *strides,
                              int scale, bc_dtype_t dtype, int label_bytes, bc_stream_t stream);
/* Block-sparse form of the same step ("x4 bilinear + argmax fused with the logits' block structure"):
 * `labels` already holds the label map of the PREVIOUS frame and `logits` is the combined dense output of this frame,
 * which differs from the previous one only inside the executed cells of grid (uint8/bool (N,1,GH,GW), 1 = executed;
 * h % GH == 0, w % GW == 0, square logit blocks).  Only the logit pixels of executed cells plus a ring of one logit
 * pixel around each are recomputed (a bilinear tap reaches one pixel into the neighbouring cell); the result equals
 * bc_upsample_argmax over the whole frame bit for bit, at num_exec / num_total of its cost.
 */
BC_API int bc_upsample_argmax_blocks(void *labels, const void *logits, const uint8_t *grid, int N, int K, int h, int w,
                                     const int64_t *strides, int scale, bc_dtype_t dtype, int label_bytes, int GH, int GW,
                                     bc_stream_t stream);

/* ---- dense pyramid pooling (the @blockcopy_noblocks module of SwiftNet, swiftnet/util.py:85-138) ----------
 * Everything between the module's first and last 1x1 conv, on dense NHWC fp16 planes; grid_h/grid_w are HOST
 * arrays of L <= 4 pooling grids that must divide (H, W).  "cells" = sum_i N*grid_h[i]*grid_w[i], level-major.
 *   bc_spp_pool   pooled[cells][C]   = average pools of x0 (N,H,W,C) for all levels (one launch)
 *   bc_spp_levels out[cells][Lc]     = conv1x1_i(relu(bn_i(pooled)))   bn: fp32 [L][4][C] = mean,invstd,weight,shift
 *                                      weights: fp16 [L][Lc][C]
 *   bc_spp_prep   y (N,H,W,Cp)       = relu(bn(cat[x0, bilinear_up(level_0), ...])), channels >= C+L*Lc are 0;
 *                                      bn: fp32 [4][Cp]
 */
BC_API int bc_spp_pool(void *pooled, const void *x0, int N, int C, int H, int W, int L, const int32_t *grid_h,
                       const int32_t *grid_w, bc_stream_t stream);
BC_API int bc_spp_levels(void *out, const void *pooled, const float *bn, const void *weights, int N, int C, int H, int W,
                         int L, const int32_t *grid_h, const int32_t *grid_w, int level_channels, bc_stream_t stream);
BC_API int bc_spp_prep(void *y, const void *x0, const void *levels, const float *bn, int N, int C, int H, int W, int L,
                       const int32_t *grid_h, const int32_t *grid_w, int level_channels, int padded_channels,
                       bc_stream_t stream);

/* Selects the implementation of bc_gather / bc_gather_halo / bc_scatter for NHWC
 * inputs: 0 = vectorised SIMT kernels, 1 = TMA-staged kernels (default when the
 * shape qualifies).  Process-wide; meant for benchmarking the two against each other. */
/* Diagnostics (tools/cta_timeline.py, profiles/): while a device buffer is registered, every CTA of
 * bc_conv_igemm / bc_conv_stem writes a 16-word (uint64) record into it at index linear_cta_id*16:
 * [0..6] SM clock at entry / after setup / first operands landed / last MMA issued / accumulator complete /
 * epilogue done / exit, [8] %globaltimer at entry, [9] SM id, [10] %globaltimer at exit.  NULL switches
 * it off (default).  The buffer must hold 16*8 bytes per CTA of the largest launch made while registered. */
BC_API int bc_debug_trace(void *device_buffer);
BC_API int bc_set_tma_enabled(int enabled);

#ifdef __cplusplus
}
#endif
#endif /* BLOCKCOPY_B200_H_ */
