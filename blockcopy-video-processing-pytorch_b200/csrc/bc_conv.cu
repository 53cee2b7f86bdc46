// bc_conv.cu -- tcgen05 / TMEM implicit-GEMM convolution on the executed blocks.
//
// Replaces, for one padded op of the wrapped CNN, the reference's   transfer_kernel -> repad_kernel
// -> cuDNN conv on the padded tile batch   (core/tensorwrapper.py:529-575): the A operand is read
// straight from the op's persistent NHWC plane with TMA box loads addressed by BLOCK INDEX
// (mapping_exec[b] -> (n, gh, gw) -> plane coordinates); the halo is simply the box reaching into
// the neighbouring cells, and the zeros at the frame border are TMA's out-of-bounds fill.  The
// gathered / padded tiles never exist in HBM.  The epilogue adds the bias, optionally the residual
// and ReLU, and writes fp16 NHWC packed tiles (E, BS_out, BS_out, Cout).
//
// GEMM view per CTA: D[128 x N_TILE] += A[128 x 64] * B[N_TILE x 64]^T for every (tap, 64-channel
// chunk); 128 rows = 128 output pixels of one block (BS_out >= 16: 128/BS_out rows of the block) or
// of 128/BS_out^2 consecutive blocks (BS_out = 4, 8).  Both operands are K-major rows of 128 bytes
// with the 128-byte swizzle, written by TMA, consumed by tcgen05.mma.kind::f16 (M = 128, K = 16,
// fp32 accumulator in TMEM).  Warp roles: warp 0 = TMA producer (one elected lane), warp 1 = TMEM
// allocator + MMA issuer (one elected lane), warps 2..5 = epilogue (tcgen05.ld, one TMEM lane
// quadrant each).  Stride 2 uses the tensor map's element strides; 1x1 convs are the 1-tap case.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "bc_conv.cuh"
#include "bc_tma.cuh"

namespace bc {

// ---------------------------------------------------------------------------------------------------
// split-K, phase 2 through the L2 scratch: CTA `rank` of the cluster owns 128/S accumulator rows and sums
// them over the S partials in rank order (bit-reproducible).  Every thread owns U units of 8 channels;
// the 2*S partial loads and the residual load of up to 8/S units are all issued before the first use, so
// the whole reduction costs ~two L2 round trips instead of one per peer and per pass.
template <int N_TILE, int S>
__device__ __forceinline__ void splitk_reduce_l2(const ConvParams &p, uint32_t rank, unsigned tile_lin, int n0,
                                                 size_t out_base, int m_valid, const long long *row_pl_s,
                                                 const float *bias_s) {
  constexpr int kTPRow = N_TILE / 8, kRows = kTileM / S;
  constexpr int U = kRows * kTPRow / 128;
  constexpr int UPB = (8 / S < U) ? 8 / S : U;
  static_assert(U >= 1 && U % UPB == 0, "unit schedule");
  const int t = threadIdx.x - 64;
#pragma unroll 1
  for (int u0 = 0; u0 < U; u0 += UPB) {
    float4 lo[UPB][S], hi[UPB][S];
    uint4 res[UPB];
    int mm[UPB], cc[UPB];
#pragma unroll
    for (int i = 0; i < UPB; ++i) {
      const int unit = (u0 + i) * 128 + t;
      mm[i] = (int)rank * kRows + unit / kTPRow;
      cc[i] = (unit % kTPRow) * 8;
      const float4 *src = reinterpret_cast<const float4 *>(p.work + ((size_t)tile_lin * S * kTileM + mm[i]) * N_TILE + cc[i]);
#pragma unroll
      for (int z = 0; z < S; ++z) {  // L2 only: the peers' stores were released by the cluster barrier
        lo[i][z] = __ldcg(src + (size_t)z * (kTileM * N_TILE / 4));
        hi[i][z] = __ldcg(src + (size_t)z * (kTileM * N_TILE / 4) + 1);
      }
      res[i] = make_uint4(0, 0, 0, 0);
      if (p.residual && mm[i] < m_valid)
        res[i] = __ldg(reinterpret_cast<const uint4 *>(p.residual + out_base + (size_t)mm[i] * p.Cout + cc[i]));
    }
#pragma unroll
    for (int i = 0; i < UPB; ++i) {
      if (mm[i] >= m_valid) continue;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
#pragma unroll
      for (int z = 0; z < S; ++z) {
        v[0] += lo[i][z].x; v[1] += lo[i][z].y; v[2] += lo[i][z].z; v[3] += lo[i][z].w;
        v[4] += hi[i][z].x; v[5] += hi[i][z].y; v[6] += hi[i][z].z; v[7] += hi[i][z].w;
      }
      uint4 o;
      __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        oh[k] = __floats2half2_rn(v[2 * k] + bias_s[cc[i] + 2 * k], v[2 * k + 1] + bias_s[cc[i] + 2 * k + 1]);
      if (p.residual) {  // fp16-rounded conv output + identity, rounded once more (see phase B above)
        const __half2 *rh = reinterpret_cast<const __half2 *>(&res[i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) oh[k] = __hadd2(oh[k], rh[k]);
      }
      if (p.relu) {
        const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) oh[k] = __hmax2(oh[k], zero);
      }
      if (p.out) *reinterpret_cast<uint4 *>(p.out + out_base + (size_t)mm[i] * p.Cout + cc[i]) = o;
      const long long rp = row_pl_s[mm[i]];
      if (rp >= 0) *reinterpret_cast<uint4 *>(p.plane_out + rp + cc[i]) = o;
    }
  }
}

// SPLITK = false compiles the split-K epilogue out (p.splits is 1 then): the single-pass variants stay within
// the register budget of 4 CTAs per SM; the split variant (2 CTAs per SM) may keep 64+ loads in flight.
template <int N_TILE, int STAGES, bool SPLITK>
__global__ void __launch_bounds__(kConvThreadsV1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap b_map,
                  const __grid_constant__ CUtensorMap bmc_map, const ConvParams p) {
  const bool is_split = SPLITK && p.splits > 1;
  constexpr uint32_t kBBytes = N_TILE * 128;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float bias_s[N_TILE];  // this CTA's slice of the bias, staged while the main loop runs
  // element offset of channel n0 of accumulator row m's pixel in the next op's plane (-1: no such pixel / no
  // plane).  Decoding a row costs two integer divisions and a mapping lookup; done once per row while the
  // main loop runs instead of once per 16-byte store (profiles/r01c_cta_timeline.md).  In the packed tile
  // batch the rows of a tile are simply consecutive pixels: offset = out_base + m * Cout.
  __shared__ long long row_pl_s[kTileM];

  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- which output pixels does this CTA own -----------------------------------------------------
  const int tile = blockIdx.x;
  const int n0 = blockIdx.y * N_TILE;
  int b0, r0, nvalid;
  if (p.blocks_per_tile == 1) {
    b0 = tile / p.tiles_per_block;
    r0 = (tile - b0 * p.tiles_per_block) * p.rows_per_tile;
    nvalid = 1;
  } else {
    b0 = tile * p.blocks_per_tile;
    r0 = 0;
    nvalid = min(p.blocks_per_tile, p.E - b0);
  }
  const int mca = (!is_split && p.mc > 1) ? p.mc : 1;    // activation multicast (cluster along blockIdx.y)
  const int mcb = (!is_split && p.mcb > 1) ? p.mcb : 1;  // weight multicast (cluster along blockIdx.x)
  const int mc = mca > mcb ? mca : mcb;                   // CTAs per cluster (one of the two is 1)
  const uint16_t mc_mask = (uint16_t)((1u << mc) - 1u);
  const int mc_rank = mc > 1 ? (int)cluster_ctarank() : 0;
  const int total_k = p.ksize * p.ksize * p.kc_per_tap;
  const int k_begin = (int)blockIdx.z * p.ksteps_per_split;
  const int num_k = (p.debug & 1) ? 1 : min(p.ksteps_per_split, total_k - k_begin);
  if (threadIdx.x == 0) { trace_wall(p, 8); trace_mark(p, 0); }

  // Plane coordinates of the (up to 32: 2-px blocks) blocks of this tile.  The mapping lookup (an L2 round trip) is issued
  // before the set-up barrier so that it overlaps barrier init and the TMEM allocation.  `mapping` is written
  // once per frame by bc_compact_mask, never by the kernel just before this one, so reading it ahead of
  // griddepcontrol.wait is safe under programmatic dependent launch as well.
  __shared__ int4 blk_coord_s[kMaxBlocksPerTile];  // (x0, y0, image, -) per block; private to warp 0 lane 0
  int cx0 = 0, cy0 = 0, cn0 = 0;   // block 0 in registers (the only one when BS_out >= 16)
  if (warp == 0 && lane == 0) {
    prefetch_map(&a_map);
    prefetch_map(&b_map);
    if (mcb > 1) prefetch_map(&bmc_map);
    for (int i = 0; i < nvalid; ++i) {
      const uint32_t cell = p.mapping ? (uint32_t)__ldg(p.mapping + b0 + i) : (uint32_t)(b0 + i);
      uint32_t n, gh, gw;
      p.cell(cell, n, gh, gw);
      const int4 c = make_int4((int)gw * p.BS_in - p.pad, (int)gh * p.BS_in + r0 * p.stride - p.pad, (int)n, 0);
      blk_coord_s[i] = c;
      if (i == 0) { cx0 = c.x; cy0 = c.y; cn0 = c.z; }
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);  // activations (warp 0) + weights (warp 6), one arrive.expect_tx each
      mbar_init(&empty_bar[s], (uint32_t)mc);  // multicast: a stage is free once EVERY CTA of the cluster has read it
    }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, N_TILE);
  tc_fence_before_sync();
  __syncthreads();
  if (mc > 1) cluster_sync_all();  // the peers' barriers exist before anything is multicast to / arrives on them
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_trigger();
  // barrier init, TMEM alloc and descriptor prefetch above overlapped the previous kernel's tail.  The weight
  // producer (warp 6) does not wait: weights are never written by a kernel of the frame, so its first ring of
  // tiles is in flight while the previous kernel drains (BC_CONV_DEBUG=4 restores the common wait)
  if (warp != 6 || (p.debug & 4)) pdl_wait();
  if (threadIdx.x == 0) trace_mark(p, 1);

  // The three single-thread loops below are the critical path of a k-step (profiles/r01c_conv_kstep_probe.md:
  // with loads and MMAs switched off a k-step still cost ~430 clk of dependent scalar instructions): no
  // divisions, no local-memory arrays, stage index and parity advanced incrementally.
  if (warp == 0) {
    // =============================== activation producer ==========================================
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)nvalid * p.box_bytes;
      const int tap0 = k_begin / p.kc_per_tap;
      int cc = k_begin - tap0 * p.kc_per_tap, kh = tap0 / p.ksize;
      int kw = tap0 - kh * p.ksize;
      int s = 0;
      uint32_t parity = 1;  // first pass over the ring: the stages are free
      uint8_t *sa = smem;
      for (int ks = 0; ks < num_k; ++ks) {
        mbar_wait(&empty_bar[s], parity);
        mbar_expect_tx(&full_bar[s], tx_bytes);
        if (mca > 1) {  // this CTA's share of the tile's boxes, delivered to every CTA of the cluster
          for (int i = mc_rank; i < nvalid; i += mca) {
            const int4 c = blk_coord_s[i];
            tma_load_4d_multicast(sa + (size_t)i * p.box_bytes, &a_map, &full_bar[s], cc * kChunkK, c.x + kw * p.dil,
                                  c.y + kh * p.dil, c.z, mc_mask);
          }
        } else if (nvalid == 1) {
          tma_load_4d(sa, &a_map, &full_bar[s], cc * kChunkK, cx0 + kw * p.dil, cy0 + kh * p.dil, cn0);
        } else {
          for (int i = 0; i < nvalid; ++i) {
            const int4 c = blk_coord_s[i];
            tma_load_4d(sa + (size_t)i * p.box_bytes, &a_map, &full_bar[s], cc * kChunkK, c.x + kw * p.dil, c.y + kh * p.dil, c.z);
          }
        }
        if (++cc == p.kc_per_tap) {
          cc = 0;
          if (++kw == p.ksize) { kw = 0; ++kh; }
        }
        sa += kStageBytes;
        if (++s == STAGES) { s = 0; parity ^= 1; sa = smem; }
      }
    }
  } else if (warp == 6) {
    // =============================== weight producer ==============================================
    if (lane == 0) {
      int s = 0, kcoord = k_begin * kChunkK;
      uint32_t parity = 1;
      uint8_t *sb = smem + kABytes;
      for (int ks = 0; ks < num_k; ++ks) {
        mbar_wait(&empty_bar[s], parity);
        mbar_expect_tx(&full_bar[s], kBBytes);
        if (mcb > 1) {  // this CTA's rows of the weight tile, delivered to every CTA of the cluster
          const int rows = N_TILE / mcb;
          tma_load_2d_multicast(sb + (size_t)mc_rank * rows * 128, &bmc_map, &full_bar[s], kcoord, n0 + mc_rank * rows, mc_mask);
        } else {
          tma_load_2d(sb, &b_map, &full_bar[s], kcoord, n0);
        }
        kcoord += kChunkK;
        sb += kStageBytes;
        if (++s == STAGES) { s = 0; parity ^= 1; sb = smem + kABytes; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===================================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kTileM, N_TILE);
      // descriptors of stage 0; a stage / a K slab of 16 is an offset in the 16-byte-unit address field
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(smem)), b_desc0 = umma_desc_sw128(smem_u32(smem) + kABytes);
      int s = 0;
      uint32_t parity = 0, stage_off = 0;
      for (int ks = 0; ks < num_k; ++ks) {
        mbar_wait(&full_bar[s], parity);
        if (ks == 0) trace_mark(p, 2);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < kChunkK / 16; ++k)
          umma_f16_ss(tmem_base, a_desc0 + stage_off + 2 * k, b_desc0 + stage_off + 2 * k, idesc, (uint32_t)((ks | k) != 0));
        // frees the stage once these MMAs have read it -- in every CTA of the cluster when the activations are multicast
        if (mc > 1) umma_commit_multicast(&empty_bar[s], mc_mask);
        else umma_commit(&empty_bar[s]);
        stage_off += kStageBytes >> 4;
        if (++s == STAGES) { s = 0; parity ^= 1; stage_off = 0; }
      }
      umma_commit(&acc_bar);  // accumulator complete
      trace_mark(p, 3);
    }
  } else {
    // =============================== epilogue =====================================================
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    for (int c = threadIdx.x - 64; c < N_TILE; c += 128) bias_s[c] = p.bias ? __half2float(__ldg(p.bias + n0 + c)) : 0.f;
    {
      const int m = q * 32 + lane;
      int blk, y, x;
      pixel_of_row(p, m, r0, blk, y, x);
      row_pl_s[m] = (blk < nvalid && p.plane_out) ? (long long)(plane_row(p, b0 + blk, y, x) - p.plane_out) + n0 : -1ll;
    }
    const size_t out_base = ((size_t)b0 * p.BS_out * p.BS_out + (size_t)r0 * p.BS_out) * p.Cout + n0;
    const int m_valid = (p.debug & 2) ? 0 : (p.blocks_per_tile == 1 ? kTileM : nvalid * p.BS_out * p.BS_out);
    asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
    mbar_wait(&acc_bar, 0);
    if (threadIdx.x == 64) trace_mark(p, 4);
    tc_fence_after_sync();
    if (!is_split) {
      // ---- phase A: TMEM -> + bias -> fp16, one accumulator row per thread, into this warp's private
      //      staging rows (the pipeline stages are free: acc_bar says every MMA and TMA load completed)
      constexpr int kRowB = N_TILE * 2 + 16;  // +16 B: rows start in different bank groups
      uint8_t *stage = smem + (size_t)q * 32 * kRowB;
#pragma unroll 1
      for (int c0 = 0; c0 < N_TILE; c0 += 32) {
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float v[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(acc[j + t]) + bias_s[c0 + j + t];
          uint4 o;
          __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t) oh[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
          *reinterpret_cast<uint4 *>(stage + (size_t)lane * kRowB + (c0 + j) * 2) = o;
        }
      }
      __syncwarp();
      // ---- phase B: coalesced: kTPR lanes cover one pixel's N_TILE channels (16 B each), so every
      //      global access of the warp is a run of full 32-byte sectors: (+ residual) -> ReLU -> tiles, plane
      constexpr int kTPR = N_TILE / 8, kRPI = 32 / kTPR;
      const int c8 = (lane % kTPR) * 8;
      constexpr int kBatch = 4;
#pragma unroll 1
      for (int i0 = 0; i0 < 32; i0 += kRPI * kBatch) {
        size_t off[kBatch];
        __half *pl[kBatch];
        uint4 res[kBatch];
        bool ok[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {  // addresses + residual loads of the whole batch first (latency overlap)
          const int m = q * 32 + i0 + u * kRPI + lane / kTPR;
          const long long rp = row_pl_s[m];
          ok[u] = m < m_valid;
          off[u] = out_base + (size_t)m * p.Cout + c8;
          pl[u] = rp >= 0 ? p.plane_out + rp + c8 : nullptr;
          res[u] = make_uint4(0, 0, 0, 0);
          if (p.residual && ok[u]) res[u] = __ldg(reinterpret_cast<const uint4 *>(p.residual + off[u]));
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          if (!ok[u]) continue;
          const int r = i0 + u * kRPI + lane / kTPR;
          // staged value = fp16-rounded conv output, as in the unfused sequence.  fp16 + fp16 rounded once
          // (HADD2) equals the float add + round of the eager op bit for bit (the float sum is exact unless
          // the exponents differ by > 13, where both round to the larger operand).
          uint4 o = *reinterpret_cast<const uint4 *>(stage + (size_t)r * kRowB + c8 * 2);
          __half2 *oh = reinterpret_cast<__half2 *>(&o);
          const __half2 *rh = reinterpret_cast<const __half2 *>(&res[u]);
          if (p.residual) {
#pragma unroll
            for (int t = 0; t < 4; ++t) oh[t] = __hadd2(oh[t], rh[t]);
          }
          if (p.relu) {
            const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
            for (int t = 0; t < 4; ++t) oh[t] = __hmax2(oh[t], zero);
          }
          if (p.out) *reinterpret_cast<uint4 *>(p.out + off[u]) = o;
          if (pl[u]) *reinterpret_cast<uint4 *>(pl[u]) = o;
        }
      }
    } else {
      // ---- split-K, phase 1 (the pipeline stages are free: acc_bar says every MMA, hence every TMA load, has
      //      completed)
      if (p.work) {
        // through the L2 scratch: half of the tile's columns at a time via per-warp staging rows (35 KB in all,
        // so the 2-stage variants can split too), whole rows per global access
        constexpr int kCP = N_TILE / 2;                      // columns per pass
        constexpr int kRowB = kCP * 4 + 16;
        constexpr int kLPR = kCP / 4, kRowsPI = 32 / kLPR;   // lanes per row (float4 each), rows per access
        uint8_t *stage = smem + (size_t)q * 32 * kRowB;
        const unsigned tile_lin = blockIdx.x + gridDim.x * blockIdx.y;
        float *ws = p.work + ((size_t)(tile_lin * p.splits + blockIdx.z) * kTileM + q * 32) * N_TILE;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
#pragma unroll 1
          for (int c0 = 0; c0 < kCP; c0 += 32) {
            uint32_t acc[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * kCP + c0), acc);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<uint4 *>(stage + (size_t)lane * kRowB + (c0 + j) * 4) =
                  make_uint4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
          }
          __syncwarp();
#pragma unroll 4
          for (int r = lane / kLPR; r < 32; r += kRowsPI) {
            const float4 v = *reinterpret_cast<const float4 *>(stage + (size_t)r * kRowB + (lane % kLPR) * 16);
            *reinterpret_cast<float4 *>(ws + (size_t)r * N_TILE + h * kCP + (lane % kLPR) * 4) = v;
          }
          __syncwarp();
        }
      } else {
        // no scratch: park the partial in this CTA's own shared memory for the DSMEM reduction
        float *part = reinterpret_cast<float *>(smem);
        const int m = q * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < N_TILE; c0 += 32) {
          uint32_t acc[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4 *>(part + (size_t)m * kPartStride<N_TILE> + c0 + j) =
                make_uint4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
      }
    }
    if (threadIdx.x == 64) trace_mark(p, 5);
    tc_fence_before_sync();
  }

  if (is_split) {
    cluster_sync_all();  // every CTA's partial is visible cluster-wide
    if (warp >= 2 && warp < 6) {
      // ---- phase 2: reduce-scatter over the cluster; this CTA owns 128/splits accumulator rows -----
      const uint32_t rank = cluster_ctarank();
      const int rows = kTileM / p.splits;
      constexpr int kThreadsPerRow = N_TILE / 8;            // 8 channels (16 bytes of fp16) per thread
      constexpr int kRowsPerPass = 128 / kThreadsPerRow;
      const int t = threadIdx.x - 64;
      const int c8 = (t % kThreadsPerRow) * 8;
      const uint32_t part_addr = smem_u32(smem);
      const size_t out_base_all = ((size_t)b0 * p.BS_out * p.BS_out + (size_t)r0 * p.BS_out) * p.Cout + n0;
      const int m_valid_all = p.blocks_per_tile == 1 ? kTileM : nvalid * p.BS_out * p.BS_out;
      const unsigned tile_lin = blockIdx.x + gridDim.x * blockIdx.y;
      if (p.work) {
        if (p.splits == 8) splitk_reduce_l2<N_TILE, 8>(p, rank, tile_lin, n0, out_base_all, m_valid_all, row_pl_s, bias_s);
        else if (p.splits == 4) splitk_reduce_l2<N_TILE, 4>(p, rank, tile_lin, n0, out_base_all, m_valid_all, row_pl_s, bias_s);
        else splitk_reduce_l2<N_TILE, 2>(p, rank, tile_lin, n0, out_base_all, m_valid_all, row_pl_s, bias_s);
      } else
      for (int rr = t / kThreadsPerRow; rr < rows; rr += kRowsPerPass) {
        const int m = (int)rank * rows + rr;
        // peers' partials in batches (one DSMEM round trip per 4 peers instead of one per peer), then the sum
        // in rank order: bit-reproducible
        const uint32_t a = part_addr + (uint32_t)(((size_t)m * kPartStride<N_TILE> + c8) * sizeof(float));
        const float *wrow = p.work ? p.work + ((size_t)tile_lin * p.splits * kTileM + m) * N_TILE + c8 : nullptr;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll 1
        for (int z0 = 0; z0 < p.splits; z0 += 4) {  // 4 peers per round trip (register budget: 4 CTAs per SM)
          float4 lo[4], hi[4];
#pragma unroll
          for (int z = 0; z < 4; ++z)
            if (z0 + z < p.splits) {
              if (wrow) {  // L2 only: the peers' stores were released by the cluster barrier
                const float *src = wrow + (size_t)(z0 + z) * kTileM * N_TILE;
                lo[z] = __ldcg(reinterpret_cast<const float4 *>(src));
                hi[z] = __ldcg(reinterpret_cast<const float4 *>(src) + 1);
              } else {
                lo[z] = ld_dsmem_f4(a, (uint32_t)(z0 + z));
                hi[z] = ld_dsmem_f4(a + 16, (uint32_t)(z0 + z));
              }
            }
#pragma unroll
          for (int z = 0; z < 4; ++z)
            if (z0 + z < p.splits) {
              v[0] += lo[z].x; v[1] += lo[z].y; v[2] += lo[z].z; v[3] += lo[z].w;
              v[4] += hi[z].x; v[5] += hi[z].y; v[6] += hi[z].z; v[7] += hi[z].w;
            }
        }
        const long long rp = row_pl_s[m];
        if (m < m_valid_all) {
          const size_t off = out_base_all + (size_t)m * p.Cout + c8;
          epilogue_store8(v, p.bias ? p.bias + n0 + c8 : nullptr, p.residual ? p.residual + off : nullptr, p.relu,
                          p.out ? p.out + off : nullptr, rp >= 0 ? p.plane_out + rp + c8 : nullptr);
        }
      }
    }
    if (!p.work) cluster_sync_all();  // nobody leaves (and frees its shared memory) while peers still read it
  }

  __syncthreads();
  if (mc > 1) cluster_sync_all();  // no CTA leaves while a peer may still arrive on its barriers
  if (threadIdx.x == 0) { trace_mark(p, 6); trace_wall(p, 10); }
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, N_TILE);
  }
}

// ---------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder();  // bc_tma.cu

template <int N_TILE, int STAGES>
static int launch_conv(const CUtensorMap &a_map, const CUtensorMap &b_map, ConvParams &p, int tiles, int ntiles_n,
                       bool allow_split, cudaStream_t s, int force_splits = 0) {
  // split-K over a cluster when the output tiles alone cannot fill the GPU (deep layers: 4..8-px blocks)
  const int total_k = p.ksize * p.ksize * p.kc_per_tap;
  const int ctas = tiles * ntiles_n;
  p.splits = 1;
  p.ksteps_per_split = total_k;
  // Measured (profiles/): the cluster reduction + the per-CTA fixed costs pay off only when the unsplit
  // grid covers less than a third of the GPU; keep the split grid within ~one wave of 2 CTAs per SM.
  static const int max_ctas = getenv("BC_SPLIT_MAX_CTAS") ? atoi(getenv("BC_SPLIT_MAX_CTAS")) : 48;   // experiments
  static const int target = getenv("BC_SPLIT_TARGET") ? atoi(getenv("BC_SPLIT_TARGET")) : 240;
  if (force_splits > 1) {
    p.splits = force_splits;
    p.ksteps_per_split = (total_k + force_splits - 1) / force_splits;
  } else if (allow_split && ctas <= max_ctas && total_k >= 8) {
    const int want = target / ctas;
    int splits = 8;                              // portable cluster size limit
    while (splits > 1 && (splits > want || splits * 2 > total_k)) splits >>= 1;
    while (splits > 1) {
      const int kpp = (total_k + splits - 1) / splits;
      if (kpp * (splits - 1) < total_k) break;   // every CTA of the cluster gets at least one k-step
      splits >>= 1;
    }
    if (splits > 1) {
      p.splits = splits;
      p.ksteps_per_split = (total_k + splits - 1) / splits;
    }
  }
  constexpr size_t smem = (size_t)STAGES * (kABytes + N_TILE * 128) + 1024;
  if (p.splits > 1 && (long long)ctas * p.splits * kTileM * N_TILE * (long long)sizeof(float) > p.work_bytes) p.work = nullptr;
  if (p.work == nullptr &&
      (size_t)kTileM * kPartStride<N_TILE> * sizeof(float) > (size_t)STAGES * (kABytes + N_TILE * 128)) {
    p.splits = 1;  // DSMEM path: the parked accumulator would not fit in this variant's pipeline stages
    p.ksteps_per_split = total_k;
  }
  if (p.splits == 1 || (long long)ctas * p.splits * kTileM * N_TILE * (long long)sizeof(float) > p.work_bytes)
    p.work = nullptr;  // scratch absent or too small: reduce through distributed shared memory
  static cudaError_t attr1 = cudaFuncSetAttribute(conv_igemm_kernel<N_TILE, STAGES, false>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  static cudaError_t attr2 = cudaFuncSetAttribute(conv_igemm_kernel<N_TILE, STAGES, true>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const cudaError_t attr = attr1 != cudaSuccess ? attr1 : attr2;
  BC_REQUIRE(attr == cudaSuccess, (int)attr, "cudaFuncSetAttribute(conv_igemm_kernel): %s", cudaGetErrorString(attr));
  // activation multicast (see ConvParams::mc): the channel slices of a pixel tile as one cluster, when the tile consists of
  // several boxes that divide evenly among them and the grid is big enough to be bound by operand delivery
  // bit 0: activation multicast (default on); bit 1: weight multicast -- bit-identical but MEASURED SLOWER (E = 320: layer #20
  // 127.8 vs 107.7 us, layer1 58.8 vs 49.6, layer2 38.6 vs 32.5: four CTAs in lockstep on a 2-3 stage ring for an operand that
  // is small and L2-hot anyway), so it is an experiment switch only (BC_CONV_MULTICAST=3)
  static const int env_mc = getenv("BC_CONV_MULTICAST") ? atoi(getenv("BC_CONV_MULTICAST")) : 1;
  p.mc = p.mcb = 1;
  // (two channel slices sharing a 2-box tile measured slower than no sharing: 32.5 vs 29.8 us on the 8-px layer at E = 320)
  if ((env_mc & 1) && p.splits == 1 && ntiles_n == 4 && p.blocks_per_tile > 1 && p.blocks_per_tile % ntiles_n == 0 && ctas >= kNumSMs)
    p.mc = ntiles_n;
  // weight multicast: pixel tiles of one channel slice as a cluster along x (tiles whose activations are a single box)
  CUtensorMap bmc_map = b_map;
  if ((env_mc & 2) && p.mc == 1 && p.splits == 1 && p.blocks_per_tile == 1 && ctas >= kNumSMs && p.w_ptr != nullptr) {
    const int mcb = tiles % 4 == 0 ? 4 : (tiles % 2 == 0 ? 2 : 1);
    if (mcb > 1) {
      cuuint64_t gdim[2] = {(cuuint64_t)p.w_K, (cuuint64_t)p.Cout};
      cuuint64_t gstr[1] = {(cuuint64_t)p.w_K * 2};
      cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)(N_TILE / mcb)};
      cuuint32_t estr[2] = {1, 1};
      const CUresult r = tensor_map_encoder()(&bmc_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(p.w_ptr), gdim, gstr,
                                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r == CUDA_SUCCESS) p.mcb = mcb;
    }
  }
  const dim3 cluster = p.mc > 1 ? dim3(1, (unsigned)p.mc, 1) : (p.mcb > 1 ? dim3((unsigned)p.mcb, 1, 1) : dim3(1, 1, (unsigned)p.splits));
  const cudaError_t e = launch_kernel_cluster(
      p.splits > 1 ? conv_igemm_kernel<N_TILE, STAGES, true> : conv_igemm_kernel<N_TILE, STAGES, false>,
      dim3((unsigned)tiles, (unsigned)ntiles_n, (unsigned)p.splits), dim3(kConvThreadsV1), smem, s, cluster, a_map, b_map,
      bmc_map, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail((int)e, "bc_conv_igemm: %s (%s)", cudaGetErrorName(e), cudaGetErrorString(e));
  }
  return check_launch("bc_conv_igemm");
}

int conv_igemm(void *out, const void *plane, const void *weight, const void *bias, const void *residual,
               const int32_t *mapping, int E, int N, int Cin, int H, int W, int BS_in, int Cout, int ksize, int stride,
               int pad, int relu, void *plane_out, const int32_t *out_mapping, int out_N, int out_GH, int out_GW,
               int allow_split_k, void *workspace, long long workspace_bytes, cudaStream_t stream) {
  BC_REQUIRE((out || plane_out) && plane && weight, BC_ERR_NULL, "bc_conv_igemm: NULL pointer");
  BC_REQUIRE(E > 0 && N > 0 && H > 0 && W > 0, BC_ERR_SHAPE, "bc_conv_igemm: empty problem");
  BC_REQUIRE(ksize == 1 || ksize == 3, BC_ERR_UNSUPPORTED, "bc_conv_igemm: kernel size %d (1 or 3)", ksize);
  BC_REQUIRE(stride == 1 || stride == 2, BC_ERR_UNSUPPORTED, "bc_conv_igemm: stride %d (1 or 2)", stride);
  // a 3x3 conv with padding p is the size-preserving DILATED conv with dilation p (taps at -p, 0, +p): the
  // Pedestron backbone's dilated stage (padding 2, dilation 2) takes the same operand path, its halo is 2 pixels
  const int dil = ksize == 3 ? pad : 1;
  BC_REQUIRE((ksize == 1 && pad == 0) || (ksize == 3 && pad >= 1 && pad <= 4), BC_ERR_UNSUPPORTED,
             "bc_conv_igemm: padding %d for kernel %d (1x1: 0; 3x3: p = dilation in 1..4)", pad, ksize);
  BC_REQUIRE(dil == 1 || stride == 1, BC_ERR_UNSUPPORTED, "bc_conv_igemm: dilation %d with stride %d", dil, stride);
  BC_REQUIRE(Cin % kChunkK == 0, BC_ERR_UNSUPPORTED, "bc_conv_igemm: Cin=%d is not a multiple of 64", Cin);
  BC_REQUIRE(Cout % 64 == 0, BC_ERR_UNSUPPORTED, "bc_conv_igemm: Cout=%d is not a multiple of 64", Cout);
  BC_REQUIRE(H % BS_in == 0 && W % BS_in == 0 && BS_in % stride == 0, BC_ERR_SHAPE,
             "bc_conv_igemm: plane %dx%d / block %d / stride %d", H, W, BS_in, stride);
  const int BS_out = BS_in / stride;
  const int px = BS_out * BS_out;
  BC_REQUIRE(BS_out >= 2 && BS_out <= 128 && (px % kTileM == 0 || kTileM % px == 0) && (BS_out & (BS_out - 1)) == 0,
             BC_ERR_UNSUPPORTED, "bc_conv_igemm: output block edge %d (power of two, 2..128)", BS_out);
  BC_REQUIRE((((uintptr_t)out | (uintptr_t)plane | (uintptr_t)weight | (uintptr_t)bias | (uintptr_t)residual) & 15) == 0,
             BC_ERR_ALIGN, "bc_conv_igemm: pointers must be 16-byte aligned");
  EncodeTiledFn enc = tensor_map_encoder();
  BC_REQUIRE(enc != nullptr, BC_ERR_NO_DEVICE, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");

  ConvParams p;
  p.mapping = mapping;
  p.cell = CellDecode(H / BS_in, W / BS_in);
  p.bias = (const __half *)bias;
  p.residual = (const __half *)residual;
  p.out = (__half *)out;
  p.E = E; p.BS_out = BS_out; p.BS_in = BS_in; p.stride = stride; p.pad = pad; p.ksize = ksize; p.Cout = Cout;
  p.dil = dil;
  p.kc_per_tap = Cin / kChunkK;
  p.relu = relu;
  {
    static const char *dbg = getenv("BC_CONV_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
    p.mc = p.mcb = 1;
    p.tma_epi = 0;
    p.w_ptr = weight;
    p.w_K = (long long)ksize * ksize * Cin;
    p.trace = debug_trace_buffer();
  }
  p.work = (((uintptr_t)workspace & 15) == 0 && workspace_bytes > 0) ? (float *)workspace : nullptr;
  p.work_bytes = workspace_bytes;
  p.plane_out = (__half *)plane_out;
  p.out_mapping = out_mapping ? out_mapping : mapping;
  p.out_cell = CellDecode(out_GH > 0 ? out_GH : 1, out_GW > 0 ? out_GW : 1);
  p.out_H = out_GH * BS_out;
  p.out_W = out_GW * BS_out;
  if (plane_out) {
    BC_REQUIRE(p.out_mapping != nullptr && out_GH > 0 && out_GW > 0 && out_N > 0, BC_ERR_NULL,
               "bc_conv_igemm: plane_out needs out_mapping and the output grid");
    BC_REQUIRE(((uintptr_t)plane_out & 15) == 0, BC_ERR_ALIGN, "bc_conv_igemm: plane_out must be 16-byte aligned");
  }
  int tiles;
  if (px >= kTileM) {
    p.blocks_per_tile = 1;
    p.tiles_per_block = px / kTileM;
    p.rows_per_tile = kTileM / BS_out;
    tiles = E * p.tiles_per_block;
  } else {
    p.blocks_per_tile = kTileM / px;
    p.tiles_per_block = 1;
    p.rows_per_tile = BS_out;
    tiles = (E + p.blocks_per_tile - 1) / p.blocks_per_tile;
  }
  p.box_bytes = (uint32_t)(p.rows_per_tile * BS_out * 128);
  p.a3_bytes = 0;
  p.a3_stages = 0;

  // A: the plane (Cin, W, H, N), box = 64 channels x BS_out x rows (sampled every `stride` pixels)
  CUtensorMap a_map, b_map;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)kChunkK, (cuuint32_t)(BS_out * stride), (cuuint32_t)(p.rows_per_tile * stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    const CUresult r = enc(&a_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(plane), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BC_REQUIRE(r == CUDA_SUCCESS, BC_ERR_UNSUPPORTED, "bc_conv_igemm: tensor map (plane) failed: CUresult %d", (int)r);
  }
  static const int env_ntile = getenv("BC_CONV_NTILE") ? atoi(getenv("BC_CONV_NTILE")) : 0;  // experiments
  static const int env_persist = getenv("BC_CONV_PERSIST") ? atoi(getenv("BC_CONV_PERSIST")) : 1;
  int n_tile = env_ntile == 64 ? 64 : (Cout % 128 == 0 ? 128 : 64);
  const int total_k_steps = ksize * ksize * (Cin / kChunkK);
  // Tiny grids (deep layers) split K over a cluster; everything else runs on the persistent kernel, with
  // 64-wide channel slices when 128-wide ones would leave more than half of the SMs without a tile.
  const bool use_split = allow_split_k != 0 && tiles * (Cout / n_tile) <= 48 && total_k_steps >= 8;
  if (!use_split && !env_ntile && n_tile == 128 && tiles * (Cout / 128) <= kNumSMs / 2) n_tile = 64;
  // Measured on B200 (profiles/r01c_conv_persistent.md): with at most one tile per SM the deep operand ring
  // of the persistent kernel wins; with 2+ tiles per SM three co-resident 2-stage CTAs still do better
  // (16.9 vs 18.3 us on the 320-tile decoder conv) -- BC_CONV_PERSIST=2 forces the persistent kernel.
  const bool persistent = !use_split && (env_persist == 2 || (env_persist == 1 && tiles * (Cout / n_tile) <= kNumSMs));
  {
    // B: weights [Cout][tap][Cin] = channels_last (Cout, Cin, k, k) memory; K-major rows
    const cuuint64_t K = (cuuint64_t)ksize * ksize * Cin;
    cuuint64_t gdim[2] = {K, (cuuint64_t)Cout};
    cuuint64_t gstr[1] = {K * 2};
    cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)n_tile};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&b_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(weight), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BC_REQUIRE(r == CUDA_SUCCESS, BC_ERR_UNSUPPORTED, "bc_conv_igemm: tensor map (weights) failed: CUresult %d", (int)r);
  }
  p.d_kc_per_tap = FastDiv((uint32_t)p.kc_per_tap);
  p.d_tiles_per_block = FastDiv((uint32_t)p.tiles_per_block);
  p.d_ntiles_n = FastDiv((uint32_t)(Cout / n_tile));
  p.d_splits = FastDiv(1u);
  // "Shared halo rows" (bc_conv_persist.cu, A3): 3x3 / stride-1 convs on blocks of 16..64 px can take ONE activation
  // box per (kw, chunk) with the three kh taps as descriptor offsets (-25 % operand bytes on 32-px blocks).  Bit-correct
  // (tests/test_gpu_conv.py::test_large_grid_all_launch_forms_agree) but MEASURED SLOWER on B200 (k-loop 560 vs 358
  // clk per k-step on the layer-2 shape, whatever the activation ring depth: profiles/r02_conv_a3.md), so it is an
  // experiment switch only: BC_CONV_A3=1 enables it.
  static const int env_a3 = getenv("BC_CONV_A3") ? atoi(getenv("BC_CONV_A3")) : 0;
  if (env_a3 && env_persist != 0 && !use_split && !env_ntile && ksize == 3 && stride == 1 && dil == 1 && p.blocks_per_tile == 1 &&
      BS_out >= 16 && BS_out <= 64) {
    CUtensorMap a3_map;
    cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)kChunkK, (cuuint32_t)BS_out, (cuuint32_t)(p.rows_per_tile + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(&a3_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(plane), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BC_REQUIRE(r == CUDA_SUCCESS, BC_ERR_UNSUPPORTED, "bc_conv_igemm: tensor map (halo-row box) failed: CUresult %d", (int)r);
    p.a3_bytes = (uint32_t)((p.rows_per_tile + 2) * BS_out * 128);
    p.a3_stages = p.a3_bytes <= 24 * 1024 ? 3 : 2;
    if (getenv("BC_A3_STAGES")) p.a3_stages = atoi(getenv("BC_A3_STAGES"));
    p.tiles_m = tiles;
    p.ntiles_n = Cout / n_tile;
    p.splits = 1;
    p.ksteps_per_split = total_k_steps;
    p.work = nullptr;
    p.tma_epi = 0;
    return launch_conv_persistent(a3_map, b_map, a3_map, a3_map, p, n_tile, stream);
  }
  if (persistent) {
    p.tiles_m = tiles;
    p.ntiles_n = Cout / n_tile;
    p.splits = 1;
    p.ksteps_per_split = total_k_steps;
    p.work = nullptr;
    // Epilogue through TMA stores (BC_CONV_TMA_EPI=0: per-thread stores): a single tile's staging -> global phase took
    // ~2800 clk of per-thread 16-byte stores (23 B/clk per SM; tools/cta_timeline.py) against ~1700 clk for TMEM -> staging
    static const int env_tma_epi = getenv("BC_CONV_TMA_EPI") ? atoi(getenv("BC_CONV_TMA_EPI")) : 1;
    CUtensorMap o_map = a_map, pl_map = a_map;
    p.tma_epi = env_tma_epi && (out || plane_out) && !(p.debug & 2);
    if (p.tma_epi) {
      const int px = BS_out * BS_out;
      p.st_px = px < 32 ? px : 32;
      p.st_bw = BS_out < 32 ? BS_out : 32;
      p.st_bh = p.st_px / p.st_bw;
      if (out) {  // the packed tile batch as (rows = E * BS_out^2, Cout): a warp's 32 accumulator rows are consecutive rows
        cuuint64_t gdim[2] = {(cuuint64_t)Cout, (cuuint64_t)E * px};
        cuuint64_t gstr[1] = {(cuuint64_t)Cout * 2};
        cuuint32_t box[2] = {64u, 32u};
        cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&o_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        BC_REQUIRE(r == CUDA_SUCCESS, BC_ERR_UNSUPPORTED, "bc_conv_igemm: tensor map (tile output) failed: CUresult %d", (int)r);
      }
      if (plane_out) {
        cuuint64_t gdim[4] = {(cuuint64_t)Cout, (cuuint64_t)p.out_W, (cuuint64_t)p.out_H, (cuuint64_t)out_N};
        cuuint64_t gstr[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)p.out_W * Cout * 2, (cuuint64_t)p.out_H * p.out_W * Cout * 2};
        cuuint32_t box[4] = {64u, (cuuint32_t)p.st_bw, (cuuint32_t)p.st_bh, 1u};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = enc(&pl_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, plane_out, gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        BC_REQUIRE(r == CUDA_SUCCESS, BC_ERR_UNSUPPORTED, "bc_conv_igemm: tensor map (plane output) failed: CUresult %d", (int)r);
      }
    }
    return launch_conv_persistent(a_map, b_map, o_map, pl_map, p, n_tile, stream);
  }
  if (use_split && env_persist != 0 && p.work != nullptr) {
    // split-K on the persistent kernel: one (tile, k-range) unit per CTA, the S CTAs of a cluster (S,1,1)
    // share a tile.  S fills ~132 SMs (clusters are GPC-local: 148 cannot be reached with every cluster
    // size) with at least 4 k-steps per CTA; the whole grid is one wave of one CTA per SM.
    static const int target = getenv("BC_SPLIT_TARGET") ? atoi(getenv("BC_SPLIT_TARGET")) : 132;  // experiments
    const int ctas = tiles * (Cout / n_tile);
    int S = target / ctas;
    if (S > 8) S = 8;
    if (S > total_k_steps / 4) S = total_k_steps / 4;
    if (S >= 2 && (long long)ctas * S * kTileM * n_tile * (long long)sizeof(float) <= p.work_bytes) {
      p.tiles_m = tiles;
      p.ntiles_n = Cout / n_tile;
      p.splits = S;
      p.d_splits = FastDiv((uint32_t)S);
      p.ksteps_per_split = (total_k_steps + S - 1) / S;
      p.tma_epi = 0;
      return launch_conv_persistent(a_map, b_map, a_map, a_map, p, n_tile, stream);
    }
  }
  // Variant selection (B200 sweep, profiles/r01b_conv_experiments.md): what counts is how many CTAs an SM can
  // keep in flight -- one CTA's operand stream tops out near 35 B/clk whatever the pipeline depth -- so
  // big grids trade stages for co-residency (2 stages -> 3-4 CTAs/SM, single wave), mid-size grids
  // take 64-wide N tiles to double the CTA count, small grids split K over a cluster.
  static const int env_stages = getenv("BC_CONV_STAGES") ? atoi(getenv("BC_CONV_STAGES")) : 0;  // experiments
  const bool split = allow_split_k != 0;
  if (env_stages || env_ntile) {
    if (n_tile == 128 && env_stages == 6) return launch_conv<128, 6>(a_map, b_map, p, tiles, Cout / 128, split, stream);
    if (n_tile == 128 && env_stages == 2) return launch_conv<128, 2>(a_map, b_map, p, tiles, Cout / 128, split, stream);
    if (n_tile == 64 && env_stages == 2) return launch_conv<64, 2>(a_map, b_map, p, tiles, Cout / 64, split, stream);
    if (n_tile == 128) return launch_conv<128, 3>(a_map, b_map, p, tiles, Cout / 128, split, stream);
    return launch_conv<64, 4>(a_map, b_map, p, tiles, Cout / 64, split, stream);
  }
  if (n_tile == 128) {
    const int ctas = tiles * (Cout / 128);
    // (Tried for the 4-px layers with 8 streams batched -- 160 tiles x 72 k-steps: two k-halves per tile on the
    //  2-stage variant.  The split variant's 100 registers allow only 2 CTAs per SM, 320 CTAs need a second wave:
    //  72 us against 60 us unsplit.  Not used.)
    if (ctas > 2 * kNumSMs) return launch_conv<128, 2>(a_map, b_map, p, tiles, Cout / 128, split, stream);
    if (ctas >= kNumSMs || (split && ctas <= 48))
      return launch_conv<128, 3>(a_map, b_map, p, tiles, Cout / 128, split, stream);
    // 48 < ctas < 148: re-encode the weight map for 64-wide tiles
    cuuint64_t gdim[2] = {(cuuint64_t)ksize * ksize * Cin, (cuuint64_t)Cout};
    cuuint64_t gstr[1] = {gdim[0] * 2};
    cuuint32_t box[2] = {(cuuint32_t)kChunkK, 64u};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&b_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(weight), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BC_REQUIRE(r == CUDA_SUCCESS, BC_ERR_UNSUPPORTED, "bc_conv_igemm: tensor map (weights) failed: CUresult %d", (int)r);
    return launch_conv<64, 4>(a_map, b_map, p, tiles, Cout / 64, split, stream);
  }
  if (tiles * (Cout / 64) > 2 * kNumSMs) return launch_conv<64, 2>(a_map, b_map, p, tiles, Cout / 64, split, stream);
  return launch_conv<64, 4>(a_map, b_map, p, tiles, Cout / 64, split, stream);
}

}  // namespace bc
