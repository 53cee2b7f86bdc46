// bc_conv_persist.cu -- persistent, warp-specialised variant of the tcgen05 implicit-GEMM convolution
// (same math, operands and epilogue semantics as conv_igemm_kernel in bc_conv.cu; see the header there).
//
// Why a second kernel: the per-CTA timelines (profiles/r01c_cta_timeline.md) show that one k-step is a
// fixed-latency round trip -- TMA issue -> L2 -> full barrier -> 4 MMAs -> commit -> empty barrier ->
// next TMA issue, ~1450 clk -- so a CTA's k-loop runs at (stages in flight) / 1450 clk, and the epilogue
// (27-45 % of a CTA's life) leaves both the tensor pipe and the operand ring idle.  This variant
//   * owns the SM alone: 5 (N_TILE 128) or 8 (N_TILE 64) operand stages = 160-192 KB in flight,
//   * loops over output tiles (static round-robin, one CTA per SM), the operand ring running across tiles,
//   * double-buffers the accumulator in TMEM (2 x N_TILE columns): the epilogue warps drain tile i
//     (TMEM -> registers -> bias -> fp16 -> staging rows -> coalesced stores to the tile batch and to
//     the next op's plane) while the MMA warp is already accumulating tile i+1.
// Warp roles: 0 = activation producer (TMA), 6 = weight producer (TMA), 1 = MMA issuer + TMEM owner,
// 2..5 = epilogue (TMEM lane quadrant = warp & 3).  All single-thread loops are division-free.
// Split-K launches (tiny grids) stay on conv_igemm_kernel<.., true>.
#include <cuda.h>
#include <cuda_fp16.h>

#include "bc_conv.cuh"

namespace bc {

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kAProd = 2;                              // activation producer warps
constexpr int kPersistThreads = 224 + 32 * (kAProd - 1);  // warps 0..6 as in conv_igemm_kernel + producers 7..

struct TileCoord {
  int b0, r0, nvalid, n0;
};

// linear tile -> (first packed block, first pixel row, blocks in the tile, first output channel); tiles
// that differ only in the channel slice are neighbours, so concurrently running CTAs share activations in L2
template <int N_TILE>
__device__ __forceinline__ TileCoord tile_coord(const ConvParams &p, int tile) {
  TileCoord t;
  const int m_idx = (int)p.d_ntiles_n.div((uint32_t)tile);
  t.n0 = (tile - m_idx * p.ntiles_n) * N_TILE;
  if (p.blocks_per_tile == 1) {
    t.b0 = (int)p.d_tiles_per_block.div((uint32_t)m_idx);
    t.r0 = (m_idx - t.b0 * p.tiles_per_block) * p.rows_per_tile;
    t.nvalid = 1;
  } else {
    t.b0 = m_idx * p.blocks_per_tile;
    t.r0 = 0;
    t.nvalid = min(p.blocks_per_tile, p.E - t.b0);
  }
  return t;
}

// ---------------------------------------------------------------------------------------------------
// split-K (cluster (S,1,1), one (tile, k-range) unit per CTA): phase 2.  The S partial accumulators of the
// tile are in the L2 scratch ([tile][z][128][N_TILE] fp32); CTA `rank` sums its share of the tile's
// 128 x N_TILE/8 units of 8 channels over the S partials IN RANK ORDER (bit-reproducible) and runs the
// epilogue for them.  All 2*S partial loads and the residual load of up to 8/S units are issued before
// the first use: the reduction costs a couple of L2 round trips.
template <int N_TILE, int S>
__device__ __forceinline__ void splitk_reduce(const ConvParams &p, int rank, int tile, int n0, size_t out_base,
                                              int m_valid, const long long *row_pl_s, const float *bias_s) {
  constexpr int kTPRow = N_TILE / 8, kUnits = kTileM * kTPRow;
  // units per thread and round: the CTA's whole share in ONE round, so that every partial load of the reduction is in
  // flight before the first use (a second round cost another L2 round trip: ~1000 clk of a ~4000 clk tail)
  constexpr int UPB = (kUnits / S + kPersistThreads) / kPersistThreads;
  const int lo_u = rank * kUnits / S, hi_u = (rank + 1) * kUnits / S;
  const int t = threadIdx.x;  // every warp of the CTA takes part: the producers and the MMA warp are idle by now
  const float *ws = p.work + (size_t)tile * S * kTileM * N_TILE;
#pragma unroll 1
  for (int u0 = lo_u + t; u0 < hi_u; u0 += kPersistThreads * UPB) {
    float4 lo[UPB][S], hi[UPB][S];
    uint4 res[UPB];
    int mm[UPB], cc[UPB];
    bool on[UPB];
#pragma unroll
    for (int i = 0; i < UPB; ++i) {
      const int unit = u0 + i * kPersistThreads;
      on[i] = unit < hi_u;
      const int uu = on[i] ? unit : lo_u;
      mm[i] = uu / kTPRow;
      cc[i] = (uu % kTPRow) * 8;
      const float4 *src = reinterpret_cast<const float4 *>(ws + (size_t)mm[i] * N_TILE + cc[i]);
#pragma unroll
      for (int z = 0; z < S; ++z) {  // L2 only: the peers' stores were released by the cluster barrier
        lo[i][z] = __ldcg(src + (size_t)z * (kTileM * N_TILE / 4));
        hi[i][z] = __ldcg(src + (size_t)z * (kTileM * N_TILE / 4) + 1);
      }
      on[i] = on[i] && mm[i] < m_valid;
      res[i] = make_uint4(0, 0, 0, 0);
      if (p.residual && on[i])
        res[i] = __ldg(reinterpret_cast<const uint4 *>(p.residual + out_base + (size_t)mm[i] * p.Cout + cc[i]));
    }
#pragma unroll
    for (int i = 0; i < UPB; ++i) {
      if (!on[i]) continue;
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
      if (p.residual) {  // fp16-rounded conv output + identity, rounded once more
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


// A3 = true: "shared halo rows" (3x3, stride 1, dilation 1, one block per tile, S = 1).  The three taps of a kernel
// column read the same pixels shifted by one image row, so ONE activation box of rows_per_tile + 2 rows per
// (kw, 64-channel chunk) feeds three k-steps: the MMA of tap kh starts kh * BS_out rows (a multiple of the 1024-byte
// swizzle atom) into the box.  Operand bytes per k-step drop from 32 KB to 16 KB + (R + 2) / 3R * 16 KB (32-px blocks:
// 24 KB), i.e. below what the ring sustains (~110 B/clk against the 128 B/clk a k-step needs at the MMA floor,
// profiles/r01c_conv_persistent.md), so the k-loop becomes tensor-pipe bound.  Activations and weights run on
// SEPARATE rings (p.a3_stages boxes of p.a3_bytes, kB3 weight tiles); k-steps are ordered (kw, chunk, kh).
template <int N_TILE> constexpr int kPersistB3 = N_TILE == 128 ? 6 : 8;
constexpr int kPersistA3Max = 3;

template <int N_TILE, int STAGES, bool A3 = false>
__global__ void __launch_bounds__(kPersistThreads, 1)
conv_igemm_persistent_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap b_map,
                             const __grid_constant__ CUtensorMap o_map, const __grid_constant__ CUtensorMap pl_map,
                             const ConvParams p) {
  constexpr uint32_t kBBytes = N_TILE * 128;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr int kRowB = N_TILE * 2 + 16;  // staged output row (+16 B: rows start in different bank groups)
  constexpr int kBStages = A3 ? kPersistB3<N_TILE> : STAGES;  // ring depth of full_bar / empty_bar
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kBStages];   // A3: the weight ring
  __shared__ __align__(8) uint64_t empty_bar[kBStages];
  __shared__ __align__(8) uint64_t a_full[kPersistA3Max];   // A3: the activation ring
  __shared__ __align__(8) uint64_t a_empty[kPersistA3Max];
  __shared__ __align__(8) uint64_t acc_full[2];   // MMA warp -> epilogue: accumulator buffer complete
  __shared__ __align__(8) uint64_t acc_empty[2];  // epilogue -> MMA warp: buffer read out (128 arrivals)
  __shared__ uint32_t tmem_base_slot;
  __shared__ float bias_s[N_TILE];
  __shared__ long long row_pl_s[kTileM];  // see conv_igemm_kernel
  __shared__ int4 blk_coord_s[kMaxBlocksPerTile * kAProd];

  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a3_stride = (p.a3_bytes + 1023u) & ~1023u;
  const int a3_stages = p.a3_stages;
  uint8_t *staging = A3 ? smem + (size_t)a3_stages * a3_stride + (size_t)kBStages * kBBytes : smem + (size_t)STAGES * kStageBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.ntiles_n;
  const int total_k = p.ksize * p.ksize * p.kc_per_tap;
  // work units: (tile, k-range).  S = 1: a unit is a tile and a CTA loops over its tiles.  S > 1 (split-K):
  // the S CTAs of a cluster own the S k-ranges of one tile; grid = units, one unit per CTA.
  const int S = p.splits;
  const int total_units = total_tiles * S;
  const bool tma_epi = !A3 && S == 1 && p.tma_epi != 0;
  if (threadIdx.x == 0) { trace_wall(p, 8); trace_mark(p, 0); }

  if (warp == 0 && lane == 0) {
    prefetch_map(&a_map);
    prefetch_map(&b_map);
    if (p.tma_epi) {
      if (p.out) prefetch_map(&o_map);
      if (p.plane_out) prefetch_map(&pl_map);
    }
    for (int s = 0; s < kBStages; ++s) {
      mbar_init(&full_bar[s], A3 ? 1 : 2);  // activations (warp 0) + weights (warp 6), one arrive.expect_tx each
      mbar_init(&empty_bar[s], 1);
    }
    if (A3)
      for (int s = 0; s < kPersistA3Max; ++s) {
        mbar_init(&a_full[s], 1);
        mbar_init(&a_empty[s], 1);
      }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
    }
    fence_mbar_init();
  }
  // The producers decode their FIRST tile here, ahead of the set-up barrier and of griddepcontrol.wait: the mapping
  // lookup (an L2 round trip) overlaps barrier init / the TMEM allocation / the previous kernel's tail.  `mapping`
  // is written once per frame by bc_compact_mask, which never triggers its dependents early, so it is complete
  // before any kernel behind it starts.  Lane i: mapping lookup + cell decode of block i of the tile.
  const bool is_producer = warp == 0 || warp >= 7;
  int4 *const my_coords = blk_coord_s + kMaxBlocksPerTile * (warp >= 7 ? warp - 6 : 0);
  auto decode_blocks = [&](const TileCoord &t) {
    if (lane < t.nvalid) {
      const uint32_t cell = p.mapping ? (uint32_t)__ldg(p.mapping + t.b0 + lane) : (uint32_t)(t.b0 + lane);
      uint32_t n, gh, gw;
      p.cell(cell, n, gh, gw);
      my_coords[lane] = make_int4((int)gw * p.BS_in - p.pad, (int)gh * p.BS_in + t.r0 * p.stride - p.pad, (int)n, 0);
    }
    __syncwarp();
  };
  struct UnitInfo { int k0, nk; TileCoord t; };
  auto unit_info = [&](int unit) {
    UnitInfo u;
    const int tile = (int)p.d_splits.div((uint32_t)unit), z = unit - tile * S;
    u.k0 = (int)p.d_splits.div((uint32_t)(z * total_k));
    u.nk = (int)p.d_splits.div((uint32_t)((z + 1) * total_k)) - u.k0;
    u.t = tile_coord<N_TILE>(p, tile);
    return u;
  };
  UnitInfo ui;
  ui.k0 = ui.nk = 0;
  ui.t.b0 = ui.t.r0 = ui.t.nvalid = ui.t.n0 = 0;
  if (is_producer) {
    __syncwarp();
    if ((int)blockIdx.x < total_units) {
      ui = unit_info((int)blockIdx.x);
      decode_blocks(ui.t);
    }
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, 2 * N_TILE);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_trigger();
  // barrier init, TMEM alloc and descriptor prefetch above overlapped the previous kernel's tail.  The weight
  // producer (warp 6) does not wait: weights are never written by a kernel of the frame, so its first ring of
  // tiles is in flight while the previous kernel drains (BC_CONV_DEBUG=4 restores the common wait)
  if (warp != 6 || (p.debug & 4)) pdl_wait();
  if (threadIdx.x == 0) trace_mark(p, 1);

  const int a3_groups = p.ksize * p.kc_per_tap;  // (kw, chunk) groups of a tile, three k-steps each
  if (A3 && (warp == 0 || warp >= 7)) {
    // =============================== activation producers (shared halo rows) ======================
    // producer j issues the groups gg = j, j + kAProd, ... of the CTA's global group sequence
    const int j = warp == 0 ? 0 : warp - 6;
    int gg = j, gg_unit0 = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      if (unit != (int)blockIdx.x) {
        ui = unit_info(unit);
        decode_blocks(ui.t);
      }
      if (lane == 0) {
        const int4 c0 = my_coords[0];
        int gi = gg - gg_unit0;  // group within the tile
        int kw = (int)p.d_kc_per_tap.div((uint32_t)gi), cc = gi - kw * p.kc_per_tap;
        for (; gi < a3_groups; gi += kAProd, gg += kAProd) {
          const int s = gg % a3_stages;
          const uint32_t parity = (uint32_t)(((gg / a3_stages) & 1) ^ 1);
          mbar_wait(&a_empty[s], parity);
          mbar_expect_tx(&a_full[s], p.a3_bytes);
          if (gg == 0) trace_mark(p, 12);
          tma_load_4d(smem + (size_t)s * a3_stride, &a_map, &a_full[s], cc * kChunkK, c0.x + kw, c0.y, c0.z);
#pragma unroll
          for (int u = 0; u < kAProd; ++u)
            if (++cc == p.kc_per_tap) { cc = 0; ++kw; }
        }
      }
      gg_unit0 += a3_groups;
      __syncwarp();
    }
  } else if (A3 && warp == 6) {
    // =============================== weight producer (k-steps ordered (kw, chunk, kh)) ============
    if (lane == 0) {
      int s = 0;
      uint32_t parity = 1;
      uint8_t *const sb0 = smem + (size_t)a3_stages * a3_stride;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const int n0 = (unit - (int)p.d_ntiles_n.div((uint32_t)unit) * p.ntiles_n) * N_TILE;
        int kw = 0, cc = 0;
        for (int gi = 0; gi < a3_groups; ++gi) {
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh) {
            mbar_wait(&empty_bar[s], parity);
            mbar_expect_tx(&full_bar[s], kBBytes);
            tma_load_2d(sb0 + (size_t)s * kBBytes, &b_map, &full_bar[s], ((kh * 3 + kw) * p.kc_per_tap + cc) * kChunkK, n0);
            if (++s == kBStages) { s = 0; parity ^= 1; }
          }
          if (++cc == p.kc_per_tap) { cc = 0; ++kw; }
        }
      }
    }
  } else if (A3 && warp == 1) {
    // =============================== MMA issuer (shared halo rows) ================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kTileM, N_TILE);
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(smem));
      const uint64_t b_desc0 = umma_desc_sw128(smem_u32(smem) + (uint32_t)a3_stages * a3_stride);
      const uint32_t tap_off = (uint32_t)(p.BS_out * 128) >> 4;  // one image row of the box, in 16-byte units
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      uint32_t buf = 0, buf_parity = 1;
      bool first = true;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        mbar_wait(&acc_empty[buf], buf_parity);
        tc_fence_after_sync();
        const uint32_t acc = tmem_base + buf * N_TILE;
        for (int gi = 0; gi < a3_groups; ++gi) {
          mbar_wait(&a_full[sa], pa);
          if (first) { trace_mark(p, 2); first = false; }
          const uint64_t a_desc = a_desc0 + (uint64_t)sa * (a3_stride >> 4);
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh) {
            mbar_wait(&full_bar[sb], pb);
            tc_fence_after_sync();
            const uint64_t b_desc = b_desc0 + (uint64_t)sb * (kBBytes >> 4);
#pragma unroll
            for (int k = 0; k < kChunkK / 16; ++k)
              umma_f16_ss(acc, a_desc + kh * tap_off + 2 * k, b_desc + 2 * k, idesc, (uint32_t)((gi | kh | k) != 0));
            umma_commit(&empty_bar[sb]);
            if (++sb == kBStages) { sb = 0; pb ^= 1; }
          }
          umma_commit(&a_empty[sa]);  // frees the activation box once its three taps have been read
          if (++sa == a3_stages) { sa = 0; pa ^= 1; }
        }
        umma_commit(&acc_full[buf]);
        if (buf == 1) buf_parity ^= 1;
        buf ^= 1;
      }
      trace_mark(p, 3);
    }
  } else if (A3 && (warp < 2 || warp >= 6)) {
  } else
  if (warp == 0 || warp >= 7) {
    // =============================== activation producers =========================================
    // kAProd warps (0, 7, ...), one elected lane each; producer j issues the k-steps g = j, j + kAProd, ...
    // of the CTA's global k-step sequence (the ring position follows from g alone)
    // Per unit the WHOLE warp decodes the tile (lane i: mapping lookup + cell decode of block i, all in flight at
    // once), then lane 0 alone issues the loads.  The decode used to be a serial single-thread prologue of ~1900-2600 clk
    // before the first load (profiles/r02_conv_prologue.md).
    const int j = warp == 0 ? 0 : warp - 6;
    int g = j;  // global k-step (over all tiles of this CTA) this producer issues next
    int g_unit0 = 0;
    int4 *coords = my_coords;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      if (unit != (int)blockIdx.x) {  // (the first unit was decoded ahead of the set-up barrier)
        ui = unit_info(unit);
        decode_blocks(ui.t);
      }
      const int k0 = ui.k0, nk = ui.nk;
      const TileCoord t = ui.t;
      if (lane == 0) {
        const int4 c0 = coords[0];
        const int cx0 = c0.x, cy0 = c0.y, cn0 = c0.z;
        const uint32_t tx_bytes = (uint32_t)t.nvalid * p.box_bytes;
        // (tap, channel chunk) of this producer's first k-step in the tile
        int ks = g - g_unit0;
        int tap = (int)p.d_kc_per_tap.div((uint32_t)(k0 + ks)), cc = (k0 + ks) - tap * p.kc_per_tap;
        int kh = p.ksize == 1 ? 0 : (tap >= 6 ? 2 : (tap >= 3 ? 1 : 0));
        int kw = tap - kh * p.ksize;
        for (; ks < nk; ks += kAProd, g += kAProd) {
          const int s = g % STAGES;
          const uint32_t parity = (uint32_t)(((g / STAGES) & 1) ^ 1);
          uint8_t *sa = smem + (size_t)s * kStageBytes;
          mbar_wait(&empty_bar[s], parity);
          mbar_expect_tx(&full_bar[s], tx_bytes);
          if (g == 0) trace_mark(p, 12);  // about to issue the first activation load
          if (t.nvalid == 1) {
            tma_load_4d(sa, &a_map, &full_bar[s], cc * kChunkK, cx0 + kw * p.dil, cy0 + kh * p.dil, cn0);
          } else {
            for (int i = 0; i < t.nvalid; ++i) {
              const int4 c = coords[i];
              tma_load_4d(sa + (size_t)i * p.box_bytes, &a_map, &full_bar[s], cc * kChunkK, c.x + kw * p.dil, c.y + kh * p.dil, c.z);
            }
          }
#pragma unroll
          for (int u = 0; u < kAProd; ++u)
            if (++cc == p.kc_per_tap) {
              cc = 0;
              if (++kw == p.ksize) { kw = 0; ++kh; }
            }
        }
      }
      g_unit0 += nk;  // (g and g_unit0 matter to lane 0 only; it advanced g inside its loop)
      __syncwarp();  // lane 0 is done with coords[] before the next unit's decode overwrites it
    }
  } else if (warp == 6) {
    // =============================== weight producer ==============================================
    if (lane == 0) {
      int s = 0;
      uint32_t parity = 1;
      uint8_t *sb = smem + kABytes;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const int tile = (int)p.d_splits.div((uint32_t)unit), z = unit - tile * S;
        const int k0 = (int)p.d_splits.div((uint32_t)(z * total_k)), nk = (int)p.d_splits.div((uint32_t)((z + 1) * total_k)) - k0;
        const int n0 = (tile - (int)p.d_ntiles_n.div((uint32_t)tile) * p.ntiles_n) * N_TILE;
        int kcoord = k0 * kChunkK;
        for (int ks = 0; ks < nk; ++ks) {
          mbar_wait(&empty_bar[s], parity);
          mbar_expect_tx(&full_bar[s], kBBytes);
          tma_load_2d(sb, &b_map, &full_bar[s], kcoord, n0);
          kcoord += kChunkK;
          sb += kStageBytes;
          if (++s == STAGES) { s = 0; parity ^= 1; sb = smem + kABytes; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===================================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kTileM, N_TILE);
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(smem)), b_desc0 = umma_desc_sw128(smem_u32(smem) + kABytes);
      int s = 0;
      uint32_t parity = 0, stage_off = 0;
      uint32_t buf = 0, buf_parity = 1;  // first use of either accumulator buffer: it is free
      bool first = true;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const int z = unit - (int)p.d_splits.div((uint32_t)unit) * S;
        const int nk = (int)p.d_splits.div((uint32_t)((z + 1) * total_k)) - (int)p.d_splits.div((uint32_t)(z * total_k));
        mbar_wait(&acc_empty[buf], buf_parity);
        tc_fence_after_sync();
        const uint32_t acc = tmem_base + buf * N_TILE;
        for (int ks = 0; ks < nk; ++ks) {
          mbar_wait(&full_bar[s], parity);
          if (first) { trace_mark(p, 2); first = false; }
          tc_fence_after_sync();
#pragma unroll
          for (int k = 0; k < kChunkK / 16; ++k)
            umma_f16_ss(acc, a_desc0 + stage_off + 2 * k, b_desc0 + stage_off + 2 * k, idesc, (uint32_t)((ks | k) != 0));
          umma_commit(&empty_bar[s]);  // frees the stage once these MMAs have read it
          stage_off += kStageBytes >> 4;
          if (++s == STAGES) { s = 0; parity ^= 1; stage_off = 0; }
        }
        umma_commit(&acc_full[buf]);  // accumulator of this tile complete
        if (buf == 1) buf_parity ^= 1;
        buf ^= 1;
      }
      trace_mark(p, 3);
    }
  } else if (warp >= 2 && warp < 6) {
    // =============================== epilogue =====================================================
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int t128 = threadIdx.x - 64;
    uint8_t *stage = staging + (size_t)q * 32 * kRowB;
    constexpr int kTPR = N_TILE / 8, kRPI = 32 / kTPR;
    const int c8 = (lane % kTPR) * 8;
    uint32_t buf = 0, buf_parity = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      const int tile = (int)p.d_splits.div((uint32_t)unit);
      const TileCoord t = tile_coord<N_TILE>(p, tile);
      // this tile's slice of the bias + this warp's 32 plane-row offsets, while the MMAs of the tile run
      asm volatile("bar.sync 1, 128;" ::: "memory");  // previous tile's readers of bias_s are done
      for (int c = t128; c < N_TILE; c += 128) bias_s[c] = p.bias ? __half2float(__ldg(p.bias + t.n0 + c)) : 0.f;
      int bx = 0, by = 0, bn = -1;  // TMA epilogue: plane coordinates of store box `lane` of this warp (bn < 0: none)
      if (tma_epi) {
        if (p.plane_out && lane * p.st_px < 32) {
          int blk, y, x;
          pixel_of_row(p, q * 32 + lane * p.st_px, t.r0, blk, y, x);
          if (blk < t.nvalid) {
            uint32_t n, gh, gw;
            p.out_cell((uint32_t)__ldg(p.out_mapping + t.b0 + blk), n, gh, gw);
            bx = (int)gw * p.BS_out + x; by = (int)gh * p.BS_out + y; bn = (int)n;
          }
        }
      } else {
        const int m = q * 32 + lane;
        int blk, y, x;
        pixel_of_row(p, m, t.r0, blk, y, x);
        row_pl_s[m] = (blk < t.nvalid && p.plane_out) ? (long long)(plane_row(p, t.b0 + blk, y, x) - p.plane_out) + t.n0 : -1ll;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const size_t out_base = ((size_t)t.b0 * p.BS_out * p.BS_out + (size_t)t.r0 * p.BS_out) * p.Cout + t.n0;
      const int m_valid = p.blocks_per_tile == 1 ? kTileM : t.nvalid * p.BS_out * p.BS_out;

      mbar_wait(&acc_full[buf], buf_parity);
      tc_fence_after_sync();
      if (S == 1 && threadIdx.x == 64) trace_mark(p, 4);  // accumulator complete (S > 1 marks slot 4 below)
      if (S > 1) {
        // ---- split-K, phase 1: this CTA's fp32 partial -> L2 scratch, whole rows per access (half of the
        //      tile's columns at a time through the staging rows)
        constexpr int kCP = N_TILE / 2;                      // columns per pass
        constexpr int kLPR = kCP / 4, kRowsPI = 32 / kLPR;   // lanes per row (float4 each), rows per access
        float *ws = p.work + ((size_t)unit * kTileM + q * 32) * N_TILE;
        if (threadIdx.x == 64) trace_mark(p, 4);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
#pragma unroll 1
          for (int c0 = 0; c0 < kCP; c0 += 32) {
            uint32_t v32[32];
            tmem_ld_32x32(tmem_base + buf * N_TILE + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * kCP + c0), v32);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<uint4 *>(stage + (size_t)lane * kRowB + (c0 + j) * 4) =
                  make_uint4(v32[j], v32[j + 1], v32[j + 2], v32[j + 3]);
          }
          __syncwarp();
#pragma unroll 4
          for (int r = lane / kLPR; r < 32; r += kRowsPI) {
            const float4 v = *reinterpret_cast<const float4 *>(stage + (size_t)r * kRowB + (lane % kLPR) * 16);
            *reinterpret_cast<float4 *>(ws + (size_t)r * N_TILE + h * kCP + (lane % kLPR) * 4) = v;
          }
          __syncwarp();
        }
        continue;  // one unit per CTA in split mode; phase 2 follows the cluster barrier below
      }
      const uint32_t acc = tmem_base + buf * N_TILE + ((uint32_t)(q * 32) << 16);
      if (tma_epi) {
        // ---- TMA epilogue: TMEM -> + bias -> fp16 (+ residual, ReLU) in registers, one accumulator row per thread, into
        //      128-byte-swizzled staging rows (per 64-channel half: 32 rows x 128 B); then one bulk tensor store per half
        //      to the tile batch and one per (half, store box) to the next op's plane: no per-thread global stores
        constexpr int kHalves = N_TILE / 64;
        uint8_t *st = staging + (size_t)q * kHalves * 4096;
        const int m = q * 32 + lane;
        const bool row_ok = m < m_valid;
        const __half *res_row = (p.residual && row_ok) ? p.residual + out_base + (size_t)m * p.Cout : nullptr;
        tma_wait_read<0>();  // this lane's stores of the previous tile have read the staging rows
        __syncwarp();
        const int row0 = t.b0 * p.BS_out * p.BS_out + t.r0 * p.BS_out + q * 32;
        // 32 accumulator columns: (+ bias) -> fp16 (+ residual, ReLU) -> swizzled staging row of this thread
        auto finish32 = [&](const uint32_t (&v32)[32], const uint4 (&rr)[4], int c0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              oh[u] = __floats2half2_rn(__uint_as_float(v32[8 * j + 2 * u]) + bias_s[c0 + 8 * j + 2 * u],
                                        __uint_as_float(v32[8 * j + 2 * u + 1]) + bias_s[c0 + 8 * j + 2 * u + 1]);
            if (p.residual) {  // fp16-rounded conv output + identity, rounded once more (HADD2 == float add + round)
              const __half2 *rh = reinterpret_cast<const __half2 *>(&rr[j]);
#pragma unroll
              for (int u = 0; u < 4; ++u) oh[u] = __hadd2(oh[u], rh[u]);
            }
            if (p.relu) {
              const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
              for (int u = 0; u < 4; ++u) oh[u] = __hmax2(oh[u], zero);
            }
            const int c = c0 + 8 * j;
            *reinterpret_cast<uint4 *>(st + (c >> 6) * 4096 + lane * 128 + ((((c & 63) >> 3) ^ (lane & 7)) << 4)) = o;
          }
        };
        auto load_res = [&](uint4 (&rr)[4], int c0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rr[j] = res_row ? __ldg(reinterpret_cast<const uint4 *>(res_row + c0 + 8 * j)) : make_uint4(0, 0, 0, 0);
        };
        // software pipeline over 32-column groups: the TMEM load of group g + 1 is in flight while group g is converted;
        // every finished 64-channel half leaves at once (its stores overlap the conversion of the next half)
        uint32_t va[32], vb[32];
        uint4 ra[4], rb[4];
        load_res(ra, 0);
        tmem_ld_32x32(acc, va);
        bool issued = false;
#pragma unroll
        for (int h = 0; h < kHalves; ++h) {
          const int c0 = 64 * h;
          tmem_ld_wait();
          load_res(rb, c0 + 32);
          tmem_ld_32x32(acc + (uint32_t)(c0 + 32), vb);
          finish32(va, ra, c0);
          tmem_ld_wait();
          if (h + 1 < kHalves) {
            load_res(ra, c0 + 64);
            tmem_ld_32x32(acc + (uint32_t)(c0 + 64), va);
          } else {
            tc_fence_before_sync();  // the accumulator buffer is in registers: hand it back to the MMA warp
            mbar_arrive(&acc_empty[buf]);
          }
          finish32(vb, rb, c0 + 32);
          fence_proxy_async_smem();
          __syncwarp();
          if (p.out && lane == 0) { tma_store_2d(&o_map, st + h * 4096, t.n0 + c0, row0); issued = true; }
          if (bn >= 0) { tma_store_4d(&pl_map, st + h * 4096 + lane * p.st_px * 128, t.n0 + c0, bx, by, bn); issued = true; }
        }
        if (issued) tma_commit();
        if (threadIdx.x == 64) trace_mark(p, 7);
        if (buf == 1) buf_parity ^= 1;
        buf ^= 1;
        continue;
      }
      // ---- phase A: TMEM -> + bias -> fp16, one accumulator row per thread, into this warp's staging rows
#pragma unroll 1
      for (int c0 = 0; c0 < N_TILE; c0 += 32) {
        uint32_t v32[32];
        tmem_ld_32x32(acc + (uint32_t)c0, v32);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 o;
          __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            oh[u] = __floats2half2_rn(__uint_as_float(v32[j + 2 * u]) + bias_s[c0 + j + 2 * u],
                                      __uint_as_float(v32[j + 2 * u + 1]) + bias_s[c0 + j + 2 * u + 1]);
          *reinterpret_cast<uint4 *>(stage + (size_t)lane * kRowB + (c0 + j) * 2) = o;
        }
      }
      // the accumulator buffer is read out: hand it back to the MMA warp before the stores
      tc_fence_before_sync();
      mbar_arrive(&acc_empty[buf]);
      __syncwarp();
      if (threadIdx.x == 64) trace_mark(p, 7);  // phase A done (TMEM -> staging rows)
      // ---- phase B: kTPR lanes cover one pixel's N_TILE channels (16 B each): (+ residual) -> ReLU ->
      //      whole 32-byte sectors to the tile batch and to the next op's plane
      constexpr int kBatch = 4;
#pragma unroll 1
      for (int i0 = 0; i0 < 32; i0 += kRPI * kBatch) {
        size_t off[kBatch];
        __half *pl[kBatch];
        uint4 res[kBatch];
        bool ok[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {  // addresses + residual loads of the whole batch first
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
          uint4 o = *reinterpret_cast<const uint4 *>(stage + (size_t)r * kRowB + c8 * 2);
          __half2 *oh = reinterpret_cast<__half2 *>(&o);
          const __half2 *rh = reinterpret_cast<const __half2 *>(&res[u]);
          if (p.residual) {  // fp16-rounded conv output + identity, rounded once more (HADD2 == float add + round)
#pragma unroll
            for (int k = 0; k < 4; ++k) oh[k] = __hadd2(oh[k], rh[k]);
          }
          if (p.relu) {
            const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
            for (int k = 0; k < 4; ++k) oh[k] = __hmax2(oh[k], zero);
          }
          if (p.out) *reinterpret_cast<uint4 *>(p.out + off[u]) = o;
          if (pl[u]) *reinterpret_cast<uint4 *>(pl[u]) = o;
        }
      }
      __syncwarp();  // staging rows are rewritten by the next tile's phase A
      if (buf == 1) buf_parity ^= 1;
      buf ^= 1;
    }
    if (tma_epi) tma_wait_read<0>();  // the staging rows must outlive the stores that read them
    if (threadIdx.x == 64) trace_mark(p, 5);
    tc_fence_before_sync();
  }

  if (S > 1) {
    if (threadIdx.x == 0) trace_mark(p, 13);
    cluster_sync_all();  // every CTA's partial is in L2 and visible cluster-wide (release / acquire)
    if (threadIdx.x == 0) trace_mark(p, 14);
    {
      const int tile = (int)p.d_splits.div(blockIdx.x), rank = (int)blockIdx.x - tile * S;  // cluster (S,1,1): rank = blockIdx.x % S
      const TileCoord t = tile_coord<N_TILE>(p, tile);
      const size_t out_base = ((size_t)t.b0 * p.BS_out * p.BS_out + (size_t)t.r0 * p.BS_out) * p.Cout + t.n0;
      const int m_valid = p.blocks_per_tile == 1 ? kTileM : t.nvalid * p.BS_out * p.BS_out;
      switch (S) {
        case 2: splitk_reduce<N_TILE, 2>(p, rank, tile, t.n0, out_base, m_valid, row_pl_s, bias_s); break;
        case 3: splitk_reduce<N_TILE, 3>(p, rank, tile, t.n0, out_base, m_valid, row_pl_s, bias_s); break;
        case 4: splitk_reduce<N_TILE, 4>(p, rank, tile, t.n0, out_base, m_valid, row_pl_s, bias_s); break;
        case 5: splitk_reduce<N_TILE, 5>(p, rank, tile, t.n0, out_base, m_valid, row_pl_s, bias_s); break;
        case 6: splitk_reduce<N_TILE, 6>(p, rank, tile, t.n0, out_base, m_valid, row_pl_s, bias_s); break;
        case 7: splitk_reduce<N_TILE, 7>(p, rank, tile, t.n0, out_base, m_valid, row_pl_s, bias_s); break;
        default: splitk_reduce<N_TILE, 8>(p, rank, tile, t.n0, out_base, m_valid, row_pl_s, bias_s); break;
      }
    }
  }

  __syncthreads();
  if (threadIdx.x == 0) { trace_mark(p, 6); trace_wall(p, 10); }
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 2 * N_TILE);
  }
}

template <int N_TILE, int STAGES>
static int launch_persistent(const CUtensorMap &a_map, const CUtensorMap &b_map, const CUtensorMap &o_map,
                             const CUtensorMap &pl_map, const ConvParams &p, cudaStream_t s) {
  constexpr size_t smem = (size_t)STAGES * (kABytes + N_TILE * 128) + 4 * 32 * (N_TILE * 2 + 16) + 1024;
  static_assert(smem <= 227 * 1024 - 4096, "operand ring + staging must fit one SM");
  static cudaError_t attr = cudaFuncSetAttribute(conv_igemm_persistent_kernel<N_TILE, STAGES>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  BC_REQUIRE(attr == cudaSuccess, (int)attr, "cudaFuncSetAttribute(conv_igemm_persistent_kernel): %s",
             cudaGetErrorString(attr));
  const int total = p.tiles_m * p.ntiles_n;
  const unsigned grid = p.splits > 1 ? (unsigned)(total * p.splits) : (unsigned)(total < kNumSMs ? total : kNumSMs);
  const cudaError_t e = launch_kernel_cluster(conv_igemm_persistent_kernel<N_TILE, STAGES>, dim3(grid), dim3(kPersistThreads),
                                              smem, s, dim3((unsigned)p.splits, 1, 1), a_map, b_map, o_map, pl_map, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail((int)e, "bc_conv_igemm: %s (%s)", cudaGetErrorName(e), cudaGetErrorString(e));
  }
  return check_launch("bc_conv_igemm");
}

// shared-halo-rows form (p.a3_bytes / p.a3_stages set by the caller; S = 1)
template <int N_TILE>
static int launch_persistent_a3(const CUtensorMap &a3_map, const CUtensorMap &b_map, const ConvParams &p, cudaStream_t s) {
  const size_t a3_stride = ((size_t)p.a3_bytes + 1023) & ~(size_t)1023;
  const size_t smem = (size_t)p.a3_stages * a3_stride + (size_t)kPersistB3<N_TILE> * N_TILE * 128 + 4 * 32 * (N_TILE * 2 + 16) + 1024;
  constexpr size_t kMax = 227 * 1024 - 4096;
  static cudaError_t attr = cudaFuncSetAttribute(conv_igemm_persistent_kernel<N_TILE, 2, true>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMax);
  BC_REQUIRE(attr == cudaSuccess, (int)attr, "cudaFuncSetAttribute(conv_igemm_persistent_kernel a3): %s", cudaGetErrorString(attr));
  BC_REQUIRE(smem <= kMax && p.a3_stages >= 1 && p.a3_stages <= kPersistA3Max && p.splits == 1, BC_ERR_UNSUPPORTED,
             "bc_conv_igemm: halo-row ring of %zu bytes", smem);
  const int total = p.tiles_m * p.ntiles_n;
  const unsigned grid = (unsigned)(total < kNumSMs ? total : kNumSMs);
  const cudaError_t e = launch_kernel_cluster(conv_igemm_persistent_kernel<N_TILE, 2, true>, dim3(grid), dim3(kPersistThreads),
                                              smem, s, dim3(1, 1, 1), a3_map, b_map, a3_map, a3_map, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail((int)e, "bc_conv_igemm: %s (%s)", cudaGetErrorName(e), cudaGetErrorString(e));
  }
  return check_launch("bc_conv_igemm");
}

int launch_conv_persistent(const CUtensorMap &a_map, const CUtensorMap &b_map, const CUtensorMap &o_map,
                           const CUtensorMap &pl_map, const ConvParams &p, int n_tile, cudaStream_t s) {
  if (p.a3_bytes) return n_tile == 128 ? launch_persistent_a3<128>(a_map, b_map, p, s) : launch_persistent_a3<64>(a_map, b_map, p, s);
  return n_tile == 128 ? launch_persistent<128, 5>(a_map, b_map, o_map, pl_map, p, s)
                       : launch_persistent<64, 8>(a_map, b_map, o_map, pl_map, p, s);
}

}  // namespace bc
