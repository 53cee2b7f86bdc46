// bc_conv.cuh -- pieces shared by the tcgen05 convolution kernels (bc_conv.cu, bc_stem.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "bc_common.cuh"
#include "bc_ptx.cuh"

namespace bc {

constexpr int kConvThreads = 192;      // stem kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int kConvThreadsV1 = 224;    // conv_igemm: + a second TMA producer warp (weights)
constexpr int kTileM = 128;
constexpr int kMaxBlocksPerTile = 32;          // 2-px output blocks: 128 / 4 blocks share an accumulator tile
constexpr int kChunkK = 64;                    // channels per k-step = one 128-byte swizzled row
constexpr uint32_t kABytes = kTileM * 128;     // 16 KB per stage

struct ConvParams {
  const int32_t *mapping;   // cell of packed tile b; nullptr = identity (input is a packed tile batch)
  CellDecode cell;          // grid of the INPUT plane
  const __half *bias;       // [Cout] or nullptr
  const __half *residual;   // same layout as out, or nullptr
  __half *out;              // (E, BS_out, BS_out, Cout) NHWC, or nullptr when only plane_out is wanted
  int E, BS_out, BS_in, stride, pad, ksize, Cout;
  int dil;                  // tap spacing (3x3: = pad, the size-preserving dilated conv; 1 otherwise)
  int kc_per_tap;           // Cin / 64
  int rows_per_tile;        // rows of BS_out pixels of ONE block in a tile (BS_out >= 16) or BS_out
  int blocks_per_tile;      // 1, or 128 / BS_out^2 for small blocks
  int tiles_per_block;      // BS_out^2 / 128 for big blocks, else 1
  int tiles_m, ntiles_n;    // persistent kernel: number of 128-pixel tiles / of N_TILE-channel slices
  // persistent kernel: exact divisions by these run-time constants as one mul.hi each (the single-thread producer
  // prologue used to spend ~1000 clk in six 32-bit integer divisions before its first load, profiles/r02_conv_prologue.md)
  FastDiv d_ntiles_n, d_tiles_per_block, d_splits, d_kc_per_tap;
  int relu;
  uint32_t box_bytes;       // bytes one A box (one block's share of the tile) occupies in smem
  uint32_t a3_bytes;        // shared-halo-rows mode: bytes of one (rows_per_tile + 2)-row activation box (0: off)
  int a3_stages;            // ... and the depth of the activation ring (persistent kernel)
  // optional second destination: the next padded op's persistent plane (N, GH*BS_out, GW*BS_out, Cout)
  __half *plane_out;
  const int32_t *out_mapping;  // cell of packed tile b in the OUTPUT grid (== mapping unless mapping is null)
  CellDecode out_cell;
  int out_H, out_W;
  // split-K: the `splits` CTAs of one thread-block CLUSTER (1,1,splits) share an output tile.  Each
  // keeps its partial accumulator (fp32) in its own shared memory; after a cluster barrier CTA r
  // reduces rows [r*128/splits, ...) over all peers through distributed shared memory, in rank
  // order (deterministic), and runs the epilogue for those rows.
  int splits, ksteps_per_split;
  // split-K through L2 instead: partials go to this global scratch ([tile][split][128][N_TILE] fp32), cluster
  // barrier, CTA r sums its rows from there.  DSMEM moves ~20 B/clk per SM, L2 several times that.
  float *work;
  long long work_bytes;
  unsigned long long *trace;  // bc_debug_trace buffer (16 words per CTA) or nullptr
  // persistent kernel, S = 1: the epilogue leaves through TMA stores (128-byte-swizzled staging rows -> o_map / pl_map)
  // instead of per-thread global stores.  One store box of the plane = st_px consecutive tile rows = st_bw x st_bh pixels.
  int tma_epi, st_px, st_bw, st_bh;
  // one-tile kernel: the `mc` CTAs of a cluster (1, mc, 1) compute the mc channel slices of ONE pixel tile; each loads
  // 1/mc of the tile's activation boxes and multicasts them to all (TMA .multicast::cluster), so the activations cross
  // the L2 -> SM path once per cluster instead of once per CTA.  1 = off.
  int mc;
  // ... or the `mcb` CTAs of a cluster (mcb, 1, 1) compute mcb PIXEL tiles of one channel slice and share the WEIGHT tile:
  // each loads N_TILE / mcb of its rows (bmc_map) and multicasts them.  At most one of mc / mcb is > 1.
  int mcb;
  const void *w_ptr;        // host side only: weights and their row length (elements), for the multicast weight map
  long long w_K;
  int debug;  // BC_CONV_DEBUG (timing experiments only): bit 0 one k-step, bit 1 no epilogue stores, bit 2 weight producer waits for the previous kernel too
};

// in-kernel timeline for profiles/ (bc_debug_trace): slot k of this CTA's 16-word record <- SM clock
__device__ __forceinline__ void trace_mark(const ConvParams &p, int k) {
  if (p.trace) {
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    p.trace[(size_t)cta * 16 + k] = (unsigned long long)clock64();
  }
}
__device__ __forceinline__ void trace_wall(const ConvParams &p, int k) {  // slot k <- %globaltimer (ns), k+1 <- SM id
  if (p.trace) {
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    unsigned long long t;
    unsigned sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    p.trace[(size_t)cta * 16 + k] = t;
    p.trace[(size_t)cta * 16 + k + 1] = sm;
  }
}
unsigned long long *debug_trace_buffer();  // bc_api.cu
int launch_conv_persistent(const CUtensorMap &a_map, const CUtensorMap &b_map, const CUtensorMap &o_map,
                           const CUtensorMap &pl_map, const ConvParams &p, int n_tile, cudaStream_t s);  // bc_conv_persist.cu

template <int N_TILE> constexpr int kPartStride = N_TILE + 4;  // floats per parked accumulator row (+4: bank spread)

// accumulator row m of a tile -> (block within the tile, y, x) of the output pixel
__device__ __forceinline__ void pixel_of_row(const ConvParams &p, int m, int r0, int &blk, int &y, int &x) {
  if (p.blocks_per_tile == 1) {
    blk = 0;
    y = r0 + m / p.BS_out;
    x = m % p.BS_out;
  } else {
    const int per = p.BS_out * p.BS_out;
    blk = m / per;
    const int rem = m - blk * per;
    y = rem / p.BS_out;
    x = rem - y * p.BS_out;
  }
}

// address of channel 0 of pixel (y, x) of packed tile b in the next op's plane
__device__ __forceinline__ __half *plane_row(const ConvParams &p, int b, int y, int x) {
  uint32_t n, gh, gw;
  p.out_cell((uint32_t)__ldg(p.out_mapping + b), n, gh, gw);
  return p.plane_out + (((size_t)n * p.out_H + gh * p.BS_out + y) * p.out_W + gw * p.BS_out + x) * p.Cout;
}

// bias -> (round, + residual) -> ReLU -> fp16, for 8 consecutive channels of one output pixel; stores
// to the packed tile batch and, if given, to the next op's plane
__device__ __forceinline__ void epilogue_store8(float (&v)[8], const __half *bias8, const __half *res8, int relu,
                                                __half *out8, __half *plane8) {
  if (bias8) {
    const uint4 bb = __ldg(reinterpret_cast<const uint4 *>(bias8));
    const __half2 *bh = reinterpret_cast<const __half2 *>(&bb);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __half22float2(bh[t]);
      v[2 * t] += f.x;
      v[2 * t + 1] += f.y;
    }
  }
  if (res8) {
    // unfused sequence: the conv output is rounded to fp16, then `out += identity` rounds again
    const uint4 rr = __ldg(reinterpret_cast<const uint4 *>(res8));
    const __half2 *rh = reinterpret_cast<const __half2 *>(&rr);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __half22float2(rh[t]);
      v[2 * t] = __half2float(__float2half_rn(v[2 * t])) + f.x;
      v[2 * t + 1] = __half2float(__float2half_rn(v[2 * t + 1])) + f.y;
    }
  }
  if (relu) {
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = fmaxf(v[t], 0.f);
  }
  uint4 o;
  __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
  for (int t = 0; t < 4; ++t) oh[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
  if (out8) *reinterpret_cast<uint4 *>(out8) = o;
  if (plane8) *reinterpret_cast<uint4 *>(plane8) = o;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(remote)
               : "memory");
  return v;
}

}  // namespace bc
