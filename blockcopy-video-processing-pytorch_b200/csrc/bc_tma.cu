// bc_tma.cu -- TMA-staged block movement: cp.async.bulk.tensor box loads (hardware zero fill
// outside the frame = the halo's frame-border zeros) -> shared memory ring -> TMA box stores.
// One elected thread per CTA drives a 4-stage mbarrier pipeline; the SM's LSU/ALUs do nothing,
// which is the point: gather / scatter are pure HBM work.
//
// Layout handling: both the dense plane and the packed tile batch are described by rank-4
// tensor maps,
//   NHWC  plane (C, W, H, N)  tile (C, TE, TE, E)  box (Cc, TE, R, 1)
//   NCHW  plane (W, H, C, N)  tile (TE, TE, C, E)  box (TE, R, Cc, 1)
// TE = tile edge (BS, or BS+2p for the halo gather), R rows and Cc channels per box.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "bc_move.cuh"
#include "bc_tma.cuh"
#include "bc_ptx.cuh"

namespace bc {

// ------------------------------------------------------------------ kernel
constexpr int kMaxStages = 8;  // ring depth is a launch parameter (p.stages <= kMaxStages)

struct TmaMoveParams {
  const int32_t *mapping;  // cell of packed tile b
  CellDecode cell;
  FastDiv items_per_tile, cchunks;  // items per tile = row_chunks * cchunks
  int n_items;
  int BS, pad, R, Cc;
  int nhwc;     // coordinate order
  int to_plane; // 0: plane -> tiles (gather), 1: tiles -> plane (scatter)
  uint32_t box_bytes, stage_bytes;
  int stages;
};

__global__ void __launch_bounds__(32)
tma_move_kernel(const __grid_constant__ CUtensorMap plane_map, const __grid_constant__ CUtensorMap tile_map,
                const TmaMoveParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  if (threadIdx.x != 0) return;

  // 1024-byte aligned staging ring
  uint8_t *ring = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  prefetch_map(&plane_map);
  prefetch_map(&tile_map);
  const int kStages = p.stages;
  for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
  fence_mbar_init();
  pdl_trigger();
  pdl_wait();

  const CUtensorMap *src = p.to_plane ? &tile_map : &plane_map;
  const CUtensorMap *dst = p.to_plane ? &plane_map : &tile_map;

  const int first = blockIdx.x, step = gridDim.x;
  const int n_mine = first < p.n_items ? (p.n_items - first + step - 1) / step : 0;

  auto coords = [&](int k, int (&pc)[4], int (&tc)[4]) {
    const uint32_t item = (uint32_t)(first + k * step);
    uint32_t b, rem, rc, cc;
    p.items_per_tile.divmod(item, b, rem);
    p.cchunks.divmod(rem, rc, cc);
    uint32_t n, gh, gw;
    p.cell((uint32_t)__ldg(p.mapping + b), n, gh, gw);
    const int r0 = (int)rc * p.R, c0 = (int)cc * p.Cc;
    const int px = (int)(gw * p.BS) - p.pad, py = (int)(gh * p.BS) - p.pad + r0;
    if (p.nhwc) {
      pc[0] = c0; pc[1] = px; pc[2] = py; pc[3] = (int)n;
      tc[0] = c0; tc[1] = 0;  tc[2] = r0; tc[3] = (int)b;
    } else {
      pc[0] = px; pc[1] = py; pc[2] = c0; pc[3] = (int)n;
      tc[0] = 0;  tc[1] = r0; tc[2] = c0; tc[3] = (int)b;
    }
  };
  auto issue_load = [&](int k) {
    int pc[4], tc[4];
    coords(k, pc, tc);
    const int s = k % kStages;
    const int(&c)[4] = p.to_plane ? tc : pc;
    mbar_expect_tx(&full_bar[s], p.box_bytes);
    tma_load_4d(ring + (size_t)s * p.stage_bytes, src, &full_bar[s], c[0], c[1], c[2], c[3]);
  };

  // prologue: kStages-1 loads in flight
  for (int k = 0; k < kStages - 1 && k < n_mine; ++k) issue_load(k);

  for (int j = 0; j < n_mine; ++j) {
    const int s = j % kStages;
    mbar_wait(&full_bar[s], (uint32_t)((j / kStages) & 1));
    int pc[4], tc[4];
    coords(j, pc, tc);
    const int(&c)[4] = p.to_plane ? pc : tc;
    tma_store_4d(dst, ring + (size_t)s * p.stage_bytes, c[0], c[1], c[2], c[3]);
    tma_commit();
    const int k = j + kStages - 1;  // reuses the stage of item j-1
    if (k < n_mine) {
      tma_wait_read<1>();  // store j-1 has finished reading its stage (store j may still be)
      issue_load(k);
    }
  }
  tma_wait_read<0>();
}

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int encode_map_4d(CUtensorMap *map, const void *base, int es, const uint64_t dims[4], const uint32_t box[4]) {
  EncodeTiledFn fn = tensor_map_encoder();
  BC_REQUIRE(fn != nullptr, BC_ERR_NO_DEVICE, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstr[3] = {dims[0] * es, dims[0] * dims[1] * es, dims[0] * dims[1] * dims[2] * es};
  cuuint32_t bdim[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(map, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 4,
                        const_cast<void *>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BC_REQUIRE(r == CUDA_SUCCESS, BC_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return BC_OK;
}

static int largest_divisor_leq(int n, int cap) {
  int best = 1;
  for (int d = 1; d <= n && d <= cap; ++d)
    if (n % d == 0) best = d;
  return best;
}

bool tma_move_eligible(const void *tiles, const void *plane, int E, int C, int W, int BS, int tile_edge, int es,
                       int layout) {
  if (E <= 0) return false;
  if (((uintptr_t)tiles | (uintptr_t)plane) & 15) return false;  // tensor-map base address
  if (tile_edge > 256) return false;                              // box dims are limited to 256 elements
  // Global strides, the inner box extent AND the byte offset of the inner-most start coordinate must be
  // multiples of 16 bytes (a misaligned inner coordinate raises "illegal instruction" on sm_100).
  if (layout == BC_NHWC) return ((int64_t)C * es) % 16 == 0;  // inner dim = channels, coordinates are chunk starts
  // NCHW: inner dim = x.  Box starts at gw*BS - pad, so a halo (tile_edge != BS) is never aligned in practice.
  return tile_edge == BS && ((int64_t)BS * es) % 16 == 0 && ((int64_t)W * es) % 16 == 0;
}

int launch_tma_move(void *tiles, void *plane, const int32_t *mapping, int E, int N, int C, int H, int W, int BS,
                    int pad, int tile_edge, int es, int layout, bool to_plane, cudaStream_t s) {
  const int TE = tile_edge;
  BC_REQUIRE(H % BS == 0 && W % BS == 0, BC_ERR_SHAPE, "plane %dx%d is not divisible by block size %d", H, W, BS);
  // Box selection: <= 32 KB per stage, inner box dim <= 256 elements and a multiple of 16 bytes.
  // tunables (env: BC_TMA_BOX_KB, BC_TMA_STAGES, BC_TMA_CTAS_PER_SM), defaults chosen on B200 (profiles/)
  static const int env_box_kb = getenv("BC_TMA_BOX_KB") ? atoi(getenv("BC_TMA_BOX_KB")) : 16;
  static const int env_stages = getenv("BC_TMA_STAGES") ? atoi(getenv("BC_TMA_STAGES")) : 4;
  static const int env_ctas = getenv("BC_TMA_CTAS_PER_SM") ? atoi(getenv("BC_TMA_CTAS_PER_SM")) : 4;
  const int64_t kBoxBudget = (int64_t)(env_box_kb < 1 ? 1 : env_box_kb > 48 ? 48 : env_box_kb) * 1024;
  const int kStages = env_stages < 2 ? 2 : env_stages > kMaxStages ? kMaxStages : env_stages;
  int Cc, R;
  if (layout == BC_NHWC) {
    int cap_c = 256;  // elements
    int64_t per_c = (int64_t)TE * es;  // bytes of one tile row per channel
    if ((int64_t)cap_c * per_c > kBoxBudget) cap_c = (int)(kBoxBudget / per_c);
    Cc = 0;
    for (int d = 1; d <= C && d <= cap_c; ++d)
      if (C % d == 0 && ((int64_t)d * es) % 16 == 0) Cc = d;
    BC_REQUIRE(Cc > 0, BC_ERR_UNSUPPORTED, "no TMA box fits C=%d, tile edge %d", C, TE);
  } else {
    // NCHW: a (c) plane of a tile is contiguous TE*TE*es bytes; prefer whole planes, few channels
    int64_t plane_bytes = (int64_t)TE * TE * es;
    int cap_c = plane_bytes <= kBoxBudget ? (int)(kBoxBudget / plane_bytes) : 1;
    if (cap_c > 256) cap_c = 256;
    Cc = largest_divisor_leq(C, cap_c);
  }
  const int64_t row_bytes = (int64_t)TE * Cc * es;  // one tile row of the box (NHWC) / one row x Cc planes (NCHW)
  int rmax = (int)(kBoxBudget / row_bytes);
  if (rmax < 1) rmax = 1;
  if (rmax > TE) rmax = TE;
  if (rmax > 256) rmax = 256;
  // a scatter must not write rows of the cell below: R has to divide the tile edge
  R = to_plane ? largest_divisor_leq(TE, rmax) : rmax;
  BC_REQUIRE((int64_t)R * row_bytes <= 56 * 1024, BC_ERR_UNSUPPORTED, "TMA box of %lld bytes is too large",
             (long long)((int64_t)R * row_bytes));

  CUtensorMap plane_map, tile_map;
  int rc;
  if (layout == BC_NHWC) {
    const uint64_t pd[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t td[4] = {(uint64_t)C, (uint64_t)TE, (uint64_t)TE, (uint64_t)E};
    const uint32_t bx[4] = {(uint32_t)Cc, (uint32_t)TE, (uint32_t)R, 1u};
    if ((rc = encode_map_4d(&plane_map, plane, es, pd, bx)) != BC_OK) return rc;
    if ((rc = encode_map_4d(&tile_map, tiles, es, td, bx)) != BC_OK) return rc;
  } else {
    const uint64_t pd[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)N};
    const uint64_t td[4] = {(uint64_t)TE, (uint64_t)TE, (uint64_t)C, (uint64_t)E};
    const uint32_t bx[4] = {(uint32_t)TE, (uint32_t)R, (uint32_t)Cc, 1u};
    if ((rc = encode_map_4d(&plane_map, plane, es, pd, bx)) != BC_OK) return rc;
    if ((rc = encode_map_4d(&tile_map, tiles, es, td, bx)) != BC_OK) return rc;
  }

  TmaMoveParams p;
  p.mapping = mapping;
  p.cell = CellDecode(H / BS, W / BS);
  const int row_chunks = (TE + R - 1) / R, cchunks = C / Cc;
  p.items_per_tile = FastDiv((uint32_t)(row_chunks * cchunks));
  p.cchunks = FastDiv((uint32_t)cchunks);
  const int64_t n_items = (int64_t)E * row_chunks * cchunks;
  BC_REQUIRE(n_items < (1ll << 30), BC_ERR_RANGE, "too many TMA work items");
  p.n_items = (int)n_items;
  p.BS = BS; p.pad = pad; p.R = R; p.Cc = Cc;
  p.nhwc = layout == BC_NHWC;
  p.to_plane = to_plane ? 1 : 0;
  p.box_bytes = (uint32_t)((int64_t)R * row_bytes);
  p.stage_bytes = (p.box_bytes + 1023u) & ~1023u;
  p.stages = kStages;
  const size_t smem = (size_t)kStages * p.stage_bytes + 1024;

  // opt in to > 48 KB of dynamic shared memory (static + dynamic must stay below 227 KB)
  static cudaError_t attr_status = cudaFuncSetAttribute(
      tma_move_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  BC_REQUIRE(attr_status == cudaSuccess, (int)attr_status, "cudaFuncSetAttribute(tma_move_kernel): %s",
             cudaGetErrorString(attr_status));
  BC_REQUIRE(smem <= 200 * 1024, BC_ERR_UNSUPPORTED, "TMA staging ring of %zu bytes is too large", smem);
  // persistent-style launch: CTAs stride over the work items; as many CTAs per SM as shared memory allows
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > env_ctas) per_sm = env_ctas > 0 ? env_ctas : 1;
  int64_t grid = (int64_t)kNumSMs * per_sm;
  if (grid > n_items) grid = n_items;
  launch_kernel(tma_move_kernel, dim3((unsigned)grid), dim3(32), smem, s, 1, plane_map, tile_map, p);
  return check_launch(to_plane ? "bc_scatter[tma]" : "bc_gather[tma]");
}

}  // namespace bc
