// bc_move.cu -- vectorised SIMT block-movement kernels (gather / scatter / copy / ring
// transfer / gather-with-halo) for both layouts.  The TMA-staged NHWC variants live in
// bc_tma.cu; the entry points in bc_api.cu choose between them.
//
// Every kernel walks a flat index of V-byte "chunks" of the packed-tile side, which is
// always dense, so tile-side accesses are perfectly coalesced 16-byte vectors; the
// plane-side address of the same chunk is a contiguous run of the same tile row (whole
// BS*C*es bytes in NHWC, BS*es bytes in NCHW).  Each thread keeps kUnroll independent
// loads in flight before its first store.  Index decoding uses mul.hi divisions by
// host-prepared constants (FastDiv); nothing is a compile-time shape, so there is no
// per-shape JIT as in the reference (utils/cuda.py:25-31).
#include "bc_move.cuh"

namespace bc {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// ---------------------------------------------------------------------------------------------
// Shared decode: chunk id -> tile-side byte offset and (b, c, y, x)
// ---------------------------------------------------------------------------------------------
struct ChunkPos {
  uint32_t b, c, y, x;  // packed-tile index, channel (NCHW only), row and pixel inside the (padded) tile
  uint32_t colb;        // byte offset inside the tile row
};

template <int V>
__device__ __forceinline__ ChunkPos decode_chunk(const MoveGeo &g, uint32_t i) {
  ChunkPos q;
  uint32_t row, col, t;
  g.row_chunks.divmod(i, row, col);
  q.colb = col * V;
  g.rows_per_tile.divmod(row, t, q.y);
  if (g.layout == BC_NCHW) {
    g.chan.divmod(t, q.b, q.c);
  } else {
    q.b = t;
    q.c = 0;
  }
  q.x = g.pix_chunks.div(col);
  return q;
}

__device__ __forceinline__ int64_t tile_offset(const MoveGeo &g, uint32_t b, uint32_t c, uint32_t y,
                                               uint32_t colb) {
  return (int64_t)b * g.tile_stride_b + (int64_t)c * g.tile_stride_c + (int64_t)y * g.tile_row_bytes + colb;
}

// byte offset in the plane of pixel (yy, xx) of image n, channel c (c = 0 for NHWC)
__device__ __forceinline__ int64_t plane_offset(const MoveGeo &g, uint32_t n, uint32_t c, int yy, int xx) {
  return (int64_t)n * g.plane_stride_n + (int64_t)c * g.plane_stride_c + (int64_t)yy * g.plane_row_bytes +
         (int64_t)xx * g.pix_bytes;
}

// ---------------------------------------------------------------------------------------------
// gather (plane -> tiles), optionally with halo + zero fill outside the frame
// ---------------------------------------------------------------------------------------------
template <int V, bool HALO>
__global__ void __launch_bounds__(kThreads)
gather_kernel(char *__restrict__ tiles, const char *__restrict__ plane,
              const int32_t *__restrict__ mapping, const MoveGeo g) {
  using T = typename Vec<V>::T;
  pdl_trigger();
  pdl_wait();
  const uint32_t stride = gridDim.x * kThreads;
  for (uint32_t base = blockIdx.x * kThreads + threadIdx.x; base < g.total; base += stride * kUnroll) {
    T v[kUnroll];
    int64_t dst[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = base + u * stride;
      dst[u] = -1;
      if (i < g.total) {
        const ChunkPos q = decode_chunk<V>(g, i);
        uint32_t n, gh, gw;
        g.cell((uint32_t)__ldg(mapping + q.b), n, gh, gw);
        const int yy = (int)(gh * g.BS + q.y) - g.pad;
        const int xx = (int)(gw * g.BS + q.x) - g.pad;
        dst[u] = tile_offset(g, q.b, q.c, q.y, q.colb);
        const uint32_t inpix = q.colb - q.x * g.pix_bytes;  // byte offset inside the pixel
        bool inside = true;
        if (HALO) inside = (yy >= 0) & (yy < g.H) & (xx >= 0) & (xx < g.W);
        v[u] = inside ? ld_stream<V>(plane + plane_offset(g, n, q.c, yy, xx) + inpix) : zero_vec<V>();
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (dst[u] >= 0) st_vec<V>(tiles + dst[u], v[u]);
  }
}

// ---------------------------------------------------------------------------------------------
// scatter (tiles -> plane), in place; no halo
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(kThreads)
scatter_kernel(const char *__restrict__ tiles, char *__restrict__ plane,
               const int32_t *__restrict__ mapping, const MoveGeo g) {
  using T = typename Vec<V>::T;
  pdl_trigger();
  pdl_wait();
  const uint32_t stride = gridDim.x * kThreads;
  for (uint32_t base = blockIdx.x * kThreads + threadIdx.x; base < g.total; base += stride * kUnroll) {
    T v[kUnroll];
    int64_t dst[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = base + u * stride;
      dst[u] = -1;
      if (i < g.total) {
        const ChunkPos q = decode_chunk<V>(g, i);
        uint32_t n, gh, gw;
        g.cell((uint32_t)__ldg(mapping + q.b), n, gh, gw);
        const uint32_t inpix = q.colb - q.x * g.pix_bytes;
        dst[u] = plane_offset(g, n, q.c, (int)(gh * g.BS + q.y), (int)(gw * g.BS + q.x)) + inpix;
        v[u] = ld_stream<V>(tiles + tile_offset(g, q.b, q.c, q.y, q.colb));
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (dst[u] >= 0) st_vec<V>(plane + dst[u], v[u]);
  }
}

// ---------------------------------------------------------------------------------------------
// copy_blocks: out[cell] = executed ? tiles[grid_idx[cell]] : prev[cell]   (whole plane, one pass)
// The flat index runs over ALL cells (b == cell id).
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(kThreads)
copy_blocks_kernel(char *__restrict__ out, const char *__restrict__ prev, const char *__restrict__ tiles,
                   const int32_t *__restrict__ grid_idx, const MoveGeo g) {
  using T = typename Vec<V>::T;
  pdl_trigger();
  pdl_wait();
  const uint32_t stride = gridDim.x * kThreads;
  for (uint32_t base = blockIdx.x * kThreads + threadIdx.x; base < g.total; base += stride * kUnroll) {
    T v[kUnroll];
    int64_t dst[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = base + u * stride;
      dst[u] = -1;
      if (i < g.total) {
        const ChunkPos q = decode_chunk<V>(g, i);
        uint32_t n, gh, gw;
        g.cell(q.b, n, gh, gw);
        const uint32_t inpix = q.colb - q.x * g.pix_bytes;
        dst[u] = plane_offset(g, n, q.c, (int)(gh * g.BS + q.y), (int)(gw * g.BS + q.x)) + inpix;
        const int32_t t = __ldg(grid_idx + q.b);
        v[u] = t >= 0 ? ld_stream<V>(tiles + tile_offset(g, (uint32_t)t, q.c, q.y, q.colb))
                      : ld_stream<V>(prev + dst[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (dst[u] >= 0) st_vec<V>(out + dst[u], v[u]);
  }
}

// ---------------------------------------------------------------------------------------------
// ring transfer between packed-tile tensors (reference protocol, block_funcs.py:201-237)
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(kThreads)
transfer_kernel(char *__restrict__ out, const char *__restrict__ prev_exec,
                const char *__restrict__ prev_transfer, const int32_t *__restrict__ transfer_idx,
                const MoveGeo g, const int G) {
  pdl_trigger();
  pdl_wait();
  const uint32_t stride = gridDim.x * kThreads;
  const int lo = g.pad, hi = g.BS - g.pad - 1;
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < g.total; i += stride) {
    const ChunkPos q = decode_chunk<V>(g, i);
    const int h = (int)q.y, w = (int)q.x;
    if (g.pad >= 0 && w >= lo && w <= hi && h >= lo && h <= hi) continue;  // interior: never needed
    int32_t bp = __ldg(transfer_idx + q.b);
    const char *src = prev_exec;
    if (bp < 0) {
      bp += G;
      src = prev_transfer;
    }
    st_vec<V>(out + tile_offset(g, q.b, q.c, q.y, q.colb),
              ld_stream<V>(src + tile_offset(g, (uint32_t)bp, q.c, q.y, q.colb)));
  }
}

// ---------------------------------------------------------------------------------------------
// gather-with-halo between packed-tile tensors (reference protocol, blockpad.py:77-156).
// The flat index walks the PADDED output; `g` describes the padded tile, `src_*` the source tile.
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(kThreads)
halo_tiles_kernel(char *__restrict__ out, const char *__restrict__ exec, const char *__restrict__ transfer,
                  const int32_t *__restrict__ grid_idx, const int32_t *__restrict__ mapping,
                  const MoveGeo g, const int64_t src_stride_b, const int64_t src_stride_c,
                  const int64_t src_row_bytes, const int G) {
  using T = typename Vec<V>::T;
  pdl_trigger();
  pdl_wait();
  const uint32_t stride = gridDim.x * kThreads;
  const int BS = g.BS, P = g.pad, BP = BS + 2 * P;
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < g.total; i += stride) {
    const ChunkPos q = decode_chunk<V>(g, i);
    const int hp = (int)q.y, wp = (int)q.x;
    const int left = wp < P, right = wp >= BP - P, top = hp < P, bottom = hp >= BP - P;
    int h = hp - P, w = wp - P;
    int32_t b = (int32_t)q.b;
    const char *src = exec;
    bool zero = false;
    if (left | right | top | bottom) {
      const uint32_t cell = (uint32_t)__ldg(mapping + q.b);
      uint32_t n, gh, gw;
      g.cell(cell, n, gh, gw);
      zero = (left & (gw == 0)) | (right & (gw == (uint32_t)g.GW - 1)) | (top & (gh == 0)) |
             (bottom & (gh == (uint32_t)g.GH - 1));
      if (!zero) {
        const int32_t nb = (int32_t)cell + (right - left) + g.GW * (bottom - top);
        b = __ldg(grid_idx + nb);
        if (b < 0) {
          b += G;
          src = transfer;
        }
        if (left) w += BS; else if (right) w -= BS;
        if (top) h += BS; else if (bottom) h -= BS;
      }
    }
    const uint32_t inpix = q.colb - q.x * g.pix_bytes;
    T v = zero_vec<V>();
    if (!zero)
      v = ld_stream<V>(src + (int64_t)b * src_stride_b + (int64_t)q.c * src_stride_c +
                       (int64_t)h * src_row_bytes + (int64_t)w * g.pix_bytes + inpix);
    st_vec<V>(out + tile_offset(g, q.b, q.c, q.y, q.colb), v);
  }
}

// ---------------------------------------------------------------------------------------------
// gather-with-halo from an NCHW plane (the reference's layout; BASELINE config 2 as stated).
// A padded NCHW tile row is (BS + 2p) elements = 68 bytes for BS 32, p 1, fp16: nothing in it is 16-byte
// aligned, and the generic kernel above degenerates to one thread per 2-byte element with a full
// div/mod chain each (24.9 us for config 2, 0.9 TB/s).  Here one WARP owns one (tile, channel): the
// (BS+2p)^2 elements of that pair are CONTIGUOUS in the output.  Rows are read with 4-byte loads over the
// aligned interior plus element loads for the 2p halo columns (zeros outside the frame), staged in shared
// memory in output order, and written back as one contiguous run of 4-byte words.
// ---------------------------------------------------------------------------------------------
template <typename E>  // element type: uint16_t (fp16) or uint32_t (fp32)
__global__ void __launch_bounds__(kThreads)
gather_halo_nchw_kernel(E *__restrict__ out, const E *__restrict__ plane, const int32_t *__restrict__ mapping,
                        const MoveGeo g, const uint32_t pairs, const int words_per_pair, const FastDiv wpr) {
  extern __shared__ uint32_t halo_smem[];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int BS = g.BS, P = g.pad, TE = BS + 2 * P;
  E *st = reinterpret_cast<E *>(halo_smem + (size_t)warp * words_per_pair);
  constexpr int kPerWord = 4 / (int)sizeof(E);  // elements per 4-byte word
  const uint32_t wstride = gridDim.x * (kThreads / 32);
  for (uint32_t pr = blockIdx.x * (kThreads / 32) + warp; pr < pairs; pr += wstride) {
    uint32_t b, c, n, gh, gw;
    g.chan.divmod(pr, b, c);
    g.cell((uint32_t)__ldg(mapping + b), n, gh, gw);
    const int y0 = (int)gh * BS - P, x0 = (int)gw * BS;
    const E *src_c = plane + ((size_t)n * g.C + c) * g.H * g.W;
    // interior: TE rows x (BS elements starting at x0 = wpr 4-byte words, aligned on the plane side); kU words
    // per lane are in flight before the first one is staged (a row-by-row loop would serialise TE round trips)
    constexpr int kU = 17;
    const int total_w = TE * wpr.d;
    for (int base = 0; base < total_w; base += 32 * kU) {
      uint32_t v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = base + u * 32 + lane;
        uint32_t r, j;
        wpr.divmod((uint32_t)min(i, total_w - 1), r, j);
        const int y = y0 + (int)r;
        v[u] = (i < total_w && y >= 0 && y < g.H)
                   ? __ldg(reinterpret_cast<const uint32_t *>(src_c + (size_t)y * g.W + x0) + j) : 0u;
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = base + u * 32 + lane;
        if (i >= total_w) continue;
        uint32_t r, j;
        wpr.divmod((uint32_t)i, r, j);
        E *row = st + r * TE;
        if (kPerWord == 2) {
          row[P + 2 * j] = (E)(v[u] & 0xffffu);
          row[P + 2 * j + 1] = (E)(v[u] >> 16);
        } else {
          row[P + j] = (E)v[u];
        }
      }
    }
    // halo columns: p on either side of every row, zeros outside the frame
    constexpr int kH = 4;
    const int total_h = TE * 2 * P;
    for (int base = 0; base < total_h; base += 32 * kH) {
      E v[kH];
#pragma unroll
      for (int u = 0; u < kH; ++u) {
        const int i = base + u * 32 + lane;
        const int ii = min(i, total_h - 1);
        const int r = ii / (2 * P), k = ii - r * 2 * P;
        const int e = k < P ? k : BS + k;
        const int y = y0 + r, x = x0 - P + e;
        v[u] = (i < total_h && y >= 0 && y < g.H && x >= 0 && x < g.W) ? __ldg(src_c + (size_t)y * g.W + x) : (E)0;
      }
#pragma unroll
      for (int u = 0; u < kH; ++u) {
        const int i = base + u * 32 + lane;
        if (i >= total_h) continue;
        const int r = i / (2 * P), k = i - r * 2 * P;
        st[r * TE + (k < P ? k : BS + k)] = v[u];
      }
    }
    __syncwarp();
    uint32_t *dst = reinterpret_cast<uint32_t *>(out + (size_t)pr * TE * TE);
    const uint32_t *sst = reinterpret_cast<const uint32_t *>(st);
    for (int w = lane; w < words_per_pair; w += 32) dst[w] = sst[w];
    __syncwarp();
  }
}

// eligible: 4-byte words work on both sides and the staging tile is small
bool gather_halo_nchw_eligible(const void *out, const void *plane, int BS, int pad, int W, int es) {
  const int TE = BS + 2 * pad;
  const long long pair_bytes = (long long)TE * TE * es;
  return (BS * es) % 4 == 0 && (W * es) % 4 == 0 && pair_bytes % 4 == 0 && pair_bytes <= 12 * 1024 && 2 * pad <= 32 &&
         (((uintptr_t)out | (uintptr_t)plane) & 3) == 0;
}

int launch_gather_halo_nchw(void *out, const void *plane, const int32_t *mapping, const MoveGeo &g, int E, int es,
                            cudaStream_t s) {
  const int TE = g.BS + 2 * g.pad;
  const int words = TE * TE * es / 4;
  const long long pairs = (long long)E * g.C;
  BC_REQUIRE(pairs < (1ll << 31), BC_ERR_RANGE, "bc_gather_halo: problem too large");
  const size_t smem = (size_t)(kThreads / 32) * words * 4;
  long long grid = (pairs + kThreads / 32 - 1) / (kThreads / 32);
  const long long cap = (long long)kNumSMs * 8;
  if (grid > cap) grid = cap;
  if (es == 2) {
    static cudaError_t attr = cudaFuncSetAttribute(gather_halo_nchw_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    BC_REQUIRE(attr == cudaSuccess, (int)attr, "cudaFuncSetAttribute(gather_halo_nchw_kernel): %s", cudaGetErrorString(attr));
    launch_kernel(gather_halo_nchw_kernel<uint16_t>, dim3((unsigned)grid), dim3(kThreads), smem, s, 1, (uint16_t *)out,
                  (const uint16_t *)plane, mapping, g, (uint32_t)pairs, words, FastDiv((uint32_t)(g.BS * 2 / 4)));
  } else {
    static cudaError_t attr = cudaFuncSetAttribute(gather_halo_nchw_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    BC_REQUIRE(attr == cudaSuccess, (int)attr, "cudaFuncSetAttribute(gather_halo_nchw_kernel): %s", cudaGetErrorString(attr));
    launch_kernel(gather_halo_nchw_kernel<uint32_t>, dim3((unsigned)grid), dim3(kThreads), smem, s, 1, (uint32_t *)out,
                  (const uint32_t *)plane, mapping, g, (uint32_t)pairs, words, FastDiv((uint32_t)g.BS));
  }
  return check_launch("bc_gather_halo");
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int make_geo(MoveGeo &g, int ntiles, int N, int C, int H, int W, int BS, int tile_edge, int pad, int es,
             int layout, bool per_pixel_chunks, const void *const *ptrs, int nptrs) {
  BC_REQUIRE(ntiles >= 0 && N > 0 && C > 0 && H > 0 && W > 0 && BS > 0 && tile_edge > 0, BC_ERR_SHAPE,
             "non-positive size (tiles=%d N=%d C=%d H=%d W=%d BS=%d)", ntiles, N, C, H, W, BS);
  BC_REQUIRE(H % BS == 0 && W % BS == 0, BC_ERR_SHAPE, "plane %dx%d is not divisible by block size %d", H, W, BS);
  BC_REQUIRE(es == 2 || es == 4, BC_ERR_DTYPE, "dtype must be BC_F16 or BC_F32");
  BC_REQUIRE(layout == BC_NCHW || layout == BC_NHWC, BC_ERR_DTYPE, "layout must be BC_NCHW or BC_NHWC");
  const int TE = tile_edge;
  g.layout = layout;
  g.BS = BS; g.pad = pad; g.C = C; g.H = H; g.W = W; g.GH = H / BS; g.GW = W / BS;
  g.cell = CellDecode(g.GH, g.GW);
  const int64_t pix = layout == BC_NHWC ? (int64_t)C * es : es;
  g.pix_bytes = (uint32_t)pix;
  g.tile_row_bytes = (int64_t)TE * pix;
  if (layout == BC_NHWC) {
    g.tile_stride_c = 0;
    g.tile_stride_b = (int64_t)TE * g.tile_row_bytes;
    g.plane_stride_c = 0;
    g.plane_row_bytes = (int64_t)W * pix;
    g.plane_stride_n = (int64_t)H * g.plane_row_bytes;
  } else {
    g.tile_stride_c = (int64_t)TE * g.tile_row_bytes;
    g.tile_stride_b = (int64_t)C * g.tile_stride_c;
    g.plane_row_bytes = (int64_t)W * es;
    g.plane_stride_c = (int64_t)H * g.plane_row_bytes;
    g.plane_stride_n = (int64_t)C * g.plane_stride_c;
  }
  // Vector width: a chunk must not straddle a pixel when pixels are classified one by one
  // (halo / ring kernels); otherwise it only has to divide the contiguous run one tile row
  // has on the plane side (BS*pix bytes).  Row starts and base pointers must be V-aligned.
  int V = gcd_pow2_bytes(per_pixel_chunks ? (uint64_t)pix : (uint64_t)BS * pix);
  V = min(V, gcd_pow2_bytes((uint64_t)g.plane_row_bytes));
  V = min(V, gcd_pow2_bytes((uint64_t)g.tile_row_bytes));
  for (int k = 0; k < nptrs; ++k) {
    BC_REQUIRE(((uintptr_t)ptrs[k] % es) == 0, BC_ERR_ALIGN, "pointer %d is not aligned to the element size", k);
    while (V > es && ((uintptr_t)ptrs[k] % V)) V >>= 1;
  }
  if (V < es) V = es;
  g.vec = V;
  const int64_t rows = (int64_t)ntiles * TE * (layout == BC_NCHW ? C : 1);
  const int64_t chunks_per_row = g.tile_row_bytes / V;
  const int64_t total = rows * chunks_per_row;
  BC_REQUIRE(total < (1ll << 31), BC_ERR_RANGE, "problem too large: %lld chunks of %d bytes", (long long)total, V);
  g.total = (uint32_t)total;
  g.row_chunks = FastDiv((uint32_t)chunks_per_row);
  g.rows_per_tile = FastDiv((uint32_t)TE);
  g.chan = FastDiv((uint32_t)C);
  // pixel index of a chunk = chunk_in_row / chunks_per_pixel.  When a chunk is wider than a
  // pixel (NCHW without halo) the whole row is addressed as one run: x == 0 for every chunk.
  g.pix_chunks = FastDiv((uint32_t)((pix % V) == 0 ? pix / V : chunks_per_row));
  return BC_OK;
}

static inline int grid_for(uint32_t total, int unroll) {
  const int64_t per_cta = (int64_t)kThreads * unroll;
  int64_t ctas = (total + per_cta - 1) / per_cta;
  const int64_t cap = (int64_t)kNumSMs * 8;  // 8 resident CTAs of 256 threads per SM
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  return (int)ctas;
}

#define BC_DISPATCH_VEC(V_, ...)                \
  switch (V_) {                                 \
    case 16: { constexpr int VV = 16; __VA_ARGS__; } break; \
    case 8:  { constexpr int VV = 8;  __VA_ARGS__; } break; \
    case 4:  { constexpr int VV = 4;  __VA_ARGS__; } break; \
    default: { constexpr int VV = 2;  __VA_ARGS__; } break; \
  }

int launch_gather_simt(void *tiles, const void *plane, const int32_t *mapping, const MoveGeo &g, bool halo,
                       cudaStream_t s) {
  if (g.total == 0) return BC_OK;
  const int grid = grid_for(g.total, kUnroll);
  BC_DISPATCH_VEC(g.vec, {
    if (halo)
      launch_kernel(gather_kernel<VV, true>, dim3(grid), dim3(kThreads), 0, s, 1, (char *)tiles, (const char *)plane, mapping, g);
    else
      launch_kernel(gather_kernel<VV, false>, dim3(grid), dim3(kThreads), 0, s, 1, (char *)tiles, (const char *)plane, mapping, g);
  });
  return check_launch(halo ? "bc_gather_halo" : "bc_gather");
}

int launch_scatter_simt(const void *tiles, void *plane, const int32_t *mapping, const MoveGeo &g, cudaStream_t s) {
  if (g.total == 0) return BC_OK;
  const int grid = grid_for(g.total, kUnroll);
  BC_DISPATCH_VEC(g.vec, { launch_kernel(scatter_kernel<VV>, dim3(grid), dim3(kThreads), 0, s, 1, (const char *)tiles, (char *)plane, mapping, g); });
  return check_launch("bc_scatter");
}

int launch_copy_blocks_simt(void *out, const void *prev, const void *tiles, const int32_t *grid_idx,
                            const MoveGeo &g, cudaStream_t s) {
  if (g.total == 0) return BC_OK;
  const int grid = grid_for(g.total, kUnroll);
  BC_DISPATCH_VEC(g.vec, {
    launch_kernel(copy_blocks_kernel<VV>, dim3(grid), dim3(kThreads), 0, s, 1, (char *)out, (const char *)prev,
                  (const char *)tiles, grid_idx, g);
  });
  return check_launch("bc_copy_blocks");
}

int launch_transfer_simt(void *out, const void *prev_exec, const void *prev_transfer, const int32_t *transfer_idx,
                         const MoveGeo &g, int G, cudaStream_t s) {
  if (g.total == 0) return BC_OK;
  const int grid = grid_for(g.total, 1);
  BC_DISPATCH_VEC(g.vec, {
    launch_kernel(transfer_kernel<VV>, dim3(grid), dim3(kThreads), 0, s, 1, (char *)out, (const char *)prev_exec,
                  (const char *)prev_transfer, transfer_idx, g, G);
  });
  return check_launch("bc_transfer");
}

int launch_halo_tiles_simt(void *out, const void *exec, const void *transfer, const int32_t *grid_idx,
                           const int32_t *mapping, const MoveGeo &g, const MoveGeo &src, int G, cudaStream_t s) {
  if (g.total == 0) return BC_OK;
  const int grid = grid_for(g.total, 1);
  BC_DISPATCH_VEC(g.vec, {
    launch_kernel(halo_tiles_kernel<VV>, dim3(grid), dim3(kThreads), 0, s, 1, (char *)out, (const char *)exec,
                  (const char *)transfer, grid_idx, mapping, g, src.tile_stride_b, src.tile_stride_c,
                  src.tile_row_bytes, G);
  });
  return check_launch("bc_gather_halo_tiles");
}

}  // namespace bc
