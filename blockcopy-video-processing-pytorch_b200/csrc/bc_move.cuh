// bc_move.cuh -- geometry descriptor shared by the SIMT and TMA movement kernels.
#pragma once
#include "bc_common.cuh"

namespace bc {

// Geometry of one movement problem; passed by value to the kernels.  "tile" is the packed
// side ((E,C,TE,TE) or (E,TE,TE,C), TE = tile edge = BS or BS+2p), "plane" the dense side.
struct MoveGeo {
  CellDecode cell;
  FastDiv row_chunks;     // V-byte chunks per tile row
  FastDiv rows_per_tile;  // TE
  FastDiv chan;           // C (NCHW decode)
  FastDiv pix_chunks;     // chunks per pixel (1 pixel = C*es bytes in NHWC, es bytes in NCHW)
  int layout, BS, pad, C, H, W, GH, GW, vec;
  uint32_t total;         // number of chunks
  uint32_t pix_bytes;
  int64_t tile_stride_b, tile_stride_c, tile_row_bytes;
  int64_t plane_stride_n, plane_stride_c, plane_row_bytes;
};

// tile_edge: BS for split/combine/transfer, BS+2p for the halo gathers.
// per_pixel_chunks: a chunk may not straddle a pixel (halo / ring classification is per pixel).
int make_geo(MoveGeo &g, int ntiles, int N, int C, int H, int W, int BS, int tile_edge, int pad, int es,
             int layout, bool per_pixel_chunks, const void *const *ptrs, int nptrs);

int launch_gather_simt(void *tiles, const void *plane, const int32_t *mapping, const MoveGeo &g, bool halo,
                       cudaStream_t s);
bool gather_halo_nchw_eligible(const void *out, const void *plane, int BS, int pad, int W, int es);
int launch_gather_halo_nchw(void *out, const void *plane, const int32_t *mapping, const MoveGeo &g, int E, int es,
                            cudaStream_t s);
int launch_scatter_simt(const void *tiles, void *plane, const int32_t *mapping, const MoveGeo &g, cudaStream_t s);
int launch_copy_blocks_simt(void *out, const void *prev, const void *tiles, const int32_t *grid_idx,
                            const MoveGeo &g, cudaStream_t s);
int launch_transfer_simt(void *out, const void *prev_exec, const void *prev_transfer, const int32_t *transfer_idx,
                         const MoveGeo &g, int G, cudaStream_t s);
int launch_halo_tiles_simt(void *out, const void *exec, const void *transfer, const int32_t *grid_idx,
                           const int32_t *mapping, const MoveGeo &g, const MoveGeo &src, int G, cudaStream_t s);

}  // namespace bc
