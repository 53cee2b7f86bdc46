// bc_tma.cuh -- TMA-staged movement path (see bc_tma.cu).
#pragma once
#include "bc_common.cuh"

namespace bc {

// True when the shape/alignment allows the cp.async.bulk.tensor path.
bool tma_move_eligible(const void *tiles, const void *plane, int E, int C, int W, int BS, int tile_edge, int es,
                       int layout);

// plane <-> packed tiles through shared memory with TMA on both sides.
// to_plane == false: gather (pad > 0 adds the halo, zero filled outside the frame);
// to_plane == true : scatter (pad must be 0).
int launch_tma_move(void *tiles, void *plane, const int32_t *mapping, int E, int N, int C, int H, int W, int BS,
                    int pad, int tile_edge, int es, int layout, bool to_plane, cudaStream_t s);

}  // namespace bc
