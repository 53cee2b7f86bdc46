// bc_index.cu -- device-side replacement of BlockFeatures._process_grid / get_grid_mappings
// (reference: core/tensorwrapper.py:108-128, :150-178, executed there on the host with a
// D2H copy of the grid, CPU TorchScript and three H2D copies).  One CTA, order preserving.
#include "bc_common.cuh"

namespace bc {

constexpr int kScanThreads = 1024;

// Each thread owns a contiguous run of cells, so that a block-wide exclusive scan of the
// per-thread counts yields row-major ranks (the reference numbers executed cells, and
// separately skipped cells, in row-major order).
__global__ void __launch_bounds__(kScanThreads)
compact_mask_kernel(const uint8_t *__restrict__ grid, int G, int32_t *__restrict__ grid_idx,
                    int32_t *__restrict__ mapping_exec, int32_t *__restrict__ counts,
                    const int32_t *__restrict__ prev_grid_idx, int32_t *__restrict__ transfer_idx) {
  __shared__ int warp_sums[kScanThreads / 32];
  __shared__ int total_exec;
  // No early griddepcontrol.launch_dependents here: the conv kernels read `mapping_exec` BEFORE their own
  // griddepcontrol.wait (the lookup then overlaps their set-up), which is only safe if a kernel launched right
  // behind this one cannot start before this one has finished (implicit trigger at grid completion).
  pdl_wait();
  const int tid = threadIdx.x;
  const int per = (G + kScanThreads - 1) / kScanThreads;
  const int lo = min(tid * per, G), hi = min(lo + per, G);

  int mine = 0;
  for (int g = lo; g < hi; ++g) mine += grid[g] != 0;

  // inclusive warp scan, then scan of warp totals
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += v;
  }
  if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = warp_sums[tid];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, w, o);
      if (tid >= o) w += v;
    }
    warp_sums[tid] = w;  // inclusive
    if (tid == 31) total_exec = w;
  }
  __syncthreads();
  int e = incl - mine + ((tid >> 5) ? warp_sums[(tid >> 5) - 1] : 0);  // executed cells before `lo`
  int k = lo - e;                                                       // skipped cells before `lo`
  for (int g = lo; g < hi; ++g) {
    if (grid[g]) {
      grid_idx[g] = e;
      mapping_exec[e] = g;
      ++e;
    } else {
      grid_idx[g] = -G + k;
      if (transfer_idx != nullptr && prev_grid_idx != nullptr) transfer_idx[k] = prev_grid_idx[g];
      ++k;
    }
  }
  if (tid == 0) {
    counts[0] = total_exec;
    counts[1] = G - total_exec;
  }
}

int launch_compact_mask(const uint8_t *grid, int G, int32_t *grid_idx, int32_t *mapping_exec, int32_t *counts,
                        const int32_t *prev_grid_idx, int32_t *transfer_idx, cudaStream_t s) {
  launch_kernel(compact_mask_kernel, dim3(1), dim3(kScanThreads), 0, s, 1, grid, G, grid_idx, mapping_exec, counts,
                prev_grid_idx, transfer_idx);
  return check_launch("bc_compact_mask");
}

}  // namespace bc
