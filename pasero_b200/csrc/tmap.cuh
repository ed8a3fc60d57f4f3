// pasero_b200 — host-side TMA tensor-map construction shared by the TMA-fed kernels (defined in gemm.cu).
#pragma once
#include <cuda.h>

namespace pb200 {

// Tensor map over a bf16 (or fp32) matrix stored as `rows` x `cols` (cols contiguous), `ld` elements between rows,
// `batch` slices `batch_stride` elements apart; box = box_cols x box_rows x 1, SWIZZLE_128B. Parts of a box outside
// [cols, rows, batch] are zero-filled on load and clipped on store. Returns 0 or a pb200 error code.
int make_tmap(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, long long batch,
              long long batch_stride, int box_cols, int box_rows, bool fp32 = false);

}  // namespace pb200
