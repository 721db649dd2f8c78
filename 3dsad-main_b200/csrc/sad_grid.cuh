// Scene-grid workspace layout and the (monotone) cell function shared by grid.cu and fps_cull.cu.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace sad {

constexpr int kGridDim = 32;                                   // cells per axis
constexpr int kGridCells = kGridDim * kGridDim * kGridDim;     // 32768
constexpr size_t kGridHeaderBytes = 64;
constexpr size_t kGridCellBytes = ((size_t)(kGridCells + 1) * 4 + 15) / 16 * 16;

// per scene: header | cell_start | N float4 sorted points | N float min-dist scratch (FPS only, padded to 16 B)
inline __host__ __device__ size_t grid_scratch_offset(int N) { return kGridHeaderBytes + kGridCellBytes + (size_t)N * 16; }
inline __host__ __device__ size_t grid_stride(int N) { return grid_scratch_offset(N) + ((size_t)N * 4 + 15) / 16 * 16; }

#ifdef __CUDACC__
// clamp(int(floor((v - mn) * inv_h)), 0, G-1): monotone non-decreasing in v (every step is).
__device__ __forceinline__ int grid_axis(float v, float mn, float inv_h) {
  const float t = floorf(__fmul_rn(__fsub_rn(v, mn), inv_h));
  return (int)fminf(fmaxf(t, 0.f), (float)(kGridDim - 1));     // NaN -> 0
}
__device__ __forceinline__ int grid_cell(float x, float y, float z, float mx, float my, float mz, float inv_h) {
  return grid_axis(x, mx, inv_h) + kGridDim * (grid_axis(y, my, inv_h) + kGridDim * grid_axis(z, mz, inv_h));
}
#endif

}  // namespace sad
