// Scene grid: a per-scene spatial sort that both the exact culled FPS (fps_cull.cu) and the
// grid-accelerated ball query below read -- SURVEY.md section 8(f) rank 2 ("cluster/DSMEM FPS +
// grid-accelerated ball query ... must remain bit-exact vs oracle").
// (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
// Layout of the caller-owned workspace, per scene (stride = sad_scene_grid_workspace_bytes / B):
//   [ 64 B header : min.x min.y min.z inv_h | ... ]
//   [ (SAD_GRID_CELLS + 1) u32 cell_start, padded to 16 B ]
//   [ N float4 : x, y, z, bits(original index) ]  sorted by cell = ix + G*(iy + G*iz), G = 32
// cell(v) = clamp(int(floor((v - min) * inv_h)), 0, G-1) per axis -- a MONOTONE function of v
// in fp32, which is all the exactness arguments below need.  The order of points inside a
// cell is unspecified (atomics); no consumer depends on it.
#include "sad_common.cuh"
#include "sad_grid.cuh"

namespace {

using namespace sad;

constexpr int GB_T = 1024;

__global__ void __launch_bounds__(GB_T, 1)
grid_build_kernel(int N, const float* __restrict__ xyz, uint8_t* __restrict__ ws, size_t stride) {
  extern __shared__ uint32_t s_hist[];                  // [kGridCells]
  __shared__ float s_red[6][GB_T / 32];
  __shared__ float s_hdr[4];
  __shared__ uint32_t s_wsum[GB_T / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float* pts = xyz + (size_t)b * N * 3;
  uint8_t* base = ws + (size_t)b * stride;
  float* hdr = reinterpret_cast<float*>(base);
  uint32_t* cell_start = reinterpret_cast<uint32_t*>(base + kGridHeaderBytes);
  float4* sorted = reinterpret_cast<float4*>(base + kGridHeaderBytes + kGridCellBytes);

  // ---- 1. bounding box
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int k = tid; k < N; k += GB_T) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = __ldg(pts + 3 * (size_t)k + a);
      lo[a] = fminf(lo[a], v);
      hi[a] = fmaxf(hi[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(FULL, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(FULL, hi[a], o));
    }
    if (lane == 0) {
      s_red[a][warp] = lo[a];
      s_red[3 + a][warp] = hi[a];
    }
  }
  for (int i = tid; i < kGridCells; i += GB_T) s_hist[i] = 0u;
  __syncthreads();
  if (warp == 0) {
    float ext = 0.f, mn[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float l = s_red[a][lane], h = s_red[3 + a][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        l = fminf(l, __shfl_xor_sync(FULL, l, o));
        h = fmaxf(h, __shfl_xor_sync(FULL, h, o));
      }
      mn[a] = l;
      ext = fmaxf(ext, h - l);
    }
    if (lane == 0) {
      // isotropic cells; a degenerate (single-point) scene maps everything to cell 0
      const float inv_h = (ext > 0.f && ext < 3.0e38f) ? (float)kGridDim / ext : 0.f;
      s_hdr[0] = mn[0];
      s_hdr[1] = mn[1];
      s_hdr[2] = mn[2];
      s_hdr[3] = inv_h;
      hdr[0] = mn[0];
      hdr[1] = mn[1];
      hdr[2] = mn[2];
      hdr[3] = inv_h;
    }
  }
  __syncthreads();
  const float mx = s_hdr[0], my = s_hdr[1], mz = s_hdr[2], inv_h = s_hdr[3];

  // ---- 2. histogram
  for (int k = tid; k < N; k += GB_T) {
    const float x = __ldg(pts + 3 * (size_t)k), y = __ldg(pts + 3 * (size_t)k + 1), z = __ldg(pts + 3 * (size_t)k + 2);
    atomicAdd(&s_hist[grid_cell(x, y, z, mx, my, mz, inv_h)], 1u);
  }
  __syncthreads();

  // ---- 3. exclusive scan of the kGridCells bins (32 per thread)
  constexpr int PER = kGridCells / GB_T;
  uint32_t loc[PER], sum = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    loc[i] = s_hist[tid * PER + i];
    sum += loc[i];
  }
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_wsum[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(FULL, winc, o);
      if (lane >= o) winc += t;
    }
    s_wsum[lane] = winc - w;
  }
  __syncthreads();
  uint32_t run = s_wsum[warp] + inc - sum;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    s_hist[tid * PER + i] = run;
    cell_start[tid * PER + i] = run;
    run += loc[i];
  }
  if (tid == GB_T - 1) cell_start[kGridCells] = run;     // == N
  __syncthreads();

  // ---- 4. scatter
  for (int k = tid; k < N; k += GB_T) {
    const float x = __ldg(pts + 3 * (size_t)k), y = __ldg(pts + 3 * (size_t)k + 1), z = __ldg(pts + 3 * (size_t)k + 2);
    const uint32_t pos = atomicAdd(&s_hist[grid_cell(x, y, z, mx, my, mz, inv_h)], 1u);
    sorted[pos] = make_float4(x, y, z, __uint_as_float((uint32_t)k));
  }
}

// ------------------------------------------------------------------ grid-accelerated ball query (a3 / a4)
// One warp per query.  The cells that can hold a point with d2 < r*r are visited row by row
// (a row = the x-run of cells for one (iy,iz), contiguous in the sorted array), every hit's
// ORIGINAL index goes to a shared-memory list (ballot + popc compaction), and the nsample
// smallest indices are then emitted in ascending order by rank counting -- exactly the set and
// order the brute-force index-order scan of the oracle produces.  A query with more hits than
// the list holds falls back to that scan itself (dense neighbourhoods: it exits after a few
// hundred candidates).
constexpr int BQG_WARPS = 8;
constexpr int BQG_CAP = 768;

__global__ void __launch_bounds__(BQG_WARPS * 32)
ball_query_grid_kernel(int N, int npoint, float radius, const float* __restrict__ radius_t, int nsample,
                       const float* __restrict__ xyz, const uint8_t* __restrict__ ws, size_t stride,
                       const float* __restrict__ new_xyz, int32_t* __restrict__ idx) {
  __shared__ uint32_t s_hits[BQG_WARPS][BQG_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int j = blockIdx.x * BQG_WARPS + warp;
  if (j >= npoint) return;
  const uint8_t* base = ws + (size_t)b * stride;
  const float* hdr = reinterpret_cast<const float*>(base);
  const uint32_t* cell_start = reinterpret_cast<const uint32_t*>(base + kGridHeaderBytes);
  const float4* sorted = reinterpret_cast<const float4*>(base + kGridHeaderBytes + kGridCellBytes);
  const float mx = __ldg(hdr), my = __ldg(hdr + 1), mz = __ldg(hdr + 2), inv_h = __ldg(hdr + 3);
  const float* c = new_xyz + ((size_t)b * npoint + j) * 3;
  const float qx = __ldg(c), qy = __ldg(c + 1), qz = __ldg(c + 2);
  const float r = radius_t ? __ldg(radius_t + (size_t)b * npoint + j) : radius;
  const float r2 = __fmul_rn(r, r);
  int32_t* out = idx + ((size_t)b * npoint + j) * nsample;
  uint32_t* hits = s_hits[warp];
  const uint32_t lt_mask = (1u << lane) - 1u;

  // conservative cell range: every p with fl(d2) < fl(r*r) has |p_a - q_a| < r * (1 + 1e-6) in exact
  // arithmetic; the extra |q_a|-relative term covers the rounding of q_a -/+ rr itself
  const float rr = fmaf(fabsf(r), 1.00001f, 1e-30f);
  const float ex = fmaf(2.4e-7f, fabsf(qx) + rr, rr), ey = fmaf(2.4e-7f, fabsf(qy) + rr, rr),
              ez = fmaf(2.4e-7f, fabsf(qz) + rr, rr);
  const int ix0 = grid_axis(qx - ex, mx, inv_h), ix1 = grid_axis(qx + ex, mx, inv_h);
  const int iy0 = grid_axis(qy - ey, my, inv_h), iy1 = grid_axis(qy + ey, my, inv_h);
  const int iz0 = grid_axis(qz - ez, mz, inv_h), iz1 = grid_axis(qz + ez, mz, inv_h);
  const int ny = iy1 - iy0 + 1, nrows = ny * (iz1 - iz0 + 1);

  int H = 0;
  bool overflow = !(r2 >= 0.f) || !(rr < 3.0e38f);      // NaN / inf radius: let the plain scan decide
  for (int row0 = 0; row0 < nrows && !overflow; row0 += 32) {
    uint32_t s = 0, e = 0;
    if (row0 + lane < nrows) {
      const int rr_ = row0 + lane;
      const int iy = iy0 + rr_ % ny, iz = iz0 + rr_ / ny;
      const int cb = kGridDim * (iy + kGridDim * iz);
      s = __ldg(cell_start + cb + ix0);
      e = __ldg(cell_start + cb + ix1 + 1);
    }
    const int nr = min(32, nrows - row0);
    for (int rw = 0; rw < nr && !overflow; ++rw) {
      const uint32_t rs = __shfl_sync(FULL, s, rw), re = __shfl_sync(FULL, e, rw);
      for (uint32_t k0 = rs; k0 < re; k0 += 32) {
        const uint32_t k = k0 + lane;
        bool hit = false;
        uint32_t oi = 0;
        if (k < re) {
          const float4 p = __ldg(sorted + k);
          hit = sqdist(p.x, p.y, p.z, qx, qy, qz) < r2;
          oi = __float_as_uint(p.w);
        }
        const uint32_t m = __ballot_sync(FULL, hit);
        if (m) {
          const int n = __popc(m);
          if (H + n > BQG_CAP) {
            overflow = true;
            break;
          }
          if (hit) hits[H + __popc(m & lt_mask)] = oi;
          H += n;
        }
      }
    }
  }
  __syncwarp();

  if (overflow) {
    // exact fallback: the oracle's own index-order scan over the original array
    const float* pts = xyz + (size_t)b * N * 3;
    int cnt = 0;
    int first = 0;
    for (int k0 = 0; k0 < N && cnt < nsample; k0 += 32) {
      const int k = k0 + lane;
      bool hit = false;
      if (k < N) hit = sqdist(__ldg(pts + 3 * (size_t)k), __ldg(pts + 3 * (size_t)k + 1), __ldg(pts + 3 * (size_t)k + 2), qx, qy, qz) < r2;
      const uint32_t m = __ballot_sync(FULL, hit);
      if (m) {
        if (cnt == 0) first = k0 + __ffs(m) - 1;
        const int pos = cnt + __popc(m & lt_mask);
        if (hit && pos < nsample) out[pos] = k;
        cnt += __popc(m);
      }
    }
    cnt = min(cnt, nsample);
    for (int l = cnt + lane; l < nsample; l += 32) out[l] = first;     // no hit: first == 0 -> zeros
    return;
  }

  // ---- emit the nsample smallest original indices in ascending order (rank counting)
  uint32_t firstv = 0xFFFFFFFFu;
  for (int i = lane; i < H; i += 32) {
    const uint32_t v = hits[i];
    int rank = 0;
    for (int t = 0; t < H; ++t) rank += (hits[t] < v) ? 1 : 0;
    if (rank < nsample) out[rank] = (int32_t)v;
    firstv = min(firstv, v);
  }
  firstv = __reduce_min_sync(FULL, firstv);
  const int fill = (H == 0) ? 0 : (int)firstv;
  for (int l = min(H, nsample) + lane; l < nsample; l += 32) out[l] = fill;
}

}  // namespace

extern "C" long long sad_scene_grid_workspace_bytes(int B, int N) {
  if (B < 0 || N < 1) return -1;
  return (long long)B * (long long)sad::grid_stride(N);
}

extern "C" int sad_scene_grid_build(int B, int N, const float* xyz, void* workspace, sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && N >= 1, "scene_grid_build: bad sizes B=%d N=%d", B, N);
  if (B == 0) return SAD_OK;
  SAD_REQUIRE(xyz && workspace, "scene_grid_build: null pointer");
  SAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "scene_grid_build: workspace must be 16-byte aligned");
  static thread_local int configured_dev = -1;
  int dev = 0;
  SAD_CUDA_OK(cudaGetDevice(&dev));
  const size_t smem = (size_t)sad::kGridCells * sizeof(uint32_t);
  if (configured_dev != dev) {
    SAD_CUDA_OK(cudaFuncSetAttribute(grid_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_dev = dev;
  }
  grid_build_kernel<<<B, GB_T, smem, (cudaStream_t)stream>>>(N, xyz, static_cast<uint8_t*>(workspace),
                                                              sad::grid_stride(N));
  SAD_LAUNCH_CHECK("grid_build_kernel");
  return SAD_OK;
}

extern "C" int sad_ball_query_grid_fwd(int B, int N, int npoint, float radius, const float* radius_t, int nsample,
                                       const float* xyz, const void* grid_ws, const float* new_xyz, int32_t* idx,
                                       sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && N >= 1 && npoint >= 0 && nsample >= 1, "ball_query_grid: bad sizes");
  if (B == 0 || npoint == 0) return SAD_OK;
  SAD_REQUIRE(xyz && grid_ws && new_xyz && idx, "ball_query_grid: null pointer");
  SAD_REQUIRE(B <= 65535, "ball_query_grid: batch too large");
  dim3 grid((unsigned)sad_ceil_div(npoint, BQG_WARPS), (unsigned)B);
  ball_query_grid_kernel<<<grid, BQG_WARPS * 32, 0, (cudaStream_t)stream>>>(
      N, npoint, radius, radius_t, nsample, xyz, static_cast<const uint8_t*>(grid_ws), sad::grid_stride(N), new_xyz, idx);
  SAD_LAUNCH_CHECK("ball_query_grid_kernel");
  return SAD_OK;
}
