// a9 three_interpolate (forward + backward) -- SURVEY.md section 8(a) row a9.
// (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
// HBM-bound: out[b,c,i] = ((w0*f[i0]) + (w1*f[i1])) + (w2*f[i2]), evaluated in exactly
// that order with no FMA contraction so the fp32 result is bit-identical to the
// oracle.  One thread owns four consecutive unknown points: 12 indices + 12 weights are
// loaded once with 128-bit loads and reused over a chunk of channels; per channel the
// thread issues 12 independent 4-byte gathers (source rows are L1/L2 resident) and one
// 128-bit streaming store (a warp writes 512 contiguous bytes).
#include <stdlib.h>

#include "sad_common.cuh"

namespace {

constexpr int TI_T = 128;
constexpr int TI_CCH = 16;

// untrusted indices: forward gathers clamp into [0, m), the backward scatter skips anything outside
__device__ __forceinline__ int clamp_idx(int i, int m) { return (int)min((unsigned)i, (unsigned)(m - 1)); }

__device__ __forceinline__ float interp3(float w0, float f0, float w1, float f1, float w2, float f2) {
  return __fadd_rn(__fadd_rn(__fmul_rn(w0, f0), __fmul_rn(w1, f1)), __fmul_rn(w2, f2));
}

template <bool VEC>
__global__ void __launch_bounds__(TI_T)
interp_fwd_kernel(int C, int m, int n, const float* __restrict__ features, const int32_t* __restrict__ idx,
                  const float* __restrict__ weight, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * TI_CCH;
  const int cn = min(TI_CCH, C - c0);
  const int t = blockIdx.x * TI_T + threadIdx.x;
  const float* f = features + ((size_t)b * C + c0) * m;
  if (VEC) {
    if (t * 4 >= n) return;
    int id[12];
    float w[12];
    const int4* ip = reinterpret_cast<const int4*>(idx + (size_t)b * n * 3) + (size_t)t * 3;
    const float4* wp = reinterpret_cast<const float4*>(weight + (size_t)b * n * 3) + (size_t)t * 3;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int4 a = __ldg(ip + u);
      const float4 ww = __ldg(wp + u);
      id[4 * u] = clamp_idx(a.x, m); id[4 * u + 1] = clamp_idx(a.y, m); id[4 * u + 2] = clamp_idx(a.z, m); id[4 * u + 3] = clamp_idx(a.w, m);
      w[4 * u] = ww.x; w[4 * u + 1] = ww.y; w[4 * u + 2] = ww.z; w[4 * u + 3] = ww.w;
    }
    float* o = out + ((size_t)b * C + c0) * n + (size_t)t * 4;
#pragma unroll 2
    for (int c = 0; c < cn; ++c) {
      const float* fc = f + (size_t)c * m;
      float g[12];
#pragma unroll
      for (int u = 0; u < 12; ++u) g[u] = __ldg(fc + id[u]);
      float4 v;
      v.x = interp3(w[0], g[0], w[1], g[1], w[2], g[2]);
      v.y = interp3(w[3], g[3], w[4], g[4], w[5], g[5]);
      v.z = interp3(w[6], g[6], w[7], g[7], w[8], g[8]);
      v.w = interp3(w[9], g[9], w[10], g[10], w[11], g[11]);
      __stcs(reinterpret_cast<float4*>(o + (size_t)c * n), v);
    }
  } else {
    if (t >= n) return;
    const int32_t* ip = idx + ((size_t)b * n + t) * 3;
    const float* wp = weight + ((size_t)b * n + t) * 3;
    const int i0 = clamp_idx(__ldg(ip), m), i1 = clamp_idx(__ldg(ip + 1), m), i2 = clamp_idx(__ldg(ip + 2), m);
    const float w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
    float* o = out + ((size_t)b * C + c0) * n + t;
#pragma unroll 4
    for (int c = 0; c < cn; ++c) {
      const float* fc = f + (size_t)c * m;
      __stcs(o + (size_t)c * n, interp3(w0, __ldg(fc + i0), w1, __ldg(fc + i1), w2, __ldg(fc + i2)));
    }
  }
}

// Row-staged variant: the chunk f[b, c0:c0+cn, :] (contiguous in the channel-first tensor) is brought into shared
// memory by one TMA bulk copy; the 12 gathers per thread and channel then hit shared-memory banks (~3 wavefronts per
// warp-wide gather instead of up to 32 L1 sector lookups).  Measured 52-58 % of the HBM peak at B >= 64; the bound
// is the shared-memory gather rate (12 four-byte gathers per 16 output bytes); a transposed [i][c] staging with
// 8-byte gathers was tried and lost more in the staging stores than it gained.
constexpr int TIS_T = 256;
__global__ void __launch_bounds__(TIS_T)
interp_fwd_staged_kernel(int C, int m, int quads, int cch, const float* __restrict__ features,
                         const int32_t* __restrict__ idx, const float* __restrict__ weight, float* __restrict__ out) {
  extern __shared__ __align__(128) float s_rows[];          // [cn][m]
  __shared__ __align__(8) uint64_t s_bar;
  using namespace sad;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * cch;
  const int cn = min(cch, C - c0);
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    const uint32_t bytes = (uint32_t)cn * (uint32_t)m * 4u;
    mbar_arrive_expect_tx(&s_bar, bytes);
    tma_bulk_g2s(s_rows, features + ((size_t)b * C + c0) * m, bytes, &s_bar);
  }
  __syncthreads();
  const int per = (quads + gridDim.x - 1) / gridDim.x;
  const int q0 = blockIdx.x * per, q1 = min(quads, q0 + per);
  const int4* ipb = reinterpret_cast<const int4*>(idx) + (size_t)b * quads * 3;
  const float4* wpb = reinterpret_cast<const float4*>(weight) + (size_t)b * quads * 3;
  float4* op = reinterpret_cast<float4*>(out) + ((size_t)b * C + c0) * quads;
  bool waited = false;
  for (int q = q0 + threadIdx.x; q < q1; q += TIS_T) {
    int id[12];
    float w[12];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int4 a = __ldg(ipb + (size_t)q * 3 + u);
      const float4 ww = __ldg(wpb + (size_t)q * 3 + u);
      id[4 * u] = clamp_idx(a.x, m); id[4 * u + 1] = clamp_idx(a.y, m); id[4 * u + 2] = clamp_idx(a.z, m); id[4 * u + 3] = clamp_idx(a.w, m);
      w[4 * u] = ww.x; w[4 * u + 1] = ww.y; w[4 * u + 2] = ww.z; w[4 * u + 3] = ww.w;
    }
    if (!waited) {
      mbar_wait(&s_bar, 0);
      waited = true;
    }
#pragma unroll 2
    for (int c = 0; c < cn; ++c) {
      const float* r = s_rows + (size_t)c * m;
      float4 v;
      v.x = interp3(w[0], r[id[0]], w[1], r[id[1]], w[2], r[id[2]]);
      v.y = interp3(w[3], r[id[3]], w[4], r[id[4]], w[5], r[id[5]]);
      v.z = interp3(w[6], r[id[6]], w[7], r[id[7]], w[8], r[id[8]]);
      v.w = interp3(w[9], r[id[9]], w[10], r[id[10]], w[11], r[id[11]]);
      __stcs(op + (size_t)c * quads + q, v);
    }
  }
  if (!waited) mbar_wait(&s_bar, 0);      // never leave with the bulk copy still in flight
}

// Point-major variant (the fast path).  What bounds the row-staged kernel above is the LSU wavefront rate (~45 per
// 128 outputs: every 4-byte gather of a warp scatters over a channel row), so this one is built to need ~27:
//   * the CTA's channel chunk is staged TRANSPOSED, s_f[k][CN] (one known point = one row), the 16-byte units of
//     every 32-channel segment XOR-swizzled by (k & 7).  Lane = (quarter Q, unit q): a quarter-warp reads one whole
//     128-byte row segment per gather -- one wavefront per quarter, conflict-free for ANY index pattern;
//   * Q owns 4 consecutive unknown points (their 12 indices / 12 weights are three 16-byte loads each, shared by the
//     quarter's lanes through L1), q owns 4 consecutive channels: a lane finishes a 4 x 4 block per half-step;
//   * the warp's 32-channel x 32-point result is transposed through a swizzled 4 KB shared-memory tile so that the
//     global stores are whole 128-byte row pieces (a quarter-warp = one row), not 16-byte slivers of 8 rows.
// Bit-exact like the other variants (same ((w0*f0)+(w1*f1))+(w2*f2), no contraction).
constexpr int TIP_T = 256;
constexpr int TIP_TILE = 32 * 32;                            // floats of one warp's transposition tile
template <int CN>
__global__ void __launch_bounds__(TIP_T, 2)
interp_fwd_pm_kernel(int C, int m, int n, int groups_per_cta, const float* __restrict__ features,
                     const int32_t* __restrict__ idx, const float* __restrict__ weight, float* __restrict__ out) {
  extern __shared__ __align__(128) float s_dyn[];
  float* s_t = s_dyn;                                       // [warps][32 channels][32 points], quads swizzled by unit
  float* s_f = s_dyn + (TIP_T / 32) * TIP_TILE;             // [m][CN], units swizzled inside 32-channel segments
  constexpr int UNITS = CN / 4;
  const int b = blockIdx.z, c0 = blockIdx.y * CN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- stage: thread = point k, 8 units (32 channels) per pass: 32 independent coalesced 4-byte loads in flight
  // (lanes = consecutive k), then eight conflict-free 16-byte stores
  const float* fb = features + ((size_t)b * C + c0) * m;
  for (int k = tid; k < m; k += TIP_T) {
#pragma unroll
    for (int u0 = 0; u0 < UNITS; u0 += 8) {
      float v[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = (c0 + 4 * u0 + e < C) ? __ldg(fb + (size_t)(4 * u0 + e) * m + k) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        *reinterpret_cast<float4*>(s_f + (size_t)k * CN + 4 * (u0 + (u ^ (k & 7)))) =
            make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
    }
  }
  __syncthreads();
  // ---- gather: a warp = 32 unknown points x one 32-channel segment per step (two half-steps of 16 points)
  constexpr int SEGS = CN / 32;
  const int seg = warp % SEGS;
  const int Q = lane >> 3, q = lane & 7;
  const int g_begin = blockIdx.x * groups_per_cta, g_end = min(g_begin + groups_per_cta, (n + 31) >> 5);
  const int4* ib = reinterpret_cast<const int4*>(idx + (size_t)b * n * 3);
  const float4* wb = reinterpret_cast<const float4*>(weight + (size_t)b * n * 3);
  float* tile = s_t + warp * TIP_TILE;
  const float* fseg = s_f + seg * 32;
  for (int g = g_begin + warp / SEGS; g < g_end; g += (TIP_T / 32) / SEGS) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int i4 = g * 32 + half * 16 + 4 * Q;              // this lane's 4 points (n % 4 == 0: all in or all out)
      int id[12];
      float w[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        id[e] = 0;
        w[e] = 0.f;
      }
      if (i4 < n) {
#pragma unroll
        for (int e = 0; e < 3; ++e) {
          const int4 a = __ldg(ib + (size_t)(i4 >> 2) * 3 + e);
          const float4 ww = __ldg(wb + (size_t)(i4 >> 2) * 3 + e);
          id[4 * e] = clamp_idx(a.x, m); id[4 * e + 1] = clamp_idx(a.y, m); id[4 * e + 2] = clamp_idx(a.z, m); id[4 * e + 3] = clamp_idx(a.w, m);
          w[4 * e] = ww.x; w[4 * e + 1] = ww.y; w[4 * e + 2] = ww.z; w[4 * e + 3] = ww.w;
        }
      }
      float acc[4][4];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        const int r0 = id[3 * pp], r1 = id[3 * pp + 1], r2 = id[3 * pp + 2];
        const float4 f0 = *reinterpret_cast<const float4*>(fseg + (size_t)r0 * CN + 4 * (q ^ (r0 & 7)));
        const float4 f1 = *reinterpret_cast<const float4*>(fseg + (size_t)r1 * CN + 4 * (q ^ (r1 & 7)));
        const float4 f2 = *reinterpret_cast<const float4*>(fseg + (size_t)r2 * CN + 4 * (q ^ (r2 & 7)));
        const float a0 = w[3 * pp], a1 = w[3 * pp + 1], a2 = w[3 * pp + 2];
        acc[pp][0] = interp3(a0, f0.x, a1, f1.x, a2, f2.x);
        acc[pp][1] = interp3(a0, f0.y, a1, f1.y, a2, f2.y);
        acc[pp][2] = interp3(a0, f0.z, a1, f1.z, a2, f2.z);
        acc[pp][3] = interp3(a0, f0.w, a1, f1.w, a2, f2.w);
      }
      // tile[channel 4q + j][point quad (Q + 4 * half) ^ q]: the 8 lanes of a quarter hit 8 different 16-byte slots
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(tile + (4 * q + j) * 32 + 4 * ((Q + 4 * half) ^ q)) =
            make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
    }
    __syncwarp();
    // ---- rows out: a quarter-warp = one channel row, 8 consecutive point quads = 128 contiguous bytes
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = Q + 4 * it;
      const float4 v = *reinterpret_cast<const float4*>(tile + row * 32 + 4 * (q ^ ((row >> 2) & 7)));
      const int c = c0 + seg * 32 + row, i = g * 32 + 4 * q;
      if (c < C && i < n) __stcs(reinterpret_cast<float4*>(out + ((size_t)b * C + c) * n + i), v);
    }
    __syncwarp();
  }
}

// Pipelined point-major variant (FP-module shapes, m <= 512): what held the two kernels above at 51-59 % of the HBM
// peak was not their gather rate but that every CTA ran stage -> barrier -> compute, with its loads exposed in the
// first phase and nothing but stores in flight in the second.  Here CTAs are persistent over (scene, 32-channel chunk)
// work items and the chunk f[b, c0:c0+32, :] -- contiguous in the channel-first tensor -- arrives by ONE TMA bulk copy
// that was issued a whole work item earlier:
//   wait(raw full) -> transpose raw[c][k] -> s_f[k][32] (shared -> shared, conflict-free both ways) -> barrier ->
//   thread 0 issues the bulk copy of the NEXT item into raw (already free) -> gather/compute/store from s_f -> barrier
// so HBM reads stream in the background of the gather phase and the read and write streams overlap inside one CTA.
// The gather phase is the point-major one of interp_fwd_pm_kernel (same lane roles, same transposition tile).
template <int T>
__global__ void __launch_bounds__(T, T == 256 ? 2 : 1)
interp_fwd_pipe_kernel(int B, int C, int m, int n, int ychunks, const float* __restrict__ features,
                       const int32_t* __restrict__ idx, const float* __restrict__ weight, float* __restrict__ out) {
  extern __shared__ __align__(128) float s_dyn[];
  __shared__ __align__(8) uint64_t s_bar;
  using namespace sad;
  constexpr int CN = 32;
  float* s_t = s_dyn;                                       // [warps][32 channels][32 points]
  float* s_f = s_t + (T / 32) * TIP_TILE;               // [m][32], 16-byte units swizzled by (k & 7)
  float* s_raw = s_f + (size_t)m * CN;                      // [32][m]: the bulk copy's landing zone
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int items = B * ychunks;
  const int per = (items + gridDim.x - 1) / gridDim.x;      // contiguous ranges: a CTA stays on one scene's idx / weights
  const int w_begin = blockIdx.x * per, w_end = min(items, w_begin + per);
  auto issue = [&](int w) {
    const int b = w / ychunks, c0 = (w - b * ychunks) * CN;
    const uint32_t bytes = (uint32_t)min(CN, C - c0) * (uint32_t)m * 4u;
    mbar_arrive_expect_tx(&s_bar, bytes);
    tma_bulk_g2s(s_raw, features + ((size_t)b * C + c0) * m, bytes, &s_bar);
  };
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    if (w_begin < w_end) issue(w_begin);
  }
  __syncthreads();
  const int Q = lane >> 3, q = lane & 7;
  float* tile = s_t + warp * TIP_TILE;
  const int groups = (n + 31) >> 5;
  uint32_t phase = 0;
  for (int w = w_begin; w < w_end; ++w, phase ^= 1u) {
    const int b = w / ychunks, c0 = (w - b * ychunks) * CN;
    mbar_wait(&s_bar, phase);
    // ---- transpose: thread = known point k; 32 conflict-free 4-byte reads, eight conflict-free 16-byte writes
    for (int k = tid; k < m; k += T) {
      float v[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = s_raw[(size_t)e * m + k];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        *reinterpret_cast<float4*>(s_f + (size_t)k * CN + 4 * (u ^ (k & 7))) =
            make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
    }
    __syncthreads();
    if (tid == 0 && w + 1 < w_end) issue(w + 1);            // raw is free: the next chunk streams in under the gathers
    // ---- gather: a warp = 32 unknown points per step (two half-steps of 16 points)
    const int4* ib = reinterpret_cast<const int4*>(idx + (size_t)b * n * 3);
    const float4* wb = reinterpret_cast<const float4*>(weight + (size_t)b * n * 3);
    for (int g = warp; g < groups; g += T / 32) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i4 = g * 32 + half * 16 + 4 * Q;
        int id[12];
        float wt[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) {
          id[e] = 0;
          wt[e] = 0.f;
        }
        if (i4 < n) {
#pragma unroll
          for (int e = 0; e < 3; ++e) {
            const int4 a = __ldg(ib + (size_t)(i4 >> 2) * 3 + e);
            const float4 ww = __ldg(wb + (size_t)(i4 >> 2) * 3 + e);
            id[4 * e] = clamp_idx(a.x, m); id[4 * e + 1] = clamp_idx(a.y, m); id[4 * e + 2] = clamp_idx(a.z, m); id[4 * e + 3] = clamp_idx(a.w, m);
            wt[4 * e] = ww.x; wt[4 * e + 1] = ww.y; wt[4 * e + 2] = ww.z; wt[4 * e + 3] = ww.w;
          }
        }
        float acc[4][4];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const int r0 = id[3 * pp], r1 = id[3 * pp + 1], r2 = id[3 * pp + 2];
          const float4 f0 = *reinterpret_cast<const float4*>(s_f + (size_t)r0 * CN + 4 * (q ^ (r0 & 7)));
          const float4 f1 = *reinterpret_cast<const float4*>(s_f + (size_t)r1 * CN + 4 * (q ^ (r1 & 7)));
          const float4 f2 = *reinterpret_cast<const float4*>(s_f + (size_t)r2 * CN + 4 * (q ^ (r2 & 7)));
          const float a0 = wt[3 * pp], a1 = wt[3 * pp + 1], a2 = wt[3 * pp + 2];
          acc[pp][0] = interp3(a0, f0.x, a1, f1.x, a2, f2.x);
          acc[pp][1] = interp3(a0, f0.y, a1, f1.y, a2, f2.y);
          acc[pp][2] = interp3(a0, f0.z, a1, f1.z, a2, f2.z);
          acc[pp][3] = interp3(a0, f0.w, a1, f1.w, a2, f2.w);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(tile + (4 * q + j) * 32 + 4 * ((Q + 4 * half) ^ q)) =
              make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = Q + 4 * it;
        const float4 v = *reinterpret_cast<const float4*>(tile + row * 32 + 4 * (q ^ ((row >> 2) & 7)));
        const int c = c0 + row, i = g * 32 + 4 * q;
        if (c < C && i < n) __stcs(reinterpret_cast<float4*>(out + ((size_t)b * C + c) * n + i), v);
      }
      __syncwarp();
    }
    __syncthreads();                                        // s_f is rewritten by the next item's transposition
  }
}

__global__ void __launch_bounds__(TI_T)
interp_bwd_kernel(int C, int n, int m, const float* __restrict__ grad_out, const int32_t* __restrict__ idx,
                  const float* __restrict__ weight, float* __restrict__ grad_features) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * TI_CCH;
  const int cn = min(TI_CCH, C - c0);
  const int t = blockIdx.x * TI_T + threadIdx.x;
  if (t >= n) return;
  const int32_t* ip = idx + ((size_t)b * n + t) * 3;
  const float* wp = weight + ((size_t)b * n + t) * 3;
  const int i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
  const float w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
  const float* go = grad_out + ((size_t)b * C + c0) * n + t;
  float* g = grad_features + ((size_t)b * C + c0) * m;
#pragma unroll 4
  for (int c = 0; c < cn; ++c) {
    const float v = __ldcs(go + (size_t)c * n);
    float* gc = g + (size_t)c * m;
    if ((unsigned)i0 < (unsigned)m) atomicAdd(gc + i0, __fmul_rn(v, w0));
    if ((unsigned)i1 < (unsigned)m) atomicAdd(gc + i1, __fmul_rn(v, w1));
    if ((unsigned)i2 < (unsigned)m) atomicAdd(gc + i2, __fmul_rn(v, w2));
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" int sad_three_interpolate_fwd(int B, int C, int m, int n, const float* features, const int32_t* idx,
                                         const float* weight, float* out, sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && C >= 0 && m >= 1 && n >= 0, "three_interpolate: bad sizes B=%d C=%d m=%d n=%d", B, C, m, n);
  if (B == 0 || C == 0 || n == 0) return SAD_OK;
  SAD_REQUIRE(features && idx && weight && out, "three_interpolate: null pointer");
  SAD_REQUIRE(B <= 65535 && sad_ceil_div(C, TI_CCH) <= 65535, "three_interpolate: B/C exceed grid limits");
  const bool vec = (n % 4 == 0) && aligned16(idx) && aligned16(weight) && aligned16(out);
  // point-major staged kernel: rows fit shared memory and are re-read (n * 3 gathers over m rows)
  // (measured, B = 256, C = 256: n=1024/m=512 58.6 % of the HBM peak vs 57.6 % row-staged; n=512/m=256 51 % vs 53 % --
  // both kernels sit at ~70 % of the SM's LSU wavefront rate; the point-major kernel takes
  // the larger-m shapes, where its lower wavefront count per output wins)
  // pipelined point-major kernel: FP-module shapes (the chunk, its transposed copy and the tiles fit one CTA)
  if (n % 4 == 0 && m % 4 == 0 && aligned16(out) && aligned16(idx) && aligned16(weight) && aligned16(features) && m >= 64 &&
      m <= 512 && 2LL * n >= m && n >= 256 && !sad_tool_env("SAD_INTERP_LEGACY")) {
    const int ychunks = sad_ceil_div(C, 32);
    const long long items = (long long)B * ychunks;
    // two 256-thread CTAs per SM while they fit (m <= 256), else one 512-thread CTA
    const bool small = ((size_t)m * 64 + 8 * TIP_TILE) * sizeof(float) <= 100 * 1024;
    const int threads = small ? 256 : 512;
    const size_t smem = ((size_t)m * 64 + (threads / 32) * TIP_TILE) * sizeof(float);
    const int ctas_per_sm = small ? 2 : 1;
    static thread_local int configured_dev_pipe = -1;
    int dev = 0;
    SAD_CUDA_OK(cudaGetDevice(&dev));
    if (configured_dev_pipe != dev) {
      SAD_CUDA_OK(cudaFuncSetAttribute(interp_fwd_pipe_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      SAD_CUDA_OK(cudaFuncSetAttribute(interp_fwd_pipe_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (512 * 64 + 16 * TIP_TILE) * 4));
      configured_dev_pipe = dev;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = (long long)sms * ctas_per_sm;
    if (grid > items) grid = items;
    if (small)
      interp_fwd_pipe_kernel<256><<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(B, C, m, n, ychunks, features, idx, weight, out);
    else
      interp_fwd_pipe_kernel<512><<<(unsigned)grid, 512, smem, (cudaStream_t)stream>>>(B, C, m, n, ychunks, features, idx, weight, out);
    SAD_LAUNCH_CHECK("three_interpolate");
    return SAD_OK;
  }
  if (n % 4 == 0 && aligned16(out) && aligned16(idx) && aligned16(weight) && m >= 384 && m <= 1408 && 2LL * n >= m &&
      !sad_tool_env("SAD_INTERP_LEGACY")) {
    const int cn = (m <= 128 && C > 32) ? 64 : 32;
    const int ychunks = sad_ceil_div(C, cn);
    SAD_REQUIRE(ychunks <= 65535, "three_interpolate: C exceeds grid limits");
    const int groups = sad_ceil_div(n, 32);
    const int gstep = (TIP_T / 32) / (cn / 32);                         // groups one pass of the CTA covers
    long long x = sad_ceil_div(444, (long long)ychunks * B);
    const long long xmax = groups / gstep > 1 ? groups / gstep : 1;     // at least one full pass per CTA
    if (x > xmax) x = xmax;
    if (x < 1) x = 1;
    const int gpc = sad_ceil_div(sad_ceil_div(groups, x), gstep) * gstep;
    x = sad_ceil_div(groups, gpc);
    const size_t smem = ((size_t)m * cn + (TIP_T / 32) * TIP_TILE) * sizeof(float);
    static thread_local int configured_dev_pm = -1;
    int dev = 0;
    SAD_CUDA_OK(cudaGetDevice(&dev));
    if (configured_dev_pm != dev) {
      SAD_CUDA_OK(cudaFuncSetAttribute(interp_fwd_pm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (1408 * 32 + (TIP_T / 32) * TIP_TILE) * 4));
      SAD_CUDA_OK(cudaFuncSetAttribute(interp_fwd_pm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (128 * 64 + (TIP_T / 32) * TIP_TILE) * 4));
      configured_dev_pm = dev;
    }
    dim3 g3((unsigned)x, (unsigned)ychunks, (unsigned)B);
    if (cn == 64)
      interp_fwd_pm_kernel<64><<<g3, TIP_T, smem, (cudaStream_t)stream>>>(C, m, n, gpc, features, idx, weight, out);
    else
      interp_fwd_pm_kernel<32><<<g3, TIP_T, smem, (cudaStream_t)stream>>>(C, m, n, gpc, features, idx, weight, out);
    SAD_LAUNCH_CHECK("three_interpolate");
    return SAD_OK;
  }
  const int cch = (m % 4 == 0 && m <= 4608) ? (18432 / m < 16 ? 18432 / m : 16) : 0;
  if (vec && cch >= 4 && 2LL * n >= m && aligned16(features)) {
    const int quads = n / 4;
    const int ychunks = sad_ceil_div(C, cch);
    long long x = sad_ceil_div(444, (long long)ychunks * B);
    const long long xmax = quads / TIS_T > 1 ? quads / TIS_T : 1;       // at least one full pass of the CTA
    if (x > xmax) x = xmax;
    if (x < 1) x = 1;
    SAD_REQUIRE(ychunks <= 65535, "three_interpolate: C exceeds grid limits");
    static thread_local int configured_dev = -1;
    int dev = 0;
    SAD_CUDA_OK(cudaGetDevice(&dev));
    if (configured_dev != dev) {
      SAD_CUDA_OK(cudaFuncSetAttribute(interp_fwd_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 73728));
      configured_dev = dev;
    }
    dim3 g2((unsigned)x, (unsigned)ychunks, (unsigned)B);
    interp_fwd_staged_kernel<<<g2, TIS_T, (size_t)cch * m * sizeof(float), (cudaStream_t)stream>>>(
        C, m, quads, cch, features, idx, weight, out);
    SAD_LAUNCH_CHECK("three_interpolate");
    return SAD_OK;
  }
  dim3 grid((unsigned)sad_ceil_div(vec ? n / 4 : n, TI_T), (unsigned)sad_ceil_div(C, TI_CCH), (unsigned)B);
  if (vec)
    interp_fwd_kernel<true><<<grid, TI_T, 0, (cudaStream_t)stream>>>(C, m, n, features, idx, weight, out);
  else
    interp_fwd_kernel<false><<<grid, TI_T, 0, (cudaStream_t)stream>>>(C, m, n, features, idx, weight, out);
  SAD_LAUNCH_CHECK("three_interpolate");
  return SAD_OK;
}

extern "C" int sad_three_interpolate_bwd(int B, int C, int n, int m, const float* grad_out, const int32_t* idx,
                                         const float* weight, float* grad_features, sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && C >= 0 && m >= 1 && n >= 0, "three_interpolate_bwd: bad sizes B=%d C=%d n=%d m=%d", B, C, n, m);
  if (B == 0 || C == 0) return SAD_OK;
  SAD_REQUIRE(grad_features, "three_interpolate_bwd: null pointer");
  SAD_CUDA_OK(cudaMemsetAsync(grad_features, 0, (size_t)B * C * m * sizeof(float), (cudaStream_t)stream));
  if (n == 0) return SAD_OK;
  SAD_REQUIRE(grad_out && idx && weight, "three_interpolate_bwd: null pointer");
  SAD_REQUIRE(B <= 65535 && sad_ceil_div(C, TI_CCH) <= 65535, "three_interpolate_bwd: B/C exceed grid limits");
  dim3 grid((unsigned)sad_ceil_div(n, TI_T), (unsigned)sad_ceil_div(C, TI_CCH), (unsigned)B);
  interp_bwd_kernel<<<grid, TI_T, 0, (cudaStream_t)stream>>>(C, n, m, grad_out, idx, weight, grad_features);
  SAD_LAUNCH_CHECK("three_interpolate_bwd");
  return SAD_OK;
}
