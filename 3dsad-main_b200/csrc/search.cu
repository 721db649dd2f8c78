// a3 ball_query, a4 ball_query_adaptive, a8 three_nn -- SURVEY.md section 8(a).
// (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
// B200 design: a tiled neighbour search.  The candidate cloud of one scene is streamed
// through shared memory in tiles by the TMA engine (1-D cp.async.bulk + mbarrier,
// double-buffered, SASS UBLKCP) while the warps of the CTA scan the previous tile.
// One warp owns QPW queries; each lane tests one candidate per step (stride-3 smem
// reads are bank-conflict free), hits are compacted IN ASCENDING INDEX ORDER with
// ballot + popc (never atomics), and a warp / CTA stops as soon as its queries are
// full.  The distance test follows the arithmetic contract H1 bit for bit.
#include "sad_common.cuh"

namespace {

using namespace sad;

constexpr int BQ_T = 256;            // 8 warps
constexpr int BQ_TILE = 2048;        // candidates per stage (24 KB)
constexpr int BQ_STAGES = 2;

// Cooperative tile load: bulk-TMA for the 16-byte-aligned body, plain loads for the tail.
// The tile is padded to a multiple of 32 points with +inf coordinates (d2 = +inf is never
// < r*r), so the scan loop needs no per-lane bounds check.
__device__ __forceinline__ void load_tile(float* dst, const float* src, int npts, bool bulk_ok,
                                          uint64_t* bar, int tid, int nthreads) {
  const uint32_t bytes = (uint32_t)npts * 12u;
  if (bulk_ok) {
    const uint32_t body = bytes & ~15u;
    if (tid == 0) {
      mbar_arrive_expect_tx(bar, body);
      if (body) tma_bulk_g2s(dst, src, body, bar);
    }
    const int tail0 = (int)(body >> 2), tail1 = (int)(bytes >> 2);
    if (tid >= 32 && tid < 32 + (tail1 - tail0)) dst[tail0 + tid - 32] = __ldg(src + tail0 + tid - 32);
  } else {
    for (int i = tid; i < npts * 3; i += nthreads) dst[i] = __ldg(src + i);
  }
  const int pad_end = ((npts + 31) & ~31) * 3;
  const int i = npts * 3 + tid - 64;
  if (tid >= 64 && i < pad_end) dst[i] = __int_as_float(0x7f800000);
}

template <int QPW>
__global__ void __launch_bounds__(BQ_T)
ball_query_kernel(int N, int npoint, float radius, const float* __restrict__ radius_t, int nsample,
                  const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                  int32_t* __restrict__ idx) {
  extern __shared__ __align__(128) float s_tile[];        // [BQ_STAGES][BQ_TILE*3]
  __shared__ __align__(8) uint64_t s_full[BQ_STAGES];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const float* pts = xyz + (size_t)b * N * 3;
  const bool bulk_ok = ((reinterpret_cast<uintptr_t>(pts) & 15u) == 0);
  const uint32_t lt_mask = (1u << lane) - 1u;

  float qx[QPW], qy[QPW], qz[QPW], r2[QPW];
  int cnt[QPW], first[QPW];
  int32_t* out[QPW];
#pragma unroll
  for (int q = 0; q < QPW; ++q) {
    const int j = (blockIdx.x * (BQ_T / 32) + warp) * QPW + q;
    first[q] = 0;
    if (j < npoint) {
      const float* c = new_xyz + ((size_t)b * npoint + j) * 3;
      qx[q] = __ldg(c);
      qy[q] = __ldg(c + 1);
      qz[q] = __ldg(c + 2);
      const float r = radius_t ? __ldg(radius_t + (size_t)b * npoint + j) : radius;
      r2[q] = __fmul_rn(r, r);
      cnt[q] = 0;
      out[q] = idx + ((size_t)b * npoint + j) * nsample;
    } else {
      qx[q] = qy[q] = qz[q] = 0.f;
      r2[q] = -1.f;         // inactive query: nothing is ever < -1
      cnt[q] = nsample;     // ... and it counts as already "full"
      out[q] = nullptr;
    }
  }

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < BQ_STAGES; ++s) mbar_init(&s_full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int ntiles = (N + BQ_TILE - 1) / BQ_TILE;
#pragma unroll
  for (int s = 0; s < BQ_STAGES; ++s)
    if (s < ntiles)
      load_tile(s_tile + s * BQ_TILE * 3, pts + (size_t)s * BQ_TILE * 3, min(BQ_TILE, N - s * BQ_TILE),
                bulk_ok, &s_full[s], tid, BQ_T);
  __syncthreads();

  int issued = min(ntiles, BQ_STAGES);     // tiles whose load has been issued
  int t = 0;
  for (; t < ntiles; ++t) {
    const int stage = t % BQ_STAGES;
    const uint32_t parity = (uint32_t)((t / BQ_STAGES) & 1);
    const float* tp = s_tile + stage * BQ_TILE * 3 + lane * 3;
    const int nchunks = (min(BQ_TILE, N - t * BQ_TILE) + 31) >> 5;
    bool warp_done = true;
#pragma unroll
    for (int q = 0; q < QPW; ++q) warp_done = warp_done && (cnt[q] >= nsample);

    if (bulk_ok) mbar_wait(&s_full[stage], parity);
    if (!warp_done) {
      int cbase = t * BQ_TILE + lane;        // candidate index of this lane in the current chunk
      for (int ch = 0; ch < nchunks; ++ch, tp += 96, cbase += 32) {
        const float x = tp[0], y = tp[1], z = tp[2];
        bool hit[QPW];
        bool anyhit = false;
#pragma unroll
        for (int q = 0; q < QPW; ++q) {
          hit[q] = sqdist(x, y, z, qx[q], qy[q], qz[q]) < r2[q];
          anyhit = anyhit || hit[q];
        }
        if (__any_sync(FULL, anyhit)) {      // rare path: ordered compaction with ballot + popc
          bool all_full = true;
#pragma unroll
          for (int q = 0; q < QPW; ++q) {
            const uint32_t m = __ballot_sync(FULL, hit[q]);
            if (m != 0u && cnt[q] < nsample) {
              const int pos = cnt[q] + __popc(m & lt_mask);
              if (hit[q] && pos < nsample) out[q][pos] = cbase;
              if (cnt[q] == 0) first[q] = cbase - lane + __ffs(m) - 1;
              cnt[q] += __popc(m);
            }
            all_full = all_full && (cnt[q] >= nsample);
          }
          if (all_full) {
            warp_done = true;
            break;
          }
        }
      }
    }
    // everyone is finished with this stage (and we learn whether the whole CTA is full)
    const int cta_done = __syncthreads_and(warp_done ? 1 : 0);
    if (cta_done) {
      ++t;
      break;
    }
    if (issued < ntiles) {
      load_tile(s_tile + stage * BQ_TILE * 3, pts + (size_t)issued * BQ_TILE * 3,
                min(BQ_TILE, N - issued * BQ_TILE), bulk_ok, &s_full[stage], tid, BQ_T);
      ++issued;
    }
  }
  // drain bulk copies still in flight before the CTA (and its shared memory) retires
  if (bulk_ok && tid == 0)
    for (int u = t; u < issued; ++u) mbar_wait(&s_full[u % BQ_STAGES], (uint32_t)((u / BQ_STAGES) & 1));

  // padding: first hit fills the unused slots; an empty ball is all zeros
#pragma unroll
  for (int q = 0; q < QPW; ++q) {
    if (out[q] == nullptr) continue;
    const int have = min(cnt[q], nsample);
    const int fill = (cnt[q] > 0) ? first[q] : 0;
    for (int s = have + lane; s < nsample; s += 32) out[q][s] = fill;
  }
}

int launch_ball_query(int B, int N, int npoint, float radius, const float* radius_t, int nsample,
                      const float* xyz, const float* new_xyz, int32_t* idx, cudaStream_t stream) {
  SAD_REQUIRE(B >= 0 && N >= 1 && npoint >= 0 && nsample >= 1, "ball_query: bad sizes B=%d N=%d npoint=%d nsample=%d",
              B, N, npoint, nsample);
  if (B == 0 || npoint == 0) return SAD_OK;
  SAD_REQUIRE(xyz && new_xyz && idx, "ball_query: null pointer");
  SAD_REQUIRE(B <= 65535, "ball_query: B=%d exceeds grid.y", B);
  const size_t smem = (size_t)BQ_STAGES * BQ_TILE * 3 * sizeof(float);
  const long long total_q = (long long)B * npoint;
  const int wpb = BQ_T / 32;
  int qpw = 8;
  if (total_q / (wpb * 8) < 296) qpw = 4;
  if (total_q / (wpb * 4) < 296) qpw = 2;
  if (total_q / (wpb * 2) < 296) qpw = 1;
#define SAD_BQ_LAUNCH(Q)                                                                            \
  {                                                                                                 \
    static thread_local int configured_dev = -1;                                                    \
    int dev = 0;                                                                                    \
    SAD_CUDA_OK(cudaGetDevice(&dev));                                                               \
    if (configured_dev != dev) {                                                                    \
      SAD_CUDA_OK(cudaFuncSetAttribute(ball_query_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)smem));                                                 \
      configured_dev = dev;                                                                         \
    }                                                                                               \
    dim3 grid((unsigned)sad_ceil_div(npoint, wpb * Q), (unsigned)B);                                \
    ball_query_kernel<Q><<<grid, BQ_T, smem, stream>>>(N, npoint, radius, radius_t, nsample, xyz,   \
                                                       new_xyz, idx);                               \
  }
  if (qpw == 8) SAD_BQ_LAUNCH(8)
  else if (qpw == 4) SAD_BQ_LAUNCH(4)
  else if (qpw == 2) SAD_BQ_LAUNCH(2)
  else SAD_BQ_LAUNCH(1)
#undef SAD_BQ_LAUNCH
  SAD_LAUNCH_CHECK("ball_query_kernel");
  return SAD_OK;
}

// ------------------------------------------------------------------ three_nn
constexpr int NN_T = 256;
constexpr int NN_TILE = 2048;

__device__ __forceinline__ void nn_insert(float d, int k, float& b0, float& b1, float& b2, int& i0, int& i1,
                                          int& i2) {
  // branch-free strict-'<' insertion into the sorted triple (selects only: no divergence)
  const bool c0 = d < b0, c1 = d < b1, c2 = d < b2;
  b2 = c1 ? b1 : (c2 ? d : b2);
  i2 = c1 ? i1 : (c2 ? k : i2);
  b1 = c0 ? b0 : (c1 ? d : b1);
  i1 = c0 ? i0 : (c1 ? k : i1);
  b0 = c0 ? d : b0;
  i0 = c0 ? k : i0;
}

// One warp per unknown point; lanes stride the known points (ascending per lane, strict
// '<' insertion), then three rounds of warp arg-min with ties to the lowest index.
__global__ void __launch_bounds__(NN_T)
three_nn_kernel(int n, int m, const float* __restrict__ unknown, const float* __restrict__ known,
                float* __restrict__ dist, int32_t* __restrict__ idx, float* __restrict__ weight) {
  __shared__ float s_known[NN_TILE * 3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int i = blockIdx.x * (NN_T / 32) + warp;
  const float* kn = known + (size_t)b * m * 3;
  const bool active = i < n;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (active) {
    const float* u = unknown + ((size_t)b * n + i) * 3;
    ux = __ldg(u);
    uy = __ldg(u + 1);
    uz = __ldg(u + 2);
  }
  const float INF = __int_as_float(0x7f800000);
  float b0 = INF, b1 = INF, b2 = INF;
  int i0 = 0x7FFFFFFF, i1 = 0x7FFFFFFF, i2 = 0x7FFFFFFF;
  for (int t0 = 0; t0 < m; t0 += NN_TILE) {
    const int npts = min(NN_TILE, m - t0);
    __syncthreads();
    for (int e = tid; e < npts * 3; e += NN_T) s_known[e] = __ldg(kn + (size_t)t0 * 3 + e);
    __syncthreads();
    if (active) {
      for (int c = lane; c < npts; c += 32) {
        const float d = sqdist(s_known[3 * c], s_known[3 * c + 1], s_known[3 * c + 2], ux, uy, uz);
        nn_insert(d, t0 + c, b0, b1, b2, i0, i1, i2);
      }
    }
  }
  if (!active) return;
  float od = 0.f;
  int oi = 0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const uint32_t v = __float_as_uint(b0);                 // d2 >= +0 (or +inf): uint order == float order
    const uint32_t mn = __reduce_min_sync(FULL, v);
    const uint32_t ck = __reduce_min_sync(FULL, v == mn ? (uint32_t)i0 : 0xFFFFFFFFu);
    if (v == mn && (uint32_t)i0 == ck) {                    // unique owner pops its head
      b0 = b1; i0 = i1; b1 = b2; i1 = i2; b2 = INF; i2 = 0x7FFFFFFF;
    }
    if (lane == r) {
      od = __fsqrt_rn(__uint_as_float(mn));
      oi = ck < (uint32_t)m ? (int)ck : 0;                  // (non-finite coordinates find no candidate: stay in range)
    }
  }
  if (lane < 3) {
    dist[((size_t)b * n + i) * 3 + lane] = od;
    idx[((size_t)b * n + i) * 3 + lane] = oi;
  }
  if (weight != nullptr) {
    // FP-module weights, the oracle's evaluation order: r_t = 1 / (dist_t + 1e-8), w_t = r_t / ((r_0 + r_1) + r_2)
    const float rc = __fdiv_rn(1.0f, __fadd_rn(od, 1e-8f));
    const float r0 = __shfl_sync(FULL, rc, 0), r1 = __shfl_sync(FULL, rc, 1), r2 = __shfl_sync(FULL, rc, 2);
    const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
    if (lane < 3) weight[((size_t)b * n + i) * 3 + lane] = __fdiv_rn(rc, norm);
  }
}

// a4 helper: predicted box size (rows,3) -> radius r = clamp(c * sqrt(((sx*sx)+(sy*sy))+(sz*sz)), r_min, r_max), c = alpha/2
__global__ void __launch_bounds__(256)
size_to_radius_kernel(long long rows, const float* __restrict__ size, float c, float r_min, float r_max, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows) return;
  const float sx = __ldg(size + 3 * i), sy = __ldg(size + 3 * i + 1), sz = __ldg(size + 3 * i + 2);
  const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy)), __fmul_rn(sz, sz));
  const float r = __fmul_rn(c, __fsqrt_rn(n2));
  out[i] = fminf(fmaxf(r, r_min), r_max);
}

}  // namespace

extern "C" int sad_ball_query_fwd(int B, int N, int npoint, float radius, int nsample, const float* xyz,
                                  const float* new_xyz, int32_t* idx, sad_stream_t stream) {
  return launch_ball_query(B, N, npoint, radius, nullptr, nsample, xyz, new_xyz, idx, (cudaStream_t)stream);
}

extern "C" int sad_ball_query_adaptive_fwd(int B, int N, int npoint, const float* radius_t, int nsample,
                                           const float* xyz, const float* new_xyz, int32_t* idx,
                                           sad_stream_t stream) {
  SAD_REQUIRE(radius_t != nullptr || B == 0 || npoint == 0, "ball_query_adaptive: null radius tensor");
  return launch_ball_query(B, N, npoint, 0.f, radius_t, nsample, xyz, new_xyz, idx, (cudaStream_t)stream);
}

extern "C" int sad_three_nn_fwd(int B, int n, int m, const float* unknown, const float* known, float* dist,
                                int32_t* idx, sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && n >= 0, "three_nn: bad sizes B=%d n=%d", B, n);
  SAD_REQUIRE(m >= 3, "three_nn: needs m >= 3 known points (m=%d)", m);
  if (B == 0 || n == 0) return SAD_OK;
  SAD_REQUIRE(unknown && known && dist && idx, "three_nn: null pointer");
  SAD_REQUIRE(B <= 65535, "three_nn: B=%d exceeds grid.y", B);
  dim3 grid((unsigned)sad_ceil_div(n, NN_T / 32), (unsigned)B);
  three_nn_kernel<<<grid, NN_T, 0, (cudaStream_t)stream>>>(n, m, unknown, known, dist, idx, nullptr);
  SAD_LAUNCH_CHECK("three_nn_kernel");
  return SAD_OK;
}

extern "C" int sad_three_nn_weights_fwd(int B, int n, int m, const float* unknown, const float* known, float* dist,
                                        int32_t* idx, float* weight, sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && n >= 0, "three_nn: bad sizes B=%d n=%d", B, n);
  SAD_REQUIRE(m >= 3, "three_nn: needs m >= 3 known points (m=%d)", m);
  if (B == 0 || n == 0) return SAD_OK;
  SAD_REQUIRE(unknown && known && dist && idx && weight, "three_nn: null pointer");
  SAD_REQUIRE(B <= 65535, "three_nn: B=%d exceeds grid.y", B);
  dim3 grid((unsigned)sad_ceil_div(n, NN_T / 32), (unsigned)B);
  three_nn_kernel<<<grid, NN_T, 0, (cudaStream_t)stream>>>(n, m, unknown, known, dist, idx, weight);
  SAD_LAUNCH_CHECK("three_nn_kernel");
  return SAD_OK;
}

extern "C" int sad_size_to_radius(long long rows, const float* size, float alpha, float r_min, float r_max, float* radius,
                                  sad_stream_t stream) {
  SAD_REQUIRE(rows >= 0, "size_to_radius: bad row count");
  if (rows == 0) return SAD_OK;
  SAD_REQUIRE(size && radius, "size_to_radius: null pointer");
  size_to_radius_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, size, alpha * 0.5f, r_min, r_max, radius);
  SAD_LAUNCH_CHECK("size_to_radius_kernel");
  return SAD_OK;
}
