// a1  furthest_point_sample -- SURVEY.md section 8(a) row a1, hard part H3.
// (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
// B200 design.  FPS is `npoint` strictly serial iterations, each a full pass over the
// scene plus an argmax; the bound is per-iteration LATENCY, not HBM.  So:
//   * every point (x,y,z,min-dist) lives in REGISTERS for the whole kernel.  Warp g
//     (g = cta_rank*NW + warp) owns the contiguous index range [g*P*32, (g+1)*P*32),
//     thread `lane` the points k = (g*P + p)*32 + lane, p < P (compile-time, unrolled);
//   * small scenes: one CTA per scene; large scenes: one thread-block CLUSTER (up to 16
//     CTAs = 16 SMs) per scene;
//   * per iteration each warp reduces its own points with redux.sync (max on the
//     non-negative distance bits, then min on the index among the maxima: ties -> lowest
//     index, H2) and pushes ONE 16-byte record {dist, x, y, z} straight into slot g of
//     every peer CTA's shared memory with st.async (DSMEM) -- the store itself signals
//     the peer's mbarrier (complete_tx), so there is no cluster barrier, no CTA barrier
//     and no global memory on the critical path;
//   * every warp then reduces the CS*NW records.  Because warps own ascending index
//     ranges, "lowest slot among equal distances" == "lowest index", so the index never
//     travels: the winning warp alone writes it to the output.
// Slots past N get min-dist 0 and an index >= N: they can only ever tie at 0 and then
// lose to a real point on the index rule.
#include "sad_common.cuh"

namespace {

using namespace sad;

__device__ __forceinline__ void st_async_v4(uint32_t raddr, float a, float b, float c, float d, uint32_t rbar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
          raddr),
      "f"(a), "f"(b), "f"(c), "f"(d), "r"(rbar)
      : "memory");
}

template <int T, int P, int CS>
__global__ void __launch_bounds__(T, 1)
fps_kernel(const float* __restrict__ xyz, int N, int npoint, int32_t* __restrict__ out, const int* __restrict__ skip) {
  constexpr int NW = T / 32;
  constexpr int NSLOT = CS * NW;
  constexpr int RPL = (NSLOT + 31) / 32;                 // records per lane in the final reduce
  extern __shared__ __align__(16) float4 s_pts[];         // [P*T] (x,y,z,-) copy for winner lookup
  __shared__ __align__(16) float4 s_rec[2][NSLOT];        // {dist bits, x, y, z} per warp of the cluster
  __shared__ __align__(8) uint64_t s_bar[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = (CS > 1) ? cluster_ctarank() : 0u;
  const int g = (int)rank * NW + warp;                    // slot id == ascending index range id
  const int b = blockIdx.x / CS;
  // prefix-ordered input whose guard passed (sad_furthest_point_sample_prefix_fwd): the answer is already written.
  // Uniform over the scene's cluster, and before its first cluster barrier.
  if (skip != nullptr && skip[b] != 0) return;
  const float* pts = xyz + (size_t)b * N * 3;
  int32_t* o = out + (size_t)b * npoint;

  float px[P], py[P], pz[P], md[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int k = (g * P + p) * 32 + lane;
    if (k < N) {
      px[p] = __ldg(pts + 3 * (size_t)k);
      py[p] = __ldg(pts + 3 * (size_t)k + 1);
      pz[p] = __ldg(pts + 3 * (size_t)k + 2);
      md[p] = 1e10f;
    } else {
      px[p] = py[p] = pz[p] = 0.f;
      md[p] = 0.f;
    }
    s_pts[p * T + tid] = make_float4(px[p], py[p], pz[p], 0.f);
  }
  float qx = __ldg(pts), qy = __ldg(pts + 1), qz = __ldg(pts + 2);
  if (rank == 0 && tid == 0) o[0] = 0;

  uint32_t r_rec[2] = {0, 0}, r_bar[2] = {0, 0};          // DSMEM addresses in peer CTA `lane`
  if (CS > 1) {
    if (tid == 0) {
      mbar_init(&s_bar[0], 1);
      mbar_init(&s_bar[1], 1);
      mbar_fence_init();
    }
    const uint32_t dst = (uint32_t)(lane % CS);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      r_rec[u] = mapa(smem_u32(&s_rec[u][g]), dst);
      r_bar[u] = mapa(smem_u32(&s_bar[u]), dst);
    }
  }
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // peers resident + mbarrier inits visible before any DSMEM store

  for (int j = 1; j < npoint; ++j) {
    const int buf = j & 1;
    if (CS > 1 && tid == 0) mbar_arrive_expect_tx(&s_bar[buf], NSLOT * 16);

    // ---- local pass over the P register-resident points (two independent best-chains)
    float bv0 = -1.f, bv1 = -1.f;
    int bp0 = 0, bp1 = 1;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float d = sqdist(px[p], py[p], pz[p], qx, qy, qz);
      const float m = fminf(md[p], d);
      md[p] = m;
      if (p & 1) {
        if (m > bv1) { bv1 = m; bp1 = p; }   // strict: ascending p == ascending index in a thread
      } else {
        if (m > bv0) { bv0 = m; bp0 = p; }
      }
    }
    // (selects, not branches: the warp must arrive converged at the redux below, or it takes the slow collective path)
    const bool second = (P > 1) & ((bv1 > bv0) | ((bv1 == bv0) & (bp1 < bp0)));
    const float bv = second ? bv1 : bv0;
    const int bp = second ? bp1 : bp0;

    const uint32_t vb = __float_as_uint(bv);
    const uint32_t bk = (uint32_t)((g * P + bp) * 32 + lane);
    const uint32_t wmax = __reduce_max_sync(FULL, vb);
    const uint32_t wk = __reduce_min_sync(FULL, vb == wmax ? bk : 0xFFFFFFFFu);
    const int lp = (int)(wk >> 5) - g * P;
    const float4 c = s_pts[lp * T + warp * 32 + (int)(wk & 31u)];

    if (CS == 1) {
      if (lane == 0) s_rec[buf][warp] = make_float4(__uint_as_float(wmax), c.x, c.y, c.z);
      __syncthreads();
    } else {
      if (lane < CS) st_async_v4(r_rec[buf], __uint_as_float(wmax), c.x, c.y, c.z, r_bar[buf]);
      // use u of s_bar[buf] is iteration j = 2u + 1 (buf 1) or 2u + 2 (buf 0): u = (j-1)/2
      mbar_wait(&s_bar[buf], (uint32_t)(((j - 1) >> 1) & 1));
    }

    // ---- every warp reduces the NSLOT records: max dist, ties -> lowest slot (== lowest index)
    uint32_t v = 0u, slot = 0xFFFFFFFFu;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
      const int s = lane + 32 * r;
      const bool in = s < NSLOT;
      const uint32_t x = __float_as_uint(s_rec[buf][in ? s : 0].x);
      const bool take = in & ((slot == 0xFFFFFFFFu) | (x > v));
      v = take ? x : v;
      slot = take ? (uint32_t)s : slot;
    }
    const uint32_t gmax = __reduce_max_sync(FULL, v);
    const uint32_t gslot = __reduce_min_sync(FULL, v == gmax ? slot : 0xFFFFFFFFu);
    const float4 w = s_rec[buf][gslot];
    qx = w.y;
    qy = w.z;
    qz = w.w;
    if ((int)gslot == g && lane == 0) o[j] = (int32_t)wk;
  }
  if (CS > 1) cluster_sync_all();   // no CTA retires while a peer's st.async may still target it
}

template <int T, int P, int CS>
int launch_fps(int B, int N, int npoint, const float* xyz, int32_t* idx, const int* skip, cudaStream_t stream) {
  auto kern = fps_kernel<T, P, CS>;
  const size_t smem = (size_t)P * T * sizeof(float4);
  static thread_local int configured_dev = -1;   // per (T,P,CS) instantiation and thread
  int dev = 0;
  SAD_CUDA_OK(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    SAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) SAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured_dev = dev;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CS));
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (CS > 1) ? 1 : 0;
  SAD_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, xyz, N, npoint, idx, skip));
  sad_count_launch(1);
  return SAD_OK;
}

constexpr int kPs[] = {1, 2, 3, 4, 6, 8, 10, 12, 16, 20, 24, 32, 40, 48};

int round_p(int p) {
  for (int a : kPs)
    if (p <= a) return a;
  return -1;
}

template <int T, int CS>
int dispatch_p(int P, int B, int N, int npoint, const float* xyz, int32_t* idx, const int* skip, cudaStream_t s) {
  switch (P) {
#define SAD_FPS_CASE(PP) \
  case PP:               \
    return launch_fps<T, PP, CS>(B, N, npoint, xyz, idx, skip, s);
    SAD_FPS_CASE(1)
    SAD_FPS_CASE(2)
    SAD_FPS_CASE(3)
    SAD_FPS_CASE(4)
    SAD_FPS_CASE(6)
    SAD_FPS_CASE(8)
    SAD_FPS_CASE(10)
    SAD_FPS_CASE(12)
    SAD_FPS_CASE(16)
    SAD_FPS_CASE(20)
    SAD_FPS_CASE(24)
    SAD_FPS_CASE(32)
    SAD_FPS_CASE(40)
    SAD_FPS_CASE(48)
#undef SAD_FPS_CASE
  }
  sad_set_error("fps: no kernel for P=%d", P);
  return SAD_EUNSUPPORTED;
}

// Large scenes only: wider CTAs at the maximum cluster size.
int dispatch_big(int T, int P, int B, int N, int npoint, const float* xyz, int32_t* idx, const int* skip, cudaStream_t s) {
  if (T == 256) {
    switch (P) {
      case 24: return launch_fps<256, 24, 16>(B, N, npoint, xyz, idx, skip, s);
      case 32: return launch_fps<256, 32, 16>(B, N, npoint, xyz, idx, skip, s);
      case 40: return launch_fps<256, 40, 16>(B, N, npoint, xyz, idx, skip, s);
      case 48: return launch_fps<256, 48, 16>(B, N, npoint, xyz, idx, skip, s);
    }
  } else if (T == 512 && P == 25) {
    return launch_fps<512, 25, 16>(B, N, npoint, xyz, idx, skip, s);
  }
  sad_set_error("fps: no large-scene kernel for T=%d P=%d", T, P);
  return SAD_EUNSUPPORTED;
}

}  // namespace

#ifdef SAD_TOOLS_ABLATE
__global__ void ablate_strided_idx_plain(int N, int npoint, int32_t* idx) {
  for (int j = threadIdx.x; j < npoint; j += blockDim.x) idx[(size_t)blockIdx.x * npoint + j] = (int32_t)((long long)j * N / npoint);
}
#endif
static int fps_plain(int B, int N, int npoint, const float* xyz, int32_t* idx, const int* skip, int force_cs, cudaStream_t stream) {
  SAD_REQUIRE(B >= 0 && N >= 1 && npoint >= 1, "furthest_point_sample: bad sizes B=%d N=%d npoint=%d", B, N,
              npoint);
  if (B == 0) return SAD_OK;
  SAD_REQUIRE(xyz && idx, "furthest_point_sample: null pointer");
#ifdef SAD_TOOLS_ABLATE
  if (sad_ablate_mask() & 4) {
    ablate_strided_idx_plain<<<B, 256, 0, stream>>>(N, npoint, idx);
    return SAD_OK;
  }
#endif
  constexpr int T = 128, PMAX = 48;
  if ((long long)N > 16LL * 512 * 25) {
    sad_set_error("furthest_point_sample: N=%d exceeds the register-resident capacity (%d)", N, 16 * 512 * 25);
    return SAD_EUNSUPPORTED;
  }
  if ((long long)N > 16LL * T * PMAX) {   // > 98304 points: wider CTAs, cluster of 16
    if ((long long)N <= 16LL * 256 * PMAX) {
      return dispatch_big(256, round_p(sad_ceil_div(N, 16 * 256)) < 24 ? 24 : round_p(sad_ceil_div(N, 16 * 256)), B,
                          N, npoint, xyz, idx, skip, stream);
    }
    return dispatch_big(512, 25, B, N, npoint, xyz, idx, skip, stream);
  }
  int cs = force_cs;
  if (cs == 0) {
    if (N <= 3072) {
      cs = 1;
    } else {
      // largest cluster that still lets every scene of the batch run in one wave
      cs = 16;
      while (cs > 2 && (long long)B * cs > 148) cs >>= 1;
    }
  }
  SAD_REQUIRE(cs == 1 || cs == 2 || cs == 4 || cs == 8 || cs == 16, "fps: bad cluster size %d", cs);
  while (cs < 16 && (long long)cs * T * PMAX < N) cs <<= 1;   // capacity: P <= 48 points / thread
  const int P = round_p(sad_ceil_div(N, (long long)cs * T));
  switch (cs) {
    case 1: return dispatch_p<T, 1>(P, B, N, npoint, xyz, idx, skip, stream);
    case 2: return dispatch_p<T, 2>(P, B, N, npoint, xyz, idx, skip, stream);
    case 4: return dispatch_p<T, 4>(P, B, N, npoint, xyz, idx, skip, stream);
    case 8: return dispatch_p<T, 8>(P, B, N, npoint, xyz, idx, skip, stream);
    default: return dispatch_p<T, 16>(P, B, N, npoint, xyz, idx, skip, stream);
  }
}

extern "C" int sad_furthest_point_sample_fwd(int B, int N, int npoint, const float* xyz, int32_t* idx,
                                             sad_stream_t stream) {
  return fps_plain(B, N, npoint, xyz, idx, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int sad_furthest_point_sample_cs_fwd(int B, int N, int npoint, const float* xyz, int32_t* idx, int cluster_size,
                                                sad_stream_t stream) {
  SAD_REQUIRE(cluster_size == 0 || cluster_size == 1 || cluster_size == 2 || cluster_size == 4 || cluster_size == 8 ||
              cluster_size == 16, "furthest_point_sample: cluster_size must be 0 (heuristic), 1, 2, 4, 8 or 16");
  return fps_plain(B, N, npoint, xyz, idx, nullptr, cluster_size, (cudaStream_t)stream);
}

// ---- a1 over PREFIX-ORDERED input (SURVEY.md H3 side note) ---------------------------------------
// If row k of `xyz` is the k-th pick of a farthest-point sampling of some superset (the previous SA stage's new_xyz),
// then sampling its first rows again returns the identity: by induction the min-distances of the rows are the
// superset's, the k-th pick of the superset is row k and every smaller row has min-distance 0 -- UNLESS a pick itself
// had min-distance 0 (it duplicates an earlier pick: the superset ran out of distinct points), where the lowest-index
// rule picks differently in the two orders.  The guard below checks exactly that, with the contract's arithmetic
// (d2(row k, row j) == 0 for some j < k < npoint), per scene; scenes that pass get arange, the others run the sampler.
namespace {
__global__ void __launch_bounds__(1024) fps_prefix_guard_kernel(const float* __restrict__ xyz, int N, int npoint,
                                                                int32_t* __restrict__ out, int* __restrict__ flags) {
  extern __shared__ __align__(16) float4 s_p[];      // the first npoint rows
  __shared__ int s_bad;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* pts = xyz + (size_t)b * N * 3;
  if (tid == 0) s_bad = 0;
  for (int k = tid; k < npoint; k += blockDim.x) s_p[k] = make_float4(__ldg(pts + 3 * k), __ldg(pts + 3 * k + 1), __ldg(pts + 3 * k + 2), 0.f);
  __syncthreads();
  // row k against every earlier row; a thread takes rows r and npoint - 1 - r (equal work per thread)
  int bad = 0;
  for (int r = tid; 2 * r < npoint; r += blockDim.x) {
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const int k = side == 0 ? r : npoint - 1 - r;
      if (side == 1 && k == r) continue;
      const float4 q = s_p[k];
      for (int j = 0; j < k; ++j) {
        const float4 a = s_p[j];
        bad |= (sqdist(q.x, q.y, q.z, a.x, a.y, a.z) == 0.f) ? 1 : 0;
      }
    }
  }
  if (bad) s_bad = 1;
  __syncthreads();
  const int ok = s_bad == 0;
  if (tid == 0) flags[b] = ok;
  if (ok)
    for (int k = tid; k < npoint; k += blockDim.x) out[(size_t)b * npoint + k] = k;
}
}  // namespace

extern "C" int sad_furthest_point_sample_prefix_fwd(int B, int N, int npoint, const float* xyz, int32_t* idx, int* flags,
                                                    sad_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SAD_REQUIRE(B >= 0 && N >= 1 && npoint >= 1 && npoint <= N, "furthest_point_sample_prefix: bad sizes B=%d N=%d npoint=%d", B, N, npoint);
  if (B == 0) return SAD_OK;
  SAD_REQUIRE(xyz && idx && flags, "furthest_point_sample_prefix: null pointer");
  if (npoint > 8192) return fps_plain(B, N, npoint, xyz, idx, nullptr, 0, stream);      // guard table would not fit: plain sampler
  const size_t smem = (size_t)npoint * sizeof(float4);
  static thread_local int configured_dev = -1;
  int dev = 0;
  SAD_CUDA_OK(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    SAD_CUDA_OK(cudaFuncSetAttribute(fps_prefix_guard_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16));
    configured_dev = dev;
  }
  fps_prefix_guard_kernel<<<B, 1024, smem, stream>>>(xyz, N, npoint, idx, flags);
  SAD_LAUNCH_CHECK("fps_prefix_guard_kernel");
  return fps_plain(B, N, npoint, xyz, idx, flags, 0, stream);
}
