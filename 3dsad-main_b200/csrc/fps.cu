// a1  furthest_point_sample -- SURVEY.md section 8(a) row a1, hard part H3.
// (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
// B200 design.  FPS is `npoint` strictly serial picks, each a full pass over the scene plus
// an argmax; the bound is the per-pick LATENCY chain, not HBM.  So:
//   * every point (x,y,z,min-dist) lives in REGISTERS for the whole kernel.  Warp g
//     (g = cta_rank*NW + warp) owns the contiguous index range [g*P*32, (g+1)*P*32),
//     thread `lane` the points k = (g*P + p)*32 + lane, p < P (compile-time, unrolled);
//     16 warps per CTA so the fp32 pipes stay busy while other warps sit in a latency chain;
//   * small scenes: one CTA per scene; large scenes: one thread-block CLUSTER (up to 16
//     CTAs = 16 SMs) per scene;
//   * the kernel works in ROUNDS.  Per round each warp reduces its points with redux.sync
//     to {best value, its coordinates, SECOND-best value}; warp 0 folds the CTA's warps and
//     pushes one 20-byte record into every peer CTA's shared memory with st.async (DSMEM) --
//     the store itself completes the peer's mbarrier: no cluster barrier, no global memory
//     on the critical path;
//   * MULTI-PICK: with one record per bin (bin = CTA, or warp in the single-CTA case) warp 0
//     replays the sequential algorithm on the records and accepts as many picks as it can
//     PROVE identical to it.  Rank the bins by (best value desc, bin asc); the rank-k
//     candidate is the true next pick iff, for every bin of lower rank, its value is
//     strictly above that bin's second-best (nothing left in a used bin can beat it) and
//     its distance to that bin's pick is >= its value (its min-dist is unchanged).  The
//     accepted prefix ends at the first failing rank; all of this is evaluated in parallel
//     (lane = bin x half of the partner bins) and one redux.min.  Accepted picks are applied
//     in the next local pass.  The exchange latency is paid once per round (typically 3-5
//     picks) instead of once per pick, and the result stays bit-identical to the oracle;
//   * ties -> lowest index everywhere (H2): warps and CTAs own ascending index ranges, so
//     "lowest bin" == "lowest index" and indices never travel.
// Slots past N get min-dist 0 and an index >= N: they can only tie at 0 and then lose.
#include <stdlib.h>

#include "sad_common.cuh"

namespace {

using namespace sad;

constexpr int FPS_T = 512;      // threads per CTA (16 warps)
constexpr int FPS_MAXB = 16;    // bins per round (<= 16 CTAs, or 16 warps)

__device__ __forceinline__ void st_async_v4(uint32_t raddr, float a, float b, float c, float d, uint32_t rbar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
          raddr),
      "f"(a), "f"(b), "f"(c), "f"(d), "r"(rbar)
      : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr),
               "r"(v), "r"(rbar)
               : "memory");
}

template <int P, int CS>
__global__ void __launch_bounds__(FPS_T, (P <= 6) ? 2 : 1)
fps_kernel(const float* __restrict__ xyz, int N, int npoint, int32_t* __restrict__ out, long long* dbg) {
  constexpr int T = FPS_T;
  constexpr int NW = T / 32;
  constexpr int NB = (CS > 1) ? CS : NW;                   // bins seen by the selection step
  static_assert(NW == 16 && CS <= 16 && NB <= FPS_MAXB, "selection maps lane = bin x half");
  extern __shared__ __align__(16) float4 s_pts[];          // [P*T] (x,y,z,-) copy for winner lookup
  __shared__ __align__(16) float4 s_loc[2][NW];            // per warp {best value bits, x, y, z}
  __shared__ __align__(8) uint2 s_loc2[2][NW];             // per warp {second-best value bits, best index}
  __shared__ __align__(16) float4 s_rec[2][FPS_MAXB];      // per CTA  {best value bits, x, y, z}   (DSMEM target)
  __shared__ uint32_t s_rec2[2][FPS_MAXB];                 // per CTA  second-best value bits       (DSMEM target)
  __shared__ __align__(16) float4 s_q[FPS_MAXB];           // picks accepted in the current round
  __shared__ int s_npick;
  __shared__ __align__(8) uint64_t s_bar[2];

  const long long t_entry = clock64();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = (CS > 1) ? cluster_ctarank() : 0u;
  const int g = (int)rank * NW + warp;
  const int b = blockIdx.x / CS;
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
  if (tid == 0) {
    s_q[0] = make_float4(__ldg(pts), __ldg(pts + 1), __ldg(pts + 2), 0.f);   // pick 0 = point 0
    s_npick = 1;
    if (rank == 0) o[0] = 0;
  }

  uint32_t r_rec[2] = {0, 0}, r_rec2[2] = {0, 0}, r_bar[2] = {0, 0};   // DSMEM addresses in peer CTA `lane`
  if (CS > 1) {
    if (tid == 0) {
      mbar_init(&s_bar[0], 1);
      mbar_init(&s_bar[1], 1);
      mbar_fence_init();
    }
    if (warp == 0) {
      const uint32_t dst = (uint32_t)(lane % CS);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        r_rec[u] = mapa(smem_u32(&s_rec[u][rank]), dst);
        r_rec2[u] = mapa(smem_u32(&s_rec2[u][rank]), dst);
        r_bar[u] = mapa(smem_u32(&s_bar[u]), dst);
      }
    }
  }
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // peers resident + mbarrier inits visible before any DSMEM store

  const bool prof = (dbg != nullptr) && blockIdx.x == 0 && tid == 0;   // optional phase timing (tools only)
  long long ph[6] = {0, 0, 0, 0, 0, 0}, tprev = 0;
#define SAD_FPS_MARK(i)                \
  if (prof) {                          \
    const long long tn = clock64();    \
    ph[i] += tn - tprev;               \
    tprev = tn;                        \
  }

  int j = 1;      // next output position
  for (uint32_t round = 0; j < npoint; ++round) {
    const int buf = (int)(round & 1u);
    if (prof) tprev = clock64();
    if (CS > 1 && tid == 0) mbar_arrive_expect_tx(&s_bar[buf], CS * 20);

    // ---- 1. apply the picks of the previous round to the register-resident points
    const int K = s_npick;
    for (int k = 0; k < K; ++k) {
      const float4 q = s_q[k];
#pragma unroll
      for (int p = 0; p < P; ++p) md[p] = fminf(md[p], sqdist(px[p], py[p], pz[p], q.x, q.y, q.z));
    }
    // best (first max: ascending p == ascending index inside a thread) and second-best value
    float b1 = md[0], b2 = 0.f;
    int bp = 0;
#pragma unroll
    for (int p = 1; p < P; ++p) {
      const float m = md[p];
      if (m > b1) {
        b2 = b1;
        b1 = m;
        bp = p;
      } else {
        b2 = fmaxf(b2, m);
      }
    }
    SAD_FPS_MARK(0)
    // ---- 2. warp: best value, lowest index among the maxima, second-best value
    const uint32_t vb = __float_as_uint(b1);
    const uint32_t bk = (uint32_t)((g * P + bp) * 32 + lane);
    const uint32_t wmax = __reduce_max_sync(FULL, vb);
    const uint32_t wk = __reduce_min_sync(FULL, vb == wmax ? bk : 0xFFFFFFFFu);
    const uint32_t w2 = __reduce_max_sync(FULL, (vb == wmax && bk == wk) ? __float_as_uint(b2) : vb);
    if (lane == 0) {
      const int lp = (int)(wk >> 5) - g * P;
      const float4 c = s_pts[lp * T + warp * 32 + (int)(wk & 31u)];
      s_loc[buf][warp] = make_float4(__uint_as_float(wmax), c.x, c.y, c.z);
      s_loc2[buf][warp] = make_uint2(w2, wk);
    }
    SAD_FPS_MARK(1)
    __syncthreads();      // (a) warp records visible; every warp is done reading s_q / s_npick
    SAD_FPS_MARK(2)

    if (warp == 0) {
      // ---- 3. one record per bin: fold the CTA and exchange over DSMEM (cluster), or use the warps
      const float4* RA;
      const uint32_t* RB;
      int rep_warp = 0;
      if (CS > 1) {
        const uint32_t lv = (lane < NW) ? __float_as_uint(s_loc[buf][lane].x) : 0u;
        const uint32_t l2 = (lane < NW) ? s_loc2[buf][lane].x : 0u;
        const uint32_t cmax = __reduce_max_sync(FULL, lv);
        rep_warp = __ffs(__ballot_sync(FULL, lane < NW && lv == cmax)) - 1;    // lowest warp == lowest index
        const uint32_t c2 = __reduce_max_sync(FULL, (lane == rep_warp) ? l2 : lv);
        const float4 r = s_loc[buf][rep_warp];
        if (lane < CS) {
          st_async_v4(r_rec[buf], r.x, r.y, r.z, r.w, r_bar[buf]);
          st_async_b32(r_rec2[buf], c2, r_bar[buf]);
        }
        mbar_wait(&s_bar[buf], (round >> 1) & 1u);
        RA = s_rec[buf];
        RB = s_rec2[buf];
      } else {
        RA = s_loc[buf];
        RB = nullptr;
      }
      SAD_FPS_MARK(3)
      // ---- 4. accepted prefix, in parallel: lane = (bin i, half h of the partner bins j)
      const int i = lane & 15, h = lane >> 4;
      const bool vi = i < NB;
      const float4 ri = vi ? RA[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      const uint32_t v1i = __float_as_uint(ri.x);
      const float fi = ri.x;
      uint32_t cnt = 0u, earlier = 0u, conflict = 0u;
      // branch-free on purpose (bitwise, not short-circuit): divergent branches here cost more
      // than the whole arithmetic of the step
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int jb = min(h * 8 + jj, NB - 1);                 // clamp: out-of-range partners are masked below
        const bool valid = (h * 8 + jj < NB) & (jb != i);
        const float4 rj = RA[jb];
        const uint32_t v1j = __float_as_uint(rj.x);
        const uint32_t v2j = (CS > 1) ? RB[jb] : s_loc2[buf][jb].x;
        const float dij = sqdist(ri.y, ri.z, ri.w, rj.y, rj.z, rj.w);
        const uint32_t before = (uint32_t)((v1j > v1i) | ((v1j == v1i) & (jb < i))) & (uint32_t)valid;   // jb outranks i
        const uint32_t conf = (uint32_t)((v2j >= v1i) | (dij < fi)) & (uint32_t)valid;
        cnt += before;
        earlier |= before << jb;
        conflict |= conf << jb;
      }
      cnt += __shfl_xor_sync(FULL, cnt, 16);
      earlier |= __shfl_xor_sync(FULL, earlier, 16);
      conflict |= __shfl_xor_sync(FULL, conflict, 16);
      const bool fail = vi && ((conflict & earlier) != 0u);
      int A = (int)__reduce_min_sync(FULL, fail ? cnt : (uint32_t)NB);          // rank 0 never fails: A >= 1
      A = min(A, npoint - j);
      if (h == 0 && vi && (int)cnt < A) {
        s_q[cnt] = make_float4(ri.y, ri.z, ri.w, 0.f);
        if (CS > 1) {
          if (i == (int)rank) o[j + (int)cnt] = (int32_t)s_loc2[buf][rep_warp].y;
        } else {
          o[j + (int)cnt] = (int32_t)s_loc2[buf][i].y;
        }
      }
      if (lane == 0) s_npick = A;
    }
    __syncthreads();      // (b) s_q / s_npick of this round visible to every warp
    j += s_npick;
    SAD_FPS_MARK(4)
    if (prof) ph[5] += 1;
  }
#undef SAD_FPS_MARK
  if (prof) {
    for (int i = 0; i < 6; ++i) dbg[i] = ph[i];
    dbg[6] = clock64() - t_entry;
  }
  if (CS > 1) cluster_sync_all();   // no CTA retires while a peer's st.async may still target it
}

thread_local long long* g_fps_dbg = nullptr;   // tools only: per-phase cycle counters of block 0

template <int P, int CS>
int launch_fps(int B, int N, int npoint, const float* xyz, int32_t* idx, cudaStream_t stream) {
  auto kern = fps_kernel<P, CS>;
  const size_t smem = (size_t)P * FPS_T * sizeof(float4);
  static thread_local int configured_dev = -1;   // per (P,CS) instantiation and thread
  int dev = 0;
  SAD_CUDA_OK(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    SAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) SAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    // several clusters must share an SM for a whole batch to run in one wave
    SAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
    configured_dev = dev;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CS));
  cfg.blockDim = dim3(FPS_T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (CS > 1) ? 1 : 0;
  if (getenv("SAD_DEBUG_OCC")) {
    int ncl = -1, nb = -1;
    if (CS > 1) cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, FPS_T, smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    fprintf(stderr, "[sad] fps_kernel<P=%d,CS=%d>: regs=%d smem=%zu+%zu maxActiveClusters=%d blocks/SM=%d\n", P, CS,
            fa.numRegs, smem, fa.sharedSizeBytes, ncl, nb);
  }
  SAD_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, xyz, N, npoint, idx, g_fps_dbg));
  sad_count_launch(1);
  return SAD_OK;
}

constexpr int kPs[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 25};
constexpr int kPMax = 25;

int round_p(int p) {
  for (int a : kPs)
    if (p <= a) return a;
  return -1;
}

template <int CS>
int dispatch_p(int P, int B, int N, int npoint, const float* xyz, int32_t* idx, cudaStream_t s) {
  switch (P) {
#define SAD_FPS_CASE(PP) \
  case PP:               \
    return launch_fps<PP, CS>(B, N, npoint, xyz, idx, s);
    SAD_FPS_CASE(1)
    SAD_FPS_CASE(2)
    SAD_FPS_CASE(3)
    SAD_FPS_CASE(4)
    SAD_FPS_CASE(5)
    SAD_FPS_CASE(6)
    SAD_FPS_CASE(8)
    SAD_FPS_CASE(10)
    SAD_FPS_CASE(12)
    SAD_FPS_CASE(16)
    SAD_FPS_CASE(20)
    SAD_FPS_CASE(25)
#undef SAD_FPS_CASE
  }
  sad_set_error("fps: no kernel for P=%d", P);
  return SAD_EUNSUPPORTED;
}

}  // namespace

// Exposed for tests/benchmarks: force a cluster size (0 = heuristic).
static thread_local int g_force_cs = 0;
extern "C" void sad_fps_force_cluster_size(int cs) { g_force_cs = cs; }
extern "C" void sad_fps_set_debug_buffer(long long* dev6) { g_fps_dbg = dev6; }

extern "C" int sad_furthest_point_sample_fwd(int B, int N, int npoint, const float* xyz, int32_t* idx,
                                             sad_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SAD_REQUIRE(B >= 0 && N >= 1 && npoint >= 1, "furthest_point_sample: bad sizes B=%d N=%d npoint=%d", B, N,
              npoint);
  if (B == 0) return SAD_OK;
  SAD_REQUIRE(xyz && idx, "furthest_point_sample: null pointer");
  constexpr int T = FPS_T;
  if ((long long)N > 16LL * T * kPMax) {
    sad_set_error("furthest_point_sample: N=%d exceeds the register-resident capacity (%d)", N, 16 * T * kPMax);
    return SAD_EUNSUPPORTED;
  }
  int cs = g_force_cs;
  if (cs == 0) {
    if (N <= 3072) {
      cs = 1;
    } else {
      // largest cluster that still lets every scene of the batch run in one wave
      cs = 16;
      while (cs > 2 && (long long)B * cs > 148) cs >>= 1;
    }
  }
  SAD_REQUIRE(cs == 1 || cs == 2 || cs == 4 || cs == 8 || cs == 12 || cs == 14 || cs == 16, "fps: bad cluster size %d",
              cs);
  while (cs < 16 && (long long)cs * T * kPMax < N) cs = (cs < 8) ? cs * 2 : (cs == 8 ? 12 : cs + 2);   // P <= 25
  const int P = round_p(sad_ceil_div(N, (long long)cs * T));
  switch (cs) {
    case 1: return dispatch_p<1>(P, B, N, npoint, xyz, idx, stream);
    case 2: return dispatch_p<2>(P, B, N, npoint, xyz, idx, stream);
    case 4: return dispatch_p<4>(P, B, N, npoint, xyz, idx, stream);
    case 8: return dispatch_p<8>(P, B, N, npoint, xyz, idx, stream);
    case 12: return dispatch_p<12>(P, B, N, npoint, xyz, idx, stream);
    case 14: return dispatch_p<14>(P, B, N, npoint, xyz, idx, stream);
    default: return dispatch_p<16>(P, B, N, npoint, xyz, idx, stream);
  }
}
