// a1  furthest_point_sample, exact CULLED variant over the scene grid -- SURVEY.md section 8(a) row
// a1, hard part H3, section 8(f) rank 2.  (No reference file exists to cite: README.md:1-2 only.)
//
// FPS is npoint strictly serial picks.  The plain kernel (fps.cu) re-tests every point against every
// pick: N distance evaluations per pick on the critical path, spread over 16 SMs to make that short.
// But a pick q only changes min-dist[p] when d2(p,q) < min-dist[p], i.e. in a small neighbourhood of q
// once a few dozen picks cover the scene.  So, with the points spatially sorted (grid.cu):
//   * a BUCKET = 32 consecutive sorted points (one per lane), with its bounding box and its current
//     maximum min-dist `bmax` (+ the lowest original index attaining it);
//   * buckets are dealt round-robin to the CTAs of the cluster and the 16 warps of each CTA, so the
//     few buckets a pick touches (spatial neighbours = consecutive buckets) spread over all warps;
//     lane s of a warp holds the box / bmax of the warp's s-th bucket in registers;
//   * per pick every warp tests its <= 32 boxes in ONE lane-parallel step:
//         d2(clamp(q, box), q) >= bmax   =>  no point of the bucket changes          (exact: every
//     fp32 operation of the contract distance is monotone, so d2(p,q) >= d2(clamp(q,box),q) holds
//     for the ROUNDED values too), and only the surviving buckets are re-evaluated from shared
//     memory (x,y,z,index as float4 + min-dist, 20 B/point, resident for the whole kernel);
//   * each warp publishes {best value, x, y, z, original index}; records are pushed straight into
//     every CTA of the cluster with st.async (DSMEM, the store completes the receiver's mbarrier), every
//     warp reduces all records itself: no cluster barrier, no CTA barrier in the cluster case;
//   * ties -> lowest ORIGINAL index at every level (value compared as bits, then index), so the
//     result is bit-identical to the oracle regardless of the sort order inside a cell.
// 40k points fit a cluster of 4 SMs (the plain kernel uses 16) and a pick costs a handful of bucket
// updates instead of N distance tests.
#include "sad_common.cuh"
#include "sad_grid.cuh"

namespace {

using namespace sad;

constexpr int FC_T = 512;
constexpr int FC_NW = FC_T / 32;
constexpr int FC_MAX_SLOTS = 21;          // buckets per warp (<= 32 lanes; 21 * 512 pts * 20 B = 215 KB)
constexpr uint32_t kInf = 0xFFFFFFFFu;

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
// order-preserving float <-> uint maps (for redux.min/max on signed floats)
__device__ __forceinline__ uint32_t f2o(float f) {
  const uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float o2f(uint32_t o) {
  return __uint_as_float(o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

template <int CS>
__global__ void __launch_bounds__(FC_T, 1)
fps_cull_kernel(int N, int npoint, const float* __restrict__ xyz, const uint8_t* __restrict__ ws, size_t stride,
                int32_t* __restrict__ out, int slots) {
  constexpr int NW = FC_NW;
  constexpr int NSLOT = CS * NW;                  // records per round
  constexpr int RPL = (NSLOT + 31) / 32;
  extern __shared__ __align__(16) uint8_t s_dyn[];
  float4* s_pts = reinterpret_cast<float4*>(s_dyn);                              // [slots*NW*32] x,y,z,bits(idx)
  float* s_md = reinterpret_cast<float*>(s_dyn + (size_t)slots * NW * 32 * 16);  // [slots*NW*32] min-dist
  __shared__ __align__(16) float4 s_rec[2][NSLOT];   // {value bits, x, y, z} per warp of the cluster
  __shared__ uint32_t s_ridx[2][NSLOT];              // original index of the record's point
  __shared__ __align__(8) uint64_t s_bar[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = (CS > 1) ? cluster_ctarank() : 0u;
  const int g = (int)rank * NW + warp;            // record slot of this warp
  const int b = blockIdx.x / CS;
  const float4* sorted = reinterpret_cast<const float4*>(ws + (size_t)b * stride + kGridHeaderBytes + kGridCellBytes);
  int32_t* o = out + (size_t)b * npoint;
  const int NBK = (N + 31) >> 5;

  // ---- load this warp's buckets; lane s keeps the box / bmax / bidx of bucket slot s
  float blx = 0.f, bly = 0.f, blz = 0.f, bhx = 0.f, bhy = 0.f, bhz = 0.f, bmax = 0.f;
  uint32_t bidx = kInf;
  int nb = 0;
  for (int s = 0; s < slots; ++s) {
    const int bkt = (s * NW + warp) * CS + (int)rank;
    const int k = bkt * 32 + lane;
    const bool ok = (bkt < NBK) && (k < N);
    float4 p = make_float4(0.f, 0.f, 0.f, __uint_as_float(kInf));
    if (ok) p = __ldg(sorted + k);
    const int i = (s * NW + warp) * 32 + lane;
    s_pts[i] = p;
    s_md[i] = ok ? 1e10f : 0.f;
    const uint32_t lx = __reduce_min_sync(FULL, ok ? f2o(p.x) : kInf), hx = __reduce_max_sync(FULL, ok ? f2o(p.x) : 0u);
    const uint32_t ly = __reduce_min_sync(FULL, ok ? f2o(p.y) : kInf), hy = __reduce_max_sync(FULL, ok ? f2o(p.y) : 0u);
    const uint32_t lz = __reduce_min_sync(FULL, ok ? f2o(p.z) : kInf), hz = __reduce_max_sync(FULL, ok ? f2o(p.z) : 0u);
    const uint32_t mi = __reduce_min_sync(FULL, ok ? __float_as_uint(p.w) : kInf);
    if (bkt < NBK) nb = s + 1;
    if (lane == s && bkt < NBK) {
      blx = o2f(lx); bly = o2f(ly); blz = o2f(lz);
      bhx = o2f(hx); bhy = o2f(hy); bhz = o2f(hz);
      bmax = 1e10f;
      bidx = mi;
    }
  }

  uint32_t r_rec[2] = {0, 0}, r_ridx[2] = {0, 0}, r_bar[2] = {0, 0};   // DSMEM addresses in peer CTA `lane`
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
      r_ridx[u] = mapa(smem_u32(&s_ridx[u][g]), dst);
      r_bar[u] = mapa(smem_u32(&s_bar[u]), dst);
    }
  }
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // peers resident + mbarrier inits visible before any DSMEM store

  const float* p0 = xyz + (size_t)b * N * 3;
  float qx = __ldg(p0), qy = __ldg(p0 + 1), qz = __ldg(p0 + 2);       // pick 0 = point 0
  if (rank == 0 && tid == 0) o[0] = 0;

  // this warp's published record (recomputed only when one of its buckets changed)
  uint32_t wmax = 0u, widx = kInf;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  const float4* wp = s_pts + warp * 32 + lane;      // + s * NW * 32
  float* wm = s_md + warp * 32 + lane;

#ifdef SAD_FPS_PROFILE
  long long ph[5] = {0, 0, 0, 0, 0}, nupd = 0, tprev = clock64();
#define SAD_MARK(i) { const long long tn = clock64(); ph[i] += tn - tprev; tprev = tn; }
#else
#define SAD_MARK(i)
#endif
  for (int j = 1; j < npoint; ++j) {
    const int buf = j & 1;
    if (CS > 1 && tid == 0) mbar_arrive_expect_tx(&s_bar[buf], NSLOT * 20);

    // ---- cull: which of my buckets can the new pick change?
    bool aff = false;
    if (lane < nb) {
      const float ccx = fminf(fmaxf(qx, blx), bhx), ccy = fminf(fmaxf(qy, bly), bhy), ccz = fminf(fmaxf(qz, blz), bhz);
      aff = sqdist(ccx, ccy, ccz, qx, qy, qz) < bmax;
    }
    uint32_t mask = __ballot_sync(FULL, aff);
    const bool changed = (mask != 0u) || (j == 1);
    SAD_MARK(0)
#ifdef SAD_FPS_PROFILE
    nupd += __popc(mask);
#endif
    while (mask) {                                     // two buckets per step: their redux chains overlap
      const int s0 = __ffs(mask) - 1;
      mask &= mask - 1;
      const bool two = mask != 0u;
      const int s1 = two ? __ffs(mask) - 1 : s0;
      if (two) mask &= mask - 1;
      const float4 pa = wp[s0 * NW * 32], pb = wp[s1 * NW * 32];
      const float ma0 = wm[s0 * NW * 32], mb0 = wm[s1 * NW * 32];
      const float ma = fminf(ma0, sqdist(pa.x, pa.y, pa.z, qx, qy, qz));
      const float mb = fminf(mb0, sqdist(pb.x, pb.y, pb.z, qx, qy, qz));
      if (ma < ma0) wm[s0 * NW * 32] = ma;
      if (two && mb < mb0) wm[s1 * NW * 32] = mb;
      const uint32_t ua = __float_as_uint(ma), ub = __float_as_uint(mb);       // >= 0: bit order == value order
      const uint32_t mxa = __reduce_max_sync(FULL, ua), mxb = __reduce_max_sync(FULL, ub);
      const uint32_t ixa = __reduce_min_sync(FULL, ua == mxa ? __float_as_uint(pa.w) : kInf);
      const uint32_t ixb = __reduce_min_sync(FULL, ub == mxb ? __float_as_uint(pb.w) : kInf);
      if (lane == s0) {
        bmax = __uint_as_float(mxa);
        bidx = ixa;
      }
      if (two && lane == s1) {
        bmax = __uint_as_float(mxb);
        bidx = ixb;
      }
    }
    SAD_MARK(1)
    if (changed) {
      const uint32_t wb = (lane < nb) ? __float_as_uint(bmax) : 0u;
      wmax = __reduce_max_sync(FULL, wb);
      widx = __reduce_min_sync(FULL, (lane < nb && wb == wmax) ? bidx : kInf);
      const uint32_t own = __ballot_sync(FULL, lane < nb && wb == wmax && bidx == widx);
      const int s = own ? __ffs(own) - 1 : 0;
      const float4 p = wp[s * NW * 32];
      const uint32_t ml = __ballot_sync(FULL, __float_as_uint(p.w) == widx);
      const int src = ml ? __ffs(ml) - 1 : 0;
      cx = __shfl_sync(FULL, p.x, src);
      cy = __shfl_sync(FULL, p.y, src);
      cz = __shfl_sync(FULL, p.z, src);
    }

    SAD_MARK(2)
    // ---- publish
    if (CS == 1) {
      if (lane == 0) {
        s_rec[buf][warp] = make_float4(__uint_as_float(wmax), cx, cy, cz);
        s_ridx[buf][warp] = widx;
      }
      __syncthreads();
    } else {
      if (lane < CS) {
        const uint32_t a_rec = buf ? r_rec[1] : r_rec[0], a_idx = buf ? r_ridx[1] : r_ridx[0],
                       a_bar = buf ? r_bar[1] : r_bar[0];
        st_async_v4(a_rec, __uint_as_float(wmax), cx, cy, cz, a_bar);
        st_async_b32(a_idx, widx, a_bar);
      }
      mbar_wait(&s_bar[buf], (uint32_t)(((j - 1) >> 1) & 1));
    }

    SAD_MARK(3)
    // ---- every warp reduces the NSLOT records: max value, ties -> lowest original index
    uint32_t v = 0u, vi = kInf, vs = 0u;
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
      const int sl = lane + 32 * r;
      const bool in = sl < NSLOT;
      const int slc = in ? sl : 0;
      const uint32_t x = __float_as_uint(s_rec[buf][slc].x);
      const uint32_t id = s_ridx[buf][slc];
      // branch-free: a divergent compare chain would leave the warp unconverged at the redux below (slow path)
      const bool take = in & ((x > v) | ((x == v) & (id < vi)));
      v = take ? x : v;
      vi = take ? id : vi;
      vs = take ? (uint32_t)sl : vs;
    }
    const uint32_t gmax = __reduce_max_sync(FULL, v);
    const uint32_t gidx = __reduce_min_sync(FULL, v == gmax ? vi : kInf);
    const uint32_t gl = __ballot_sync(FULL, v == gmax && vi == gidx);
    const uint32_t gs = __shfl_sync(FULL, vs, gl ? __ffs(gl) - 1 : 0);
    const float4 w = s_rec[buf][gs];
    qx = w.y;
    qy = w.z;
    qz = w.w;
    if (rank == 0 && tid == 0) o[j] = (int32_t)gidx;
    SAD_MARK(4)
  }
#ifdef SAD_FPS_PROFILE
  if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 7))
    printf("[fps_cull CS=%d N=%d warp %d] per pick: cull %lld  update %lld (%.2f buckets)  record %lld  exchange %lld  reduce %lld cycles\n",
           CS, N, warp, ph[0] / (npoint - 1), ph[1] / (npoint - 1), (double)nupd / (npoint - 1), ph[2] / (npoint - 1),
           ph[3] / (npoint - 1), ph[4] / (npoint - 1));
#endif
#undef SAD_MARK
  if (CS > 1) cluster_sync_all();   // no CTA retires while a peer's st.async may still target it
}

// ---------------------------------------------------------------------------------------------
// Single-CTA variant: one SM per scene, any N.  The points stay in the (L2-resident) sorted array and
// the min-distances in the workspace scratch; the CTA keeps only per-bucket state on chip: box / bmax /
// bidx in registers (slot s of a warp lives in lane s % 32, register set s / 32) and each bucket's
// best point in shared memory.  A pick costs: each warp's record merged into a packed 64-bit key (shared-memory
// max), ONE block barrier, the winner's record read by every thread, the lane-parallel box test, and an L2 round
// trip for the handful of buckets that survive it (four at a time per warp for latency overlap) -- no cluster, no
// DSMEM exchange, and 1/4 .. 1/16 of the SMs of the other kernels.  (Two scenes per CTA, SC == 2, keep the earlier
// barrier -> warp-0 reduce -> barrier sequence on a named barrier per scene.)
constexpr int FC1_NW = 16;                // warps of the capacity bound (sad_fps_grid_max_points)
constexpr int FC1_DEPTH = 4;              // bucket updates (independent L2 round trips) in flight per warp
constexpr int FC1_OUT = 2048;             // picks buffered in shared memory between flushes to the output
constexpr int FC1_MDS_MAX = 46000;        // min-distances in shared memory up to this many points (4.5 B per point)

// MDS: the min-distances live in shared memory instead of the workspace scratch (scenes up to FC1_MDS_MAX points), and
// the picks are buffered in shared memory and written out once per FC1_OUT picks: no global store sits between the
// two barriers of a pick.  Measured on 8 x 40k points -> 2048: 2.99 ms (3.05 ms with both in global memory).  What a
// pick costs (ablation build -DSAD_FPS_ABLATE, no bucket updates): ~1.5 k cycles FIXED -- 32 warps x ~100
// instructions of box tests and bookkeeping issue-bound on the four schedulers, plus the record -> barrier -> one-warp
// reduce -> barrier chain -- and ~1.5 k cycles for the L2 round trip of the surviving buckets' points.
// SC: scenes per CTA.  SC == 2 puts two independent scenes (NW warps each, one named barrier each) on one SM, so one
// scene's L2 round trip hides behind the other's barrier / reduce / record chain; it needs the per-thread state under
// 64 registers (fewer updates in flight) and keeps the min-distances in the workspace scratch (MDS == false).
template <int R, int NW, bool MDS, int SC>
__global__ void __launch_bounds__(SC * NW * 32, 1)
fps_cull1_kernel(int B, int N, int npoint, const float* __restrict__ xyz, uint8_t* __restrict__ ws, size_t stride,
                 int32_t* __restrict__ out, size_t scene_smem) {
  extern __shared__ __align__(16) uint8_t s_dyn1[];
  __shared__ __align__(16) float4 s_wrec_[SC][NW];      // per warp: best point {x,y,z,bits(idx)}
  __shared__ uint32_t s_wval_[SC][NW];                  // per warp: its min-dist bits
  __shared__ __align__(16) float4 s_pick_[SC];          // the pick of this round {x,y,z,bits(idx)}, written by warp 0
  // SC == 1: ONE barrier per pick.  Every warp merges {min-dist bits | inverted original index | warp} into a packed
  // 64-bit key with a shared-memory max before the barrier; after it every thread reads the winning key and the winning
  // warp's record itself -- no reduce by warp 0 and no second barrier (ncu: 4 of 20 warps sat at those barriers per
  // issued instruction).  Keys are triple-buffered (the key of pick j+2 is cleared after pick j's barrier), records
  // double-buffered (a slow warp may still read pick j's while a fast one writes pick j+1's).
  __shared__ unsigned long long s_key[3];
  __shared__ __align__(16) float4 s_wrec2[2][NW];
  static_assert(!(MDS && SC > 1), "two scenes per CTA keep the min-distances in global memory");

  const int half = SC > 1 ? (int)threadIdx.x / (NW * 32) : 0;
  const int tid = SC > 1 ? (int)threadIdx.x % (NW * 32) : (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x * SC + half;
  if (b >= B) return;                                   // odd batch: the second half of the last CTA has no scene
  auto cta_sync = [&]() {
    if (SC > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "r"(NW * 32) : "memory");
    else __syncthreads();
  };
  float4* s_best = reinterpret_cast<float4*>(s_dyn1 + (size_t)half * scene_smem);   // [slot * NW + warp] best point of the bucket
  // then: int32 s_out[FC1_OUT]; then (MDS) float s_md[N]
  float4* s_wrec = s_wrec_[half];
  uint32_t* s_wval = s_wval_[half];
  float4& s_pick = s_pick_[half];
  uint8_t* base = ws + (size_t)b * stride;
  const float4* sorted = reinterpret_cast<const float4*>(base + kGridHeaderBytes + kGridCellBytes);
  int32_t* o = out + (size_t)b * npoint;
  const int NB = (N + 31) >> 5;
  int32_t* s_out = reinterpret_cast<int32_t*>(s_best + (size_t)((NB + NW - 1) / NW) * NW);
  float* mind = MDS ? reinterpret_cast<float*>(s_out + FC1_OUT) : reinterpret_cast<float*>(base + grid_scratch_offset(N));
  const int nslots = NB > warp ? (NB - warp + NW - 1) / NW : 0;      // buckets warp, warp+NW, ...

  const float* p0 = xyz + (size_t)b * N * 3;
  float qx = __ldg(p0), qy = __ldg(p0 + 1), qz = __ldg(p0 + 2);       // pick 0 = point 0
  if (tid == 0) {
    s_out[0] = 0;
    if (npoint == 1) o[0] = 0;            // no pick loop, no flush
  }

  float blx[R], bly[R], blz[R], bhx[R], bhy[R], bhz[R], bmax[R];
  uint32_t bidx[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    blx[r] = bly[r] = blz[r] = bhx[r] = bhy[r] = bhz[r] = bmax[r] = 0.f;
    bidx[r] = kInf;
  }
  // ---- load pass: boxes, and pick 0 applied on the fly (min-dist = min(1e10, d2(p, p0)))
  // (register set r as the OUTER, unrolled loop: written as one loop over s with an `if (s >> 5 == r)` chain the
  // compiler turns the eight state arrays into dynamically indexed local memory)
#pragma unroll
  for (int r = 0; r < R; ++r)
  for (int sl = 0; sl < 32; ++sl) {
    const int s = r * 32 + sl;
    if (s >= nslots) break;
    const int k = (s * NW + warp) * 32 + lane;
    const bool ok = k < N;
    float4 p = make_float4(0.f, 0.f, 0.f, __uint_as_float(kInf));
    float m = 0.f;
    if (ok) {
      p = __ldg(sorted + k);
      m = fminf(1e10f, sqdist(p.x, p.y, p.z, qx, qy, qz));
      mind[k] = m;
    }
    const uint32_t lx = __reduce_min_sync(FULL, ok ? f2o(p.x) : kInf), hx = __reduce_max_sync(FULL, ok ? f2o(p.x) : 0u);
    const uint32_t ly = __reduce_min_sync(FULL, ok ? f2o(p.y) : kInf), hy = __reduce_max_sync(FULL, ok ? f2o(p.y) : 0u);
    const uint32_t lz = __reduce_min_sync(FULL, ok ? f2o(p.z) : kInf), hz = __reduce_max_sync(FULL, ok ? f2o(p.z) : 0u);
    const uint32_t mb = __float_as_uint(m);
    const uint32_t mx = __reduce_max_sync(FULL, mb);
    const uint32_t ix = __reduce_min_sync(FULL, (ok && mb == mx) ? __float_as_uint(p.w) : kInf);
    if (ok && __float_as_uint(p.w) == ix) s_best[s * NW + warp] = p;
    if (lane == sl) {
      blx[r] = o2f(lx); bly[r] = o2f(ly); blz[r] = o2f(lz);
      bhx[r] = o2f(hx); bhy[r] = o2f(hy); bhz[r] = o2f(hz);
      bmax[r] = __uint_as_float(mx);
      bidx[r] = ix;
    }
  }
  bool changed = true;
  unsigned long long ckey = 0ull;                       // this warp's key (warp-uniform), kept while its buckets do not change
  int k3 = 1;                                           // j % 3
  if (SC == 1) {
    if (tid < 3) s_key[tid] = 0ull;
    cta_sync();
  }
#ifdef SAD_FPS_PROFILE
  long long ph[5] = {0, 0, 0, 0, 0}, nupd = 0, nround = 0, tprev = clock64();
#define SAD_MARK1(i) { const long long tn = clock64(); ph[i] += tn - tprev; tprev = tn; }
#else
#define SAD_MARK1(i)
#endif

  for (int j = 1; j < npoint; ++j) {
    if constexpr (SC == 1) {
      // ---- this warp's record and key (recomputed only when one of its buckets changed)
      __syncwarp();
      const int p = j & 1;
      if (changed) {
        uint32_t v = 0u, vi = kInf;
        int vr = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const bool valid = (r * 32 + lane) < nslots;
          const uint32_t x = valid ? __float_as_uint(bmax[r]) : 0u;
          const uint32_t id = valid ? bidx[r] : kInf;
          // branch-free on purpose: a divergent compare chain leaves the warp unconverged at the redux below, which
          // then takes its slow collective path (measured: ~900 instead of ~150 cycles for this block)
          const bool take = (x > v) | ((x == v) & (id < vi));
          v = take ? x : v;
          vi = take ? id : vi;
          vr = take ? r : vr;
        }
        const uint32_t wmax = __reduce_max_sync(FULL, v);
        const uint32_t widx = __reduce_min_sync(FULL, v == wmax ? vi : kInf);
        if (widx == kInf) {                               // no bucket: loses against every real record (key < 32)
          ckey = (unsigned long long)warp;
          if (lane == 0) s_wrec2[p][warp] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kInf));
        } else {                                          // max min-dist, ties -> lowest original index (< 2^19)
          ckey = ((unsigned long long)wmax << 32) | ((unsigned long long)(0x7FFFFu - widx) << 5) | (unsigned long long)warp;
          if (v == wmax && vi == widx) s_wrec2[p][warp] = s_best[(vr * 32 + lane) * NW + warp];
        }
      } else if (lane == 0) {
        s_wrec2[p][warp] = s_wrec2[p ^ 1][warp];
      }
      if (lane == 0) atomicMax(&s_key[k3], ckey);
      SAD_MARK1(0)
      cta_sync();
      SAD_MARK1(1)
      const unsigned long long key = s_key[k3];
      const float4 w = s_wrec2[p][(int)(key & 31ull)];
      qx = w.x;
      qy = w.y;
      qz = w.z;
      if (tid == 0) {
        s_out[j & (FC1_OUT - 1)] = (int32_t)__float_as_uint(w.w);
        s_key[k3 == 0 ? 2 : k3 - 1] = 0ull;               // the key of pick j + 2
      }
      k3 = k3 == 2 ? 0 : k3 + 1;
      if (((j + 1) & (FC1_OUT - 1)) == 0 || j == npoint - 1) {     // flush the buffered picks (coalesced, rare)
        cta_sync();
        const int j0 = j & ~(FC1_OUT - 1);
        for (int i = tid; j0 + i <= j; i += NW * 32) o[j0 + i] = s_out[i];
      }
    } else {
      // ---- this warp's record: its best bucket (rewritten only when one of its buckets changed; records are read
      // by warp 0 alone, between the two barriers, so one buffer suffices and an unchanged warp does nothing)
      __syncwarp();
      if (changed) {
        uint32_t v = 0u, vi = kInf;
        int vr = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const bool valid = (r * 32 + lane) < nslots;
          const uint32_t x = valid ? __float_as_uint(bmax[r]) : 0u;
          const uint32_t id = valid ? bidx[r] : kInf;
          // branch-free on purpose: a divergent compare chain leaves the warp unconverged at the redux below, which
          // then takes its slow collective path (measured: ~900 instead of ~150 cycles for this block)
          const bool take = (x > v) | ((x == v) & (id < vi));
          v = take ? x : v;
          vi = take ? id : vi;
          vr = take ? r : vr;
        }
        const uint32_t wmax = __reduce_max_sync(FULL, v);
        const uint32_t widx = __reduce_min_sync(FULL, v == wmax ? vi : kInf);
        if (widx == kInf) {
          if (lane == 0) {
            s_wval[warp] = 0u;
            s_wrec[warp] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kInf));
          }
        } else if (v == wmax && vi == widx) {
          s_wval[warp] = wmax;
          s_wrec[warp] = s_best[(vr * 32 + lane) * NW + warp];
        }
      }
      SAD_MARK1(0)
      cta_sync();
      SAD_MARK1(1)

      // ---- warp 0 reduces the NW records (max value, ties -> lowest original index) and publishes the pick: one
      // warp's two redux instead of every warp's (the redux unit is shared; NW redundant reductions queue on it)
#if defined(SAD_FPS_ABLATE) && SAD_FPS_ABLATE >= 4
      if (false) {
#else
      if (warp == 0) {
#endif
        const uint32_t x = lane < NW ? s_wval[lane] : 0u;
        const uint32_t id = lane < NW ? __float_as_uint(s_wrec[lane].w) : kInf;
        const uint32_t gmax = __reduce_max_sync(FULL, x);
        const uint32_t gidx = __reduce_min_sync(FULL, x == gmax ? id : kInf);
        if (x == gmax && id == gidx && lane < NW) {
          s_pick = s_wrec[lane];
          s_out[j & (FC1_OUT - 1)] = (int32_t)gidx;
        }
      }
      cta_sync();
      if (((j + 1) & (FC1_OUT - 1)) == 0 || j == npoint - 1) {     // flush the buffered picks (coalesced, rare)
        const int j0 = j & ~(FC1_OUT - 1);
        for (int i = tid; j0 + i <= j; i += NW * 32) o[j0 + i] = s_out[i];
      }
      {
        const float4 w = s_pick;
        qx = w.x;
        qy = w.y;
        qz = w.z;
      }
    }
    if (j == npoint - 1) break;
    SAD_MARK1(2)

    // ---- cull + update: only buckets whose box the pick can reach
    changed = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      bool aff = false;
#if !defined(SAD_FPS_ABLATE) || SAD_FPS_ABLATE < 3
      if (r * 32 + lane < nslots) {
        const float ccx = fminf(fmaxf(qx, blx[r]), bhx[r]), ccy = fminf(fmaxf(qy, bly[r]), bhy[r]),
                    ccz = fminf(fmaxf(qz, blz[r]), bhz[r]);
        aff = sqdist(ccx, ccy, ccz, qx, qy, qz) < bmax[r];
      }
#endif
      uint32_t mask = __ballot_sync(FULL, aff);
      if (mask) changed = true;
#ifdef SAD_FPS_PROFILE
      nupd += __popc(mask);
#endif
#ifdef SAD_FPS_ABLATE
      mask = 0u;                                       // tools only: no bucket updates at all (wrong picks): the fixed cost per pick
#if SAD_FPS_ABLATE >= 2
      changed = false;                                 // ... and no record rewrite
#endif
#endif
      while (mask) {                                   // FC1_DEPTH buckets in flight per step (independent L2 round trips)
#ifdef SAD_FPS_PROFILE
        ++nround;
#endif
        constexpr int DEPTH = SC > 1 ? 2 : FC1_DEPTH;
        int bsel[DEPTH], kk[DEPTH];
        bool okk[DEPTH], use[DEPTH];
        float4 pt[DEPTH];
        float md[DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
          use[u] = mask != 0u;
          bsel[u] = use[u] ? __ffs(mask) - 1 : 0;
          if (use[u]) mask &= mask - 1;
          kk[u] = ((r * 32 + bsel[u]) * NW + warp) * 32 + lane;
          okk[u] = use[u] && kk[u] < N;
          pt[u] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kInf));
          md[u] = 0.f;
          if (okk[u]) {
            pt[u] = __ldg(sorted + kk[u]);
            md[u] = mind[kk[u]];
          }
        }
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
          if (!use[u]) break;                            // warp-uniform
          const float nd = fminf(md[u], sqdist(pt[u].x, pt[u].y, pt[u].z, qx, qy, qz));
          if (okk[u] && nd < md[u]) mind[kk[u]] = nd;
          const uint32_t un = __float_as_uint(nd);       // >= 0: bit order == value order
          const uint32_t mx = __reduce_max_sync(FULL, un);
          const uint32_t ix = __reduce_min_sync(FULL, (okk[u] && un == mx) ? __float_as_uint(pt[u].w) : kInf);
          if (okk[u] && __float_as_uint(pt[u].w) == ix) s_best[(r * 32 + bsel[u]) * NW + warp] = pt[u];
          if (lane == bsel[u]) {
            bmax[r] = __uint_as_float(mx);
            bidx[r] = ix;
          }
        }
      }
    }
    SAD_MARK1(3)
  }
#ifdef SAD_FPS_PROFILE
  if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 7 || warp == NW - 1))
    printf("[fps_cull1 R=%d NW=%d N=%d warp %d] per pick: record %lld  barrier %lld  reduce %lld  box+update %lld cycles  (%.2f buckets, %.2f rounds)\n",
           R, NW, N, warp, ph[0] / (npoint - 1), ph[1] / (npoint - 1), ph[2] / (npoint - 1), ph[3] / (npoint - 1),
           (double)nupd / (npoint - 1), (double)nround / (npoint - 1));
#endif
#undef SAD_MARK1
}

template <int R, int NW, bool MDS, int SC = 1>
int launch_cull1(int B, int N, int npoint, const float* xyz, void* ws, int32_t* idx, cudaStream_t stream) {
  auto kern = fps_cull1_kernel<R, NW, MDS, SC>;
  const int nb = (N + 31) / 32;
  const size_t scene_smem = ((size_t)sad_ceil_div(nb, NW) * NW * sizeof(float4) + FC1_OUT * sizeof(int32_t) +
                             (MDS ? (size_t)N * sizeof(float) : 0) + 15) / 16 * 16;
  const size_t smem = SC * scene_smem;
  static thread_local int configured_dev = -1;
  int dev = 0;
  SAD_CUDA_OK(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    SAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     MDS ? 227 * 1024 - 2048 : SC * (R * 32 * NW * 16 + FC1_OUT * 4 + 16)));
    configured_dev = dev;
  }
  kern<<<sad_ceil_div(B, SC), SC * NW * 32, smem, stream>>>(B, N, npoint, xyz, static_cast<uint8_t*>(ws), sad::grid_stride(N), idx,
                                                       scene_smem);
  SAD_LAUNCH_CHECK("fps_cull1_kernel");
  return SAD_OK;
}

template <int CS>
int launch_cull(int B, int N, int npoint, const float* xyz, const void* ws, int32_t* idx, int slots, cudaStream_t stream) {
  auto kern = fps_cull_kernel<CS>;
  const size_t smem = (size_t)slots * FC_NW * 32 * 20;
  static thread_local int configured_dev = -1;
  int dev = 0;
  SAD_CUDA_OK(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    SAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     FC_MAX_SLOTS * FC_NW * 32 * 20));
    if (CS > 8) SAD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured_dev = dev;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CS));
  cfg.blockDim = dim3(FC_T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (CS > 1) ? 1 : 0;
  SAD_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, N, npoint, xyz, static_cast<const uint8_t*>(ws), sad::grid_stride(N), idx,
                                 slots));
  sad_count_launch(1);
  return SAD_OK;
}

}  // namespace

// Tests / benchmarks: 0 = default (shared-memory-resident cluster kernel with the fewest CTAs that hold the
// scene; single-CTA L2-resident kernel beyond its capacity); 1,2,4,8,16 = cluster kernel with at least that
// many CTAs per scene; -1 = single-CTA kernel.
#ifdef SAD_TOOLS_ABLATE
__global__ void ablate_strided_idx(int N, int npoint, int32_t* idx) {
  for (int j = threadIdx.x; j < npoint; j += blockDim.x) idx[(size_t)blockIdx.x * npoint + j] = (int32_t)((long long)j * N / npoint);
}
#endif

// Largest scene the culled kernels accept.
extern "C" int sad_fps_grid_max_points(void) { return 16 * 32 * FC1_NW * 32; }

extern "C" int sad_furthest_point_sample_grid_fwd(int B, int N, int npoint, const float* xyz, void* grid_ws,
                                                  int32_t* idx, sad_stream_t stream_) {
  return sad_furthest_point_sample_grid_policy_fwd(B, N, npoint, xyz, grid_ws, idx, SAD_FPS_LATENCY, 0, stream_);
}

extern "C" int sad_furthest_point_sample_grid_policy_fwd(int B, int N, int npoint, const float* xyz, void* grid_ws,
                                                         int32_t* idx, int policy, int variant, sad_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SAD_REQUIRE(policy == SAD_FPS_LATENCY || policy == SAD_FPS_THROUGHPUT || policy == SAD_FPS_THROUGHPUT_PAIRED,
              "furthest_point_sample_grid: bad policy %d", policy);
  SAD_REQUIRE(variant == 0 || variant == -1 || variant == -2 || variant == 1 || variant == 2 || variant == 4 || variant == 8 || variant == 16,
              "furthest_point_sample_grid: bad variant %d", variant);
  const int force = variant != 0 ? (variant == -2 ? -1 : variant) : (policy == SAD_FPS_LATENCY ? 0 : -1);
  SAD_REQUIRE(B >= 0 && N >= 1 && npoint >= 1, "furthest_point_sample_grid: bad sizes B=%d N=%d npoint=%d", B, N, npoint);
  if (B == 0) return SAD_OK;
  SAD_REQUIRE(xyz && grid_ws && idx, "furthest_point_sample_grid: null pointer");
#ifdef SAD_TOOLS_ABLATE
  if (sad_ablate_mask() & 1) {                           // tools only: what does the pipeline cost without this kernel?
    ablate_strided_idx<<<B, 256, 0, stream>>>(N, npoint, idx);
    return SAD_OK;
  }
#endif
  if (N > 16 * (FC_MAX_SLOTS * FC_NW * 32) && N <= 204800 && force == 0)
    return sad_furthest_point_sample_fwd(B, N, npoint, xyz, idx, stream_);   // beyond the cluster capacity the register-resident kernel wins
  if (N > sad_fps_grid_max_points()) {
    sad_set_error("furthest_point_sample_grid: N=%d exceeds the capacity (%d)", N, sad_fps_grid_max_points());
    return SAD_EUNSUPPORTED;
  }
  const int nbk = (N + 31) / 32;
  const int cap = FC_MAX_SLOTS * FC_NW * 32;           // points per CTA of the cluster kernel
  if (force >= 0 && N <= 16 * cap) {
    int cs = 1;
    while (cs < force && cs < 16) cs <<= 1;
    while (cs * cap < N) cs <<= 1;
    const int slots = sad_ceil_div(sad_ceil_div(nbk, cs), FC_NW);
    switch (cs) {
      case 1: return launch_cull<1>(B, N, npoint, xyz, grid_ws, idx, slots, stream);
      case 2: return launch_cull<2>(B, N, npoint, xyz, grid_ws, idx, slots, stream);
      case 4: return launch_cull<4>(B, N, npoint, xyz, grid_ws, idx, slots, stream);
      case 8: return launch_cull<8>(B, N, npoint, xyz, grid_ws, idx, slots, stream);
      default: return launch_cull<16>(B, N, npoint, xyz, grid_ws, idx, slots, stream);
    }
  }
  // one SM per scene: 32 warps while two register sets of bucket state per lane suffice (N <= 65536), else 16 warps
  const int per_lane32 = sad_ceil_div(sad_ceil_div(nbk, 32), 32);
  if (N <= FC1_MDS_MAX) {                 // min-distances in shared memory
    // 16 warps per scene: a pick's fixed cost is issue-bound (every warp runs ~180 bookkeeping + box-test
    // instructions per pick), so 16 warps x 3 register sets beat 32 x 2 (2.83 vs 2.99 ms at 40k points)
    if (policy == SAD_FPS_THROUGHPUT_PAIRED && B >= 2) {      // two scenes per SM (16 warps each)
      const int pl = sad_ceil_div(sad_ceil_div(nbk, 16), 32);
      if (pl <= 1) return launch_cull1<1, 16, false, 2>(B, N, npoint, xyz, grid_ws, idx, stream);
      if (pl <= 2) return launch_cull1<2, 16, false, 2>(B, N, npoint, xyz, grid_ws, idx, stream);
      return launch_cull1<3, 16, false, 2>(B, N, npoint, xyz, grid_ws, idx, stream);
    }
    const char* e_nw = sad_tool_env("SAD_FPS1_NW");                                      // tools: warps per scene
    if (!e_nw || atoi(e_nw) != 32) {
      // (register sets R, warps NW) with the fewest lane-rounds of box tests per pick (R x NW) that holds the scene's
      // buckets: a partly filled last register set costs a full round, so 20 / 24 warps with full sets beat 16 warps
      // with one more set (40k points: <2,20> 1.10 us per pick, <2,28> 1.11, <2,24> 1.22, <3,16> 1.26, <2,32> 1.34).
      // variant -2 (tests / tools) keeps to the 16-warp instances.
      const bool wide = variant != -2;
      if (nbk <= 16 * 32) return launch_cull1<1, 16, true>(B, N, npoint, xyz, grid_ws, idx, stream);
      if (wide && nbk <= 20 * 32) return launch_cull1<1, 20, true>(B, N, npoint, xyz, grid_ws, idx, stream);
      if (nbk <= 16 * 64) return launch_cull1<2, 16, true>(B, N, npoint, xyz, grid_ws, idx, stream);
      if (wide && nbk <= 20 * 64) return launch_cull1<2, 20, true>(B, N, npoint, xyz, grid_ws, idx, stream);
      if (wide && nbk <= 24 * 64) return launch_cull1<2, 24, true>(B, N, npoint, xyz, grid_ws, idx, stream);
      return launch_cull1<3, 16, true>(B, N, npoint, xyz, grid_ws, idx, stream);
    }
    if (per_lane32 <= 1) return launch_cull1<1, 32, true>(B, N, npoint, xyz, grid_ws, idx, stream);
    return launch_cull1<2, 32, true>(B, N, npoint, xyz, grid_ws, idx, stream);
  }
  if (per_lane32 <= 2) return launch_cull1<2, 32, false>(B, N, npoint, xyz, grid_ws, idx, stream);
  const int per_lane = sad_ceil_div(sad_ceil_div(nbk, FC1_NW), 32);   // bucket slots per lane
  if (per_lane <= 6) return launch_cull1<6, FC1_NW, false>(B, N, npoint, xyz, grid_ws, idx, stream);
  if (per_lane <= 8) return launch_cull1<8, FC1_NW, false>(B, N, npoint, xyz, grid_ws, idx, stream);
  return launch_cull1<16, FC1_NW, false>(B, N, npoint, xyz, grid_ws, idx, stream);
}
