// a6 (fast path)  shape-specialised fused SA stage: neighbourhood gather -> 3-layer shared MLP -> max-pool over nsample
// -- SURVEY.md section 8(a) rows a5/a6, section 8(f) rank 1, hard parts H4/H5; VERDICT r1 item 1.
// (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
// Same math as csrc/mlp.cu (which stays as the general-shape kernel); this one is compiled per stage shape
// <CG, NF, H, C3, S> so that every loop bound, descriptor and barrier index is a constant, and is built around what
// bounded the general kernel: L2 -> shared-memory traffic and single-warp latency chains.
//
//   * ALL weights are pinned in shared memory for the whole kernel.  The 128-wide stages only fit because a CTA PAIR
//     shares them: cta_group::2 MMAs (M = 256 = two 128-row tiles, one per CTA) read half of every hidden-layer
//     weight matrix from each CTA, and the last layer -- evaluated transposed, D^T = W3 . H^T, so that the max-pool is
//     an in-thread reduction -- gives each CTA 128 of the 256 output channels for BOTH tiles' rows.
//   * The layer-1 operand of a tile is built in a TILE BUFFER that later holds the tile's hidden activations (the
//     gathered rows are dead once layer 1 has completed): channel-last bf16 feature rows by 16-byte cp.async straight
//     into the tcgen05 K-major SWIZZLE_128B layout (a warp copies its own 32 rows, 8 lanes per row, source rows
//     exchanged by shuffle: no shared-memory index table, no block barrier), one tile of loads in flight behind the
//     tile being published; the relative xyz / scalar features / constant 1 form one 16-wide K step in the no-swizzle
//     K-major layout (4 KB per tile instead of a 16 KB chunk).  NB >= G buffers: the gather runs up to NB tiles ahead.
//     (TMA gather4 was tried first: ~100 cycles of issue per 4-row instruction, 5x slower than the cp.async path.)
//   * Biases of layers 1 and 2 ride on the constant-1 column (K index 15) of the tile's special K step: layer 1 has it
//     as part of its operand, layer 2 issues one extra K = 16 step of the same block against a weight piece that is
//     zero except for that column, so the hidden epilogues are TMEM -> ReLU/bf16 -> smem only.
//   * G = 2 or 4 tile contexts in TMEM, walked in two half-groups in anti-phase (L1 A, L1 B, L2 A, L2 B, L3 A, L3 B):
//     the epilogue of one half runs under the MMAs of the other; per step the MMA warp waits for every barrier of the
//     half, fences once and issues everything.
//   * 16 epilogue warps (4 per scheduler): one warpgroup per context, or two splitting its columns.
//   * cluster rank 1 has no MMAs to issue: its MMA warp RELAYS the local "operand ready" barriers to the leader with
//     one remote arrive each, in exactly the order the leader consumes them; tcgen05.commit multicasts the
//     "accumulator ready" / "buffer free" barriers to both CTAs.
#include <string.h>

#include "sad_tc.cuh"

namespace {

using namespace sad;

constexpr int kWarpGather = 16;          // warps 16-19: layer-1 operand (feature rows, special K step)
constexpr int kWarpMma = 20;             // warps 20, 21: one per half of the contexts; leader: MMA issue, rank 1: relay
constexpr int kWarpProd = 22;            // tile scheduler + pinned weights
constexpr int kThreads = 23 * 32;        // (registers are allocated for 24 warps either way)
constexpr int kGatherThreads = 128;
constexpr int kChunk = 16384;            // 128 rows x 128 B
constexpr int kSpBytes = 4096;           // 128 rows x 32 B (one 16-wide bf16 K step)
constexpr int kRing = 32;                // tile-id ring
constexpr uint32_t kNoRow = 0xFFFFFFFFu;

template <int CG_, int NF_, int H_, int C3_, int S_>
struct Cfg {
  static constexpr int CG = CG_, NF = NF_, H = H_, C3 = C3_, S = S_;
  static constexpr int HC = H / 64;                      // 64-wide chunks of a hidden activation
  static constexpr int WROWS = H / CG;                   // hidden-layer weight rows held by one CTA
  static constexpr int NQ = C3 / 128;                    // 128-column blocks of the transposed last layer per CTA
  static constexpr int NBLK = C3 / (128 * CG);           // MMA groups of the last layer (M = 128 * CG channels each)
  static constexpr int CW = (H > C3 ? H : C3) <= 128 ? 128 : 256;   // TMEM columns of one context
  static constexpr int G = 512 / CW;                     // contexts
  static constexpr int HG = G / 2;                       // contexts per half-group
  static constexpr int NP = 4 / G;                       // epilogue warpgroups per context
  static constexpr int HP = H / NP;                      // hidden columns per warpgroup
  static constexpr int NQP = NQ / NP;                    // last-layer blocks per warpgroup
  static constexpr int BUFCH = NF > HC ? NF : HC;        // 16 KB chunks of a tile buffer
  // weight image of one CTA: W1 feature pieces | W1 special K step | W2 pieces | W2 bias K step | W3 pieces
  static constexpr int W1F = NF * WROWS * 128, WK16 = WROWS * 32, W2B = HC * WROWS * 128, W3B = NBLK * HC * kChunk;
  static constexpr int OFF_W1SP = W1F, OFF_W2 = OFF_W1SP + WK16, OFF_W2BIAS = OFF_W2 + W2B, OFF_W3 = OFF_W2BIAS + WK16;
  static constexpr int WBYTES = OFF_W3 + W3B;
  static constexpr int BUFBYTES = BUFCH * kChunk;
  static constexpr int kMisc = 2048;
  static constexpr int kBudget = 227 * 1024 - 1024 - kMisc - WBYTES;
  static constexpr int NB_fit = kBudget / (BUFBYTES + kSpBytes);
  static constexpr int NB = NB_fit > 8 ? 8 : NB_fit;     // tile buffers
  static constexpr int SMEM = 1024 + WBYTES + NB * (BUFBYTES + kSpBytes) + kMisc;
  static constexpr int PTS = 128 / S;                    // points per tile
  static_assert(H == 64 || H == 128, "hidden width");
  static_assert(C3 % (128 * CG) == 0 && C3 <= 256, "output width");
  static_assert(HP % 64 == 0 && NQ % NP == 0 && NQP >= 1, "epilogue split");
  static_assert(NB >= G, "not enough tile buffers for the contexts");
  static_assert(S == 8 || S == 16 || S == 32 || S == 64, "nsample");
  static_assert(WBYTES % 1024 == 0, "weight image alignment");
};

struct SaParams {
  int N, P, log2P, log2S;
  long long total_rows;
  uint32_t total_points;
  int num_tiles, num_units;       // unit = tile (CG = 1) or tile pair (CG = 2)
  const __nv_bfloat16* feat_cl;   // (B*N rows) x C0 bf16 channel-last (NF > 0)
  const float* xyz;               // (B,N,3)
  const float4* xyzw;             // (B,N) {x, y, z, scalar feature}: one 16-byte load per gathered row instead of 3 + E
                                  // scattered 4-byte loads (optional; needs E <= 1)
  const float* new_xyz;           // (B,P,3)
  const int32_t* idx;             // (B,P,S)
  const float* radius_t;          // (B,P) or null
  float inv_radius;               // scalar radius: 1/r (or 1 when not normalising)
  int normalize;
  const float* extra;             // (B,N,E) fp32 scalar features, E <= 4
  int E;
  const uint8_t* w_img;           // CG images of WBYTES
  const float* bias3;             // (C3), zero padded
  int c3_real;                    // channels actually stored
  __nv_bfloat16* out_cl;          // (B,P,c3_real) bf16 or null
  float* out_cf;                  // (B,c3_real,P) f32 or null
  // Duplicate-free ("planned") mode, see sad_sa_mlp_dedup_fwd: the launch works on SLOTS of S consecutive samples
  // instead of points.  idx then holds GLOBAL source rows (b * N + neighbour) per slot, q4 the slot's query point
  // {x, y, z, radius}, pid its real point id; consecutive slots with the same pid are one point (max-combined in the
  // epilogue).  The slot count is only known on the device: plan_counts = {runs of 4, runs of 2, single slots}.
  const int* plan_counts;
  const float4* q4;
  const int32_t* pid;
  uint32_t plan_points;           // real points B * P: the three run classes live in regions of 4, 2 and 1 x plan_points slots
  int* sched;                     // [0] next unit (zero between launches: the kernel resets it), [1] clusters done
  int static_sched;               // few units per cluster: unit(ordinal) = cluster + ordinal * clusters, no scheduler
  int claim;                      // dynamic scheduling: units claimed per atomic
};

struct Misc {
  uint64_t wfull, p_wfull;
  uint64_t tfull[kRing];
  uint64_t full[8], bfree[8], p_ready[8];
  uint64_t dfull[4], actfull[4], p_actfull[4];
  int units[kRing];
  uint32_t tmem_base;
  int gprog;                      // ordinal the gather warps have reached (monotonic: the scheduler's flow control)
  alignas(16) float bias3[256];
};

#ifdef SAD_MLP_PROFILE
// tools only: (event, clock) log of one warp per role of CTA 0 (role: 0 epilogue warp 0, 1 gather warp, 2 MMA warp,
// 3 scheduler warp, 4 epilogue warp 4, 5 second MMA warp)
__device__ long long g_sa_log[6][2 * 4096];
__device__ int g_sa_logn[6];
#define SALOG(role, ev)                                                         \
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && sa_logn < 4095) {           \
    g_sa_log[role][2 * sa_logn] = (ev);                                         \
    g_sa_log[role][2 * sa_logn + 1] = clock64();                                \
    g_sa_logn[role] = ++sa_logn;                                                \
  }
#else
#define SALOG(role, ev)
#endif

#ifdef SAD_MLP_PROFILE
__device__ unsigned long long* g_sa_dbg;     // host-mapped words that survive a trap (tools only)
#endif
// mbarrier wait that traps instead of hanging the GPU when a protocol bug leaves it unsatisfied (~2 s)
__device__ __noinline__ void bar_timeout(uint32_t bar_addr, uint32_t parity) {
#ifdef SAD_MLP_PROFILE
  if (g_sa_dbg) {
    volatile unsigned long long* d = g_sa_dbg;
    const unsigned slot = atomicAdd((unsigned*)(g_sa_dbg + 63), 1u);
    if (slot < 15) {
      d[4 * slot + 0] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x;
      d[4 * slot + 1] = ((unsigned long long)bar_addr << 32) | parity;
      d[4 * slot + 2] = clock64();
      d[4 * slot + 3] = 0xDEADBEEF;
    }
    __threadfence_system();
    // give the other stuck waiters a moment to report too
    const long long t = clock64();
    while (clock64() - t < 200000000LL) {
    }
  }
#endif
#ifdef SAD_MLP_DEBUG
  printf("[sad] sa_mlp: barrier timeout (block %d thread %d bar +%u parity %u)\n", (int)blockIdx.x, (int)threadIdx.x,
         bar_addr & 0xFFFFu, parity);
#endif
  __trap();
}
template <int CG>
__device__ __forceinline__ bool bar_try(uint64_t* bar, uint32_t parity) {
  return (CG == 2) ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity);
}
template <int CG>
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  long long t0 = 0;
  for (;;) {
    if ((CG == 2) ? mbar_try_wait_cluster_sleep(bar, parity) : mbar_try_wait_sleep(bar, parity)) return;
    if ((++spins & 0xFFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) bar_timeout(smem_u32(bar), parity);
    }
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <class C>
__global__ void __launch_bounds__(kThreads, 1) sa_mlp_kernel(const __grid_constant__ SaParams p) {
#ifdef SAD_MLP_PROFILE
  const long long sa_t_entry = clock64();
#endif
  constexpr int CG = C::CG, NF = C::NF, H = C::H, G = C::G, HG = C::HG, NB = C::NB, NP = C::NP;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  constexpr uint32_t off_buf = C::WBYTES;
  constexpr uint32_t off_sp = off_buf + NB * C::BUFBYTES;
  constexpr uint32_t off_misc = off_sp + NB * kSpBytes;
  Misc* ms = reinterpret_cast<Misc*>(gbase + off_misc);
  static_assert(sizeof(Misc) <= C::kMisc, "misc area");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef SAD_MLP_PROFILE
  int sa_logn = 0;
#endif
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int cluster_id = (int)blockIdx.x / CG, num_clusters = (int)gridDim.x / CG;
  const bool leader = (rank == 0);
  // problem size: launch arguments, or (planned mode) the slot counts the plan kernels left on the device
  long long total_rows = p.total_rows;
  uint32_t total_points = p.total_points;
  int num_units = p.num_units;
  constexpr bool kCanPlan = (CG == 1 && C::S <= 16);      // only the 8- / 16-sample single-CTA instances run plans
  const bool planned = kCanPlan && p.plan_counts != nullptr;
  uint32_t n4x = 0, n42x = 0;      // slots of the 4-runs, of the 4- and 2-runs
  if (planned) {
    n4x = 4u * (uint32_t)__ldg(p.plan_counts);
    n42x = n4x + 2u * (uint32_t)__ldg(p.plan_counts + 1);
    const int n = (int)n42x + __ldg(p.plan_counts + 2);
    total_points = (uint32_t)n;
    total_rows = (long long)n << p.log2S;
    num_units = (int)(((total_rows + 127) / 128 + CG - 1) / CG);
  }
  // few units per cluster: fixed assignment (see the host side); decided here when the size is only known on the device
  const bool static_sched = planned ? (long long)num_clusters * 8 >= num_units : p.static_sched != 0;

  // planned mode: logical slot (dense: 4-runs, then 2-runs, then singles) -> slot in the plan's three regions
  auto slot_addr = [&](uint32_t sl) -> uint32_t {
    return sl < n4x ? sl : (sl < n42x ? 4u * p.plan_points + (sl - n4x) : 6u * p.plan_points + (sl - n42x));
  };

  if (tid == 0) {
    mbar_init(&ms->wfull, 1);
    mbar_init(&ms->p_wfull, 1);
    for (int i = 0; i < kRing; ++i) mbar_init(&ms->tfull[i], 1);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&ms->full[i], kGatherThreads / 32);   // one arrival per gather warp (32 same-address arrivals serialise)
      mbar_init(&ms->bfree[i], 1);
      mbar_init(&ms->p_ready[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&ms->dfull[i], 1);
      mbar_init(&ms->actfull[i], 4 * NP);        // one arrival per epilogue warp of the context
      mbar_init(&ms->p_actfull[i], 1);
    }
    ms->gprog = 0;
    mbar_fence_init();
  }
  if (warp == kWarpMma) tmem_alloc<CG>(&ms->tmem_base, 512);
  for (int c = tid; c < C::NQ * 128; c += kThreads) {
    // channel of (block q, lane r) of this CTA: CG = 2 -> rank * 128 + r (one block spans both tiles' rows)
    const int q = c >> 7, r = c & 127;
    const int ch = (CG == 2) ? (int)rank * 128 + r : q * 128 + r;
    ms->bias3[c] = __ldg(p.bias3 + ch);
  }
  tc_fence_before_sync();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&ms->tmem_base);
#ifdef SAD_MLP_PROFILE
  if (blockIdx.x == 0 && tid == kWarpProd * 32) {
    g_sa_log[3][0] = 0;
    g_sa_log[3][1] = sa_t_entry;
    g_sa_logn[3] = sa_logn = 1;
  }
#endif
  if (warp == kWarpProd) { SALOG(3, 1) }

  // unit of ordinal o (tile / tile pair), -1 = no more work
  auto get_unit = [&](int o) -> int {
    if (static_sched) {
      const long long u = (long long)cluster_id + (long long)o * num_clusters;
      return u < num_units ? (int)u : -1;
    }
    bar_wait<CG>(&ms->tfull[o & (kRing - 1)], (uint32_t)((o / kRing) & 1));
    return *reinterpret_cast<volatile int*>(&ms->units[o & (kRing - 1)]);
  };
  auto tile_of = [&](int unit) -> long long { return (CG == 2) ? 2LL * unit + rank : (long long)unit; };

  if (warp == kWarpProd) {
    // ============================================================ tile scheduler + pinned weights
    if (lane == 0) {      // this CTA's weight image
      mbar_arrive_expect_tx(&ms->wfull, (uint32_t)C::WBYTES);
      const uint8_t* src = p.w_img + (size_t)rank * C::WBYTES;
      for (int o = 0; o < C::WBYTES; o += 32768) {
        const int n = C::WBYTES - o < 32768 ? C::WBYTES - o : 32768;
        tma_bulk_g2s(gbase + o, src + o, (uint32_t)n, &ms->wfull);
      }
    }
    auto publish = [&](int o, int unit) {      // lane 0 of the leader
      const int s = o & (kRing - 1);
      ms->units[s] = unit;
      mbar_arrive(&ms->tfull[s]);
      if constexpr (CG == 2) {
        st_cluster_b32(mapa(smem_u32(&ms->units[s]), 1), (uint32_t)unit);
        mbar_arrive_remote(&ms->tfull[s], 1);
      }
    };
    // Only lane 0 of the leader schedules.  Flow control is a monotonic counter, never a parity wait: the scheduler
    // may fall arbitrarily far behind the pipeline (a slow atomic), and a parity wait that is two phases late
    // deadlocks.  Every consumer, in both CTAs of a pair, has read ordinal x once the gather warps have passed the
    // "buffer free" wait of ordinal x + NB, so the ring slot of x may be rewritten (ordinal x + kRing) then.
    if (leader && lane == 0 && !static_sched) {
      constexpr int kEndMarks = G > 7 ? G : 7;      // -1 marks past the last unit (every role's look-ahead)
      int pub = 0, n_end = 0;
      bool ended = false;
      auto push = [&](int u) {
        while (pub > *reinterpret_cast<volatile int*>(&ms->gprog) + (kRing - 1 - NB)) __nanosleep(64);
        publish(pub++, u);
        if (u < 0) {
          ended = true;
          ++n_end;
        }
      };
      // The cluster's own first unit, then units claimed `claim` at a time from the global counter.  One atomic per
      // unit would cap the whole GPU at one tile per ~11 cycles (same-address atomics serialise in L2: 8192 tiles of
      // SA1 = 47 us) and every cluster at one unit per atomic round trip; two claims are kept in flight.
      int claim = p.claim;
      if (planned) {      // sized on the device like the problem itself
        const int fair = num_units / num_clusters / 6;
        claim = fair < 1 ? 1 : (fair > 8 ? 8 : fair);
      }
      int raw_a = atomicAdd(p.sched, claim), raw_b = atomicAdd(p.sched, claim);
      push(cluster_id < num_units ? cluster_id : -1);
      long long cur = (long long)num_clusters + raw_a;
      int rem = claim;
      while (n_end < kEndMarks) {
        if (!ended && rem == 0) {
          cur = (long long)num_clusters + raw_b;
          rem = claim;
          raw_b = atomicAdd(p.sched, claim);      // next round trip under the publications of this chunk
        }
        push((ended || cur >= num_units) ? -1 : (int)cur);
        ++cur;
        --rem;
      }
      const int raw_next = raw_b;
      // every atomic of this cluster has returned before the cluster reports "done" (the last cluster out re-zeroes
      // the scheduler words)
      asm volatile("" ::"r"(raw_next) : "memory");
    }
  } else if (warp == kWarpMma || warp == kWarpMma + 1) {
    // ============================================================ MMA issue (leader) / barrier relay (rank 1)
    // One warp per half of the contexts: the two halves are independent pipelines (each waits only for its own
    // contexts' operands and epilogues), so their wait -> fence -> issue sequences overlap and the tensor pipe is fed
    // from two instruction streams.
    const int half = warp - kWarpMma;
    const bool issuer = elect_one();
    // wait for a local barrier; the leader of a pair also waits for the peer's relayed copy, the peer relays
    // (The generic -> async proxy fence for operands written with st.shared / cp.async sits HERE, on the consumer
    // side: the waiting thread has observed the writes through the barrier and fences before it issues the MMAs that
    // read them -- or, in the peer CTA, before it relays the barrier.  A fence on the writer side costs every gather
    // thread a full L2 round trip per tile, because it also waits for the thread's prefetch loads in flight.)
    auto ready = [&](uint64_t* local, uint64_t* mirror, uint32_t parity) {
      bar_wait<CG>(local, parity);
      if constexpr (CG == 2) {
        if (leader) {
          bar_wait<CG>(mirror, parity);
        } else {
          fence_proxy_async_smem();
          if (issuer) mbar_arrive_remote(mirror, 0);
        }
      }
    };
    if (half == 0) ready(&ms->wfull, &ms->p_wfull, 0);
    else bar_wait<CG>(&ms->wfull, 0);              // (the pair's second copy is relayed once, by half 0)
    if constexpr (CG == 2) {
      if (half == 1 && leader) bar_wait<CG>(&ms->p_wfull, 0);
    }
    constexpr uint32_t idesc_h = uidesc_bf16(128 * CG, H);            // hidden layers: N = H
    constexpr uint32_t idesc_t = uidesc_bf16(128 * CG, 128 * CG);     // transposed last layer: N = rows of the unit
    const uint32_t w_base = base;
    const int c_lo = half * HG;
    int last_valid = -1;
    for (int round = 0;; ++round) {
      const int o0 = round * G;
      int nact = 0;                                 // contexts of this half with a unit in this round
#pragma unroll
      for (int cc = 0; cc < HG; ++cc)
        if (nact == cc && get_unit(o0 + c_lo + cc) >= 0) nact = cc + 1;
      if (nact == 0) break;
      last_valid = o0 + c_lo + nact - 1;
#pragma unroll
      for (int phase = 0; phase < 3; ++phase) {
        // one context per step (wait -> fence -> issue): the epilogue of a context runs under the step of the other
#pragma unroll
        for (int cc = 0; cc < HG; ++cc) {
          if (cc >= nact) break;
          const int c = c_lo + cc;
          const int o = o0 + c, b = o % NB;
          SALOG(2 + 3 * half, 10000 * (phase + 1) + 10 * round + cc)
          if (phase == 0) {
            if (round > 0) ready(&ms->actfull[c], &ms->p_actfull[c], (uint32_t)(3 * round - 1) & 1u);   // TMEM drained
            ready(&ms->full[b], &ms->p_ready[b], (uint32_t)(o / NB) & 1u);                              // layer-1 operand
          } else {
            ready(&ms->actfull[c], &ms->p_actfull[c], (uint32_t)(3 * round + phase - 1) & 1u);          // activations
          }
          fence_proxy_async_smem();
          tc_fence_after_sync();
          SALOG(2 + 3 * half, 10000 * (phase + 1) + 10 * round + cc + 2)
          if (leader && issuer) {
            const uint32_t buf = base + off_buf + (uint32_t)b * C::BUFBYTES;
            if (phase == 0) {
              // layer 1: gathered feature chunks + the special K step (relative xyz, scalar features, bias)
              const uint32_t d = tmem_base + (uint32_t)(c * C::CW);
#pragma unroll
              for (int kc = 0; kc < NF; ++kc) {
                const uint64_t ad = udesc_sw128(buf + kc * kChunk), bd = udesc_sw128(w_base + kc * (C::WROWS * 128));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16<CG>(d, ad + 2u * k, bd + 2u * k, idesc_h, (kc | k) ? 1u : 0u);
              }
              umma_f16<CG>(d, udesc_k16(base + off_sp + (uint32_t)b * kSpBytes), udesc_k16(w_base + C::OFF_W1SP), idesc_h,
                           NF > 0 ? 1u : 0u);
              umma_commit_to<CG>(&ms->dfull[c]);
            } else if (phase == 1) {
              // layer 2 (+ bias: the special K step again, against a piece that is zero except for the ones column)
              const uint32_t d = tmem_base + (uint32_t)(c * C::CW);
#pragma unroll
              for (int kc = 0; kc < C::HC; ++kc) {
                const uint64_t ad = udesc_sw128(buf + kc * kChunk), bd = udesc_sw128(w_base + C::OFF_W2 + kc * (C::WROWS * 128));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16<CG>(d, ad + 2u * k, bd + 2u * k, idesc_h, (kc | k) ? 1u : 0u);
              }
              umma_f16<CG>(d, udesc_k16(base + off_sp + (uint32_t)b * kSpBytes), udesc_k16(w_base + C::OFF_W2BIAS), idesc_h, 1u);
              umma_commit_to<CG>(&ms->dfull[c]);
            } else {
              // layer 3, transposed: D^T (channels x rows) = W3 . H^T
#pragma unroll
              for (int blk = 0; blk < C::NBLK; ++blk) {
                const uint32_t d = tmem_base + (uint32_t)(c * C::CW + blk * 128 * CG);
#pragma unroll
                for (int kc = 0; kc < C::HC; ++kc) {
                  const uint64_t ad = udesc_sw128(w_base + C::OFF_W3 + (blk * C::HC + kc) * kChunk), bd = udesc_sw128(buf + kc * kChunk);
#pragma unroll
                  for (int k = 0; k < 4; ++k) umma_f16<CG>(d, ad + 2u * k, bd + 2u * k, idesc_t, (kc | k) ? 1u : 0u);
                }
              }
              umma_commit_to<CG>(&ms->dfull[c]);
              umma_commit_to<CG>(&ms->bfree[b]);      // the tile buffer (and its special chunk) may be refilled
            }
          }
          __syncwarp();
          SALOG(2 + 3 * half, 10000 * (phase + 1) + 10 * round + cc + 4)
        }
      }
      if (nact < HG) break;
    }
    if (half == 0) {
      // tail: the last "buffer free" commits are multicast to both CTAs of a pair; each CTA sees its own copies
      // complete before it may leave (an arrival must never target the shared memory of a CTA that has exited).  This
      // warp has never waited on these barriers, and their final completion is the one asked for: no phase is skipped.
      int n_total = last_valid + 1;                  // valid ordinals form a prefix
      while (get_unit(n_total) >= 0) ++n_total;
      for (int b = 0; b < NB; ++b) {
        const int uses = (n_total - b + NB - 1) / NB;      // ordinals o < n_total with o % NB == b
        if (n_total > b && uses >= 1) bar_wait<CG>(&ms->bfree[b], (uint32_t)((uses - 1) & 1));
      }
    }
  } else if (warp >= kWarpGather) {
    // ============================================================ layer-1 operand: feature rows + special K step
    const int gt = tid - kWarpGather * 32;      // row of the tile this thread owns
    const int wrow0 = (gt >> 5) * 32;           // first row of this warp
    const int unit = lane & 7, rsub = lane >> 3;
    struct Sp {
      float x, y, z, qx, qy, qz, r, e[4];
    };
    auto issue_idx = [&](int unit_id) -> int {
      if (unit_id < 0) return 0;
      const long long R = tile_of(unit_id) * 128 + gt;
      if (R >= total_rows) return 0;
      if (planned) return ldg_nc_s32(p.idx + ((size_t)slot_addr((uint32_t)(R >> p.log2S)) << p.log2S) + (R & ((1 << p.log2S) - 1)));
      return ldg_nc_s32(p.idx + R);
    };
    // source row (b * N + idx) of this thread's row, kNoRow past the end; starts the loads of the special K step
    auto load_sp = [&](int unit_id, int raw, Sp& s) -> uint32_t {
      s.x = s.y = s.z = s.qx = s.qy = s.qz = 0.f;
      s.r = 1.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) s.e[e] = 0.f;
      if (unit_id < 0) return kNoRow;
      const long long R = tile_of(unit_id) * 128 + gt;
      if (R >= total_rows) return kNoRow;
      const uint32_t pt = (uint32_t)(R >> p.log2S);
      const uint32_t src = planned ? (uint32_t)raw : (pt >> p.log2P) * (uint32_t)p.N + (uint32_t)raw;
      const float* q = p.new_xyz + (size_t)pt * 3;
      if (p.xyzw) {
        const float4 v = ldg_nc_f32x4(p.xyzw + src);
        s.x = v.x;
        s.y = v.y;
        s.z = v.z;
        if (p.E > 0) s.e[0] = v.w;
      } else {
        const float* a = p.xyz + (size_t)src * 3;
        s.x = ldg_nc_f32(a);
        s.y = ldg_nc_f32(a + 1);
        s.z = ldg_nc_f32(a + 2);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (e < p.E) s.e[e] = ldg_nc_f32(p.extra + (size_t)src * p.E + e);
      }
      if (planned) {
        const float4 qq = ldg_nc_f32x4(p.q4 + slot_addr(pt));
        s.qx = qq.x;
        s.qy = qq.y;
        s.qz = qq.z;
        s.r = qq.w;
      } else {
        s.qx = ldg_nc_f32(q);
        s.qy = ldg_nc_f32(q + 1);
        s.qz = ldg_nc_f32(q + 2);
        if (p.radius_t) s.r = ldg_nc_f32(p.radius_t + pt);
      }
      return src;
    };
    auto publish = [&](int b) {                 // this warp's stores / copies of buffer b are complete
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->full[b]); // (release; the proxy fence is the consumer's, see the MMA warp)
    };
    // Software pipeline over tiles: two dependent L2 round trips per row (neighbour index, then xyz / features at
    // that index), each requested kDepth tiles before its result is needed -- the inputs of the special K step of
    // tiles o .. o + 2 sit in registers, the indices of tiles o + 3 .. o + 5 are in flight.  Nothing in the loop waits
    // for a load issued less than three iterations ago.
    constexpr int kDepth = 3;
    Sp sp[kDepth];
    uint32_t src[kDepth];
    int uq[kDepth];                             // units of tiles o .. o + 2 (slot = tile % 3)
    int ui[kDepth], idxq[kDepth];               // units / neighbour indices (in flight) of tiles o + 3 .. o + 5
    {
      int raw[kDepth];
#pragma unroll
      for (int k = 0; k < kDepth; ++k) {
        uq[k] = get_unit(k);
        raw[k] = issue_idx(uq[k]);
      }
#pragma unroll
      for (int k = 0; k < kDepth; ++k) {
        ui[k] = get_unit(kDepth + k);
        idxq[k] = issue_idx(ui[k]);
      }
#pragma unroll
      for (int k = 0; k < kDepth; ++k) src[k] = load_sp(uq[k], raw[k], sp[k]);
    }
    int pend_b = -1;                            // buffer whose cp.async group is still in flight (published one tile later)
    int o = 0;
    for (bool more = true; more;) {
#pragma unroll
      for (int k = 0; k < kDepth; ++k, ++o) {
        if (uq[k] < 0) {
          more = false;
          break;
        }
        const int b = o % NB;
        if (o >= NB) {
          const uint32_t par = (uint32_t)((o / NB - 1) & 1);
          if (!__all_sync(FULL, bar_try<CG>(&ms->bfree[b], par))) {
            // about to block: publish what is in flight first (the MMAs that free this buffer may be waiting for it)
            if (pend_b >= 0) {
              cp_async_wait<0>();
              publish(pend_b);
              pend_b = -1;
            }
            bar_wait<CG>(&ms->bfree[b], par);
          }
        }
        if (gt == 0) *reinterpret_cast<volatile int*>(&ms->gprog) = o;
        SALOG(1, 1000 + o)
        // ---- special K step: [dx, dy, dz, e0..e3, 0 x 8, 1]
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (src[k] != kNoRow) {
          float dx = __fsub_rn(sp[k].x, sp[k].qx), dy = __fsub_rn(sp[k].y, sp[k].qy), dz = __fsub_rn(sp[k].z, sp[k].qz);
          if (p.normalize) {
            const float inv = p.radius_t ? __frcp_rn(sp[k].r) : p.inv_radius;
            dx *= inv;
            dy *= inv;
            dz *= inv;
          }
          v[0] = dx;
          v[1] = dy;
          v[2] = dz;
#pragma unroll
          for (int e = 0; e < 4; ++e) v[3 + e] = sp[k].e[e];
        }
        const uint32_t dsp = base + off_sp + (uint32_t)b * kSpBytes;
        sts_v4(dsp + k16_off(gt, 0), bf16x2_rn(v[0], v[1]), bf16x2_rn(v[2], v[3]), bf16x2_rn(v[4], v[5]), bf16x2_rn(v[6], v[7]));
        sts_v4(dsp + k16_off(gt, 1), 0u, 0u, 0u, 0x3F800000u);   // K index 15 = 1.0 (bf16, high half): the bias column
        SALOG(1, 2000 + o)
        if constexpr (NF > 0) {
          // ---- feature rows: this warp's 32 rows, 8 lanes per row (one 128-byte line per row and chunk)
          const uint32_t dbuf = base + off_buf + (uint32_t)b * C::BUFBYTES;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t sj = __shfl_sync(FULL, src[k], 4 * j + rsub);
            const int row = wrow0 + 4 * j + rsub;
            const __nv_bfloat16* g = p.feat_cl + (size_t)(sj == kNoRow ? 0u : sj) * (NF * 64) + unit * 8;
            const uint32_t nbytes = sj == kNoRow ? 0u : 16u;
            const uint32_t d = dbuf + swz128(row, unit);
#pragma unroll
            for (int kc = 0; kc < NF; ++kc) cp_async16(d + kc * kChunk, g + kc * 64, nbytes);
          }
          cp_async_commit();
          if (pend_b >= 0) {
            cp_async_wait<1>();                   // the previous tile's group has landed
            publish(pend_b);
          }
          pend_b = b;
        } else {
          publish(b);
        }
        SALOG(1, 100 + o)
        // refill this slot with tile o + 3 (its index landed long ago), request the index of tile o + 6
        uq[k] = ui[k];
        src[k] = load_sp(ui[k], idxq[k], sp[k]);
        SALOG(1, 3000 + o)
        ui[k] = get_unit(o + 2 * kDepth);
        SALOG(1, 4000 + o)
        idxq[k] = issue_idx(ui[k]);
      }
    }
    if (pend_b >= 0) {
      cp_async_wait<0>();
      publish(pend_b);
    }
  } else {
    // ============================================================ epilogue warpgroups
    const int wg = warp >> 2;
    const int c = wg % G, part = wg / G;
    const int row = (warp & 3) * 32 + lane;                         // TMEM lane
    const uint32_t tctx = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c * C::CW);
    auto done_phase = [&]() {
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->actfull[c]);
    };
    // hidden layer: TMEM -> ReLU -> bf16 -> swizzled activation chunk (= the next layer's operand); the bias is
    // already in the accumulator.  16 columns at a time, the next load in flight under the conversion.
    auto hidden = [&](uint32_t buf) {
      constexpr int NG = C::HP / 16;
      const int c0 = part * C::HP;
      uint32_t v[2][16];
      tmem_ld_x16(tctx + (uint32_t)c0, v[0]);
      tmem_ld_fence();
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if (g + 1 < NG) tmem_ld_x16(tctx + (uint32_t)(c0 + (g + 1) * 16), v[(g + 1) & 1]);
        const uint32_t* w = v[g & 1];
        const int col = c0 + g * 16;
        const uint32_t chunk = buf + (uint32_t)(col >> 6) * kChunk;
        const int u0 = (col & 63) >> 3;
#pragma unroll
        for (int u = 0; u < 2; ++u)
          sts_v4(chunk + swz128(row, u0 + u),
                 bf16x2_relu(__uint_as_float(w[u * 8 + 0]), __uint_as_float(w[u * 8 + 1])),
                 bf16x2_relu(__uint_as_float(w[u * 8 + 2]), __uint_as_float(w[u * 8 + 3])),
                 bf16x2_relu(__uint_as_float(w[u * 8 + 4]), __uint_as_float(w[u * 8 + 5])),
                 bf16x2_relu(__uint_as_float(w[u * 8 + 6]), __uint_as_float(w[u * 8 + 7])));
        if (g + 1 < NG) tmem_ld_fence();
      }
    };
    for (int t = 0;; ++t) {
      const int o = t * G + c;
      const int unit = get_unit(o);
      if (unit < 0) break;
      const uint32_t buf = base + off_buf + (uint32_t)(o % NB) * C::BUFBYTES;
      bar_wait<CG>(&ms->dfull[c], (uint32_t)(3 * t) & 1u);
      tc_fence_after_sync();
      if (warp == 0 || warp == 4) { SALOG(warp, 1000 + 10 * t) }
      hidden(buf);
      done_phase();
      if (warp == 0 || warp == 4) { SALOG(warp, 1500 + 10 * t) }
      bar_wait<CG>(&ms->dfull[c], (uint32_t)(3 * t + 1) & 1u);
      tc_fence_after_sync();
      if (warp == 0 || warp == 4) { SALOG(warp, 2000 + 10 * t) }
      hidden(buf);
      done_phase();
      if (warp == 0 || warp == 4) { SALOG(warp, 2500 + 10 * t) }
      // planned mode: the tile's point ids, requested under the wait for the last layer's MMAs
      int id[C::PTS];
      if (planned) {
        const uint32_t s0 = (uint32_t)(tile_of(unit) * C::PTS);
#pragma unroll
        for (int i = 0; i < C::PTS; ++i) id[i] = s0 + i < total_points ? __ldg(p.pid + slot_addr(s0 + i)) : -1;
      }
      bar_wait<CG>(&ms->dfull[c], (uint32_t)(3 * t + 2) & 1u);
      tc_fence_after_sync();
      if (warp == 0 || warp == 4) { SALOG(warp, 3000 + 10 * t) }
      // ---- last layer: lane = output channel, columns = rows of a tile; max over each point's S rows
#pragma unroll
      for (int qi = 0; qi < C::NQP; ++qi) {
        const int q = part * C::NQP + qi;
        const long long tile = (CG == 2) ? 2LL * unit + q : (long long)unit;      // whose rows these columns are
        const int ch = (CG == 2) ? (int)rank * 128 + row : q * 128 + row;
        const float bias = ms->bias3[q * 128 + row];
        float y[C::PTS];
        uint32_t v[2][16];
        tmem_ld_x16(tctx + (uint32_t)(q * 128), v[0]);
        tmem_ld_fence();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (g + 1 < 8) tmem_ld_x16(tctx + (uint32_t)(q * 128 + (g + 1) * 16), v[(g + 1) & 1]);
          if constexpr (C::S == 8) {                   // two points per 16-column group
            y[2 * g] = vmax_tree<8>(v[g & 1]);
            y[2 * g + 1] = vmax_tree<8>(v[g & 1] + 8);
          } else {
            const float m = vmax_tree<16>(v[g & 1]);
            constexpr int GP = C::S / 16;               // 16-column groups per point
            if (g % GP == 0) y[g / GP] = m;
            else y[g / GP] = fmaxf(y[g / GP], m);
          }
          if (g + 1 < 8) tmem_ld_fence();
        }
        const uint32_t pt0 = (uint32_t)(tile * C::PTS);
        if (planned) {
          // slots -> points: consecutive slots with one pid are one point (runs never straddle a tile); the run's last
          // slot carries the combined maximum.  The plan keeps neighbouring points in neighbouring slots, so the usual
          // tile is 8 consecutive points (1-slot runs) or 4 (2-slot runs): those store float4s like the plain path.
          if (ch < p.c3_real && pt0 < total_points) {
#pragma unroll
            for (int i = 1; i < C::PTS; ++i)
              if (id[i] == id[i - 1]) y[i] = fmaxf(y[i], y[i - 1]);
#pragma unroll
            for (int i = 0; i < C::PTS; ++i) y[i] = fmaxf(y[i] + bias, 0.f);
            bool seq = id[0] >= 0 && (id[0] & 3) == 0, pair = seq && C::PTS >= 8;
#pragma unroll
            for (int i = 1; i < C::PTS; ++i) {
              seq = seq && id[i] == id[0] + i;
              pair = pair && id[i] == id[0] + (i >> 1);
            }
            seq = seq && (uint32_t)(id[0] & (p.P - 1)) + (uint32_t)C::PTS <= (uint32_t)p.P;
            pair = pair && (uint32_t)(id[0] & (p.P - 1)) + (uint32_t)(C::PTS / 2) <= (uint32_t)p.P;
            if (seq || pair) {
              const uint32_t rp = (uint32_t)id[0], bb = rp >> p.log2P, jj = rp - (bb << p.log2P);
              if (p.out_cf) {
                float* o_cf = p.out_cf + ((size_t)bb * p.c3_real + ch) * p.P + jj;
                if (seq) {
#pragma unroll
                  for (int i = 0; i < C::PTS; i += 4) *reinterpret_cast<float4*>(o_cf + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
                } else {
                  if constexpr (C::PTS >= 8) {
#pragma unroll
                    for (int i = 0; i < C::PTS / 2; i += 4)
                      *reinterpret_cast<float4*>(o_cf + i) = make_float4(y[2 * i + 1], y[2 * i + 3], y[2 * i + 5], y[2 * i + 7]);
                  }
                }
              }
              if (p.out_cl) {
                if (seq) {
#pragma unroll
                  for (int i = 0; i < C::PTS; ++i) p.out_cl[(size_t)(rp + i) * p.c3_real + ch] = __float2bfloat16_rn(y[i]);
                } else {
#pragma unroll
                  for (int i = 0; i < C::PTS / 2; ++i) p.out_cl[(size_t)(rp + i) * p.c3_real + ch] = __float2bfloat16_rn(y[2 * i + 1]);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < C::PTS; ++i) {
                const bool last = i + 1 == C::PTS || id[i + 1] != id[i];
                if (last && id[i] >= 0) {
                  const uint32_t rp = (uint32_t)id[i], bb = rp >> p.log2P, jj = rp - (bb << p.log2P);
                  if (p.out_cf) p.out_cf[((size_t)bb * p.c3_real + ch) * p.P + jj] = y[i];
                  if (p.out_cl) p.out_cl[(size_t)rp * p.c3_real + ch] = __float2bfloat16_rn(y[i]);
                }
              }
            }
          }
        } else if (ch < p.c3_real && pt0 < total_points) {
#pragma unroll
          for (int i = 0; i < C::PTS; ++i) y[i] = fmaxf(y[i] + bias, 0.f);
          const uint32_t b0 = pt0 >> p.log2P, j0 = pt0 - (b0 << p.log2P);
          const bool whole = pt0 + C::PTS <= total_points && j0 + C::PTS <= (uint32_t)p.P;
          if (p.out_cf) {
            float* o_cf = p.out_cf + ((size_t)b0 * p.c3_real + ch) * p.P + j0;
            if (whole) {
              if constexpr (C::PTS == 2) {
                *reinterpret_cast<float2*>(o_cf) = make_float2(y[0], y[1]);
              } else {
#pragma unroll
                for (int i = 0; i < C::PTS; i += 4) *reinterpret_cast<float4*>(o_cf + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < C::PTS; ++i) {
                const uint32_t pt = pt0 + i;
                if (pt < total_points) {
                  const uint32_t bb = pt >> p.log2P, jj = pt - (bb << p.log2P);
                  p.out_cf[((size_t)bb * p.c3_real + ch) * p.P + jj] = y[i];
                }
              }
            }
          }
          if (p.out_cl) {
#pragma unroll
            for (int i = 0; i < C::PTS; ++i)
              if (pt0 + i < total_points) p.out_cl[(size_t)(pt0 + i) * p.c3_real + ch] = __float2bfloat16_rn(y[i]);
          }
        }
      }
      done_phase();
      if (warp == 0 || warp == 4) { SALOG(warp, 3500 + 10 * t) }
    }
  }

  if (warp == kWarpProd) { SALOG(3, 9000) }
  tc_fence_before_sync();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == kWarpProd) { SALOG(3, 9001) }
  if (warp == kWarpMma) tmem_dealloc<CG>(tmem_base, 512);
  if (tid == 0 && leader) {
    // last cluster out re-arms the scheduler words for the next launch that uses them
    __threadfence();
    if (atomicAdd(p.sched + 1, 1) == num_clusters - 1) {
      p.sched[0] = 0;
      p.sched[1] = 0;
      if (planned) {
        int* c = const_cast<int*>(p.plan_counts);
        c[0] = c[1] = c[2] = 0;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
struct Instance {
  int id, CG, NF, H, C3, S, wbytes, smem;
  void (*launch)(const SaParams&, int grid, cudaStream_t);
  cudaError_t (*configure)();
};

template <class C>
void launch_inst(const SaParams& p, int grid, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C::CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, sa_mlp_kernel<C>, p);
}
template <class C>
cudaError_t configure_inst() {
  return cudaFuncSetAttribute(sa_mlp_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
}
template <class C>
Instance make_inst(int id) {
  return Instance{id, C::CG, C::NF, C::H, C::C3, C::S, C::WBYTES, C::SMEM, &launch_inst<C>, &configure_inst<C>};
}

using CfgSA1 = Cfg<1, 0, 64, 128, 64>;      // SA1: xyz + 1 scalar feature -> 64 -> 64 -> 128, nsample 64
using CfgSA2 = Cfg<2, 2, 128, 256, 32>;     // SA2: 128 features + xyz -> 128 -> 128 -> 256, nsample 32 (CTA pair)
using CfgSA3 = Cfg<2, 4, 128, 256, 16>;     // SA3 / SA4 / vote aggregation (128 outputs zero-padded), nsample 16 (CTA pair)
using CfgSA2s = Cfg<1, 2, 128, 256, 32>;    // SA2 on one CTA (no cluster): bring-up / comparison
using CfgSA1d = Cfg<1, 0, 64, 128, 16>;     // SA1 / SA2 on 16-sample slots: the duplicate-free mode (sad_sa_mlp_dedup_fwd)
using CfgSA2d = Cfg<1, 2, 128, 256, 16>;
// (An 8-sample-slot sibling <1, 2, 128, 256, 8> was built and measured for SA2: 61 % of the 16-slot rows, but 16 output
// points per tile make its epilogue slower than the rows it saves -- 41 us against 38 us -- so it is not compiled in.)

const Instance* instances(int* n) {
  // order = preference among instances that fit the same stage (SA2: the single-CTA instance measures 51 us against
  // 58 us for the pair at 8 x 1024 x 32 rows -- the relayed barriers cost more than the halved weight footprint buys)
  static const Instance tab[] = {make_inst<CfgSA1>(0), make_inst<CfgSA2s>(3), make_inst<CfgSA2>(1), make_inst<CfgSA3>(2),
                                 make_inst<CfgSA1d>(4), make_inst<CfgSA2d>(5)};
  *n = 6;
  return tab;
}
const Instance* instance_by_id(int id) {
  int n;
  const Instance* t = instances(&n);
  for (int i = 0; i < n; ++i)
    if (t[i].id == id) return &t[i];
  return nullptr;
}

uint16_t bf16_bits(float w) {
  uint32_t u;
  memcpy(&u, &w, 4);
  return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);      // RNE (finite inputs)
}
void put_sw128(uint8_t* piece, int r, int kk, float w) {           // element (row r, K index kk < 64) of a SWIZZLE_128B piece
  const int unit = kk >> 3;
  const size_t byte = (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + (size_t)((unit ^ (r & 7)) << 4) + (size_t)(kk & 7) * 2;
  const uint16_t b = bf16_bits(w);
  memcpy(piece + byte, &b, 2);
}
void put_k16(uint8_t* piece, int r, int kk, float w) {             // element (row r, K index kk < 16) of the no-swizzle K step
  const size_t byte = (size_t)(r >> 3) * 256 + (size_t)(kk >> 3) * 128 + (size_t)(r & 7) * 16 + (size_t)(kk & 7) * 2;
  const uint16_t b = bf16_bits(w);
  memcpy(piece + byte, &b, 2);
}

}  // namespace

#ifdef SAD_MLP_PROFILE
extern "C" SAD_API int sad_sa_debug_buffer(void* host_mapped_64_words) {
  return (int)cudaMemcpyToSymbol(g_sa_dbg, &host_mapped_64_words, sizeof(void*));
}
extern "C" SAD_API int sad_sa_profile_dump(long long* host_log /*6*8192*/, int* host_n /*6*/) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host_log, g_sa_log, sizeof(long long) * 6 * 8192);
  cudaMemcpyFromSymbol(host_n, g_sa_logn, sizeof(int) * 6);
  int z[6] = {0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_sa_logn, z, sizeof(z));
  return 0;
}
#endif

// Instance that runs (C0 gathered channels, hidden widths h1 == h2, c3 outputs, nsample S, E scalar features), or -1.
// `prefer`: 0 = the library's choice, 1 = an instance that runs on single CTAs, 2 = on CTA pairs, where one exists
// (tools / tests).
extern "C" int sad_sa_mlp_query(int C0, int h1, int h2, int c3, int S, int E, int has_xyz, int prefer) {
  if (C0 % 64 || h1 != h2 || E < 0 || E > 4 || !has_xyz || c3 < 8 || c3 % 8) return -1;
  int n, first = -1;
  const Instance* t = instances(&n);
  for (int i = 0; i < n; ++i) {
    if (t[i].NF != C0 / 64 || t[i].H != h1 || t[i].S != S || c3 > t[i].C3) continue;
    if (prefer == 0 || (prefer == 1) == (t[i].CG == 1)) return t[i].id;      // the table is in order of preference
    if (first < 0) first = t[i].id;
  }
  return first;
}

extern "C" int sad_sa_mlp_instance_info(int instance, int* out5) {
  const Instance* in = instance_by_id(instance);
  SAD_REQUIRE(in && out5, "sa_mlp_instance_info: unknown instance %d", instance);
  out5[0] = in->CG; out5[1] = in->NF; out5[2] = in->H; out5[3] = in->C3; out5[4] = in->S;
  return SAD_OK;
}

extern "C" long long sad_sa_mlp_image_bytes(int instance) {
  const Instance* in = instance_by_id(instance);
  return in ? (long long)in->CG * in->wbytes : -1;
}

// Pack the stage's weights into the per-CTA shared-memory images (bf16).  W1 (h x cin1) with
//   perm_feat[C0]: source column of gathered-feature K index k (-1 = zero), perm_sp[7]: source columns of
//   [dx, dy, dz, e0..e3] (-1 = zero); b1 and b2 go into constant-1 K columns.  W2 (h x h), W3 (c3 x h).
extern "C" int sad_sa_mlp_pack(int instance, const float* W1, int cin1, const int32_t* perm_feat, const int32_t* perm_sp,
                               const float* b1, const float* W2, const float* b2, const float* W3, int c3, void* out_image) {
  const Instance* in = instance_by_id(instance);
  SAD_REQUIRE(in, "sa_mlp_pack: unknown instance %d", instance);
  SAD_REQUIRE(W1 && perm_sp && b1 && W2 && b2 && W3 && out_image && (in->NF == 0 || perm_feat), "sa_mlp_pack: null pointer");
  SAD_REQUIRE(c3 >= 1 && c3 <= in->C3, "sa_mlp_pack: c3 out of range");
  const int CG = in->CG, NF = in->NF, H = in->H, HC = H / 64, WROWS = H / CG, NBLK = in->C3 / (128 * CG);
  uint8_t* img = static_cast<uint8_t*>(out_image);
  memset(img, 0, (size_t)CG * in->wbytes);
  const int off_w1sp = NF * WROWS * 128, off_w2 = off_w1sp + WROWS * 32, off_w2b = off_w2 + HC * WROWS * 128,
            off_w3 = off_w2b + WROWS * 32;
  for (int r = 0; r < CG; ++r) {
    uint8_t* im = img + (size_t)r * in->wbytes;
    for (int nl = 0; nl < WROWS; ++nl) {
      const int n = r * WROWS + nl;
      for (int k = 0; k < NF * 64; ++k) {
        const int src = perm_feat[k];
        if (src < 0) continue;
        SAD_REQUIRE(src < cin1, "sa_mlp_pack: perm_feat[%d]=%d out of range", k, src);
        put_sw128(im + (size_t)(k >> 6) * WROWS * 128, nl, k & 63, W1[(size_t)n * cin1 + src]);
      }
      for (int k = 0; k < 7; ++k) {
        const int src = perm_sp[k];
        if (src < 0) continue;
        SAD_REQUIRE(src < cin1, "sa_mlp_pack: perm_sp[%d]=%d out of range", k, src);
        put_k16(im + off_w1sp, nl, k, W1[(size_t)n * cin1 + src]);
      }
      put_k16(im + off_w1sp, nl, 15, b1[n]);
      for (int k = 0; k < H; ++k) put_sw128(im + off_w2 + (size_t)(k >> 6) * WROWS * 128, nl, k & 63, W2[(size_t)n * H + k]);
      put_k16(im + off_w2b, nl, 15, b2[n]);
    }
    for (int blk = 0; blk < NBLK; ++blk)
      for (int row = 0; row < 128; ++row) {
        const int ch = (blk * CG + r) * 128 + row;
        if (ch >= c3) continue;
        for (int k = 0; k < H; ++k)
          put_sw128(im + off_w3 + (size_t)(blk * HC + (k >> 6)) * kChunk, row, k & 63, W3[(size_t)ch * H + k]);
      }
  }
  return SAD_OK;
}

// (B,N,3) coordinates (+ one scalar feature per point) -> (B,N) float4 {x, y, z, feature | 0}: the gathered source of the
// special K step as one aligned 16-byte row
__global__ void __launch_bounds__(256) pack_xyzw_kernel(long long rows, const float* __restrict__ xyz,
                                                        const float* __restrict__ extra, float4* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows) return;
  out[i] = make_float4(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2), extra ? __ldg(extra + i) : 0.f);
}

extern "C" int sad_pack_xyzw(int B, int N, const float* xyz, const float* extra1, void* out_xyzw, sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && N >= 1, "pack_xyzw: bad sizes");
  if (B == 0) return SAD_OK;
  SAD_REQUIRE(xyz && out_xyzw && (reinterpret_cast<uintptr_t>(out_xyzw) & 15) == 0, "pack_xyzw: null / misaligned pointer");
  const long long rows = (long long)B * N;
  pack_xyzw_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, xyz, extra1, static_cast<float4*>(out_xyzw));
  SAD_LAUNCH_CHECK("pack_xyzw_kernel");
  return SAD_OK;
}

// ---- duplicate-free mode: the plan.  A ball query pads a neighbourhood that has fewer than nsample hits with copies of
// its first hit, and max-pooling ignores copies, so only the leading samples need to go through the MLP.  A point
// becomes a RUN of r slots of 16 consecutive samples, r the smallest of {1, 2, 4} (nsample 64) / {1, 2} (nsample 32) such
// that every sample from 16 r on equals sample 0 (checked, not assumed: any idx is handled exactly).  The three run
// classes are compacted into three regions (4 / 2 / 1 x points slots): per slot 16 GLOBAL source rows, the query point
// {x, y, z, radius} and the point id.  One warp per point, 32 points per block, one atomic per class and block.
constexpr int kPlanWarps = 32;
__global__ void __launch_bounds__(kPlanWarps * 32)
sa_plan_kernel(uint32_t total_points, int N, int log2P, int S, int SL, const int32_t* __restrict__ idx,
               const float* __restrict__ new_xyz, const float* __restrict__ radius_t, int* __restrict__ counts,
               int32_t* __restrict__ idxc, float4* __restrict__ q4, int32_t* __restrict__ pid) {
  __shared__ int s_cls[kPlanWarps], s_base[3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t pt = blockIdx.x * kPlanWarps + warp;
  int r = 0, cls = 3;
  int v0 = 0, v1 = 0;
  if (pt < total_points) {
    const int32_t* row = idx + (size_t)pt * S;
    v0 = lane < S ? __ldg(row + lane) : 0;
    v1 = S > 32 ? __ldg(row + 32 + lane) : 0;
    const int first = __shfl_sync(0xFFFFFFFFu, v0, 0);
    const unsigned e0 = __ballot_sync(0xFFFFFFFFu, v0 == first || lane >= S);
    const unsigned e1 = S > 32 ? __ballot_sync(0xFFFFFFFFu, v1 == first) : 0xFFFFFFFFu;
    // padding from sample SL on -> 1 slot, from 2 SL on -> 2 slots, else 4 (nsample / SL is 2 or 4)
    const unsigned long long eq = ((unsigned long long)e1 << 32) | e0;      // bit j: sample j equals sample 0 (or is past nsample)
    const bool from1 = (eq >> SL) == (~0ull >> SL), from2 = (eq >> (2 * SL)) == (~0ull >> (2 * SL));
    r = from1 ? 1 : ((from2 || S <= 2 * SL) ? 2 : 4);
    cls = r == 4 ? 0 : (r == 2 ? 1 : 2);
  }
  if (lane == 0) s_cls[warp] = cls;
  __syncthreads();
  // rank of this point among the block's points of its class, in point order: the 32 consecutive points of a block
  // stay consecutive inside their class, which is what lets the MLP kernel store float4s of neighbouring points
  const int other = s_cls[lane];
  const unsigned same = __ballot_sync(0xFFFFFFFFu, other == cls);
  const int rank = __popc(same & ((1u << warp) - 1u));
  if (warp == 0) {
    const unsigned m0 = __ballot_sync(0xFFFFFFFFu, other == 0), m1 = __ballot_sync(0xFFFFFFFFu, other == 1),
                   m2 = __ballot_sync(0xFFFFFFFFu, other == 2);
    if (lane < 3) {
      const int c = __popc(lane == 0 ? m0 : (lane == 1 ? m1 : m2));
      s_base[lane] = c ? atomicAdd(counts + lane, c) : 0;
    }
  }
  __syncthreads();
  if (pt >= total_points) return;
  const uint32_t region = cls == 0 ? 0u : (cls == 1 ? 4u * total_points : 6u * total_points);
  const uint32_t slot0 = region + (uint32_t)(s_base[cls] + rank) * (uint32_t)r;
  const int src_base = (int)(pt >> log2P) * N;
  // samples 0 .. 16 r - 1 as global source rows; out-of-range neighbours are clamped like every other gather
  auto fix = [&](int v) { return src_base + (int)min((unsigned)v, (unsigned)(N - 1)); };
  if (lane < SL * r) idxc[(size_t)slot0 * SL + lane] = fix(v0);
  if (32 + lane < SL * r) idxc[(size_t)slot0 * SL + 32 + lane] = fix(v1);
  if (lane < r) {
    const float* q = new_xyz + (size_t)pt * 3;
    q4[slot0 + lane] = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), radius_t ? __ldg(radius_t + pt) : 1.f);
    pid[slot0 + lane] = (int32_t)pt;
  }
}

extern "C" long long sad_sa_mlp_dedup_workspace_bytes(int B, int P) {
  if (B < 0 || P < 0) return SAD_EINVAL;
  return (long long)B * P * 7 * (16 * 4 + 16 + 4) + 256;      // 7 slots per point at most: idx rows, q4, pid
}

namespace {
struct PlanArgs {
  void* workspace;      // null = plain launch
  int s_full;           // nsample of idx
};
int sa_mlp_launch(int instance, int B, int N, int P, const void* feat_cl, const float* xyz, const void* xyzw,
                  const float* new_xyz, const int32_t* idx, float radius, const float* radius_t, int normalize_xyz,
                  const float* extra, int E, const void* w_image, const float* bias3_padded, int c3, void* out_cl_bf16,
                  float* out_cf_f32, int* sched, int tiles_per_cta, cudaStream_t stream, PlanArgs plan);
}  // namespace

extern "C" int sad_sa_mlp_fwd(int instance, int B, int N, int P, const void* feat_cl, const float* xyz, const void* xyzw,
                              const float* new_xyz, const int32_t* idx, float radius, const float* radius_t,
                              int normalize_xyz, const float* extra, int E, const void* w_image, const float* bias3_padded, int c3, void* out_cl_bf16,
                              float* out_cf_f32, int* sched, int tiles_per_cta, sad_stream_t stream_) {
  return sa_mlp_launch(instance, B, N, P, feat_cl, xyz, xyzw, new_xyz, idx, radius, radius_t, normalize_xyz, extra, E, w_image,
                       bias3_padded, c3, out_cl_bf16, out_cf_f32, sched, tiles_per_cta, (cudaStream_t)stream_, PlanArgs{nullptr, 0});
}

// Same stage, duplicate-free: `instance` is the stage's ordinary instance (nsample 32 or 64, single-CTA); the launch runs
// its sibling with 8- or 16-sample slots (slot_samples; 0 = the smallest available) over the plan's slots.  workspace: sad_sa_mlp_dedup_workspace_bytes(B, P) bytes, 16-byte
// aligned.  sched: 32 zero-initialised ints that the kernel re-zeroes ([0..1] scheduler, [16..18] the plan's counters).
// Results are bit-identical to sad_sa_mlp_fwd.
extern "C" int sad_sa_mlp_dedup_fwd(int instance, int B, int N, int P, const void* feat_cl, const float* xyz, const void* xyzw,
                                    const float* new_xyz, const int32_t* idx, float radius, const float* radius_t,
                                    int normalize_xyz, const float* extra, int E, const void* w_image, const float* bias3_padded,
                                    int c3, void* out_cl_bf16, float* out_cf_f32, int* sched, void* workspace,
                                    int slot_samples, int tiles_per_cta, sad_stream_t stream_) {
  const Instance* full = instance_by_id(instance);
  SAD_REQUIRE(full, "sa_mlp_dedup: unknown instance %d", instance);
  SAD_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "sa_mlp_dedup: workspace null / misaligned");
  int n;
  const Instance* t = instances(&n);
  const Instance* sib = nullptr;
  for (int i = 0; i < n; ++i)
    if (t[i].CG == 1 && t[i].S * 4 >= full->S && t[i].S * 2 <= full->S && t[i].NF == full->NF && t[i].H == full->H &&
        t[i].C3 == full->C3 && (slot_samples ? t[i].S == slot_samples : (!sib || t[i].S < sib->S)))
      sib = &t[i];      // the sibling with the smallest slot whose runs of 1 / 2 / 4 slots still cover nsample
  if (!sib || full->CG != 1 || (full->S != 32 && full->S != 64)) {
    sad_set_error("sa_mlp_dedup: instance %d has no sibling with %d-sample slots", instance, slot_samples);
    return SAD_EUNSUPPORTED;
  }
  return sa_mlp_launch(sib->id, B, N, P, feat_cl, xyz, xyzw, new_xyz, idx, radius, radius_t, normalize_xyz, extra, E, w_image,
                       bias3_padded, c3, out_cl_bf16, out_cf_f32, sched, tiles_per_cta, (cudaStream_t)stream_,
                       PlanArgs{workspace, full->S});
}

namespace {
int sa_mlp_launch(int instance, int B, int N, int P, const void* feat_cl, const float* xyz, const void* xyzw,
                  const float* new_xyz, const int32_t* idx, float radius, const float* radius_t, int normalize_xyz,
                  const float* extra, int E, const void* w_image, const float* bias3_padded, int c3, void* out_cl_bf16,
                  float* out_cf_f32, int* sched, int tiles_per_cta, cudaStream_t stream, PlanArgs plan) {
  const Instance* in = instance_by_id(instance);
  SAD_REQUIRE(in, "sa_mlp: unknown instance %d", instance);
  SAD_REQUIRE(B >= 0 && N >= 1 && P >= 0, "sa_mlp: bad sizes B=%d N=%d P=%d", B, N, P);
  SAD_REQUIRE(xyz && new_xyz && idx && w_image && bias3_padded && sched, "sa_mlp: null pointer");
  SAD_REQUIRE(!xyzw || (E <= 1 && (reinterpret_cast<uintptr_t>(xyzw) & 15) == 0), "sa_mlp: xyzw needs E <= 1 and 16-byte alignment");
  SAD_REQUIRE((in->NF == 0) == (feat_cl == nullptr), "sa_mlp: gathered source / instance mismatch");
  SAD_REQUIRE((reinterpret_cast<uintptr_t>(feat_cl) & 15) == 0, "sa_mlp: feat_cl must be 16-byte aligned");
  SAD_REQUIRE(E >= 0 && E <= 4 && (E == 0 || extra), "sa_mlp: 0..4 scalar features");
  SAD_REQUIRE(c3 >= 1 && c3 <= in->C3 && c3 % 8 == 0, "sa_mlp: bad output width %d", c3);
  SAD_REQUIRE(out_cl_bf16 || out_cf_f32, "sa_mlp: no output requested");
  SAD_REQUIRE((long long)B * N < 0x7FFFFFFFLL && (long long)B * P * in->S < (1LL << 37), "sa_mlp: problem too large");
  if (B == 0 || P == 0) return SAD_OK;
  int log2P = -1;
  for (int k = 0; k < 31; ++k)
    if ((1 << k) == P) log2P = k;
  SAD_REQUIRE(log2P >= 0, "sa_mlp: npoint must be a power of two (got %d)", P);

  SaParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.P = P; p.log2P = log2P;
  for (p.log2S = 0; (1 << p.log2S) < in->S; ++p.log2S) {}
  const int s_idx = plan.workspace ? plan.s_full : in->S;      // nsample of the idx tensor
  p.total_rows = (long long)B * P * s_idx;                      // (planned mode: the upper bound; the kernel reads the counts)
  p.total_points = (uint32_t)((long long)B * P);
  const long long tiles = (p.total_rows + 127) / 128;
  SAD_REQUIRE(tiles < 0x3FFFFFFFLL, "sa_mlp: too many rows");
  p.num_tiles = (int)tiles;
  p.num_units = (int)((tiles + in->CG - 1) / in->CG);
  p.feat_cl = static_cast<const __nv_bfloat16*>(feat_cl);
  p.xyz = xyz; p.xyzw = static_cast<const float4*>(xyzw); p.new_xyz = new_xyz; p.idx = idx; p.radius_t = radius_t;
  p.normalize = normalize_xyz;
  p.inv_radius = (normalize_xyz && !radius_t) ? 1.0f / radius : 1.0f;
  p.extra = extra; p.E = E;
  p.w_img = static_cast<const uint8_t*>(w_image);
  p.bias3 = bias3_padded; p.c3_real = c3;
  p.out_cl = static_cast<__nv_bfloat16*>(out_cl_bf16); p.out_cf = out_cf_f32;
  p.sched = sched;
  if (plan.workspace) {
    // workspace: idx rows (7 * points slots x 16) | q4 | pid
    const size_t slots = (size_t)B * P * 7;
    int32_t* idxc = static_cast<int32_t*>(plan.workspace);
    float4* q4 = reinterpret_cast<float4*>(idxc + slots * 16);      // (sized for 16-sample slots; 8-sample slots use half)
    int32_t* pid = reinterpret_cast<int32_t*>(q4 + slots);
    int* counts = sched + 16;
    sa_plan_kernel<<<(unsigned)((p.total_points + kPlanWarps - 1) / kPlanWarps), kPlanWarps * 32, 0, stream>>>(
        p.total_points, N, log2P, plan.s_full, in->S, idx, new_xyz, (normalize_xyz && radius_t) ? radius_t : nullptr, counts, idxc, q4, pid);
    SAD_LAUNCH_CHECK("sa_plan_kernel");
    p.idx = idxc;
    p.q4 = q4;
    p.pid = pid;
    p.plan_counts = counts;
    p.plan_points = p.total_points;
  }

  int dev = 0, sms = 0;
  SAD_CUDA_OK(cudaGetDevice(&dev));
  SAD_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static thread_local unsigned configured = 0;       // bit per instance
  static thread_local int configured_dev = -1;
  if (configured_dev != dev) {
    configured = 0;
    configured_dev = dev;
  }
  if (!(configured & (1u << in->id))) {
    SAD_CUDA_OK(in->configure());
    configured |= 1u << in->id;
  }
  const int tpc = tiles_per_cta < 1 ? 1 : tiles_per_cta;
  int clusters = sad_ceil_div(p.num_units, tpc);
  const int max_clusters = sms / in->CG;
  if (clusters > max_clusters) clusters = max_clusters;
  // Few units per cluster: a fixed assignment (no scheduler round trips, and no cluster that claims units ahead of
  // need while others idle).  Many: units are handed out dynamically, a few ahead of use.
  p.static_sched = (!plan.workspace && (long long)clusters * 8 >= p.num_units) ? 1 : 0;
  {
    const long long fair = p.num_units / clusters;      // units per cluster if all ran equally fast
    p.claim = (int)(fair / 6 < 1 ? 1 : (fair / 6 > 8 ? 8 : fair / 6));
  }
  in->launch(p, clusters * in->CG, stream);
  SAD_LAUNCH_CHECK("sa_mlp_kernel");
  return SAD_OK;
}
}  // namespace
