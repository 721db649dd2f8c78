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
//     gathered rows are dead once layer 1 has been issued): channel-last bf16 feature rows land by TMA gather4
//     (one instruction = 4 rows x 128 B, swizzled by the TMA unit, completion on an mbarrier -- no thread ever touches
//     them), the relative xyz / scalar features / constant-1 (bias) column form one 16-wide K step written by the
//     gather warps in the no-swizzle K-major layout (4 KB per tile instead of a 16 KB chunk).
//   * G = 2 or 4 tile contexts in TMEM; the MMA warp walks them in lockstep (L1 of every context, then L2, then L3),
//     so the epilogue of context c runs under the MMAs of the others.
//   * 16 epilogue warps (4 per scheduler): one warpgroup per context, or two splitting its columns.
//   * cluster rank 1 has no MMAs to issue: its MMA warp RELAYS the local "operand ready" barriers to the leader with
//     one remote arrive each, in exactly the order the leader consumes them; tcgen05.commit multicasts the
//     "accumulator ready" / "buffer free" barriers to both CTAs.
#include <cuda.h>
#include <string.h>

#include "sad_tc.cuh"

namespace {

using namespace sad;

constexpr int kWarpGather = 16;          // warps 16-19: special chunk (relative xyz, scalar features)
constexpr int kWarpMma = 20;             // leader: MMA issue; rank 1: relay
constexpr int kWarpProd = 21;            // tile scheduler + TMA gather
constexpr int kThreads = 22 * 32;
constexpr int kChunk = 16384;            // 128 rows x 128 B
constexpr int kSpBytes = 4096;           // 128 rows x 32 B (one 16-wide bf16 K step)
constexpr int kRing = 16;                // tile-id ring
constexpr int kAhead = 4;                // tile ids published ahead of the load cursor

template <int CG_, int NF_, int H_, int C3_, int S_>
struct Cfg {
  static constexpr int CG = CG_, NF = NF_, H = H_, C3 = C3_, S = S_;
  static constexpr int HC = H / 64;                      // 64-wide chunks of a hidden activation
  static constexpr int WROWS = H / CG;                   // hidden-layer weight rows held by one CTA
  static constexpr int NQ = C3 / 128;                    // 128-column blocks of the transposed last layer per CTA
  static constexpr int NBLK = C3 / (128 * CG);           // MMA groups of the last layer (M = 128 * CG channels each)
  static constexpr int CW = (H > C3 ? H : C3) <= 128 ? 128 : 256;   // TMEM columns of one context
  static constexpr int G = 512 / CW;                     // contexts
  static constexpr int NP = 4 / G;                       // epilogue warpgroups per context
  static constexpr int HP = H / NP;                      // hidden columns per warpgroup
  static constexpr int NQP = NQ / NP;                    // last-layer blocks per warpgroup
  static constexpr int BUFCH = NF > HC ? NF : HC;        // 16 KB chunks of a tile buffer
  static constexpr int W1F = NF * WROWS * 128, W1SP = WROWS * 32, W2B = HC * WROWS * 128, W3B = NBLK * HC * kChunk;
  static constexpr int OFF_W1SP = W1F, OFF_W2 = W1F + W1SP, OFF_W3 = OFF_W2 + W2B;
  static constexpr int WBYTES = OFF_W3 + W3B;            // weight image of one CTA
  static constexpr int BUFBYTES = BUFCH * kChunk;
  static constexpr int kMisc = 4096;
  static constexpr int kBudget = 227 * 1024 - 1024 - kMisc - WBYTES;
  static constexpr int NB_fit = kBudget / (BUFBYTES + kSpBytes);
  static constexpr int NB = NB_fit > 8 ? 8 : NB_fit;     // tile buffers
  static constexpr int SMEM = 1024 + WBYTES + NB * (BUFBYTES + kSpBytes) + kMisc;
  static constexpr int PTS = 128 / S;                    // points per tile
  static_assert(H == 64 || H == 128, "hidden width");
  static_assert(C3 % (128 * CG) == 0 && C3 <= 256, "output width");
  static_assert(HP % 64 == 0 && NQ % NP == 0 && NQP >= 1, "epilogue split");
  static_assert(NB >= G, "not enough tile buffers for the contexts");
  static_assert(S == 16 || S == 32 || S == 64, "nsample");
  static_assert(WBYTES % 1024 == 0, "weight image alignment");
};

struct alignas(64) SaParams {
  CUtensorMap tmap;               // (B*N rows) x C0 bf16, box {64, 1}, SWIZZLE_128B (NF > 0)
  int N, P, log2P, log2S;
  long long total_rows;
  uint32_t total_points;
  int num_tiles, num_units;       // unit = tile (CG = 1) or tile pair (CG = 2)
  const float* xyz;               // (B,N,3)
  const float* new_xyz;           // (B,P,3)
  const int32_t* idx;             // (B,P,S)
  const float* radius_t;          // (B,P) or null
  float inv_radius;               // scalar radius: 1/r (or 1 when not normalising)
  int normalize;
  const float* extra;             // (B,N,E) fp32 scalar features, E <= 4
  int E;
  const uint8_t* w_img;           // CG images of WBYTES
  const float* bias2;             // (H)
  const float* bias3;             // (C3), zero padded
  int c3_real;                    // channels actually stored
  __nv_bfloat16* out_cl;          // (B,P,c3_real) bf16 or null
  float* out_cf;                  // (B,c3_real,P) f32 or null
  int* sched;                     // [0] next unit (zero between launches: the kernel resets it), [1] clusters done
};

template <class C>
struct Misc {
  uint64_t wfull, p_wfull;
  uint64_t tfull[kRing];
  uint64_t bfull[8], sfull[8], bfree[8], p_ready[8];
  uint64_t dfull[4], actfull[4], p_actfull[4];
  int units[kRing];
  uint32_t tmem_base, pad_;
  alignas(16) float bias2[128];
  alignas(16) float bias3[256];
};

#ifdef SAD_MLP_PROFILE
// tools only: (event, clock) log of one warp per role of CTA 0 (role: 0 epilogue warp 0, 1 gather warp, 2 MMA warp,
// 3 producer warp, 4 epilogue warp 4)
__device__ long long g_sa_log[5][2 * 4096];
__device__ int g_sa_logn[5];
#define SALOG(role, ev)                                                         \
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && sa_logn < 4095) {           \
    g_sa_log[role][2 * sa_logn] = (ev);                                         \
    g_sa_log[role][2 * sa_logn + 1] = clock64();                                \
    g_sa_logn[role] = ++sa_logn;                                                \
  }
#else
#define SALOG(role, ev)
#endif

// mbarrier wait that traps instead of hanging the GPU when a protocol bug leaves it unsatisfied (~2 s)
__device__ __noinline__ void bar_timeout(uint32_t bar_addr, uint32_t parity) {
#ifdef SAD_MLP_DEBUG
  printf("[sad] sa_mlp: barrier timeout (block %d thread %d bar +%u parity %u)\n", (int)blockIdx.x, (int)threadIdx.x,
         bar_addr & 0xFFFFu, parity);
#endif
  __trap();
}
template <int CG>
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  long long t0 = 0;
  for (;;) {
    const bool ok = (CG == 2) ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity);
    if (ok) return;
    if ((++spins & 0xFFFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) bar_timeout(smem_u32(bar), parity);
    }
  }
}

template <class C>
__global__ void __launch_bounds__(kThreads, 1) sa_mlp_kernel(const __grid_constant__ SaParams p) {
#ifdef SAD_MLP_PROFILE
  const long long sa_t_entry = clock64();
#endif
  constexpr int CG = C::CG, NF = C::NF, H = C::H, G = C::G, NB = C::NB, NP = C::NP;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  constexpr uint32_t off_buf = C::WBYTES;
  constexpr uint32_t off_sp = off_buf + NB * C::BUFBYTES;
  constexpr uint32_t off_misc = off_sp + NB * kSpBytes;
  Misc<C>* ms = reinterpret_cast<Misc<C>*>(gbase + off_misc);
  static_assert(sizeof(Misc<C>) <= C::kMisc, "misc area");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef SAD_MLP_PROFILE
  int sa_logn = 0;
#endif
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int cluster_id = (int)blockIdx.x / CG, num_clusters = (int)gridDim.x / CG;
  const bool leader = (rank == 0);

  if (tid == 0) {
    mbar_init(&ms->wfull, 1);
    mbar_init(&ms->p_wfull, 1);
    for (int i = 0; i < kRing; ++i) mbar_init(&ms->tfull[i], 1);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&ms->bfull[i], 1);
      mbar_init(&ms->sfull[i], 4);          // one arrival per gather warp
      mbar_init(&ms->bfree[i], 1);
      mbar_init(&ms->p_ready[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&ms->dfull[i], 1);
      mbar_init(&ms->actfull[i], 4 * NP);   // one arrival per epilogue warp of the context
      mbar_init(&ms->p_actfull[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == kWarpMma) tmem_alloc<CG>(&ms->tmem_base, 512);
  for (int c = tid; c < H; c += kThreads) ms->bias2[c] = __ldg(p.bias2 + c);
  for (int c = tid; c < C::NQ * 128; c += kThreads) {
    // channel of (block q, lane r) of this CTA: CG = 2 -> rank * 128 + r (one block spans both tiles' rows)
    const int q = c >> 7, r = c & 127;
    const int ch = (CG == 2) ? (int)rank * 128 + r : q * 128 + r;
    ms->bias3[c] = __ldg(p.bias3 + ch);
  }
  if (warp == kWarpProd && lane == 0) {
    if constexpr (NF > 0) tma_prefetch_desc(&p.tmap);
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
    bar_wait<CG>(&ms->tfull[o & (kRing - 1)], (uint32_t)((o >> 4) & 1));
    return *reinterpret_cast<volatile int*>(&ms->units[o & (kRing - 1)]);
  };
  auto tile_of = [&](int unit) -> long long { return (CG == 2) ? 2LL * unit + rank : (long long)unit; };

  if (warp == kWarpProd) {
    // ============================================================ tile scheduler + weight / gather TMA
    if (lane == 0) {      // pinned weights: this CTA's image
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
    auto unit_of_raw = [&](int raw) -> int {
      const long long u = (long long)num_clusters + raw;
      return u < p.num_units ? (int)u : -1;
    };
    int pub = 0;            // next ordinal to publish (leader, lane 0)
    bool ended = false;
    int raw_next = 0;
    if (leader && lane == 0) {
      // ordinals 0..kAhead: the cluster's own first unit, then independent atomics (their round trips overlap)
      int raw[kAhead + 1];
#pragma unroll
      for (int i = 0; i <= kAhead; ++i) raw[i] = atomicAdd(p.sched, 1);
      publish(0, cluster_id < p.num_units ? cluster_id : -1);
      ended = !(cluster_id < p.num_units);
      pub = 1;
#pragma unroll
      for (int i = 0; i < kAhead; ++i) {
        const int u = ended ? -1 : unit_of_raw(raw[i]);
        publish(pub++, u);
        if (u < 0) ended = true;
      }
      raw_next = raw[kAhead];
    }
    int L = 0;
    for (;; ++L) {
      if (leader && lane == 0) {
        // ordinal L + kAhead + 1 from the atomic issued one iteration ago; the next atomic goes out now
        const int u = ended ? -1 : unit_of_raw(raw_next);
        publish(pub++, u);
        if (u < 0) ended = true;
        if (!ended) raw_next = atomicAdd(p.sched, 1);
      }
      const int unit = get_unit(L);
      if (unit < 0) break;
      const int b = L % NB;
      SALOG(3, 100 + L)
      if (L >= NB) bar_wait<CG>(&ms->bfree[b], (uint32_t)((L / NB - 1) & 1));
      SALOG(3, 200 + L)
      if constexpr (NF > 0) {
        const long long R0 = tile_of(unit) * 128 + 4 * lane;
        int4 id = make_int4(0, 0, 0, 0);
        if (R0 + 3 < p.total_rows) {
          id = ldg_nc_s32x4(p.idx + R0);
        } else {
          if (R0 + 0 < p.total_rows) id.x = __ldg(p.idx + R0);
          if (R0 + 1 < p.total_rows) id.y = __ldg(p.idx + R0 + 1);
          if (R0 + 2 < p.total_rows) id.z = __ldg(p.idx + R0 + 2);
        }
        const long long Rc = R0 < p.total_rows ? R0 : 0;
        const int rowbase = (int)((uint32_t)(Rc >> p.log2S) >> p.log2P) * p.N;
        if (lane == 0) mbar_arrive_expect_tx(&ms->bfull[b], (uint32_t)(NF * kChunk));
        __syncwarp();
        const uint32_t dst = base + off_buf + (uint32_t)b * C::BUFBYTES + (uint32_t)lane * 512u;
#pragma unroll
        for (int kc = 0; kc < NF; ++kc)
          tma_gather4(dst + kc * kChunk, &p.tmap, kc * 64, rowbase + id.x, rowbase + id.y, rowbase + id.z, rowbase + id.w,
                      &ms->bfull[b]);
      }
    }
    // tail: the last "buffer free" commits are multicast to both CTAs of a pair; each CTA sees its own copies complete
    // before it may leave (an arrival must never target the shared memory of a CTA that has exited)
    for (int b = 0; b < NB; ++b) {
      const int uses = (L - b + NB - 1) / NB;          // ordinals o < L with o % NB == b
      if (L > b && uses >= 1) bar_wait<CG>(&ms->bfree[b], (uint32_t)((uses - 1) & 1));
    }
  } else if (warp == kWarpMma) {
    // ============================================================ MMA issue (leader) / barrier relay (rank 1)
    const bool issuer = elect_one();
    // wait for a local barrier; the leader of a pair also waits for the peer's relayed copy, the peer relays
    auto ready = [&](uint64_t* local, uint64_t* mirror, uint32_t parity) {
      bar_wait<CG>(local, parity);
      if constexpr (CG == 2) {
        if (leader) bar_wait<CG>(mirror, parity);
        else if (issuer) mbar_arrive_remote(mirror, 0);
      }
    };
    ready(&ms->wfull, &ms->p_wfull, 0);
    constexpr uint32_t idesc_h = uidesc_bf16(128 * CG, H);            // hidden layers: N = H
    constexpr uint32_t idesc_t = uidesc_bf16(128 * CG, 128 * CG);     // transposed last layer: N = rows of the unit
    const uint32_t w_base = base;
    for (int round = 0;; ++round) {
      const int o0 = round * G;
      int nact = 0;
#pragma unroll
      for (int c = 0; c < G; ++c)
        if (nact == c && get_unit(o0 + c) >= 0) nact = c + 1;
      if (nact == 0) break;
      // ---- layer 1: gathered feature chunks + the special K step
#pragma unroll
      for (int c = 0; c < G; ++c) {
        if (c >= nact) break;
        const int o = o0 + c, b = o % NB;
        const uint32_t use = (uint32_t)(o / NB) & 1u;
        if (round > 0) ready(&ms->actfull[c], &ms->p_actfull[c], (uint32_t)(3 * round - 1) & 1u);   // TMEM drained
        bar_wait<CG>(&ms->sfull[b], use);
        if constexpr (NF > 0) bar_wait<CG>(&ms->bfull[b], use);
        if constexpr (CG == 2) {
          if (leader) bar_wait<CG>(&ms->p_ready[b], use);
          else if (issuer) mbar_arrive_remote(&ms->p_ready[b], 0);
        }
        tc_fence_after_sync();
        SALOG(2, 1000 + 10 * round + c)
        if (leader && issuer) {
          const uint32_t d = tmem_base + (uint32_t)(c * C::CW);
          const uint32_t buf = base + off_buf + (uint32_t)b * C::BUFBYTES;
#pragma unroll
          for (int kc = 0; kc < NF; ++kc) {
            const uint64_t ad = udesc_sw128(buf + kc * kChunk), bd = udesc_sw128(w_base + kc * (C::WROWS * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16<CG>(d, ad + 2u * k, bd + 2u * k, idesc_h, (kc | k) ? 1u : 0u);
          }
          umma_f16<CG>(d, udesc_k16(base + off_sp + (uint32_t)b * kSpBytes), udesc_k16(w_base + C::OFF_W1SP), idesc_h,
                       NF > 0 ? 1u : 0u);
          umma_commit_to<CG>(&ms->dfull[c]);
        }
        __syncwarp();
      }
      // ---- layer 2
#pragma unroll
      for (int c = 0; c < G; ++c) {
        if (c >= nact) break;
        const int o = o0 + c, b = o % NB;
        ready(&ms->actfull[c], &ms->p_actfull[c], (uint32_t)(3 * round) & 1u);
        tc_fence_after_sync();
        SALOG(2, 2000 + 10 * round + c)
        if (leader && issuer) {
          const uint32_t d = tmem_base + (uint32_t)(c * C::CW);
          const uint32_t buf = base + off_buf + (uint32_t)b * C::BUFBYTES;
#pragma unroll
          for (int kc = 0; kc < C::HC; ++kc) {
            const uint64_t ad = udesc_sw128(buf + kc * kChunk), bd = udesc_sw128(w_base + C::OFF_W2 + kc * (C::WROWS * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16<CG>(d, ad + 2u * k, bd + 2u * k, idesc_h, (kc | k) ? 1u : 0u);
          }
          umma_commit_to<CG>(&ms->dfull[c]);
        }
        __syncwarp();
      }
      // ---- layer 3, transposed: D^T (channels x rows) = W3 . H^T
#pragma unroll
      for (int c = 0; c < G; ++c) {
        if (c >= nact) break;
        const int o = o0 + c, b = o % NB;
        SALOG(2, 30000 + 10 * round + c)
        ready(&ms->actfull[c], &ms->p_actfull[c], (uint32_t)(3 * round + 1) & 1u);
        SALOG(2, 31000 + 10 * round + c)
        tc_fence_after_sync();
        SALOG(2, 3000 + 10 * round + c)
        if (leader && issuer) {
          const uint32_t buf = base + off_buf + (uint32_t)b * C::BUFBYTES;
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
          SALOG(2, 32000 + 10 * round + c)
          umma_commit_to<CG>(&ms->dfull[c]);
          umma_commit_to<CG>(&ms->bfree[b]);      // the tile buffer (and its special chunk) may be refilled
        }
        SALOG(2, 33000 + 10 * round + c)
        __syncwarp();
        SALOG(2, 3500 + 10 * round + c)
      }
      if (nact < G) break;
    }
  } else if (warp >= kWarpGather) {
    // ============================================================ special K step: relative xyz, scalar features, 1
    const int gt = tid - kWarpGather * 32;      // row of the tile
    struct Sp {
      float x, y, z, qx, qy, qz, r, e[4];
    };
    auto issue_idx = [&](int unit) -> int {
      if (unit < 0) return 0;
      const long long R = tile_of(unit) * 128 + gt;
      return R < p.total_rows ? ldg_nc_s32(p.idx + R) : 0;
    };
    auto load_sp = [&](int unit, int raw, Sp& s, bool& valid) {
      s.x = s.y = s.z = s.qx = s.qy = s.qz = 0.f;
      s.r = 1.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) s.e[e] = 0.f;
      valid = false;
      if (unit < 0) return;
      const long long R = tile_of(unit) * 128 + gt;
      if (R >= p.total_rows) return;
      valid = true;
      const uint32_t pt = (uint32_t)(R >> p.log2S);
      const size_t src = (size_t)(pt >> p.log2P) * p.N + (uint32_t)raw;
      const float* a = p.xyz + src * 3;
      const float* q = p.new_xyz + (size_t)pt * 3;
      s.x = ldg_nc_f32(a);
      s.y = ldg_nc_f32(a + 1);
      s.z = ldg_nc_f32(a + 2);
      s.qx = ldg_nc_f32(q);
      s.qy = ldg_nc_f32(q + 1);
      s.qz = ldg_nc_f32(q + 2);
      if (p.radius_t) s.r = ldg_nc_f32(p.radius_t + pt);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (e < p.E) s.e[e] = ldg_nc_f32(p.extra + src * p.E + e);
    };
    int u_cur = get_unit(0), u_nxt = get_unit(1);
    Sp sp;
    bool valid;
    load_sp(u_cur, issue_idx(u_cur), sp, valid);
    int idraw = issue_idx(u_nxt);
    for (int o = 0; u_cur >= 0; ++o) {
      const int b = o % NB;
      if (o >= NB) bar_wait<CG>(&ms->bfree[b], (uint32_t)((o / NB - 1) & 1));
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
      if (valid) {
        float dx = __fsub_rn(sp.x, sp.qx), dy = __fsub_rn(sp.y, sp.qy), dz = __fsub_rn(sp.z, sp.qz);
        if (p.normalize) {
          const float inv = p.radius_t ? __frcp_rn(sp.r) : p.inv_radius;
          dx *= inv;
          dy *= inv;
          dz *= inv;
        }
        v[0] = dx;
        v[1] = dy;
        v[2] = dz;
#pragma unroll
        for (int e = 0; e < 4; ++e) v[3 + e] = sp.e[e];
      }
      const uint32_t dst = base + off_sp + (uint32_t)b * kSpBytes;
      sts_v4(dst + k16_off(gt, 0), bf16x2_rn(v[0], v[1]), bf16x2_rn(v[2], v[3]), bf16x2_rn(v[4], v[5]), bf16x2_rn(v[6], v[7]));
      sts_v4(dst + k16_off(gt, 1), 0u, 0u, 0u, 0x3F800000u);   // K index 15 = 1.0 (bf16, high half): the bias column
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->sfull[b]);
      SALOG(1, 100 + o)
      // loads for the next tiles, issued after this tile's publication (each has an iteration to land)
      const int u_n2 = get_unit(o + 2);
      load_sp(u_nxt, idraw, sp, valid);
      idraw = issue_idx(u_n2);
      u_cur = u_nxt;
      u_nxt = u_n2;
    }
  } else {
    // ============================================================ epilogue warpgroups
    const int wg = warp >> 2;
    const int c = wg % G, part = wg / G;
    const int row = (warp & 3) * 32 + lane;                         // TMEM lane
    const uint32_t tctx = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c * C::CW);
    auto done_phase = [&]() {
      if (warp == 0) { SALOG(0, 70000) }
      tc_fence_before_sync();
      if (warp == 0) { SALOG(0, 71000) }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->actfull[c]);
    };
    // hidden layer: TMEM -> (+bias) -> ReLU -> bf16 -> swizzled activation chunk (= the next layer's operand)
    auto hidden = [&](uint32_t buf, const float* bias) {
#pragma unroll
      for (int i = 0; i < C::HP / 64; ++i) {
        const int c0 = part * C::HP + i * 64;
        const uint32_t chunk = buf + (uint32_t)(c0 >> 6) * kChunk;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_x32(tctx + (uint32_t)(c0 + h * 32), v);
          tmem_ld_fence();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[u * 8 + k]);
            if (bias) {
              const float4 ba = *reinterpret_cast<const float4*>(bias + c0 + h * 32 + u * 8);
              const float4 bb = *reinterpret_cast<const float4*>(bias + c0 + h * 32 + u * 8 + 4);
              f[0] += ba.x; f[1] += ba.y; f[2] += ba.z; f[3] += ba.w;
              f[4] += bb.x; f[5] += bb.y; f[6] += bb.z; f[7] += bb.w;
            }
            sts_v4(chunk + swz128(row, h * 4 + u), bf16x2_relu(f[0], f[1]), bf16x2_relu(f[2], f[3]), bf16x2_relu(f[4], f[5]),
                   bf16x2_relu(f[6], f[7]));
          }
        }
      }
      if (warp == 0) { SALOG(0, 60000) }
      fence_proxy_async_smem();
    };
    for (int t = 0;; ++t) {
      const int o = t * G + c;
      const int unit = get_unit(o);
      if (unit < 0) break;
      const uint32_t buf = base + off_buf + (uint32_t)(o % NB) * C::BUFBYTES;
      bar_wait<CG>(&ms->dfull[c], (uint32_t)(3 * t) & 1u);
      if (warp == 0) { SALOG(0, 50000 + 10 * t) }
      tc_fence_after_sync();
      if (warp == 0 || warp == 4) { SALOG(warp, 1000 + 10 * t) }
      hidden(buf, nullptr);                     // layer-1 bias rides on the constant-1 column of the special K step
      done_phase();
      if (warp == 0 || warp == 4) { SALOG(warp, 1500 + 10 * t) }
      bar_wait<CG>(&ms->dfull[c], (uint32_t)(3 * t + 1) & 1u);
      tc_fence_after_sync();
      if (warp == 0 || warp == 4) { SALOG(warp, 2000 + 10 * t) }
      hidden(buf, ms->bias2);
      done_phase();
      if (warp == 0 || warp == 4) { SALOG(warp, 2500 + 10 * t) }
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
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t v[32];
          tmem_ld_x32(tctx + (uint32_t)(q * 128 + g * 32), v);
          tmem_ld_fence();
          if constexpr (C::S == 64) {
            const float m = vmax_tree<32>(v);
            if (g & 1) y[g >> 1] = fmaxf(y[g >> 1], m);
            else y[g >> 1] = m;
          } else if constexpr (C::S == 32) {
            y[g] = vmax_tree<32>(v);
          } else {
            y[2 * g] = vmax_tree<16>(v);
            y[2 * g + 1] = vmax_tree<16>(v + 16);
          }
        }
        const uint32_t pt0 = (uint32_t)(tile * C::PTS);
        if (ch < p.c3_real && pt0 < p.total_points) {
#pragma unroll
          for (int i = 0; i < C::PTS; ++i) y[i] = fmaxf(y[i] + bias, 0.f);
          const uint32_t b0 = pt0 >> p.log2P, j0 = pt0 - (b0 << p.log2P);
          const bool whole = pt0 + C::PTS <= p.total_points && j0 + C::PTS <= (uint32_t)p.P;
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
                if (pt < p.total_points) {
                  const uint32_t bb = pt >> p.log2P, jj = pt - (bb << p.log2P);
                  p.out_cf[((size_t)bb * p.c3_real + ch) * p.P + jj] = y[i];
                }
              }
            }
          }
          if (p.out_cl) {
#pragma unroll
            for (int i = 0; i < C::PTS; ++i)
              if (pt0 + i < p.total_points) p.out_cl[(size_t)(pt0 + i) * p.c3_real + ch] = __float2bfloat16_rn(y[i]);
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
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = []() -> EncodeFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeFn>(f);
  }();
  return fn;
}

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

const Instance* instances(int* n) {
  static const Instance tab[] = {make_inst<CfgSA1>(0), make_inst<CfgSA2>(1), make_inst<CfgSA3>(2), make_inst<CfgSA2s>(3)};
  *n = 4;
  return tab;
}
const Instance* instance_by_id(int id) {
  int n;
  const Instance* t = instances(&n);
  return (id >= 0 && id < n) ? &t[id] : nullptr;
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
extern "C" SAD_API int sad_sa_profile_dump(long long* host_log /*5*8192*/, int* host_n /*5*/) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host_log, g_sa_log, sizeof(long long) * 5 * 8192);
  cudaMemcpyFromSymbol(host_n, g_sa_logn, sizeof(int) * 5);
  int z[5] = {0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_sa_logn, z, sizeof(z));
  return 0;
}
#endif

// Instance that runs (NF gathered chunks, hidden widths h1 == h2, c3 outputs, nsample S, E scalar features), or -1.
// `single_cta` != 0 asks for a non-cluster instance where one exists (tools / tests).
extern "C" int sad_sa_mlp_query(int C0, int h1, int h2, int c3, int S, int E, int has_xyz, int single_cta) {
  if (C0 % 64 || h1 != h2 || E < 0 || E > 4 || !has_xyz || c3 < 8 || c3 % 8) return -1;
  int n, fallback = -1;
  const Instance* t = instances(&n);
  for (int i = 0; i < n; ++i) {
    if (t[i].NF != C0 / 64 || t[i].H != h1 || t[i].S != S || c3 > t[i].C3) continue;
    if ((t[i].CG == 1) == (single_cta != 0)) return t[i].id;
    if (fallback < 0) fallback = t[i].id;
  }
  return fallback;
}

extern "C" long long sad_sa_mlp_image_bytes(int instance) {
  const Instance* in = instance_by_id(instance);
  return in ? (long long)in->CG * in->wbytes : -1;
}

// Pack the stage's weights into the per-CTA shared-memory images (bf16).  W1 (h x cin1) with
//   perm_feat[NF*64]: source column of gathered-feature K index k (-1 = zero), perm_sp[7]: source columns of
//   [dx, dy, dz, e0..e3] (-1 = zero); b1 goes into the constant-1 column.  W2 (h x h), W3 (c3 x h).
extern "C" int sad_sa_mlp_pack(int instance, const float* W1, int cin1, const int32_t* perm_feat, const int32_t* perm_sp,
                               const float* b1, const float* W2, const float* W3, int c3, void* out_image) {
  const Instance* in = instance_by_id(instance);
  SAD_REQUIRE(in, "sa_mlp_pack: unknown instance %d", instance);
  SAD_REQUIRE(W1 && perm_sp && b1 && W2 && W3 && out_image && (in->NF == 0 || perm_feat), "sa_mlp_pack: null pointer");
  SAD_REQUIRE(c3 >= 1 && c3 <= in->C3, "sa_mlp_pack: c3 out of range");
  const int CG = in->CG, NF = in->NF, H = in->H, HC = H / 64, WROWS = H / CG, NBLK = in->C3 / (128 * CG);
  uint8_t* img = static_cast<uint8_t*>(out_image);
  memset(img, 0, (size_t)CG * in->wbytes);
  const int off_w1sp = NF * WROWS * 128, off_w2 = off_w1sp + WROWS * 32, off_w3 = off_w2 + HC * WROWS * 128;
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

// tools / tests: grid width override is not needed here; the scheduling hint below mirrors sad_mlp_set_tiles_per_cta
extern "C" int sad_sa_mlp_fwd(int instance, int B, int N, int P, const void* feat_cl, const float* xyz, const float* new_xyz,
                              const int32_t* idx, float radius, const float* radius_t, int normalize_xyz, const float* extra,
                              int E, const void* w_image, const float* bias2, const float* bias3_padded, int c3,
                              void* out_cl_bf16, float* out_cf_f32, int* sched, int tiles_per_cta, sad_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const Instance* in = instance_by_id(instance);
  SAD_REQUIRE(in, "sa_mlp: unknown instance %d", instance);
  SAD_REQUIRE(B >= 0 && N >= 1 && P >= 0, "sa_mlp: bad sizes B=%d N=%d P=%d", B, N, P);
  SAD_REQUIRE(xyz && new_xyz && idx && w_image && bias2 && bias3_padded && sched, "sa_mlp: null pointer");
  SAD_REQUIRE((in->NF == 0) == (feat_cl == nullptr), "sa_mlp: gathered source / instance mismatch");
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
  p.total_rows = (long long)B * P * in->S;
  p.total_points = (uint32_t)((long long)B * P);
  const long long tiles = (p.total_rows + 127) / 128;
  SAD_REQUIRE(tiles < 0x3FFFFFFFLL, "sa_mlp: too many rows");
  p.num_tiles = (int)tiles;
  p.num_units = (int)((tiles + in->CG - 1) / in->CG);
  p.xyz = xyz; p.new_xyz = new_xyz; p.idx = idx; p.radius_t = radius_t;
  p.normalize = normalize_xyz;
  p.inv_radius = (normalize_xyz && !radius_t) ? 1.0f / radius : 1.0f;
  p.extra = extra; p.E = E;
  p.w_img = static_cast<const uint8_t*>(w_image);
  p.bias2 = bias2; p.bias3 = bias3_padded; p.c3_real = c3;
  p.out_cl = static_cast<__nv_bfloat16*>(out_cl_bf16); p.out_cf = out_cf_f32;
  p.sched = sched;
  if (in->NF > 0) {
    EncodeFn enc = encode_fn();
    SAD_REQUIRE(enc, "sa_mlp: cuTensorMapEncodeTiled unavailable");
    SAD_REQUIRE((reinterpret_cast<uintptr_t>(feat_cl) & 15) == 0, "sa_mlp: feat_cl must be 16-byte aligned");
    const cuuint64_t dims[2] = {(cuuint64_t)(in->NF * 64), (cuuint64_t)((long long)B * N)};
    const cuuint64_t strides[1] = {(cuuint64_t)(in->NF * 64 * 2)};
    const cuuint32_t box[2] = {64, 1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(feat_cl), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SAD_REQUIRE(r == CUDA_SUCCESS, "sa_mlp: cuTensorMapEncodeTiled failed (%d)", (int)r);
  }

  int dev = 0, sms = 0;
  SAD_CUDA_OK(cudaGetDevice(&dev));
  SAD_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static thread_local unsigned configured = 0;       // bit per instance (per host thread; devices share the attribute call)
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
  in->launch(p, clusters * in->CG, stream);
  SAD_LAUNCH_CHECK("sa_mlp_kernel");
  return SAD_OK;
}
