// a6  shared point-wise MLP (+ max-pool over nsample), fused with the neighbourhood gather
// -- SURVEY.md section 8(a) rows a5/a6, section 8(f) rank 1, hard parts H4/H5.
// (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
// One launch per SA / FP / voting stage, persistent CTAs (one per SM), 128-row tiles
// (row = (query point, sample)), warp-specialised and software-pipelined:
//
//   warps 8-11 GATHER   build the layer-1 operand of tile t+1, t+2, ... straight in shared memory
//                       in the tcgen05 K-major SWIZZLE_128B layout while tile t is still being
//                       computed: channel-last bf16 feature rows fetched by `idx` with 16-byte
//                       cp.async through an A-ring of 16 KB stages (several groups in flight per
//                       thread; the grouped tensor never exists in HBM); the relative,
//                       radius-normalised xyz (+ fp32 scalar features) form one extra 16-wide K
//                       step whose dependent loads (idx -> xyz) are prefetched one tile ahead.
//   warp 14    WEIGHTS  pre-swizzled bf16 weight images by TMA bulk copy (cp.async.bulk): as many
//                       pieces as fit stay PINNED in shared memory for the whole kernel, the rest
//                       stream through an mbarrier ring.
//   warps 12,13 MMA     one warp per tile context; one elected thread issues tcgen05.mma (bf16 x bf16 -> fp32 in TMEM).  Two tile
//                       contexts (A, B) ping-pong: while the epilogue warps drain layer l of tile
//                       A the tensor core runs layer l of tile B, so neither side waits for the
//                       other's latency chain.
//   (SA1-type stages -- all weights pinned, TMEM room for 2 x 2 tiles -- run SUPER-TILES: a context handles two
//   tile ordinals per phase and the gather a pair of tiles per iteration)
//   warps 0-7  EPILOGUE two warpgroups, one per tile context (they drain concurrently); thread == TMEM lane.  Hidden layers: tcgen05.ld -> +bias (shared memory)
//                       -> ReLU fused into cvt.rn.relu.bf16x2 -> swizzled shared memory = the next
//                       layer's operand.  Last layer, pooled stages (S > 1): evaluated TRANSPOSED
//                       (D^T = W . H^T) so a lane is an output channel and the nsample rows of a
//                       point are consecutive TMEM columns: the max-pool is an in-register
//                       reduction.  Last layer, S == 1: plain orientation (lane == row), so both
//                       the channel-first f32 and the channel-last bf16 outputs store coalesced.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "sad_common.cuh"

namespace {

using namespace sad;

constexpr int kEpiWG = 128;              // threads of one epilogue warpgroup (thread == TMEM lane)
constexpr int kEpi = 2 * kEpiWG;         // two epilogue warpgroups: warps 0-3 drain tile context 0, warps 4-7 context 1
constexpr int kGather = 128;             // gather threads   (warps 8-11)
constexpr int kWarpMma = 12, kWarpTma = 14;
constexpr int kThreads = 480;            // + MMA warps (12: context 0, 13: context 1) + weight-TMA warp (14)
constexpr int kChunkBytes = 128 * 128;   // one 128-row x 64-bf16 K chunk (A operand / activations)
constexpr int kMaxLayers = 3;
constexpr int kMaxA = 8;                 // A-ring stages
constexpr int kMaxPin = 32;              // pinned weight pieces
constexpr int kMaxRing = 8;              // streamed weight ring slots
// "full" barriers of the two rings are indexed by ring POSITION modulo kFullBars, not by stage: with one consumer
// warp per tile context a consumer skips the other context's share of the ring, and a parity wait is only sound
// while the waiter is less than one barrier period away from the last fill it knows to be complete
// (plan_launch checks share + 1 + stages <= kFullBars before it allows two contexts).
constexpr int kFullBars = 32;
constexpr int kBiasPad = 544;            // floats per layer in the shared bias table
constexpr int kMiscBytes = 12288;
constexpr uint32_t kNoRow = 0xFFFFFFFFu;

struct MlpParams {
  int B, N, P, S, log2S;
  int log2P;                      // >= 0 when P is a power of two (batch index by shift), else -1
  long long total_rows;
  int num_tiles;
  uint32_t total_points;
  const __nv_bfloat16* feat_cl;   // (B,N,C0) channel-last source gathered by idx (or identity)
  int C0;
  const __nv_bfloat16* feat2_cl;  // (B,P,C1in) rows aligned with the output points (S == 1)
  int C1in;
  const float* xyz;               // (B,N,3)      } special chunk: (xyz[idx] - new_xyz) / r, extras
  const float* new_xyz;           // (B,P,3)
  const int32_t* idx;             // (B,P,S) or null (identity: row i of batch b)
  const float* radius_t;          // (B,P) or null
  float radius;
  int normalize;
  const float* extra;             // (B,N,E) fp32 scalar features appended after xyz
  int E;
  int has_special;
  int n_layers;
  int c[kMaxLayers];              // output channels per layer
  int kpad[kMaxLayers];           // K per layer, multiple of 64 (layer 0: C0 + C1in + 64*has_special)
  const uint8_t* w_img[kMaxLayers];
  const float* bias[kMaxLayers];
  int last_relu;
  __nv_bfloat16* out_cl;          // (B,P,c_last) bf16 or null
  float* out_cf;                  // (B,c_last,P) f32 or null
  int* tile_counter;              // zeroed by the caller: dynamic tile scheduling; null: static round-robin
  // ---- plan (host)
  int transposed;                 // last layer evaluated transposed (S > 1)
  int nslot;                      // tile contexts in flight (2 = ping-pong)
  int T;                          // tiles per context and phase (super-tile: 2 when every weight piece is pinned and TMEM allows)
  int region1;                    // TMEM columns of one tile (region_cols = T * region1)
  int na;                         // A-ring stages
  int depth;                      // cp.async groups in flight per gather thread
  int act_chunks;                 // 16 KB chunks per activation buffer (hidden_max / 64)
  int region_cols;                // TMEM columns per tile context
  int tmem_cols;                  // TMEM allocation (power of two)
  int nblk;                       // 128-channel blocks of a transposed last layer
  int cpad_last;                  // plain last layer: c_last rounded up to 32
  int piece_bytes[kMaxLayers];    // weight piece size per layer
  int pieces[kMaxLayers];         // pieces per tile per layer
  int first_piece[kMaxLayers];    // running piece index of the layer's first piece
  int pin_off[kMaxLayers];        // byte offset of the layer's first piece inside the pinned area
  int n_pieces, n_pinned, pinned_bytes, nr, ring_slot_bytes;
};

#ifdef SAD_MLP_PROFILE
__device__ long long g_mlp_log[4][2048];        // role (0 epilogue, 1 gather, 2 mma, 3 producer) x (event, clock)...
__device__ int g_mlp_logn[4];
#define SAD_LOG(role, ev)                                                                  \
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) == (role == 0 ? 0 : role == 1 ? 8 : role == 2 ? kWarpMma : kWarpTma)) { \
    if (sad_logn < 1023) {                                                                 \
      g_mlp_log[role][2 * sad_logn] = (ev);                                                \
      g_mlp_log[role][2 * sad_logn + 1] = clock64();                                       \
      g_mlp_logn[role] = ++sad_logn;                                                       \
    }                                                                                      \
  }
#else
#define SAD_LOG(role, ev)
#endif

struct Misc {
  uint64_t afull[kFullBars], afree[kMaxA];
  uint64_t wpin[kMaxPin];
  uint64_t wfull[kFullBars], wfree[kMaxRing];
  uint64_t dfull[2][5], actfull[2];
  uint64_t tfull[16];             // tile ordinal k of this CTA published in tiles[k & 15]
  int tiles[16];
  int boot[4], fetch[2];
  int fetch2[2][2];               // paired gather: tile ids of the pair two iterations ahead
  uint32_t tmem_base, pad_;
  uint32_t src[2][128];
  alignas(16) float bias[kMaxLayers][kBiasPad];
};
static_assert(sizeof(Misc) <= kMiscBytes, "misc area too small");

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// Prefetch loads: `volatile` pins them where they are written (one / two tiles ahead of their use);
// a plain __ldg gets sunk by the compiler to just before the first use, exposing the L2 round trip.
__device__ __forceinline__ float ldg_f32_pinned(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ldg_s32_pinned(const int32_t* p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// one lane of a converged warp (the warp runs the role's loop uniformly; only the tcgen05 / TMA
// instructions are predicated, so loop state stays in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, bf16 x bf16 -> f32, both operands K-major.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start >> 4, LBO = 1
// (unused for swizzled K-major), SBO = 1024 B between 8-row groups, version 1, layout type 2.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor: c=f32 (1<<4), a=b=bf16 (1<<7, 1<<10), K-major both, N>>3 at 17, M>>4 at 24.
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// bf16x2 of (max(lo,0), max(hi,0)): the ReLU rides on the conversion
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// byte offset of (row, 16-byte unit) inside a 128-row x 128-byte SWIZZLE_128B chunk
__device__ __forceinline__ uint32_t swz(int row, int unit) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((unit ^ (row & 7)) << 4));
}

// max of N accumulator words as a ternary tree (depth 4 for 32 values instead of a 31-long dependent chain;
// max.f32 takes three operands on sm_100)
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
template <int N>
__device__ __forceinline__ float vmax(const uint32_t* v) {
  if constexpr (N == 1) {
    return __uint_as_float(v[0]);
  } else if constexpr (N == 2) {
    return fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
  } else {
    constexpr int A = N / 3 + (N % 3 > 0), B = N / 3 + (N % 3 > 1);
    return max3(vmax<A>(v), vmax<B>(v + A), vmax<N - A - B>(v + A + B));
  }
}

// ---- pooled outputs of one thread (= output channel `ch`) for the 32 consecutive rows held in v
template <int S>
__device__ __forceinline__ void emit_group(const MlpParams& p, const uint32_t (&v)[32], float& run, int g, int ch,
                                           int c_last, float bias, uint32_t pt0, uint32_t b0, uint32_t j0,
                                           bool ch_ok) {
  auto store = [&](uint32_t q, float m) {       // q = point index inside the tile
    const uint32_t pt = pt0 + q;
    if (!ch_ok || pt >= p.total_points) return;
    uint32_t b = b0, j = j0 + q;
    while (j >= (uint32_t)p.P) {
      j -= (uint32_t)p.P;
      ++b;
    }
    float y = m + bias;
    if (p.last_relu) y = fmaxf(y, 0.f);
    if (p.out_cf) p.out_cf[((size_t)b * c_last + ch) * p.P + j] = y;
    if (p.out_cl) p.out_cl[(size_t)pt * c_last + ch] = __float2bfloat16_rn(y);
  };
  if constexpr (S >= 32) {
    const float m = vmax<32>(&v[0]);
    constexpr int GP = S / 32;                  // 32-column groups per point
    run = (g % GP == 0) ? m : fmaxf(run, m);
    if (g % GP == GP - 1) store((uint32_t)(g / GP), run);
  } else {
#pragma unroll
    for (int q = 0; q < 32 / S; ++q) {
      const float m = vmax<S>(&v[q * S]);
      store((uint32_t)(g * (32 / S) + q), m);
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) fused_mlp_kernel(const __grid_constant__ MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B atoms need 1024-B alignment
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  // carve-up (all offsets multiples of 1024)
  const uint32_t off_a = 0;
  const uint32_t off_act = off_a + (uint32_t)p.na * kChunkBytes;
  const uint32_t off_pin = off_act + (uint32_t)(p.nslot * p.T * p.act_chunks) * kChunkBytes;
  const uint32_t off_ring = off_pin + (uint32_t)p.pinned_bytes;
  const uint32_t off_misc = off_ring + (uint32_t)(p.nr * p.ring_slot_bytes);
  Misc* ms = reinterpret_cast<Misc*>(gbase + off_misc);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef SAD_MLP_PROFILE
  int sad_logn = 0;      // events logged by this thread (register counter: the log must not stall the role)
#endif
  const int nl = p.n_layers;
  const int c_last = p.c[nl - 1];
  const int chunks0 = p.kpad[0] / 64;
  // tile ordinal k of this CTA -> tile id (or -1 = no more work); published by the gather warps, which
  // run ahead of every other role
  auto get_tile = [&](int ord) -> int {
    mbar_wait(&ms->tfull[ord & 15], (uint32_t)((ord >> 4) & 1));
    return *reinterpret_cast<volatile int*>(&ms->tiles[ord & 15]);
  };

  if (tid == 0) {
    for (int i = 0; i < kFullBars; ++i) {
      mbar_init(&ms->afull[i], kGather);
      mbar_init(&ms->wfull[i], 1);
    }
    for (int i = 0; i < kMaxA; ++i) mbar_init(&ms->afree[i], 1);
    for (int i = 0; i < kMaxPin; ++i) mbar_init(&ms->wpin[i], 1);
    for (int i = 0; i < kMaxRing; ++i) mbar_init(&ms->wfree[i], 1);
    for (int i = 0; i < 16; ++i) mbar_init(&ms->tfull[i], 1);
    for (int i = 0; i < 2; ++i) {
      for (int k = 0; k < 5; ++k) mbar_init(&ms->dfull[i][k], 1);   // [0] hidden / plain last, [1+blk] transposed blocks
      mbar_init(&ms->actfull[i], kEpiWG);
    }
    mbar_fence_init();
  }
  if (warp == kWarpMma) {   // TMEM allocation: one warp, power-of-two columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ms->tmem_base)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int li = 0; li < nl; ++li)     // bias table (zero padded)
    for (int c = tid; c < kBiasPad; c += kThreads) ms->bias[li][c] = (c < p.c[li]) ? __ldg(p.bias[li] + c) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&ms->tmem_base);

  if (warp == kWarpTma) {
    // ================================================================== weight TMA producer
    if (lane == 0) {
      for (int li = 0; li < nl; ++li)
        for (int i = 0; i < p.pieces[li]; ++i) {
          const int pc = p.first_piece[li] + i;
          if (pc >= p.n_pinned) break;
          const uint32_t bytes = (uint32_t)p.piece_bytes[li];
          mbar_arrive_expect_tx(&ms->wpin[pc], bytes);
          tma_bulk_g2s(gbase + off_pin + p.pin_off[li] + (size_t)i * bytes, p.w_img[li] + (size_t)i * bytes, bytes,
                       &ms->wpin[pc]);
        }
      if (p.n_pinned < p.n_pieces) {
        uint32_t cnt = 0;
        for (int t0 = 0;; t0 += p.nslot) {
          int ns = p.nslot;
          for (int li = 0; li < nl; ++li)
            for (int s = 0; s < ns; ++s) {
              if (li == 0 && get_tile(t0 + s) < 0) {
                ns = s;
                break;
              }
              for (int i = 0; i < p.pieces[li]; ++i) {
                if (p.first_piece[li] + i < p.n_pinned) continue;
                const uint32_t r = cnt % (uint32_t)p.nr, use = cnt / (uint32_t)p.nr;
                if (use > 0) mbar_wait(&ms->wfree[r], (use - 1) & 1u);
                const uint32_t bytes = (uint32_t)p.piece_bytes[li];
                uint64_t* full = &ms->wfull[cnt & (kFullBars - 1)];
                mbar_arrive_expect_tx(full, bytes);
                tma_bulk_g2s(gbase + off_ring + (size_t)r * p.ring_slot_bytes, p.w_img[li] + (size_t)i * bytes, bytes, full);
                ++cnt;
              }
            }
          if (ns < p.nslot) break;
        }
      }
    }
  } else if (warp == kWarpMma || warp == kWarpMma + 1) {
    // ================================================================== MMA issuers (one warp per tile context)
    // Each context has its own issuing warp, so the two contexts' wait -> issue -> commit sequences (a few hundred
    // single-warp instructions per layer) run concurrently.  The A ring and the streamed-weight ring are filled in one
    // global order (round, layer, context, piece); each warp keeps the global position and skips the other context's
    // share, so no hand-off between the two warps is needed.
    const int s = warp - kWarpMma;
    if (s < p.nslot) {
      uint32_t w_pos = 0, w_slot = 0;      // streamed-weight ring: global position + slot wrap counter
      auto skip_w = [&](int n) {
        w_pos += (uint32_t)n;
        for (int i = 0; i < n; ++i)
          if (++w_slot == (uint32_t)p.nr) w_slot = 0;
      };
      uint32_t act_cnt = 0;
      bool pinned_ready = false;                                        // pinned pieces are waited for once
      const bool leader = elect_one();
      const int T = p.T;
      const uint32_t region = tmem_base + (uint32_t)(s * p.region_cols);
      const uint32_t act_s = base + off_act + (uint32_t)(s * T * p.act_chunks) * kChunkBytes;
      const int chunks0 = p.kpad[0] / 64;
      // weights of piece (li, i): pinned address or the next ring slot; returns the smem address
      // tcgen05.fence::after_thread_sync is needed only after this thread actually waited on a barrier; measured
      // (CTA-0 timeline) it costs a few hundred cycles on the issuing warp, so it is not issued per chunk
      bool waited = false;
      auto weights = [&](int li, int i, bool& streamed, uint32_t& slot) -> uint32_t {
        const int pc = p.first_piece[li] + i;
        if (pc < p.n_pinned) {
          if (!pinned_ready) {
            mbar_wait(&ms->wpin[pc], 0);
            waited = true;
          }
          streamed = false;
          return base + off_pin + (uint32_t)p.pin_off[li] + (uint32_t)i * (uint32_t)p.piece_bytes[li];
        }
        slot = w_slot;
        mbar_wait(&ms->wfull[w_pos & (kFullBars - 1)], (w_pos / kFullBars) & 1u);
        waited = true;
        skip_w(1);
        streamed = true;
        return base + off_ring + slot * (uint32_t)p.ring_slot_bytes;
      };
      // A context handles T consecutive tile ordinals per round and phase (a SUPER-TILE): ordinal (round * nslot +
      // context) * T + t.  T == 2 halves the hand-offs per row (one actfull / dfull round trip per layer for 256
      // rows); the planner allows it only when every weight piece is pinned (no streamed ring to share).
      for (int rnd = 0;; ++rnd) {
        const int t0 = rnd * p.nslot * T;            // first ordinal of the round (context 0)
        const int k0 = t0 + s * T;                   // this context's first ordinal
        // contexts active in this round (the last round of a CTA may have only context 0).  Context 0 must not ask
        // for context 1's ordinal before its own layer-0 chunks are consumed: with more K chunks than A-ring stages
        // the gather warps publish that ordinal only after this warp has freed stages (ns < 0: not known yet).
        if (get_tile(t0) < 0) break;
        int ns = -1;
        if (s == 1) {
          if (get_tile(k0) < 0) break;
          ns = 2;
        }
        int nt = 1;                                  // tiles of this super-tile (the CTA's last one may be short)
        while (nt < T && get_tile(k0 + nt) >= 0) ++nt;
#pragma unroll
        for (int li = 0; li < kMaxLayers; ++li) {
          if (li >= nl) break;
          const bool last = (li == nl - 1);
          const int chunks = p.kpad[li] / 64;
          // streamed pieces of this layer per context (the other context's share is skipped, before or after mine)
          const int lo = p.first_piece[li] > p.n_pinned ? p.first_piece[li] : p.n_pinned;
          const int nstream = p.first_piece[li] + p.pieces[li] > lo ? p.first_piece[li] + p.pieces[li] - lo : 0;
          if (s == 1) skip_w(nstream);
          // the context's TMEM region / activation buffer must have been drained by the epilogue of
          // the previous layer (li > 0) or of the previous tile in this context (li == 0)
          SAD_LOG(2, 100 + li * 10 + s)
          if (li > 0 || rnd > 0) {
            mbar_wait(&ms->actfull[s], act_cnt & 1u);
            ++act_cnt;
          }
          tc_fence_after();
          SAD_LOG(2, 200 + li * 10 + s)
          if (!(last && p.transposed)) {
            // D (128 rows x N) = A (rows x K) . W^T ; N = layer width (plain last layer: cpad, split at 256)
            const int ncols = last ? p.cpad_last : p.c[li];
            const uint32_t idesc_l = umma_idesc(128, ncols <= 256 ? ncols : 256);
            for (int t = 0; t < nt; ++t) {
            const uint32_t region_t = region + (uint32_t)(t * p.region1);
            const uint32_t act_t = act_s + (uint32_t)(t * p.act_chunks) * kChunkBytes;
            for (int kc = 0; kc < chunks; ++kc) {
              uint32_t a_addr;
              int ksteps = 4;
              uint32_t stage = 0;
              if (li == 0) {
                // A-ring position of chunk kc of ordinal k0 + t (the ring is filled in ordinal order)
                const uint32_t a_pos = (uint32_t)(k0 + t) * (uint32_t)chunks0 + (uint32_t)kc;
                stage = (p.na == 4) ? (a_pos & 3u) : (a_pos % (uint32_t)p.na);
                mbar_wait(&ms->afull[a_pos & (kFullBars - 1)], (a_pos / kFullBars) & 1u);
                waited = true;
                a_addr = base + off_a + stage * kChunkBytes;
                if (p.has_special && kc == chunks - 1) ksteps = 1;
              } else {
                a_addr = act_t + (uint32_t)kc * kChunkBytes;
              }
              bool streamed;
              uint32_t slot = 0;
              const uint32_t b_addr = weights(li, kc, streamed, slot);
              if (waited) {
                tc_fence_after();
                waited = false;
              }
              SAD_LOG(2, 400 + li * 10 + s)
              if (leader) {
                if (ncols <= 256) {                    // every hidden layer: one N pass, K steps unrolled
                  const uint64_t ad = umma_desc(a_addr), bd = umma_desc(b_addr);
                  umma_bf16(region_t, ad, bd, idesc_l, kc > 0 ? 1u : 0u);
                  if (ksteps == 4) {
                    umma_bf16(region_t, ad + 2u, bd + 2u, idesc_l, 1u);     // +2 per 32-byte K step in the (addr >> 4) field
                    umma_bf16(region_t, ad + 4u, bd + 4u, idesc_l, 1u);
                    umma_bf16(region_t, ad + 6u, bd + 6u, idesc_l, 1u);
                  }
                } else {
                for (int n0 = 0; n0 < ncols; n0 += 256) {
                  const int nn = min(256, ncols - n0);
                  const uint32_t idesc = umma_idesc(128, nn);
                  const uint64_t ad = umma_desc(a_addr), bd = umma_desc(b_addr + (uint32_t)n0 * 128u);
                  for (int k = 0; k < ksteps; ++k)
                    umma_bf16(region_t + (uint32_t)n0, ad + 2u * k, bd + 2u * k, idesc, (kc > 0 || k > 0) ? 1u : 0u);
                }
                }
                SAD_LOG(2, 500 + li * 10 + s)
                if (li == 0) umma_commit(&ms->afree[stage]);
                if (streamed) umma_commit(&ms->wfree[slot]);
                if (kc == chunks - 1 && t == nt - 1) umma_commit(&ms->dfull[s][0]);   // once per super-tile
                SAD_LOG(2, 600 + li * 10 + s)
              }
            }
            }
            __syncwarp();
            SAD_LOG(2, 300 + li * 10 + s)
          } else {
            // transposed last layer: D^T (128 channels x 128 rows) = W_blk . H^T, one commit per block
            const uint32_t idesc = umma_idesc(128, 128);
            for (int blk = 0; blk < p.nblk; ++blk) {
              for (int t = 0; t < nt; ++t) {
              const uint32_t region_t = region + (uint32_t)(t * p.region1);
              const uint32_t act_t = act_s + (uint32_t)(t * p.act_chunks) * kChunkBytes;
              for (int kc = 0; kc < chunks; ++kc) {
                bool streamed;
                uint32_t slot = 0;
                const uint32_t a_addr = weights(li, blk * chunks + kc, streamed, slot);   // W block rows = M
                const uint32_t b_addr = act_t + (uint32_t)kc * kChunkBytes;               // activations rows = N
                if (waited) {
                  tc_fence_after();
                  waited = false;
                }
                SAD_LOG(2, 400 + li * 10 + s)
                if (leader) {
                  const uint64_t ad = umma_desc(a_addr), bd = umma_desc(b_addr);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16(region_t + (uint32_t)(blk * 128), ad + 2u * k, bd + 2u * k, idesc, (kc > 0 || k > 0) ? 1u : 0u);
                  if (streamed) umma_commit(&ms->wfree[slot]);
                  // one barrier per block: never two phases outstanding
                  if (kc == chunks - 1 && t == nt - 1) umma_commit(&ms->dfull[s][1 + blk]);
                }
              }
              }
              __syncwarp();
              SAD_LOG(2, 300 + li * 10 + s)
            }
          }
          if (s == 0) {
            if (ns < 0) ns = (p.nslot > 1 && get_tile(t0 + T) >= 0) ? 2 : 1;
            if (ns == 2) skip_w(nstream);
          }
        }
        pinned_ready = true;
        if (nt < T) break;                           // a short super-tile is the CTA's last
      }
    }
  } else if (warp >= kEpi / 32) {
    // ================================================================== gather producers
    const int gt = tid - kEpi;                 // row of the tile this thread owns for bookkeeping / special chunk
    const int unit = gt & 7, rbase = gt >> 3;
    const int nf0 = p.C0 / 64, nf1 = p.C1in / 64;
    uint32_t g_pos = 0;                                // ring position of the next chunk (its full barrier: g_pos % kFullBars)
    uint32_t g_stage = 0, g_phase = 0;                 // afree parity of the NEXT wait (the first pass over the ring never waits)
    bool g_first = true;
    int pend = 0;
    uint32_t ps0 = 0, ps1 = 0, ps2 = 0;        // full barriers of the cp.async groups still in flight (oldest first)

    struct Special {
      float x, y, z, qx, qy, qz, r, e[4];
    };
    // Source row of this thread in `tile`, in two steps so that no instruction ever waits for a load inside the
    // iteration that issued it: issue_idx() starts the index load, resolve_src() (one iteration later) turns the
    // landed value into the global row b * N + idx.
    auto issue_idx = [&](int tile) -> int {
      if (tile < 0 || !p.idx) return 0;
      const long long R = (long long)tile * 128 + gt;
      return R < p.total_rows ? ldg_s32_pinned(p.idx + R) : 0;
    };
    // point -> (batch, point in batch): a shift when P is a power of two (every stage of the detector), else a division
    auto batch_of = [&](uint32_t pt) -> uint32_t {
      return p.log2P >= 0 ? (pt >> p.log2P) : pt / (uint32_t)p.P;
    };
    auto resolve_src = [&](int tile, int raw) -> uint32_t {
      if (tile < 0) return kNoRow;
      const long long R = (long long)tile * 128 + gt;
      if (R >= p.total_rows) return kNoRow;
      const uint32_t pt = (uint32_t)(R >> p.log2S);
      const uint32_t b = batch_of(pt);
      const uint32_t id = p.idx ? (uint32_t)raw : (pt - b * (uint32_t)p.P);
      return b * (uint32_t)p.N + id;
    };
    // 1 / radius of a scalar-radius stage, once per kernel (the per-cluster radius is inverted per row); the
    // normalised offsets feed a bf16 operand, so multiplying by the reciprocal instead of dividing is far inside
    // the rounding of the storage format
    const float inv_radius = (p.normalize && !p.radius_t) ? __frcp_rn(p.radius) : 1.f;
    auto load_special = [&](uint32_t src, int tile, Special& s) {
      s.x = s.y = s.z = s.qx = s.qy = s.qz = 0.f;
      s.r = 1.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) s.e[e] = 0.f;
      if (src == kNoRow || !p.has_special) return;
      if (p.xyz) {
        const uint32_t pt = (uint32_t)(((long long)tile * 128 + gt) >> p.log2S);
        const float* a = p.xyz + (size_t)src * 3;
        const float* q = p.new_xyz + (size_t)pt * 3;
        s.x = ldg_f32_pinned(a);
        s.y = ldg_f32_pinned(a + 1);
        s.z = ldg_f32_pinned(a + 2);
        s.qx = ldg_f32_pinned(q);
        s.qy = ldg_f32_pinned(q + 1);
        s.qz = ldg_f32_pinned(q + 2);
        if (p.normalize && p.radius_t) s.r = ldg_f32_pinned(p.radius_t + pt);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (e < p.E) s.e[e] = ldg_f32_pinned(p.extra + (size_t)src * p.E + e);
    };
    auto retire_oldest = [&]() {               // oldest cp.async group has landed: publish its stage
      fence_proxy_async();
      mbar_arrive(&ms->afull[ps0]);
      ps0 = ps1;
      ps1 = ps2;
      --pend;
    };

    // next tile of this CTA: static round-robin, or one atomic on the caller-zeroed counter (dynamic: CTAs
    // that start late -- SMs held by another stream's kernel -- simply find less work left)
    // (raw value now, bounds check when it is consumed: the atomic's round trip stays off this iteration's path)
    // (fetch_raw returns the atomic's result untouched -- even adding gridDim.x here would make this iteration wait
    // for the round trip; tile_of_raw finishes the job where the value is consumed)
    auto fetch_raw = [&](int ord) -> int {
      if (p.tile_counter) return atomicAdd(p.tile_counter, 1);
      return ord;
    };
    auto tile_of_raw = [&](int raw) -> int {
      const long long t = p.tile_counter ? (long long)gridDim.x + raw : (long long)blockIdx.x + (long long)raw * gridDim.x;
      return (raw >= 0 && t < p.num_tiles) ? (int)t : -1;
    };
    if (p.T == 2 && chunks0 == 1 && nf0 + nf1 == 0 && p.E <= 4) {
      // ---------------------------------------------------------------- paired gather (SA1: special chunk only)
      // Two tiles per iteration, one row of each per thread: the iteration's fixed cost (publication, barrier, ring
      // bookkeeping, the publication fence) is paid once per 256 rows, which is what lets the gather keep up with
      // the super-tile consumers.  Same load discipline as the single-tile loop: every value consumed here was
      // requested one iteration earlier.
      int ta[2], tb[2], tc[2];                 // tile ids of the current pair, the next, the one after
      int rawp[2] = {-1, -1};                  // gt == 0: raw fetches of the pair two iterations ahead
      if (gt == 0) {
        ms->boot[0] = tile_of_raw(fetch_raw(1));
        ms->boot[1] = tile_of_raw(fetch_raw(2));
        ms->boot[2] = tile_of_raw(fetch_raw(3));
        rawp[0] = fetch_raw(4);
        rawp[1] = fetch_raw(5);
      }
      named_bar_sync(1, kGather);
      ta[0] = (int)blockIdx.x;
      ta[1] = ms->boot[0];
      tb[0] = ms->boot[1];
      tb[1] = ms->boot[2];
      uint32_t srcp[2];
      int idr[2];
      Special spp[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        srcp[u] = resolve_src(ta[u], issue_idx(ta[u]));
        load_special(srcp[u], ta[u], spp[u]);
        idr[u] = issue_idx(tb[u]);
      }
      int it2 = 0;
      for (; ta[0] >= 0; ++it2) {
        if (gt == 0) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            ms->tiles[(2 * it2 + u) & 15] = ta[u];
            mbar_arrive(&ms->tfull[(2 * it2 + u) & 15]);
            ms->fetch2[it2 & 1][u] = tile_of_raw(rawp[u]);       // pair it2 + 2 (fetched one iteration ago)
          }
        }
        SAD_LOG(1, 100)
        named_bar_sync(1, kGather);
        SAD_LOG(1, 200)
        tc[0] = ms->fetch2[it2 & 1][0];
        tc[1] = ms->fetch2[it2 & 1][1];
        uint32_t fb[2] = {0, 0};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (ta[u] < 0) continue;                               // warp-uniform: the pair's second tile may not exist
          const uint32_t stage = g_stage;
          if (!g_first) mbar_wait(&ms->afree[stage], g_phase);
          if (++g_stage == (uint32_t)p.na) {
            g_stage = 0;
            if (g_first) g_first = false;
            else g_phase ^= 1u;
          }
          const uint32_t dst = base + off_a + stage * kChunkBytes;
          fb[u] = g_pos++ & (kFullBars - 1);
          float vals[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) vals[i] = 0.f;
          if (srcp[u] != kNoRow) {
            if (p.xyz) {
              float dx = __fsub_rn(spp[u].x, spp[u].qx), dy = __fsub_rn(spp[u].y, spp[u].qy),
                    dz = __fsub_rn(spp[u].z, spp[u].qz);
              if (p.normalize) {
                const float inv = p.radius_t ? __frcp_rn(spp[u].r) : inv_radius;
                dx *= inv;
                dy *= inv;
                dz *= inv;
              }
              vals[0] = dx;
              vals[1] = dy;
              vals[2] = dz;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) vals[3 + e] = spp[u].e[e];
          }
          st_shared_v4(dst + swz(gt, 0), pack_bf16(vals[0], vals[1]), pack_bf16(vals[2], vals[3]),
                       pack_bf16(vals[4], vals[5]), pack_bf16(vals[6], vals[7]));
          st_shared_v4(dst + swz(gt, 1), 0u, 0u, 0u, 0u);
        }
        fence_proxy_async();                                     // one publication fence for both tiles
        if (ta[0] >= 0) mbar_arrive(&ms->afull[fb[0]]);
        if (ta[1] >= 0) mbar_arrive(&ms->afull[fb[1]]);
        SAD_LOG(1, 600)
        if (gt == 0) {
          rawp[0] = fetch_raw(2 * it2 + 6);
          rawp[1] = fetch_raw(2 * it2 + 7);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          srcp[u] = resolve_src(tb[u], idr[u]);                  // next pair's source rows (their indices have landed)
          idr[u] = issue_idx(tc[u]);                             // indices two pairs ahead
          load_special(srcp[u], tb[u], spp[u]);                  // special-chunk inputs one pair ahead
          ta[u] = tb[u];
          tb[u] = tc[u];
        }
        SAD_LOG(1, 300)
      }
      if (gt == 0) {                                             // end markers (see the single-tile loop)
        for (int e = 0; e <= p.nslot * p.T; ++e) {
          ms->tiles[(2 * it2 + e) & 15] = -1;
          mbar_arrive(&ms->tfull[(2 * it2 + e) & 15]);
        }
      }
    } else {
    int tq0 = (int)blockIdx.x, tq1, tq2;
    int raw_f = -1;                            // gt == 0: raw tile fetch of ordinal it + 2
    if (gt == 0) {
      ms->boot[0] = tile_of_raw(fetch_raw(1));
      raw_f = fetch_raw(2);
      ms->boot[1] = tile_of_raw(raw_f);
    }
    named_bar_sync(1, kGather);
    tq1 = ms->boot[0];
    tq2 = ms->boot[1];
    uint32_t src_cur = resolve_src(tq0, issue_idx(tq0));
    int idraw = issue_idx(tq1);                // index of this thread's row in tile tq1 (in flight)
    Special sp_cur;                            // special-chunk inputs of the tile about to be gathered
    load_special(src_cur, tq0, sp_cur);
    int it = 0;
    for (; tq0 >= 0; ++it) {
      const long long R0 = (long long)tq0 * 128;
      uint32_t* s_src = ms->src[it & 1];
      s_src[gt] = src_cur;
      if (gt == 0) {
        ms->tiles[it & 15] = tq0;
        mbar_arrive(&ms->tfull[it & 15]);
        ms->fetch[it & 1] = tile_of_raw(raw_f);           // tile id of ordinal it + 2 (fetched one iteration ago)
      }
      SAD_LOG(1, 100)
      named_bar_sync(1, kGather);
      SAD_LOG(1, 200)
      tq2 = ms->fetch[it & 1];
      for (int kc = 0; kc < chunks0; ++kc) {
        const uint32_t stage = g_stage;
        if (!g_first && !mbar_try_wait(&ms->afree[stage], g_phase)) {
          // ring full: publish everything already issued before blocking, or the MMA warp (which frees
          // the stage) would be waiting for a chunk that only the next issue would have published
          if (pend > 0) {
            cp_async_wait<0>();
            while (pend > 0) retire_oldest();
          }
          mbar_wait(&ms->afree[stage], g_phase);
        }
        if (++g_stage == (uint32_t)p.na) {
          g_stage = 0;
          if (g_first) g_first = false;
          else g_phase ^= 1u;
        }
        const uint32_t dst = base + off_a + stage * kChunkBytes;
        const uint32_t fbar = g_pos++ & (kFullBars - 1);
        SAD_LOG(1, 400)
        if (kc < nf0 + nf1) {
          if (kc < nf0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = rbase + 16 * i;
              const uint32_t src = s_src[r];
              const __nv_bfloat16* g = p.feat_cl + (size_t)(src == kNoRow ? 0u : src) * p.C0 + kc * 64 + unit * 8;
              cp_async16(dst + swz(r, unit), g, src == kNoRow ? 0u : 16u);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = rbase + 16 * i;
              const long long R = R0 + r;
              const bool ok = R < p.total_rows;
              const __nv_bfloat16* g = p.feat2_cl + (size_t)(ok ? R : 0) * p.C1in + (kc - nf0) * 64 + unit * 8;
              cp_async16(dst + swz(r, unit), g, ok ? 16u : 0u);
            }
          }
          cp_async_commit();
          if (pend == 0) ps0 = fbar;
          else if (pend == 1) ps1 = fbar;
          else ps2 = fbar;
          ++pend;
          if (pend == p.depth) {
            if (p.depth == 3) cp_async_wait<2>();
            else if (p.depth == 2) cp_async_wait<1>();
            else cp_async_wait<0>();
            retire_oldest();
          }
        } else {
          // special 16-wide K step: [dx, dy, dz, extras..., 0]
          float vals[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) vals[i] = 0.f;
          if (src_cur != kNoRow) {
            if (p.xyz) {
              float dx = __fsub_rn(sp_cur.x, sp_cur.qx), dy = __fsub_rn(sp_cur.y, sp_cur.qy),
                    dz = __fsub_rn(sp_cur.z, sp_cur.qz);
              if (p.normalize) {
                const float inv = p.radius_t ? __frcp_rn(sp_cur.r) : inv_radius;
                dx *= inv;
                dy *= inv;
                dz *= inv;
              }
              vals[0] = dx;
              vals[1] = dy;
              vals[2] = dz;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) vals[3 + e] = sp_cur.e[e];
#pragma unroll
            for (int e = 4; e < 13; ++e)
              if (e < p.E) vals[3 + e] = __ldg(p.extra + (size_t)src_cur * p.E + e);   // rare
          }
          st_shared_v4(dst + swz(gt, 0), pack_bf16(vals[0], vals[1]), pack_bf16(vals[2], vals[3]),
                       pack_bf16(vals[4], vals[5]), pack_bf16(vals[6], vals[7]));
          SAD_LOG(1, 500)
          st_shared_v4(dst + swz(gt, 1), pack_bf16(vals[8], vals[9]), pack_bf16(vals[10], vals[11]),
                       pack_bf16(vals[12], vals[13]), pack_bf16(vals[14], vals[15]));
          fence_proxy_async();
          mbar_arrive(&ms->afull[fbar]);
          SAD_LOG(1, 600)
        }
      }
      // Prefetches for the next tiles are issued AFTER this tile's chunks are published: the publication fence
      // (fence.proxy.async = MEMBAR + proxy fence) waits for every load the thread has in flight, so loads issued
      // before it would put a full L2 / atomic round trip on the path of every chunk.  Issued here they have a whole
      // iteration to land before the next fence.
      // Every value consumed here was requested one iteration ago; every request made here is consumed one
      // iteration later (raw_f at the top, idraw here, sp_cur in the special chunk).
      if (gt == 0) raw_f = fetch_raw(it + 3);
      src_cur = resolve_src(tq1, idraw);                  // next tile's source row (its index has landed)
      idraw = issue_idx(tq2);                             // index two tiles ahead
      load_special(src_cur, tq1, sp_cur);                 // special-chunk inputs one tile ahead
      SAD_LOG(1, 300)
      tq0 = tq1;
      tq1 = tq2;
    }
    if (gt == 0) {                                        // end markers for the other roles: a context asks for ordinals
      for (int e = 0; e <= p.nslot * p.T; ++e) {          // up to one round (nslot * T) past the last tile
        ms->tiles[(it + e) & 15] = -1;
        mbar_arrive(&ms->tfull[(it + e) & 15]);
      }
    }
    cp_async_wait<0>();
    while (pend > 0) retire_oldest();
    }
  } else {
    // ================================================================== epilogue warpgroups (thread == TMEM lane)
    // warpgroup s drains tile context s only: the two contexts' epilogues run concurrently, and the MMA warp
    // (which alternates between the contexts) never finds both of them queued behind one set of warps
    const int s = warp >> 2;                    // tile context of this warpgroup
    const int et = tid & (kEpiWG - 1);          // TMEM lane == row of the tile (plain) / output channel (transposed)
    if (s < p.nslot) {
      uint32_t d_cnt = 0;            // uses of dfull[s][0]
      uint32_t t_cnt = 0;            // tiles finished in this context (= uses of each dfull[s][1+blk])
      const int T = p.T;
      const uint32_t region_s = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(s * p.region_cols);
      const uint32_t act_ctx = base + off_act + (uint32_t)(s * T * p.act_chunks) * kChunkBytes;
      for (int rnd = 0;; ++rnd) {
        const int k0 = (rnd * p.nslot + s) * T;      // this context's first ordinal of the round (super-tile of T tiles)
        const int tile_a = get_tile(k0);
        if (tile_a < 0) break;
        const int tile_b = T > 1 ? get_tile(k0 + 1) : -1;
        const int nt = tile_b >= 0 ? 2 : 1;
        for (int li = 0; li < nl; ++li) {
          const bool last = (li == nl - 1);
          SAD_LOG(0, 100 + li * 10 + s)
          if (!last) {
            // ---- hidden layer: TMEM -> +bias -> ReLU -> bf16 -> swizzled ACT (next layer's operand)
            mbar_wait(&ms->dfull[s][0], d_cnt & 1u);
            ++d_cnt;
            tc_fence_after();
            SAD_LOG(0, 200 + li * 10 + s)
            const float* sb = ms->bias[li];
            for (int t = 0; t < nt; ++t) {
            const uint32_t region = region_s + (uint32_t)(t * p.region1);
            const uint32_t act_s = act_ctx + (uint32_t)(t * p.act_chunks) * kChunkBytes;
            for (int c0 = 0; c0 < p.c[li]; c0 += 64) {
              uint32_t v0[32], v1[32];
              tmem_ld32_issue(region + (uint32_t)c0, v0);
              tmem_ld32_issue(region + (uint32_t)c0 + 32u, v1);
              tmem_ld_wait();
              SAD_LOG(0, 400 + li * 10 + s)
              const uint32_t chunk = act_s + (uint32_t)(c0 >> 6) * kChunkBytes;
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const uint32_t* v = (u < 4) ? v0 : v1;
                const int o = (u & 3) * 8;
                const float4 ba = *reinterpret_cast<const float4*>(sb + c0 + u * 8);
                const float4 bb = *reinterpret_cast<const float4*>(sb + c0 + u * 8 + 4);
                st_shared_v4(chunk + swz(et, u),
                             pack_bf16_relu(__uint_as_float(v[o + 0]) + ba.x, __uint_as_float(v[o + 1]) + ba.y),
                             pack_bf16_relu(__uint_as_float(v[o + 2]) + ba.z, __uint_as_float(v[o + 3]) + ba.w),
                             pack_bf16_relu(__uint_as_float(v[o + 4]) + bb.x, __uint_as_float(v[o + 5]) + bb.y),
                             pack_bf16_relu(__uint_as_float(v[o + 6]) + bb.z, __uint_as_float(v[o + 7]) + bb.w));
              }
            }
            }
            SAD_LOG(0, 500 + li * 10 + s)
            fence_proxy_async();
          } else if (p.transposed) {
            // ---- last layer, transposed: thread == output channel, columns == rows; pool over S
            const uint32_t npt = 128u >> p.log2S;
            for (int blk = 0; blk < p.nblk; ++blk) {
              mbar_wait(&ms->dfull[s][1 + blk], t_cnt & 1u);
              tc_fence_after();
              SAD_LOG(0, 200 + li * 10 + s)
              for (int t = 0; t < nt; ++t) {
              const uint32_t region = region_s + (uint32_t)(t * p.region1);
              const uint32_t pt0 = (uint32_t)(t == 0 ? tile_a : tile_b) * npt;
              const uint32_t b0 = p.log2P >= 0 ? (pt0 >> p.log2P) : pt0 / (uint32_t)p.P, j0 = pt0 - b0 * (uint32_t)p.P;
              const int ch = blk * 128 + et;
              const bool ch_ok = ch < c_last;
              const float bias = ms->bias[li][ch_ok ? ch : 0];
              float run = 0.f;
#pragma unroll 1
              for (int g = 0; g < 4; g += 2) {
                uint32_t v0[32], v1[32];
                tmem_ld32_issue(region + (uint32_t)(blk * 128 + g * 32), v0);
                tmem_ld32_issue(region + (uint32_t)(blk * 128 + g * 32 + 32), v1);
                tmem_ld_wait();
                SAD_LOG(0, 400 + li * 10 + s)
                switch (p.S) {
#define SAD_EMIT(SS)                                                           \
  case SS:                                                                     \
    emit_group<SS>(p, v0, run, g, ch, c_last, bias, pt0, b0, j0, ch_ok);       \
    emit_group<SS>(p, v1, run, g + 1, ch, c_last, bias, pt0, b0, j0, ch_ok);   \
    break;
                  SAD_EMIT(2)
                  SAD_EMIT(4)
                  SAD_EMIT(8)
                  SAD_EMIT(16)
                  SAD_EMIT(32)
                  SAD_EMIT(64)
                  default:
                    emit_group<128>(p, v0, run, g, ch, c_last, bias, pt0, b0, j0, ch_ok);
                    emit_group<128>(p, v1, run, g + 1, ch, c_last, bias, pt0, b0, j0, ch_ok);
                    break;
#undef SAD_EMIT
                }
              }
              }
            }
            ++t_cnt;
          } else {
            // ---- last layer, plain orientation (S == 1): thread == row; both outputs store coalesced (T == 1 here)
            mbar_wait(&ms->dfull[s][0], d_cnt & 1u);
            ++d_cnt;
            tc_fence_after();
            SAD_LOG(0, 200 + li * 10 + s)
            const uint32_t region = region_s;
            const long long tile = tile_a;
            const long long R = tile * 128 + et;
            const bool ok = R < p.total_rows;
            const uint32_t pt = ok ? (uint32_t)R : 0u;
            const uint32_t b = p.log2P >= 0 ? (pt >> p.log2P) : pt / (uint32_t)p.P, j = pt - b * (uint32_t)p.P;
            const float* sb = ms->bias[li];
            float* ocf = p.out_cf ? p.out_cf + (size_t)b * c_last * p.P + j : nullptr;
            __nv_bfloat16* ocl = p.out_cl ? p.out_cl + (size_t)pt * c_last : nullptr;
            const bool vec_cl = (c_last % 8) == 0;
            for (int c0 = 0; c0 < p.cpad_last; c0 += 32) {
              uint32_t v[32];
              tmem_ld32_issue(region + (uint32_t)c0, v);
              tmem_ld_wait();
              SAD_LOG(0, 400 + li * 10 + s)
              float y[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                y[i] = __uint_as_float(v[i]) + sb[c0 + i];
                if (p.last_relu) y[i] = fmaxf(y[i], 0.f);
              }
              if (ok) {
                if (ocf) {
#pragma unroll
                  for (int i = 0; i < 32; ++i)
                    if (c0 + i < c_last) ocf[(size_t)(c0 + i) * p.P] = y[i];
                }
                if (ocl) {
                  if (vec_cl) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                      if (c0 + u * 8 < c_last)
                        *reinterpret_cast<uint4*>(ocl + c0 + u * 8) =
                            make_uint4(pack_bf16(y[u * 8], y[u * 8 + 1]), pack_bf16(y[u * 8 + 2], y[u * 8 + 3]),
                                       pack_bf16(y[u * 8 + 4], y[u * 8 + 5]), pack_bf16(y[u * 8 + 6], y[u * 8 + 7]));
                  } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                      if (c0 + i < c_last) ocl[c0 + i] = __float2bfloat16_rn(y[i]);
                  }
                }
              }
            }
          }
          tc_fence_before();
          mbar_arrive(&ms->actfull[s]);
          SAD_LOG(0, 300 + li * 10 + s)
        }
        if (nt < T) break;                           // a short super-tile is the CTA's last
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------ three_interpolate, channel-last bf16
// out[b,i,:] = w0*f[b,i0,:] + w1*f[b,i1,:] + w2*f[b,i2,:]  (fp32 math, bf16 storage); one warp per output
// row, 16-byte (8-channel) vectors per lane: every load/store instruction of a warp moves 512 contiguous bytes.
__global__ void __launch_bounds__(256)
interp_cl_kernel(long long rows, int n, int m, int C, const __nv_bfloat16* __restrict__ feat,
                 const int32_t* __restrict__ idx, const float* __restrict__ weight, __nv_bfloat16* __restrict__ out) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long b = row / n;
  const int i0 = __ldg(idx + row * 3), i1 = __ldg(idx + row * 3 + 1), i2 = __ldg(idx + row * 3 + 2);
  const float w0 = __ldg(weight + row * 3), w1 = __ldg(weight + row * 3 + 1), w2 = __ldg(weight + row * 3 + 2);
  const uint4* f0 = reinterpret_cast<const uint4*>(feat + (b * m + i0) * C);
  const uint4* f1 = reinterpret_cast<const uint4*>(feat + (b * m + i1) * C);
  const uint4* f2 = reinterpret_cast<const uint4*>(feat + (b * m + i2) * C);
  uint4* o = reinterpret_cast<uint4*>(out + row * C);
  for (int v = lane; v < C / 8; v += 32) {
    const uint4 a = __ldg(f0 + v), bq = __ldg(f1 + v), c = __ldg(f2 + v);
    const uint32_t* pa = &a.x;
    const uint32_t* pb = &bq.x;
    const uint32_t* pc = &c.x;
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 xa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pa + k));
      const float2 xb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pb + k));
      const float2 xc = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pc + k));
      r[k] = pack_bf16(w0 * xa.x + w1 * xb.x + w2 * xc.x, w0 * xa.y + w1 * xb.y + w2 * xc.y);
    }
    o[v] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

// (B,C,N) f32 channel-first -> (B,N,C) bf16 channel-last through a padded smem tile.
__global__ void __launch_bounds__(256)
cf_to_cl_kernel(int C, int N, const float* __restrict__ in, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, n = n0 + tx;
    tile[r][tx] = (c < C && n < N) ? __ldg(in + ((size_t)b * C + c) * N + n) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, c = c0 + tx;
    if (n < N && c < C) out[((size_t)b * N + n) * C + c] = __float2bfloat16_rn(tile[tx][r]);
  }
}

}  // namespace

// ------------------------------------------------------------------------------ host side
// Weight image formats (all K-major SWIZZLE_128B, one piece per 64-wide K chunk):
//   mode 0  hidden layer          piece = cout rows                      (cout % 16 == 0, <= 256)
//   mode 1  last layer, S  > 1    transposed evaluation: pieces of 128 rows, blocked [blk][kc]
//   mode 2  last layer, S == 1    plain evaluation: piece = cout rounded up to 32 rows (<= 512)
static int image_rows(int cout, int mode) { return mode == 1 ? 128 : (mode == 2 ? (cout + 31) / 32 * 32 : cout); }

#ifdef SAD_MLP_PROFILE
// tools only: copy the CTA-0 timeline of the last launch to the host and reset it
extern "C" SAD_API int sad_mlp_profile_dump(long long* host_log /*4*2048*/, int* host_n /*4*/) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host_log, g_mlp_log, sizeof(long long) * 4 * 2048);
  cudaMemcpyFromSymbol(host_n, g_mlp_logn, sizeof(int) * 4);
  int z[4] = {0, 0, 0, 0};
  cudaMemcpyToSymbol(g_mlp_logn, z, sizeof(z));
  return 0;
}
#endif

extern "C" long long sad_mlp_weight_image_bytes(int cout, int kpad, int is_last) {
  if (cout < 1 || kpad < 64 || (kpad % 64) || is_last < 0 || is_last > 2) return -1;
  const long long chunks = kpad / 64;
  const long long nblk = is_last == 1 ? (cout + 127) / 128 : 1;
  return nblk * chunks * image_rows(cout, is_last) * 128;
}

// Host-side packer: fp32 W (cout x cin, row-major) -> bf16 image in the exact shared-memory byte
// layout the kernel consumes.  perm[k] = source column of packed K index k, or -1 for a zero column.
extern "C" int sad_mlp_pack_weights(const float* W, int cout, int cin, const int32_t* perm, int kpad, int is_last,
                                    void* out_image) {
  SAD_REQUIRE(W && perm && out_image, "mlp_pack_weights: null pointer");
  SAD_REQUIRE(cout >= 1 && cin >= 1 && kpad >= 64 && kpad % 64 == 0, "mlp_pack_weights: bad sizes");
  SAD_REQUIRE(is_last >= 0 && is_last <= 2, "mlp_pack_weights: is_last must be 0, 1 (pooled) or 2 (S == 1)");
  SAD_REQUIRE(is_last != 0 || (cout % 16 == 0 && cout <= 256),
              "mlp_pack_weights: hidden width must be a multiple of 16, <= 256");
  SAD_REQUIRE(is_last != 2 || cout <= 512, "mlp_pack_weights: S == 1 last layer supports <= 512 channels");
  const int chunks = kpad / 64;
  const int rows_per_piece = image_rows(cout, is_last);
  const int nblk = is_last == 1 ? (cout + 127) / 128 : 1;
  uint16_t* img = static_cast<uint16_t*>(out_image);
  const size_t piece_elems = (size_t)rows_per_piece * 64;
  for (int blk = 0; blk < nblk; ++blk)
    for (int kc = 0; kc < chunks; ++kc) {
      uint16_t* piece = img + ((size_t)blk * chunks + kc) * piece_elems;
      for (int r = 0; r < rows_per_piece; ++r) {
        const int n = blk * rows_per_piece + r;
        for (int kk = 0; kk < 64; ++kk) {
          const int k = kc * 64 + kk;
          float w = 0.f;
          if (n < cout && perm[k] >= 0) {
            SAD_REQUIRE(perm[k] < cin, "mlp_pack_weights: perm[%d]=%d out of range", k, perm[k]);
            w = W[(size_t)n * cin + perm[k]];
          }
          uint32_t u;
          memcpy(&u, &w, 4);
          const uint32_t rounded = (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;     // RNE fp32 -> bf16 (finite inputs)
          const int unit = kk >> 3;
          const size_t byte = (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + (size_t)((unit ^ (r & 7)) << 4) +
                              (size_t)(kk & 7) * 2;
          piece[byte / 2] = (uint16_t)rounded;
        }
      }
    }
  return SAD_OK;
}

// Shared-memory / TMEM plan of one launch.  Returns false when nothing fits.
static bool plan_launch(MlpParams& p, int hidden_max, int super_tiles, size_t& smem_out) {
  const int nl = p.n_layers;
  const int c_last = p.c[nl - 1];
  p.transposed = p.S > 1 ? 1 : 0;
  p.nblk = p.transposed ? (c_last + 127) / 128 : 0;
  p.cpad_last = p.transposed ? 0 : (c_last + 31) / 32 * 32;
  p.act_chunks = hidden_max / 64;
  p.region1 = p.transposed ? (hidden_max > 128 * p.nblk ? hidden_max : 128 * p.nblk)
                           : (hidden_max > p.cpad_last ? hidden_max : p.cpad_last);
  if (p.region1 > 512) return false;
  int np = 0, max_piece = 0;
  long long total_w = 0;
  for (int li = 0; li < nl; ++li) {
    const bool last = (li == nl - 1);
    const int chunks = p.kpad[li] / 64;
    p.piece_bytes[li] = (last ? (p.transposed ? 128 : p.cpad_last) : p.c[li]) * 128;
    p.pieces[li] = (last && p.transposed) ? p.nblk * chunks : chunks;
    p.first_piece[li] = np;
    np += p.pieces[li];
    total_w += (long long)p.pieces[li] * p.piece_bytes[li];
    if (p.piece_bytes[li] > max_piece) max_piece = p.piece_bytes[li];
  }
  p.n_pieces = np;
  // shared-memory budget per CTA: everything the SM has by default; SAD_MLP_SMEM_KB (tuning hook) leaves room for
  // another stream's CTAs (e.g. a latency-bound FPS cluster) to share the SM
  long long budget_kb = 227;
  if (const char* e_kb = sad_tool_env("SAD_MLP_SMEM_KB")) budget_kb = atoi(e_kb) < 64 ? 64 : (atoi(e_kb) > 227 ? 227 : atoi(e_kb));
  const long long avail = budget_kb * 1024 - 1024 - kMiscBytes;
  // tuning hooks (benchmarks only): cap the tile contexts / A-ring stages, force a minimum weight ring
  const char* e_slot = sad_tool_env("SAD_MLP_NSLOT");
  const char* e_na = sad_tool_env("SAD_MLP_NA");
  const char* e_nr = sad_tool_env("SAD_MLP_NR");
  const int max_slot = e_slot ? atoi(e_slot) : 2, max_na = e_na ? atoi(e_na) : 4, want_nr = e_nr ? atoi(e_nr) : 2;
  int max_pieces = 0;
  for (int li = 0; li < nl; ++li) max_pieces = p.pieces[li] > max_pieces ? p.pieces[li] : max_pieces;
  // two contexts = two consumers per ring: the skipped share + 1 + the ring depth must stay inside one period of
  // the position-indexed full barriers (see kFullBars)
  const bool two_ok = p.kpad[0] / 64 + 1 + kMaxA <= kFullBars && max_pieces + 1 + kMaxRing <= kFullBars;
  // super-tiles (T = 2 tiles per context and phase): pooled stages with enough tiles per CTA, two contexts of two
  // tiles in TMEM, two tiles of layer-1 chunks in the A ring, and EVERY weight piece pinned (SAD_MLP_T: tuning / tests)
  const bool try_T2 = p.transposed && 4 * p.region1 <= 512 && max_slot >= 2 && two_ok &&
                      (super_tiles == 0 ? p.num_tiles >= 1024 : super_tiles == 2);
  for (int T = try_T2 ? 2 : 1; T >= 1; --T)
  for (int nslot = (2 * T * p.region1 <= 512 && p.num_tiles > 1 && max_slot >= 2 && two_ok) ? 2 : 1; nslot >= 1; --nslot) {
    if (T == 2 && nslot != 2) continue;
    p.T = T;
    p.region_cols = T * p.region1;
    const long long act = (long long)nslot * T * p.act_chunks * kChunkBytes;
    for (int na = (max_na < 2 ? 2 : (max_na > 4 ? 4 : max_na)); na >= 2; --na) {
      const long long rest = avail - act - (long long)na * kChunkBytes;
      if (rest < 0) continue;
      if (T == 2 && !(total_w <= rest && np <= kMaxPin && T * (p.kpad[0] / 64) <= na)) continue;
      int n_pinned = 0, nr = 0;
      long long pinned = 0;
      if (total_w <= rest && np <= kMaxPin) {
        n_pinned = np;
        pinned = total_w;
      } else {
        nr = want_nr < 2 ? 2 : (want_nr > kMaxRing ? kMaxRing : want_nr);
        long long budget = rest - (long long)nr * max_piece;
        if (budget < 0) continue;
        for (int li = 0; li < nl && n_pinned == p.first_piece[li]; ++li)
          for (int i = 0; i < p.pieces[li] && n_pinned < kMaxPin; ++i) {
            if (pinned + p.piece_bytes[li] > budget) break;
            pinned += p.piece_bytes[li];
            ++n_pinned;
          }
      }
      p.nslot = nslot;
      p.na = na;
      // cp.async groups a gather thread keeps in flight before it publishes the oldest chunk.  Measured (SA2: 80 us
      // at 1, 84 at 2, 92 at 3): publishing every chunk as soon as it lands beats deeper prefetch, because the MMA warp
      // consumes chunks in order and the four gather warps already overlap each other's round trips.
      // Exception: the wide point-wise stages (S == 1, K0 >= 384: FP layers) run one tile per CTA, nothing overlaps the
      // gather, and three groups in flight cut its eight serial L2 round trips (CTA span 20.6 -> 15.9 us).
      p.depth = (p.S == 1 && p.kpad[0] >= 384) ? 3 : 1;
      if (const char* e_d = sad_tool_env("SAD_MLP_DEPTH")) p.depth = atoi(e_d) < 1 ? 1 : (atoi(e_d) > 3 ? 3 : atoi(e_d));
      if (p.depth > na - 1) p.depth = na - 1;
      p.n_pinned = n_pinned;
      p.pinned_bytes = (int)pinned;
      p.nr = nr;
      p.ring_slot_bytes = nr ? max_piece : 0;
      long long off = 0;
      for (int li = 0; li < nl; ++li) {
        p.pin_off[li] = (int)off;
        off += (long long)p.pieces[li] * p.piece_bytes[li];
      }
      int cols = nslot * p.region_cols, pow2 = 32;
      while (pow2 < cols) pow2 <<= 1;
      p.tmem_cols = pow2;
      smem_out = (size_t)(1024 + act + (long long)na * kChunkBytes + pinned + (long long)nr * max_piece + kMiscBytes);
      return true;
    }
  }
  return false;
}

extern "C" int sad_shared_mlp_fwd(int B, int N, int P, int S, const void* feat_cl, int C0, const void* feat2_cl,
                                  int C1in, const float* xyz, const float* new_xyz, const int32_t* idx, float radius,
                                  const float* radius_t, int normalize_xyz, const float* extra, int E, int n_layers,
                                  const void* const* w_img, const float* const* bias, const int* c_out, int last_relu,
                                  void* out_cl_bf16, float* out_cf_f32, int* tile_counter, const sad_mlp_opts* opts,
                                  sad_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
#ifdef SAD_TOOLS_ABLATE
  if (sad_ablate_mask() & 2) return SAD_OK;
#endif
  SAD_REQUIRE(B >= 0 && N >= 1 && P >= 0 && S >= 1, "shared_mlp: bad sizes B=%d N=%d P=%d S=%d", B, N, P, S);
  SAD_REQUIRE(S == 1 || S == 2 || S == 4 || S == 8 || S == 16 || S == 32 || S == 64 || S == 128,
              "shared_mlp: nsample must be a power of two <= 128 (got %d)", S);
  SAD_REQUIRE(n_layers >= 2 && n_layers <= kMaxLayers, "shared_mlp: 2 or 3 layers supported (got %d)", n_layers);
  SAD_REQUIRE(w_img && bias && c_out, "shared_mlp: null layer tables");
  SAD_REQUIRE(C0 >= 0 && C0 % 64 == 0 && C1in >= 0 && C1in % 64 == 0, "shared_mlp: source widths must be multiples of 64");
  SAD_REQUIRE((C0 == 0) == (feat_cl == nullptr) && (C1in == 0) == (feat2_cl == nullptr), "shared_mlp: source/width mismatch");
  SAD_REQUIRE(E >= 0 && E <= 13 && (E == 0 || extra), "shared_mlp: 0..13 extra scalar features");
  SAD_REQUIRE(C1in == 0 || S == 1, "shared_mlp: the row-aligned second source needs S == 1");
  SAD_REQUIRE(idx || (S == 1 && N == P), "shared_mlp: identity rows need S == 1 and N == P");
  SAD_REQUIRE(out_cl_bf16 || out_cf_f32, "shared_mlp: no output requested");
  const int has_special = (xyz != nullptr || E > 0) ? 1 : 0;
  SAD_REQUIRE(!xyz || new_xyz, "shared_mlp: xyz needs new_xyz");
  SAD_REQUIRE(C0 + C1in + has_special > 0, "shared_mlp: no input");
  SAD_REQUIRE((long long)B * N < 0x7FFFFFFFLL && (long long)B * P < 0x7FFFFFFFLL, "shared_mlp: B*N and B*P must fit 31 bits");
  if (B == 0 || P == 0) return SAD_OK;

  MlpParams p = {};
  p.B = B; p.N = N; p.P = P; p.S = S;
  for (p.log2S = 0; (1 << p.log2S) < S; ++p.log2S) {}
  p.log2P = -1;
  for (int k = 0; k < 31; ++k)
    if ((1 << k) == P) p.log2P = k;
  p.total_rows = (long long)B * P * S;
  p.total_points = (uint32_t)((long long)B * P);
  const long long tiles = (p.total_rows + 127) / 128;
  SAD_REQUIRE(tiles < 0x7FFFFFFFLL, "shared_mlp: too many rows");
  p.num_tiles = (int)tiles;
  p.feat_cl = static_cast<const __nv_bfloat16*>(feat_cl); p.C0 = C0;
  p.feat2_cl = static_cast<const __nv_bfloat16*>(feat2_cl); p.C1in = C1in;
  p.xyz = xyz; p.new_xyz = new_xyz; p.idx = idx; p.radius_t = radius_t; p.radius = radius;
  p.normalize = normalize_xyz; p.extra = extra; p.E = E; p.has_special = has_special;
  p.n_layers = n_layers; p.last_relu = last_relu;
  p.out_cl = static_cast<__nv_bfloat16*>(out_cl_bf16); p.out_cf = out_cf_f32;
  p.tile_counter = tile_counter;
  int hidden_max = 0;
  for (int li = 0; li < n_layers; ++li) {
    SAD_REQUIRE(w_img[li] && bias[li] && c_out[li] >= 1, "shared_mlp: layer %d incomplete", li);
    p.c[li] = c_out[li];
    p.kpad[li] = (li == 0) ? (C0 + C1in + 64 * has_special) : c_out[li - 1];
    p.w_img[li] = static_cast<const uint8_t*>(w_img[li]);
    p.bias[li] = bias[li];
    if (li < n_layers - 1) {
      SAD_REQUIRE(c_out[li] % 64 == 0 && c_out[li] <= 256, "shared_mlp: hidden width %d must be a multiple of 64, <= 256",
                  c_out[li]);
      hidden_max = c_out[li] > hidden_max ? c_out[li] : hidden_max;
    }
  }
  SAD_REQUIRE(c_out[n_layers - 1] <= 512, "shared_mlp: at most 512 output channels");
  size_t smem = 0;
  const int super_tiles = opts ? opts->super_tiles : 0;
  SAD_REQUIRE(super_tiles >= 0 && super_tiles <= 2, "shared_mlp: opts.super_tiles must be 0 (auto), 1 (off) or 2 (on)");
  SAD_REQUIRE(plan_launch(p, hidden_max, super_tiles, smem), "shared_mlp: shared-memory / TMEM budget exceeded");

  int dev = 0, sms = 0;
  SAD_CUDA_OK(cudaGetDevice(&dev));
  SAD_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static thread_local int configured_dev = -1;
  if (configured_dev != dev) {
    SAD_CUDA_OK(cudaFuncSetAttribute(fused_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured_dev = dev;
  }
  if (sad_tool_env("SAD_DEBUG_MLP"))
    fprintf(stderr, "[sad] fused_mlp: rows=%lld tiles=%d S=%d K0=%d c=[%d,%d,%d] T=%d nslot=%d na=%d depth=%d pinned=%d/%d "
            "(%d B) ring=%dx%d tmem=%d smem=%zu\n", p.total_rows, p.num_tiles, S, p.kpad[0], p.c[0], p.c[1], p.c[2], p.T, p.nslot,
            p.na, p.depth, p.n_pinned, p.n_pieces, p.pinned_bytes, p.nr, p.ring_slot_bytes, p.tmem_cols, smem);
  // CTAs: one per SM, but never fewer than opts.tiles_per_cta tiles per CTA.  The per-CTA prologue -- TMEM allocation, barrier init, pinned weight loads -- is paid per
  // CTA, and under a pipelined caller a narrower grid leaves SMs to the other streams' kernels: the small stages take
  // longer alone but cost less SM-time (bench: 15.7k -> 16.5k scenes/s at 6 tiles per CTA).
  const int tpc = (opts && opts->tiles_per_cta > 0) ? opts->tiles_per_cta : 1;
  int grid = sad_ceil_div(p.num_tiles, tpc < 1 ? 1 : tpc);
  if (grid > sms) grid = sms;
  fused_mlp_kernel<<<grid, kThreads, smem, stream>>>(p);
  SAD_LAUNCH_CHECK("fused_mlp_kernel");
  return SAD_OK;
}

extern "C" int sad_three_interpolate_cl_fwd(int B, int C, int m, int n, const void* feat_cl_bf16, const int32_t* idx,
                                            const float* weight, void* out_cl_bf16, sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && m >= 1 && n >= 0 && C >= 8 && C % 8 == 0, "three_interpolate_cl: bad sizes (C must be a multiple of 8)");
  if (B == 0 || n == 0) return SAD_OK;
  SAD_REQUIRE(feat_cl_bf16 && idx && weight && out_cl_bf16, "three_interpolate_cl: null pointer");
  const long long rows = (long long)B * n;
  interp_cl_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      rows, n, m, C, static_cast<const __nv_bfloat16*>(feat_cl_bf16), idx, weight,
      static_cast<__nv_bfloat16*>(out_cl_bf16));
  SAD_LAUNCH_CHECK("interp_cl_kernel");
  return SAD_OK;
}

extern "C" int sad_cf_to_cl_bf16(int B, int C, int N, const float* in_cf, void* out_cl_bf16, sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && C >= 1 && N >= 1, "cf_to_cl: bad sizes");
  if (B == 0) return SAD_OK;
  SAD_REQUIRE(in_cf && out_cl_bf16 && B <= 65535, "cf_to_cl: null pointer / batch too large");
  dim3 grid((unsigned)sad_ceil_div(N, 32), (unsigned)sad_ceil_div(C, 32), (unsigned)B);
  cf_to_cl_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(C, N, in_cf, static_cast<__nv_bfloat16*>(out_cl_bf16));
  SAD_LAUNCH_CHECK("cf_to_cl_kernel");
  return SAD_OK;
}
