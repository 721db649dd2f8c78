// a6 (tf32 mode)  fused gather / interpolation -> shared MLP -> max-pool with fp32 activations and kind::tf32 MMAs
// -- SURVEY.md section 8(a) row a6, BASELINE.json north_star ("tcgen05 bf16/tf32 GEMM with fused max-pool epilogue"),
// VERDICT r1 item 6.  (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
// Same fusion as the bf16 kernels (mlp_sa.cu / mlp_pw.cu), one general kernel instead of shape-specialised ones:
// features travel between stages as fp32 channel-last rows, the layer-1 operand is built in shared memory straight
// from HBM (neighbour gather through idx, identity rows, or the three-nearest-neighbour interpolation itself, plus the
// relative-xyz / scalar-feature K step), every layer's activations stay on chip in fp32 and the last layer is evaluated
// transposed (lane = output channel) so the max over nsample is an in-register reduction.  Operands are rounded to
// tf32 (cvt.rna, 10-bit mantissa) when they are written; accumulation, bias, ReLU and max are fp32.
//
//   shared memory   activations 128 rows x widest hidden layer x 4 B (<= 128 KB; K-major SWIZZLE_128B, 32-channel
//                   chunks of 16 KB)
//                   W ring  2-8 stages of one weight piece (<= 256 rows x 32 K values = 32 KB), one TMA bulk copy
//                   each: the pieces are re-streamed from L2 for every tile, the ring depth is the prefetch distance
//                   A ring  2-4 x 16 KB   layer-1 operand chunks (128 rows x 32 channels)
//   TMEM            per tile context max(widest hidden layer, 128 x last-layer blocks) columns: hidden accumulators and
//                   the last layer's 128-channel x 128-row blocks share them (a layer's accumulators are drained
//                   before the next layer's MMAs are issued)
//   contexts        TWO tiles in flight per CTA when two activation buffers and 2 x those columns fit (all SA stages):
//                   the roles walk (layer, context) in the same order, so one context's epilogue runs under the other's
//                   MMAs; the 256-wide point-wise stages (FP, voting) run one
//   warps           0-15 epilogue (4 column groups x 4 TMEM lane quarters), 16-19 operand producers (thread = row),
//                   20 MMA issue, 21 weight TMA
//
// Persistent CTAs take 128-row tiles (row = (query point, sample)) round-robin.
// Roofline: tensor (kind::tf32 peak = half the bf16 peak); algorithmic flops = 2 * rows * sum(Cin*Cout).
#include <string.h>

#include "sad_tc.cuh"

namespace {

using namespace sad;

constexpr int kChunk = 16384;                 // 128 rows x 128 B = 32 fp32 channels per row
constexpr int kMaxNSA = 4, kMaxNSW = 8;       // ring depths are chosen per launch from what the activations leave free
constexpr int kMisc = 1024;
constexpr int kSmemMax = 227 * 1024;
constexpr int kEpiWarps = 16, kWarpProd = 16, kWarpMma = 20, kWarpTma = 21, kThreads = 22 * 32;
constexpr int kMaxLayers = 3;

struct TfParams {
  int B, N, P, S, log2S;             // rows = B * P * S; source points per scene N (gather) / P (identity rows)
  long long total_rows, total_points;
  int num_tiles;
  // layer-1 operand, K order [interp | feat | special]
  const float* known_cl;             // (B, m, CI) fp32 channel-last, interpolation source, or null
  const int32_t* nn_idx;             // (B*P, 3)
  const float* nn_w;                 // (B*P, 3)
  int m, CI;
  const float* feat_cl;              // (B, N, CF) fp32 channel-last gathered through idx, or identity rows when idx == null
  int CF;
  const int32_t* idx;                // (B, P, S) or null
  const float* xyz;                  // (B, N, 3): the special K step [dx,dy,dz,e0..e(E-1),0..] exists when non-null
  const float* new_xyz;              // (B, P, 3)
  const float* radius_t;             // (B, P) per-cluster radius or null
  float radius;
  int normalize;
  const float* extra;                // (B, N, E) fp32 scalar features, E <= 4
  int E;
  // layers
  int n_layers;
  int c_out[kMaxLayers];             // real widths; hidden widths are multiples of 32 and <= 256
  int kc[kMaxLayers];                // K chunks per layer (layer 1: interp + feat + special chunks)
  const uint8_t* w_img[kMaxLayers];  // pieces in consumption order (see sad_mlp_tf32_pack)
  const float* bias[kMaxLayers];     // last layer: padded to 128 * nblk
  int nblk;                          // 128-channel blocks of the last layer (<= 4)
  int last_relu;
  float* out_cf;                     // (B, c_last, P) f32 or null
  float* out_cl;                     // (B, P, c_last) f32 or null
  // shared-memory carve-up
  int act_bytes, nsa, nsw, wstage;
  // tile contexts in flight per CTA (1 or 2: own activation buffer and TMEM columns each) and TMEM columns of one
  int nctx, ctx_cols;
};

struct Misc {
  uint64_t wfull[kMaxNSW], wfree[kMaxNSW], afull[kMaxNSA], afree[kMaxNSA];
  uint64_t dfull[2][1 + 4];          // per context: [0] hidden layers, [1 + blk] last-layer blocks
  uint64_t actfull[2], tfree[2];
  uint32_t tmem_base;
};

__device__ __noinline__ void bar_timeout(uint32_t bar, uint32_t parity) {
  printf("mlp_tf32: barrier %u parity %u timed out (block %d warp %d)\n", bar, parity, blockIdx.x, threadIdx.x >> 5);
  __trap();
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  long long t0 = 0;
  for (;;) {
#ifdef SAD_TF32_SPIN
    if (mbar_try_wait(bar, parity)) return;
#else
    if (mbar_try_wait_sleep(bar, parity)) return;
#endif
    if ((++spins & 0xFFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) bar_timeout(smem_u32(bar), parity);
    }
  }
}
#ifdef SAD_TF32_PROFILE
#define TFP_DECL long long tfp_wait[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const long long tfp_t0 = clock64();
#define TFP_WAIT(slot, bar, par) do { const long long t_ = clock64(); bar_wait(bar, par); tfp_wait[slot] += clock64() - t_; } while (0)
#define TFP_T(var) const long long var = clock64()
#define TFP_ADD(slot, a, b) tfp_wait[slot] += (b) - (a)
#define TFP_REPORT(role) do { if (blockIdx.x == 0 && lane == 0) printf("tf32 %s total %lld waits %lld %lld %lld %lld %lld %lld | %lld %lld %lld %lld tiles %d\n", role, clock64() - tfp_t0, tfp_wait[0], tfp_wait[1], tfp_wait[2], tfp_wait[3], tfp_wait[4], tfp_wait[5], tfp_wait[6], tfp_wait[7], tfp_wait[8], tfp_wait[9], my_tiles); } while (0)
#else
#define TFP_DECL
#define TFP_WAIT(slot, bar, par) bar_wait(bar, par)
#define TFP_T(var)
#define TFP_ADD(slot, a, b)
#define TFP_REPORT(role)
#endif

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor kind::tf32: D f32 (bit 4), A = B = tf32 (format 2 at bits 7 and 10), both K-major
__device__ __forceinline__ uint32_t uidesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Output cursor of one lane (= one output channel) of the transposed last layer.
struct OutCursor {
  float* pcl;          // out_cl + pt * c_last + ch, or null
  float* pcf;          // out_cf + (b * c_last + ch) * P + p, or null
  long long rem;       // outputs this lane may still write (0 for padded channels / past the end)
  int op, P, c_last;
  size_t cf_wrap;      // (c_last - 1) * P: from the end of one scene's row to the next scene's row of this channel
  float bias;
  int relu;
  __device__ __forceinline__ void emit(float mx) {
    float y = __fadd_rn(mx, bias);
    if (relu) y = fmaxf(y, 0.f);
    if (rem > 0) {
      if (pcl) *pcl = y;
      if (pcf) *pcf = y;
    }
    --rem;
    if (pcl) pcl += c_last;
    if (pcf) {
      ++pcf;
      if (++op == P) {
        op = 0;
        pcf += cf_wrap;
      }
    }
  }
};
// One 32-column group of the transposed last layer: max over windows of S columns (compile-time S: the reduction is
// an in-register tree, no per-column branch), one emit per finished window.
template <int S>
__device__ __forceinline__ void pool_group(const uint32_t (&v)[32], float& mx, bool window_ends, OutCursor& oc) {
  if constexpr (S <= 32) {
#pragma unroll
    for (int k = 0; k < 32 / S; ++k) oc.emit(vmax_tree<S>(v + k * S));
  } else {
    mx = fmaxf(mx, vmax_tree<32>(v));
    if (window_ends) {
      oc.emit(mx);
      mx = -INFINITY;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) mlp_tf32_kernel(const TfParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* act = base;
  const int kNSA = p.nsa, kNSW = p.nsw, kWStage = p.wstage;
  uint8_t* aring = act + p.nctx * p.act_bytes;
  uint8_t* wring = aring + kNSA * kChunk;
  Misc* ms = reinterpret_cast<Misc*>(wring + kNSW * kWStage);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NL = p.n_layers;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxNSW; ++i) { mbar_init(&ms->wfull[i], 1); mbar_init(&ms->wfree[i], 1); }
    for (int i = 0; i < kMaxNSA; ++i) { mbar_init(&ms->afull[i], 4); mbar_init(&ms->afree[i], 1); }
    for (int x = 0; x < 2; ++x) {
      for (int i = 0; i < 5; ++i) mbar_init(&ms->dfull[x][i], 1);
      mbar_init(&ms->actfull[x], kEpiWarps);
      mbar_init(&ms->tfree[x], kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<1>(&ms->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = ms->tmem_base;
  const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int kcI = (p.CI + 31) >> 5, kcF = (p.CF + 31) >> 5;
  const bool special = p.xyz != nullptr;
  // Tiles are taken in rounds of nctx: tile ordinal t = round * nctx + context.  Within a round the roles walk
  // (layer 1, ctx 0), (layer 1, ctx 1), (layer 2, ctx 0), ... so one context's epilogue runs under the other's MMAs.
  const int NX = p.nctx;
  const int rounds = (my_tiles + NX - 1) / NX;
  TFP_DECL

  if (warp == kWarpTma) {
    // ============================================================ weight pieces through the W ring
    if (lane == 0) {
      uint32_t w = 0;
      for (int r = 0; r < rounds; ++r) {
        for (int l = 0; l < NL; ++l) {
          const bool last = l == NL - 1;
          const int pieces = last ? p.nblk * p.kc[l] : p.kc[l];
          const uint32_t bytes = last ? 128u * 128u : (uint32_t)((p.c_out[l] + 15) & ~15) * 128u;
          const uint8_t* src = p.w_img[l];
          for (int x = 0; x < NX; ++x) {
            if (r * NX + x >= my_tiles) break;
            for (int i = 0; i < pieces; ++i, ++w) {
              const int st = w % kNSW;
              TFP_WAIT(0, &ms->wfree[st], ((w / kNSW) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&ms->wfull[st], bytes);
              tma_bulk_g2s(wring + st * kWStage, src + (size_t)i * bytes, bytes, &ms->wfull[st]);
            }
          }
        }
      }
      TFP_REPORT("loader  ");
    }
  } else if (warp == kWarpMma) {
    // ============================================================ MMA issue
    uint32_t w = 0, a = 0;
    uint32_t n_act[2] = {0, 0};
    for (int r = 0; r < rounds; ++r) {
      for (int l = 0; l < NL; ++l) {
        const bool last = l == NL - 1;
        const int KC = p.kc[l];
        const int nb = last ? p.nblk : 1;
        const uint32_t idesc = last ? uidesc_tf32(128, 128) : uidesc_tf32(128, (p.c_out[l] + 15) & ~15);
        for (int x = 0; x < NX; ++x) {
          if (r * NX + x >= my_tiles) break;
          const uint32_t act_addr = smem_u32(act) + (uint32_t)(x * p.act_bytes);
          const uint32_t tctx = tmem + (uint32_t)(x * p.ctx_cols);
          // the context's TMEM columns are reused by every layer: its previous tile must have been drained
          if (l == 0 && r > 0) TFP_WAIT(3, &ms->tfree[x], (uint32_t)(r - 1) & 1u);
          if (l > 0) TFP_WAIT(2, &ms->actfull[x], (n_act[x]++) & 1u);
          for (int blk = 0; blk < nb; ++blk) {
            const uint32_t d = tctx + (last ? 128u * blk : 0u);
            for (int c = 0; c < KC; ++c, ++w) {
              const int ws = w % kNSW;
              TFP_WAIT(0, &ms->wfull[ws], (w / kNSW) & 1u);
              int as = 0;
              int ksteps = 4;
              if (l == 0) {
                as = a % kNSA;
                TFP_WAIT(1, &ms->afull[as], (a / kNSA) & 1u);
                if (special && c == KC - 1) ksteps = 1;
              }
              tc_fence_after_sync();
              if (elect_one()) {
                const uint32_t wa = smem_u32(wring + ws * kWStage);
                const uint32_t xa = l == 0 ? smem_u32(aring + as * kChunk) : act_addr + (uint32_t)c * kChunk;
                const uint64_t dw = udesc_sw128(wa), dx = udesc_sw128(xa);
                for (int k = 0; k < ksteps; ++k) {
                  // hidden: D[rows, cout] = X . W^T; last: D[cout, rows] = W . X^T
                  if (last) umma_tf32(d, dw + 2 * k, dx + 2 * k, idesc, (c | k) != 0);
                  else umma_tf32(d, dx + 2 * k, dw + 2 * k, idesc, (c | k) != 0);
                }
                umma_commit_to<1>(&ms->wfree[ws]);
                if (l == 0) umma_commit_to<1>(&ms->afree[as]);
                if (c == KC - 1) umma_commit_to<1>(&ms->dfull[x][last ? 1 + blk : 0]);
              }
              __syncwarp();
              if (l == 0) ++a;
            }
          }
        }
      }
    }
    TFP_REPORT("mma     ");
  } else if (warp >= kWarpProd && warp < kWarpProd + 4) {
    // ============================================================ layer-1 operand producers: thread = row of the tile
    const int r = (warp - kWarpProd) * 32 + lane;
    uint32_t a = 0;
    const long long PS = (long long)p.P * p.S;
    for (int t = 0; t < my_tiles; ++t) {
      const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
      const long long row = tile * 128 + r;
      const bool live = row < p.total_rows;
      int b = 0, j = 0;
      long long pt = 0;
      if (live) {
        b = (int)(row / PS);
        pt = row >> p.log2S;
        j = p.idx ? __ldg(p.idx + row) : (int)(row - (long long)b * PS);
      }
      const int KC = p.kc[0];
      for (int c = 0; c < KC; ++c, ++a) {
        const int st = a % kNSA;
        TFP_WAIT(1, &ms->afree[st], ((a / kNSA) & 1u) ^ 1u);
        const uint32_t dst = smem_u32(aring + st * kChunk);
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
        if (live) {
          if (c < kcI) {
            const int c0 = c * 32;
            const int i0 = __ldg(p.nn_idx + pt * 3), i1 = __ldg(p.nn_idx + pt * 3 + 1), i2 = __ldg(p.nn_idx + pt * 3 + 2);
            const float w0 = __ldg(p.nn_w + pt * 3), w1 = __ldg(p.nn_w + pt * 3 + 1), w2 = __ldg(p.nn_w + pt * 3 + 2);
            const float* k0 = p.known_cl + ((size_t)b * p.m + i0) * p.CI + c0;
            const float* k1 = p.known_cl + ((size_t)b * p.m + i1) * p.CI + c0;
            const float* k2 = p.known_cl + ((size_t)b * p.m + i2) * p.CI + c0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (c0 + u * 4 < p.CI) {
                const float4 f0 = __ldg(reinterpret_cast<const float4*>(k0) + u);
                const float4 f1 = __ldg(reinterpret_cast<const float4*>(k1) + u);
                const float4 f2 = __ldg(reinterpret_cast<const float4*>(k2) + u);
                // ((w0*f0)+(w1*f1))+(w2*f2), one rounding per operation: the oracle's three_interpolate
                v[u * 4 + 0] = tf32_rna(__fadd_rn(__fadd_rn(__fmul_rn(w0, f0.x), __fmul_rn(w1, f1.x)), __fmul_rn(w2, f2.x)));
                v[u * 4 + 1] = tf32_rna(__fadd_rn(__fadd_rn(__fmul_rn(w0, f0.y), __fmul_rn(w1, f1.y)), __fmul_rn(w2, f2.y)));
                v[u * 4 + 2] = tf32_rna(__fadd_rn(__fadd_rn(__fmul_rn(w0, f0.z), __fmul_rn(w1, f1.z)), __fmul_rn(w2, f2.z)));
                v[u * 4 + 3] = tf32_rna(__fadd_rn(__fadd_rn(__fmul_rn(w0, f0.w), __fmul_rn(w1, f1.w)), __fmul_rn(w2, f2.w)));
              }
            }
          } else if (c < kcI + kcF) {
            const int c0 = (c - kcI) * 32;
            const float* src = p.feat_cl + ((size_t)b * p.N + j) * p.CF + c0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (c0 + u * 4 < p.CF) {
                const float4 f = __ldg(reinterpret_cast<const float4*>(src) + u);
                v[u * 4 + 0] = tf32_rna(f.x);
                v[u * 4 + 1] = tf32_rna(f.y);
                v[u * 4 + 2] = tf32_rna(f.z);
                v[u * 4 + 3] = tf32_rna(f.w);
              }
            }
          } else {      // special K step: relative (optionally radius-normalised) xyz, then the scalar features
            const float* q = p.xyz + ((size_t)b * p.N + j) * 3;
            const float* o = p.new_xyz + (size_t)pt * 3;
            float dx = __fsub_rn(__ldg(q), __ldg(o)), dy = __fsub_rn(__ldg(q + 1), __ldg(o + 1)), dz = __fsub_rn(__ldg(q + 2), __ldg(o + 2));
            if (p.normalize) {
              const float rad = p.radius_t ? __ldg(p.radius_t + pt) : p.radius;
              dx = __fdiv_rn(dx, rad);
              dy = __fdiv_rn(dy, rad);
              dz = __fdiv_rn(dz, rad);
            }
            v[0] = tf32_rna(dx);
            v[1] = tf32_rna(dy);
            v[2] = tf32_rna(dz);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (e < p.E) v[3 + e] = tf32_rna(__ldg(p.extra + ((size_t)b * p.N + j) * p.E + e));
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) sts_v4(dst + swz128(r, u), v[u * 4], v[u * 4 + 1], v[u * 4 + 2], v[u * 4 + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ms->afull[st]);
      }
    }
    if (warp == kWarpProd) TFP_REPORT("producer");
  } else if (warp < kEpiWarps) {
    // ============================================================ epilogue: 16 warps = 4 column groups x 4 lane quarters
    const int eg = warp >> 2, q = warp & 3;       // a warp may only read TMEM lanes 32 * (warp % 4) .. + 31
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int row = q * 32 + lane;
    uint32_t n_d0[2] = {0, 0};
    const int c_last = p.c_out[NL - 1];
    // last layer: a unit = the 32-column groups one max-pool window spans (1 for nsample <= 32, 2 for 64, 4 for 128)
    const int ug = p.S <= 32 ? 1 : p.S >> 5;
    const int upb = 4 / ug;
    const int items = p.nblk * upb;
    for (int r = 0; r < rounds; ++r) {
      for (int l = 0; l < NL - 1; ++l) {
       for (int x = 0; x < NX; ++x) {
        if (r * NX + x >= my_tiles) break;
        const uint32_t lane_x = lane_addr + (uint32_t)(x * p.ctx_cols);
        const uint32_t act_x = smem_u32(act) + (uint32_t)(x * p.act_bytes);
        TFP_WAIT(4, &ms->dfull[x][0], (n_d0[x]++) & 1u);
        tc_fence_after_sync();
        const int groups = p.c_out[l] >> 5;
        const float* bias = p.bias[l];
        for (int g = eg; g < groups; g += 4) {
          uint32_t v[32];
          TFP_T(e0);
          tmem_ld_x32(lane_x + (uint32_t)(g * 32), v);
          float bv[32];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + g * 32) + u);
            bv[4 * u] = b4.x; bv[4 * u + 1] = b4.y; bv[4 * u + 2] = b4.z; bv[4 * u + 3] = b4.w;
          }
          tmem_ld_fence();
          TFP_T(e1);
          TFP_ADD(0, e0, e1);
          const uint32_t dst = act_x + (uint32_t)g * kChunk;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
              o[k] = tf32_rna(fmaxf(__fadd_rn(__uint_as_float(v[u * 4 + k]), bv[u * 4 + k]), 0.f));
            sts_v4(dst + swz128(row, u), o[0], o[1], o[2], o[3]);
          }
          TFP_T(e2);
          TFP_ADD(1, e1, e2);
        }
        TFP_T(e3);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ms->actfull[x]);
        TFP_T(e4);
        TFP_ADD(2, e3, e4);
       }
      }
      // last layer, transposed: lane = output channel, TMEM column = row of the tile
     for (int x = 0; x < NX; ++x) {
      const int t = r * NX + x;
      if (t >= my_tiles) break;
      const uint32_t lane_x = lane_addr + (uint32_t)(x * p.ctx_cols);
      const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
      const long long pt0 = (tile * 128) >> p.log2S;       // first point of the tile (128 % S == 0)
      for (int blk = 0; blk < p.nblk; ++blk) {
        TFP_WAIT(5, &ms->dfull[x][1 + blk], (uint32_t)r & 1u);
      }
      tc_fence_after_sync();
      TFP_T(e5);
      for (int item = eg; item < items; item += 4) {
        const int blk = item / upb, g0 = (item - blk * upb) * ug;
        const int ch = blk * 128 + row;
        const float bias = __ldg(p.bias[NL - 1] + ch);
        const bool ch_ok = ch < c_last;
        TFP_T(f0);
        // (scene, point) of the unit's first output, advanced incrementally: one division per unit.  The emit is kept
        // to a dozen instructions because it is inlined 32 times (instruction-cache footprint of the column loop).
        const long long pt = pt0 + ((g0 * 32) >> p.log2S);
        const int ob = (int)(pt / p.P);
        OutCursor oc;
        oc.op = (int)(pt - (long long)ob * p.P);
        oc.P = p.P;
        oc.c_last = c_last;
        oc.rem = ch_ok ? p.total_points - pt : 0;
        oc.pcl = p.out_cl ? p.out_cl + (size_t)pt * c_last + ch : nullptr;
        oc.pcf = p.out_cf ? p.out_cf + ((size_t)ob * c_last + ch) * p.P + oc.op : nullptr;
        oc.cf_wrap = (size_t)(c_last - 1) * p.P;
        oc.bias = bias;
        oc.relu = p.last_relu;
        float mx = -INFINITY;
        TFP_T(f1);
        TFP_ADD(6, f0, f1);
        for (int g = g0; g < g0 + ug; ++g) {
          uint32_t v[32];
          TFP_T(f2);
          tmem_ld_x32(lane_x + (uint32_t)(blk * 128 + g * 32), v);
          tmem_ld_fence();
          TFP_T(f3);
          TFP_ADD(7, f2, f3);
          const bool ends = g == g0 + ug - 1;
          switch (p.log2S) {
            case 0: pool_group<1>(v, mx, ends, oc); break;
            case 1: pool_group<2>(v, mx, ends, oc); break;
            case 2: pool_group<4>(v, mx, ends, oc); break;
            case 3: pool_group<8>(v, mx, ends, oc); break;
            case 4: pool_group<16>(v, mx, ends, oc); break;
            case 5: pool_group<32>(v, mx, ends, oc); break;
            default: pool_group<64>(v, mx, ends, oc); break;
          }
        }
      }
      TFP_T(e6);
      TFP_ADD(3, e5, e6);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->tfree[x]);
     }
    }
    if (warp == 0) TFP_REPORT("epilogue");
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem, 512);
}

int fail(int code, const char* msg) {
  sad_set_error("%s", msg);
  return code;
}

int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI
// Weight image of ONE layer in the order the kernel consumes it.  W is (c_out, c_in) row-major fp32 with the input
// channels already in operand order and zero-padded by the caller to kc * 32 columns (c_in == kc * 32).
//   hidden layer  kc pieces, piece = pad16(c_out) rows x 32 K values, K-major SWIZZLE_128B
//   last layer    nblk * kc pieces (block-major), piece = 128 rows (output channels, zero padded) x 32 K values
// Values are rounded to tf32 (round to nearest, ties away -- the cvt.rna the kernel applies to activations).
extern "C" long long sad_mlp_tf32_image_bytes(int c_out, int kc, int last) {
  if (c_out <= 0 || kc <= 0) return SAD_EINVAL;
  if (last) return (long long)((c_out + 127) / 128) * kc * 128 * 128;
  return (long long)kc * ((c_out + 15) & ~15) * 128;
}

extern "C" int sad_mlp_tf32_pack(const float* W, int c_out, int kc, int last, void* out) {
  if (!W || !out || c_out <= 0 || kc <= 0) return fail(SAD_EINVAL, "mlp_tf32_pack: bad arguments");
  const int c_in = kc * 32;
  const long long bytes = sad_mlp_tf32_image_bytes(c_out, kc, last);
  memset(out, 0, (size_t)bytes);
  uint8_t* o = static_cast<uint8_t*>(out);
  const int rows_piece = last ? 128 : ((c_out + 15) & ~15);
  const int nblk = last ? (c_out + 127) / 128 : 1;
  for (int blk = 0; blk < nblk; ++blk)
    for (int c = 0; c < kc; ++c) {
      uint8_t* piece = o + (size_t)(blk * kc + c) * rows_piece * 128;
      for (int r = 0; r < rows_piece; ++r) {
        const int n = blk * 128 + r;
        if (n >= c_out) continue;
        for (int k = 0; k < 32; ++k) {
          float w = W[(size_t)n * c_in + c * 32 + k];
          uint32_t u;
          memcpy(&u, &w, 4);
          if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & ~0x1FFFu;      // rna to 10 mantissa bits
          const size_t off = (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + (size_t)(((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4;
          memcpy(piece + off, &u, 4);
        }
      }
    }
  return SAD_OK;
}

extern "C" int sad_mlp_tf32_fwd(int B, int N, int P, int S, const float* known_cl, int m, int CI, const int32_t* nn_idx,
                                const float* nn_w, const float* feat_cl, int CF, const int32_t* idx, const float* xyz,
                                const float* new_xyz, float radius, const float* radius_t, int normalize_xyz,
                                const float* extra, int E, int n_layers, const void* const* w_img, const float* const* bias,
                                const int* c_out, int last_relu, float* out_cf, float* out_cl, sad_stream_t stream) {
  if (B <= 0 || P <= 0 || S <= 0 || N <= 0) return fail(SAD_EINVAL, "mlp_tf32: bad shape");
  if (n_layers < 2 || n_layers > kMaxLayers) return fail(SAD_EUNSUPPORTED, "mlp_tf32: 2-3 layers");
  if ((S & (S - 1)) != 0 || S > 128) return fail(SAD_EUNSUPPORTED, "mlp_tf32: nsample must be a power of two <= 128");
  if (!w_img || !bias || !c_out || (!out_cf && !out_cl)) return fail(SAD_EINVAL, "mlp_tf32: null pointer");
  if ((CI & 3) || (CF & 3) || CI < 0 || CF < 0) return fail(SAD_EINVAL, "mlp_tf32: channel counts must be multiples of 4");
  if (CI && (!known_cl || !nn_idx || !nn_w || m < 3 || S != 1)) return fail(SAD_EINVAL, "mlp_tf32: interpolation source");
  if (CF && !feat_cl) return fail(SAD_EINVAL, "mlp_tf32: feat_cl is null");
  if (!idx && (S != 1 || N != P)) return fail(SAD_EINVAL, "mlp_tf32: identity rows need nsample == 1 and N == P");
  if (xyz && (!new_xyz || !idx || E < 0 || E > 4 || (E && !extra))) return fail(SAD_EINVAL, "mlp_tf32: special K step");
  if (xyz && normalize_xyz && !radius_t && !(radius > 0.f)) return fail(SAD_EINVAL, "mlp_tf32: radius");
  for (int l = 0; l < n_layers - 1; ++l)
    if (c_out[l] <= 0 || c_out[l] > 256 || (c_out[l] & 31)) return fail(SAD_EUNSUPPORTED, "mlp_tf32: hidden widths % 32 == 0, <= 256");
  if (c_out[n_layers - 1] <= 0 || c_out[n_layers - 1] > 512) return fail(SAD_EUNSUPPORTED, "mlp_tf32: last width <= 512");
  TfParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.N = N; p.P = P; p.S = S; p.log2S = ilog2(S);
  p.total_points = (long long)B * P;
  p.total_rows = p.total_points * S;
  p.num_tiles = (int)((p.total_rows + 127) / 128);
  p.known_cl = CI ? known_cl : nullptr; p.nn_idx = nn_idx; p.nn_w = nn_w; p.m = m; p.CI = CI;
  p.feat_cl = feat_cl; p.CF = CF; p.idx = idx;
  p.xyz = xyz; p.new_xyz = new_xyz; p.radius_t = radius_t; p.radius = radius; p.normalize = normalize_xyz;
  p.extra = extra; p.E = xyz ? E : 0;
  p.n_layers = n_layers;
  const int k1 = (CI + 31) / 32 + (CF + 31) / 32 + (xyz ? 1 : 0);
  if (k1 == 0) return fail(SAD_EINVAL, "mlp_tf32: empty layer-1 operand");
  for (int l = 0; l < n_layers; ++l) {
    p.c_out[l] = c_out[l];
    p.kc[l] = l == 0 ? k1 : c_out[l - 1] / 32;
    p.w_img[l] = static_cast<const uint8_t*>(w_img[l]);
    p.bias[l] = bias[l];
    if (!w_img[l] || !bias[l]) return fail(SAD_EINVAL, "mlp_tf32: null layer");
  }
  p.nblk = (c_out[n_layers - 1] + 127) / 128;
  p.last_relu = last_relu;
  p.out_cf = out_cf; p.out_cl = out_cl;
  // shared memory: activations sized by the widest hidden layer; what is left goes to the weight ring (every piece
  // is re-streamed from L2 per tile, so its depth is the prefetch distance), then to the layer-1 operand ring
  int max_h = 0, piece = 16384;
  for (int l = 0; l < n_layers - 1; ++l) {
    if (c_out[l] > max_h) max_h = c_out[l];
    const int b = ((c_out[l] + 15) & ~15) * 128;
    if (b > piece) piece = b;
  }
  p.act_bytes = (max_h / 32) * kChunk;
  p.wstage = (piece + 1023) & ~1023;
  const int usable = kSmemMax - 1024 - kMisc;
  // two tile contexts when two activation buffers leave room for the rings and both fit the 512 TMEM columns: one
  // context's epilogue then runs under the other's MMAs
  p.ctx_cols = max_h > 128 * p.nblk ? max_h : 128 * p.nblk;
  p.nctx = (2 * p.ctx_cols <= 512 && usable - 2 * p.act_bytes - 2 * kChunk >= 3 * p.wstage && p.num_tiles > 1) ? 2 : 1;
  if (p.ctx_cols > 512) return fail(SAD_EUNSUPPORTED, "mlp_tf32: last layer too wide for TMEM");
  const int act_total = p.nctx * p.act_bytes;
  p.nsw = (usable - act_total - 2 * kChunk) / p.wstage;
  if (p.nsw > kMaxNSW) p.nsw = kMaxNSW;
  if (p.nsw < 2) return fail(SAD_EUNSUPPORTED, "mlp_tf32: layers too wide for shared memory");
  p.nsa = 2 + (usable - act_total - p.nsw * p.wstage - 2 * kChunk) / kChunk;
  if (p.nsa > kMaxNSA) p.nsa = kMaxNSA;
  const int kSmem = 1024 + act_total + p.nsa * kChunk + p.nsw * p.wstage + kMisc;
  static bool configured = false;
  if (!configured) {
    SAD_CUDA_OK(cudaFuncSetAttribute(mlp_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    configured = true;
  }
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  mlp_tf32_kernel<<<grid, kThreads, kSmem, (cudaStream_t)stream>>>(p);
  SAD_LAUNCH_CHECK("mlp_tf32_kernel");
  return SAD_OK;
}
