// a6 / a9 (fast path)  shape-specialised fused POINT-WISE stage (nsample == 1): the FP modules and the voting module
// -- SURVEY.md section 8(a) rows a6 / a9, section 3 call stack 2; VERDICT r1 items 1(b,c), 4, 5.
// (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
//   FP module      [three_interpolate(known) | skip] (512) -> 256 -> 256          <K0C = 8, NL = 2, NBLK = 2>
//   voting module  seed features (256) -> 256 -> 256 -> 3 + 256, vote = seed + y  <K0C = 4, NL = 3, NBLK = 3>
//
// 256-wide layers: the weights (384 KB) cannot be pinned, so they STREAM through a ring of 48 KB stages
// (one 64-wide K chunk of the layer-1 operand, 16 KB, plus one weight piece, <= 32 KB), one TMA bulk copy per piece,
// while the tile's activations stay in a 64 KB buffer.  One 128-row tile is in flight per CTA (these stages have 32-64
// tiles: one per CTA), so what matters is the length of the tile's chain:
//   * 16 epilogue warps split every accumulator four ways by columns (64 columns per warp: the hidden-layer drain is
//     ~0.5 k cycles instead of ~2.6 k on one warpgroup);
//   * the last layer is evaluated TRANSPOSED (D^T = W . H^T, lane = output channel, column = row), so the
//     channel-first f32 output is 32 consecutive floats per thread and block, the channel-last bf16 twin 64 B per warp;
//   * the layer-1 operand is built by ALL 20 worker warps (the 16 epilogue warps have nothing else to do while the
//     tile's first layer is being fed) straight in the swizzled operand layout: skip / seed rows by
//     16-byte cp.async, and -- FP modules -- the three-nearest-neighbour INTERPOLATION itself (3 weighted rows of the
//     known features, fp32 math, bf16 operand): three_interpolate never runs as a kernel and its output never exists;
//   * the voting module's epilogue adds the seed coordinates / features (vote = seed + y) and writes vote_xyz,
//     vote_features (f32 channel-first) and their bf16 channel-last twin directly.
// The generic -> async proxy fence for thread-written operands is on the consumer side (see mlp_sa.cu).
#include <string.h>

#include "sad_tc.cuh"

namespace {

using namespace sad;

constexpr int kWorkers = 20 * 32;        // warps 0-19 build the layer-1 operand; warps 0-15 are also the epilogue
constexpr int kWarpMma = 20;
constexpr int kWarpTma = 21;             // weight pieces
constexpr int kThreads = 22 * 32;
constexpr int kChunk = 16384;            // 128 rows x 128 B
constexpr int kStages = 3;
constexpr int kStageBytes = kChunk + 2 * kChunk;      // operand chunk + weight piece (256 rows x 128 B)
constexpr int kActBytes = 4 * kChunk;                 // 128 rows x 256 channels bf16
constexpr int kMisc = 12288;
constexpr int kSmem = 1024 + kStages * kStageBytes + kActBytes + kMisc;
constexpr int H = 256, HC = 4;
static_assert(kSmem <= 227 * 1024, "shared-memory budget");

struct PwParams {
  int n, m;                       // rows per batch element (multiple of 128), known points per batch element (interp)
  long long total_rows;
  int num_tiles;
  const __nv_bfloat16* src_cl;    // (rows, src_chunks * 64) bf16 channel-last: skip / seed features (identity rows)
  int src_chunks;                 // 64-wide chunks taken from src_cl (the operand's last chunks)
  const __nv_bfloat16* known_cl;  // (B*m, 256) bf16: interpolation source (the operand's first 4 chunks), or null
  const int32_t* nn_idx;          // (rows, 3)
  const float* nn_w;              // (rows, 3)
  const uint8_t* w_img;           // streamed pieces in consumption order
  const float* bias1;             // (256)
  const float* bias2;             // (256) (NL == 3)
  const float* bias_last;         // (NBLK * 128), zero padded
  int c_last;                     // real output channels (256 | 259)
  int last_relu;
  float* out_cf;                  // (B, C_out, n) f32
  __nv_bfloat16* out_cl;          // (rows, C_out) bf16 or null
  // voting epilogue (vote = seed + y): channels 0..2 are xyz offsets, 3.. the feature residual (C_out = 256)
  int vote;
  const float* seed_xyz;          // (rows, 3)
  const float* seed_cf;           // (B, 256, n) f32
  float* vote_xyz;                // (rows, 3)
};

struct NnRow {
  int i0, i1, i2, pad0;
  float w0, w1, w2, pad1;
};

struct Misc {
  uint64_t wfull[kStages], afull[kStages], sfree[kStages], afree[kStages];
  uint64_t dfull, actfull;
  uint32_t tmem_base, pad_;
  alignas(16) float bias1[H], bias2[H], bias_last[384];
  alignas(16) NnRow nn[2][128];   // neighbour indices (absolute rows of known_cl) and weights of the tile's rows
};
static_assert(sizeof(Misc) <= kMisc, "misc area");

__device__ __noinline__ void pw_timeout() { __trap(); }
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  long long t0 = 0;
  for (;;) {
    if (mbar_try_wait_sleep(bar, parity)) return;
    if ((++spins & 0xFFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) pw_timeout();
    }
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// ring uses of one tile, in order: K0C layer-1 uses (operand chunk + 32 KB piece), HC layer-2 uses (NL == 3),
// NBLK * HC last-layer uses (16 KB pieces)
template <int K0C, int NL, int NBLK>
__global__ void __launch_bounds__(kThreads, 1) pw_mlp_kernel(const __grid_constant__ PwParams p) {
  constexpr int U1 = K0C, U2 = (NL == 3) ? HC : 0, U3 = NBLK * HC, UPT = U1 + U2 + U3;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  constexpr uint32_t off_act = kStages * kStageBytes;
  constexpr uint32_t off_misc = off_act + kActBytes;
  Misc* ms = reinterpret_cast<Misc*>(gbase + off_misc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&ms->wfull[i], 1);
      mbar_init(&ms->afull[i], kWorkers / 32);      // one arrival per worker warp
      mbar_init(&ms->sfree[i], 1);
      mbar_init(&ms->afree[i], 1);
    }
    mbar_init(&ms->dfull, 1);
    mbar_init(&ms->actfull, 16);
    mbar_fence_init();
  }
  if (warp == kWarpMma) tmem_alloc<1>(&ms->tmem_base, 512);
  for (int c = tid; c < H; c += kThreads) {
    ms->bias1[c] = __ldg(p.bias1 + c);
    ms->bias2[c] = (NL == 3) ? __ldg(p.bias2 + c) : 0.f;
  }
  for (int c = tid; c < NBLK * 128; c += kThreads) ms->bias_last[c] = __ldg(p.bias_last + c);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&ms->tmem_base);
  const int my_tiles =
      p.num_tiles > (int)blockIdx.x ? (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto tile_of = [&](int t) -> long long { return (long long)blockIdx.x + (long long)t * gridDim.x; };

  if (warp == kWarpTma) {
    // ============================================================ weight pieces, in consumption order
    if (lane == 0) {
      uint32_t u = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const uint8_t* src = p.w_img;
        for (int i = 0; i < UPT; ++i, ++u) {
          const uint32_t st = u % kStages;
          if (u >= kStages) bar_wait(&ms->sfree[st], (u / kStages - 1) & 1u);
          const uint32_t bytes = (i < U1 + U2) ? 2u * kChunk : (uint32_t)kChunk;
          mbar_arrive_expect_tx(&ms->wfull[st], bytes);
          tma_bulk_g2s(gbase + st * kStageBytes + kChunk, src, bytes, &ms->wfull[st]);
          src += bytes;
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ============================================================ MMA issue
    const bool issuer = elect_one();
    constexpr uint32_t idesc_h = uidesc_bf16(128, 256);     // hidden layers: rows x 256 channels
    constexpr uint32_t idesc_t = uidesc_bf16(128, 128);     // transposed last layer: 128 channels x 128 rows
    const uint32_t act = base + off_act;
    uint32_t u = 0, n_act = 0;
    uint32_t acnt[kStages] = {0, 0, 0};                      // layer-1 fills seen per stage
    for (int t = 0; t < my_tiles; ++t) {
      // ---- layer 1 (TMEM free: the previous tile's last epilogue has drained it)
      if (t > 0) bar_wait(&ms->actfull, (n_act++) & 1u);
#pragma unroll 1
      for (int kc = 0; kc < U1; ++kc, ++u) {
        const uint32_t st = u % kStages;
        bar_wait(&ms->wfull[st], (u / kStages) & 1u);
        bar_wait(&ms->afull[st], acnt[st] & 1u);
        ++acnt[st];
        fence_proxy_async_smem();
        tc_fence_after_sync();
        if (issuer) {
          const uint64_t ad = udesc_sw128(base + st * kStageBytes), bd = udesc_sw128(base + st * kStageBytes + kChunk);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16<1>(tmem_base, ad + 2u * k, bd + 2u * k, idesc_h, (kc | k) ? 1u : 0u);
          umma_commit_to<1>(&ms->sfree[st]);
          umma_commit_to<1>(&ms->afree[st]);      // the operand half of the stage: the builders wait on this one only
          if (kc == U1 - 1) umma_commit_to<1>(&ms->dfull);
        }
        __syncwarp();
      }
      // ---- layer 2
      if constexpr (NL == 3) {
        bar_wait(&ms->actfull, (n_act++) & 1u);
#pragma unroll 1
        for (int kc = 0; kc < HC; ++kc, ++u) {
          const uint32_t st = u % kStages;
          bar_wait(&ms->wfull[st], (u / kStages) & 1u);
          fence_proxy_async_smem();
          tc_fence_after_sync();
          if (issuer) {
            const uint64_t ad = udesc_sw128(act + kc * kChunk), bd = udesc_sw128(base + st * kStageBytes + kChunk);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16<1>(tmem_base, ad + 2u * k, bd + 2u * k, idesc_h, (kc | k) ? 1u : 0u);
            umma_commit_to<1>(&ms->sfree[st]);
            if (kc == HC - 1) umma_commit_to<1>(&ms->dfull);
          }
          __syncwarp();
        }
      }
      // ---- last layer, transposed: D^T (128 channels x 128 rows) per block
      bar_wait(&ms->actfull, (n_act++) & 1u);
#pragma unroll 1
      for (int i = 0; i < U3; ++i, ++u) {
        const int blk = i / HC, kc = i % HC;
        const uint32_t st = u % kStages;
        bar_wait(&ms->wfull[st], (u / kStages) & 1u);
        fence_proxy_async_smem();
        tc_fence_after_sync();
        if (issuer) {
          const uint64_t ad = udesc_sw128(base + st * kStageBytes + kChunk), bd = udesc_sw128(act + kc * kChunk);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16<1>(tmem_base + (uint32_t)(blk * 128), ad + 2u * k, bd + 2u * k, idesc_t, (kc | k) ? 1u : 0u);
          umma_commit_to<1>(&ms->sfree[st]);
          if (i == U3 - 1) umma_commit_to<1>(&ms->dfull);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================================================ workers: operand builders (20 warps) + epilogue (16)
    // The builders wait for a stage's operand half on `afree` (committed by the MMAs of layer-1 uses only), counted per
    // stage by the builders themselves: a parity wait must never skip completions (the weight-only uses of a stage
    // complete `sfree` in between, and a waiter that skips phases aliases -- found as a hang with two tiles per CTA).
    const int interp_chunks = p.known_cl ? 4 : 0;
    const int srcC = p.src_chunks * 64;
    const int wg = warp >> 2;
    const int row_t = (warp & 3) * 32 + lane;                       // TMEM lane (epilogue)
    const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t act = base + off_act;
    uint32_t n_d = 0;
    uint32_t fills[kStages] = {0, 0, 0};                            // layer-1 fills of each stage so far
    auto done_phase = [&]() {
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->actfull);
    };
    auto publish = [&](int st) {                // this warp's part of the stage's operand chunk is complete
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->afull[st]);
    };
    // hidden layer: columns [64 wg, 64 wg + 64) of the 256: TMEM -> +bias -> ReLU -> bf16 -> activation chunk wg
    auto hidden = [&](const float* bias) {
      const int c0 = wg * 64;
      const uint32_t chunk = act + (uint32_t)wg * kChunk;
      uint32_t v[2][16];
      tmem_ld_x16(tl + (uint32_t)c0, v[0]);
      tmem_ld_fence();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (g + 1 < 4) tmem_ld_x16(tl + (uint32_t)(c0 + (g + 1) * 16), v[(g + 1) & 1]);
        const uint32_t* w = v[g & 1];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 ba = *reinterpret_cast<const float4*>(bias + c0 + g * 16 + u * 8);
          const float4 bb = *reinterpret_cast<const float4*>(bias + c0 + g * 16 + u * 8 + 4);
          sts_v4(chunk + swz128(row_t, g * 2 + u),
                 bf16x2_relu(__uint_as_float(w[u * 8 + 0]) + ba.x, __uint_as_float(w[u * 8 + 1]) + ba.y),
                 bf16x2_relu(__uint_as_float(w[u * 8 + 2]) + ba.z, __uint_as_float(w[u * 8 + 3]) + ba.w),
                 bf16x2_relu(__uint_as_float(w[u * 8 + 4]) + bb.x, __uint_as_float(w[u * 8 + 5]) + bb.y),
                 bf16x2_relu(__uint_as_float(w[u * 8 + 6]) + bb.z, __uint_as_float(w[u * 8 + 7]) + bb.w));
        }
        if (g + 1 < 4) tmem_ld_fence();
      }
    };
    for (int t = 0; t < my_tiles; ++t) {
      const long long R0 = tile_of(t) * 128;
      // ------------------------------------------------ layer-1 operand of this tile
      if (interp_chunks) {
        if (tid < 128) {
          NnRow r;
          const long long R = R0 + tid;
          if (R < p.total_rows) {
            const int kb = (int)(R / p.n) * p.m;
            r.i0 = kb + __ldg(p.nn_idx + 3 * R);
            r.i1 = kb + __ldg(p.nn_idx + 3 * R + 1);
            r.i2 = kb + __ldg(p.nn_idx + 3 * R + 2);
            r.w0 = __ldg(p.nn_w + 3 * R);
            r.w1 = __ldg(p.nn_w + 3 * R + 1);
            r.w2 = __ldg(p.nn_w + 3 * R + 2);
          } else {
            r.i0 = r.i1 = r.i2 = 0;
            r.w0 = r.w1 = r.w2 = 0.f;
          }
          r.pad0 = 0;
          r.pad1 = 0.f;
          ms->nn[t & 1][tid] = r;
        }
        named_bar_sync(1, kWorkers);
      }
      int pend_st = -1;
      for (int kc = 0; kc < U1; ++kc) {
        const uint32_t u = (uint32_t)t * UPT + (uint32_t)kc;
        const uint32_t st = u % kStages;
        if (fills[st] > 0) {
          const uint32_t par = (fills[st] - 1) & 1u;
          if (!__all_sync(FULL, mbar_try_wait(&ms->afree[st], par))) {
            if (pend_st >= 0) {                 // about to block: publish what is in flight first
              cp_async_wait<0>();
              publish(pend_st);
              pend_st = -1;
            }
            bar_wait(&ms->afree[st], par);
          }
        }
        ++fills[st];
        const uint32_t dst = base + st * kStageBytes;
        if (kc < interp_chunks) {
          // three_interpolate: (row, 8 channels) pairs, 8 consecutive threads per row; both of a thread's pairs in flight
          uint4 a[2], b[2], c[2];
          NnRow q[2];
          int rw[2], un[2];
          bool on[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int pair = tid + i * kWorkers;
            on[i] = pair < 1024;
            rw[i] = on[i] ? pair >> 3 : 0;
            un[i] = pair & 7;
            q[i] = ms->nn[t & 1][rw[i]];
            const size_t off = (size_t)kc * 64 + un[i] * 8;
            if (on[i]) {
              a[i] = ldg_nc_u4(p.known_cl + (size_t)q[i].i0 * 256 + off);
              b[i] = ldg_nc_u4(p.known_cl + (size_t)q[i].i1 * 256 + off);
              c[i] = ldg_nc_u4(p.known_cl + (size_t)q[i].i2 * 256 + off);
            }
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (!on[i]) continue;
            const uint32_t* pa = &a[i].x;
            const uint32_t* pb = &b[i].x;
            const uint32_t* pc = &c[i].x;
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float lo = q[i].w0 * bf_lo(pa[k]) + q[i].w1 * bf_lo(pb[k]) + q[i].w2 * bf_lo(pc[k]);
              const float hi = q[i].w0 * bf_hi(pa[k]) + q[i].w1 * bf_hi(pb[k]) + q[i].w2 * bf_hi(pc[k]);
              o[k] = bf16x2_rn(lo, hi);
            }
            sts_v4(dst + swz128(rw[i], un[i]), o[0], o[1], o[2], o[3]);
          }
          publish((int)st);
        } else {
          const int sc = kc - interp_chunks;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int pair = tid + i * kWorkers;
            if (pair < 1024) {
              const int rw = pair >> 3, un = pair & 7;
              const long long R = R0 + rw < p.total_rows ? R0 + rw : p.total_rows - 1;
              cp_async16(dst + swz128(rw, un), p.src_cl + (size_t)R * srcC + sc * 64 + un * 8);
            }
          }
          cp_async_commit();
          if (pend_st >= 0) {
            cp_async_wait<1>();
            publish(pend_st);
          }
          pend_st = (int)st;
        }
      }
      if (pend_st >= 0) {                        // these warps turn into the epilogue now: nothing may stay pending
        cp_async_wait<0>();
        publish(pend_st);
      }
      if (warp >= 16) continue;
      // ------------------------------------------------ epilogue: accumulator split 4 ways by columns
      bar_wait(&ms->dfull, (n_d++) & 1u);
      tc_fence_after_sync();
      hidden(ms->bias1);
      done_phase();
      if constexpr (NL == 3) {
        bar_wait(&ms->dfull, (n_d++) & 1u);
        tc_fence_after_sync();
        hidden(ms->bias2);
        done_phase();
      }
      bar_wait(&ms->dfull, (n_d++) & 1u);
      tc_fence_after_sync();
      // ---- last layer: lane = output channel of the block, columns [32 wg, 32 wg + 32) = rows of the tile
      const int col0 = wg * 32;
      const long long Rc = R0 + col0;                       // first row of this thread's columns
      const int b = (int)(R0 / p.n);
      const int j0 = (int)(R0 - (long long)b * p.n) + col0;
#pragma unroll 1
      for (int blk = 0; blk < NBLK; ++blk) {
        const int ch = blk * 128 + row_t;
        const float bias = ms->bias_last[ch];
        float y[32];
        {
          uint32_t v[2][16];
          tmem_ld_x16(tl + (uint32_t)(blk * 128 + col0), v[0]);
          tmem_ld_x16(tl + (uint32_t)(blk * 128 + col0 + 16), v[1]);
          tmem_ld_fence();
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float f = __uint_as_float(v[c >> 4][c & 15]) + bias;
            y[c] = p.last_relu ? fmaxf(f, 0.f) : f;
          }
        }
        if (ch >= p.c_last) continue;
        if (!p.vote) {
          float* o = p.out_cf + ((size_t)b * p.c_last + ch) * p.n + j0;
#pragma unroll
          for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(o + c) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
          if (p.out_cl) {
#pragma unroll
            for (int c = 0; c < 32; ++c) p.out_cl[(size_t)(Rc + c) * p.c_last + ch] = __float2bfloat16_rn(y[c]);
          }
        } else if (ch < 3) {
#pragma unroll
          for (int c = 0; c < 32; ++c) p.vote_xyz[(size_t)(Rc + c) * 3 + ch] = __ldg(p.seed_xyz + (size_t)(Rc + c) * 3 + ch) + y[c];
        } else {
          const int f = ch - 3;
          const float* sd = p.seed_cf + ((size_t)b * 256 + f) * p.n + j0;
          float* o = p.out_cf + ((size_t)b * 256 + f) * p.n + j0;
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(sd + c));
            y[c] += s4.x;
            y[c + 1] += s4.y;
            y[c + 2] += s4.z;
            y[c + 3] += s4.w;
            *reinterpret_cast<float4*>(o + c) = make_float4(y[c], y[c + 1], y[c + 2], y[c + 3]);
          }
          if (p.out_cl) {
#pragma unroll
            for (int c = 0; c < 32; ++c) p.out_cl[(size_t)(Rc + c) * 256 + f] = __float2bfloat16_rn(y[c]);
          }
        }
      }
      done_phase();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kWarpMma) tmem_dealloc<1>(tmem_base, 512);
}

uint16_t bf16_bits(float w) {
  uint32_t u;
  memcpy(&u, &w, 4);
  return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}
void put_sw128(uint8_t* piece, int r, int kk, float w) {
  const int unit = kk >> 3;
  const size_t byte = (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + (size_t)((unit ^ (r & 7)) << 4) + (size_t)(kk & 7) * 2;
  const uint16_t b = bf16_bits(w);
  memcpy(piece + byte, &b, 2);
}

struct PwShape {
  int k0c, nl, nblk;
};
bool pw_shape(int kind, PwShape& s) {
  if (kind == 0) s = {8, 2, 2};          // FP module: 512 -> 256 -> 256
  else if (kind == 1) s = {4, 3, 3};     // voting module: 256 -> 256 -> 256 -> 259
  else return false;
  return true;
}

}  // namespace

// kind 0 = FP module (K0 = 512: interpolated 256 | skip 256; layers 512 -> 256 -> c_last <= 256),
// kind 1 = voting module (K0 = 256; layers 256 -> 256 -> 256 -> c_last <= 384).
extern "C" long long sad_pw_mlp_image_bytes(int kind) {
  PwShape s;
  if (!pw_shape(kind, s)) return -1;
  return (long long)(s.k0c + (s.nl == 3 ? HC : 0)) * 2 * kChunk + (long long)s.nblk * HC * kChunk;
}

// HOST: W1 (256 x K0), [W2 (256 x 256)], Wlast (c_last x 256), row-major fp32 -> streamed bf16 pieces
extern "C" int sad_pw_mlp_pack(int kind, const float* W1, const float* W2, const float* Wlast, int c_last, void* out_image) {
  PwShape s;
  SAD_REQUIRE(pw_shape(kind, s), "pw_mlp_pack: unknown kind %d", kind);
  SAD_REQUIRE(W1 && Wlast && out_image && (s.nl == 2 || W2), "pw_mlp_pack: null pointer");
  SAD_REQUIRE(c_last >= 1 && c_last <= s.nblk * 128, "pw_mlp_pack: c_last out of range");
  uint8_t* img = static_cast<uint8_t*>(out_image);
  memset(img, 0, (size_t)sad_pw_mlp_image_bytes(kind));
  const int K0 = s.k0c * 64;
  for (int kc = 0; kc < s.k0c; ++kc)
    for (int r = 0; r < 256; ++r)
      for (int kk = 0; kk < 64; ++kk) put_sw128(img + (size_t)kc * 2 * kChunk, r, kk, W1[(size_t)r * K0 + kc * 64 + kk]);
  uint8_t* q = img + (size_t)s.k0c * 2 * kChunk;
  if (s.nl == 3) {
    for (int kc = 0; kc < HC; ++kc)
      for (int r = 0; r < 256; ++r)
        for (int kk = 0; kk < 64; ++kk) put_sw128(q + (size_t)kc * 2 * kChunk, r, kk, W2[(size_t)r * 256 + kc * 64 + kk]);
    q += (size_t)HC * 2 * kChunk;
  }
  for (int blk = 0; blk < s.nblk; ++blk)
    for (int kc = 0; kc < HC; ++kc)
      for (int r = 0; r < 128; ++r) {
        const int ch = blk * 128 + r;
        if (ch >= c_last) continue;
        for (int kk = 0; kk < 64; ++kk) put_sw128(q + (size_t)(blk * HC + kc) * kChunk, r, kk, Wlast[(size_t)ch * 256 + kc * 64 + kk]);
      }
  return SAD_OK;
}

// Launch.  rows = B * n (n a multiple of 128).
//   kind 0: known_cl (B*m,256) bf16 + nn_idx / nn_w (B*n,3) -> interpolated half; src_cl (B*n,256) bf16 = skip half;
//           out_cf (B,c_last,n) f32 and / or out_cl (B*n,c_last) bf16, ReLU on the output.
//   kind 1: src_cl (B*n,256) bf16 seed features; c_last = 3 + 256; vote_xyz (B*n,3) = seed_xyz + y[0:3],
//           out_cf (B,256,n) = seed_cf + y[3:], out_cl (B*n,256) bf16 twin; no ReLU on the output.
extern "C" int sad_pw_mlp_fwd(int kind, int B, int n, int m, const void* src_cl, const void* known_cl, const int32_t* nn_idx,
                              const float* nn_w, const void* w_image, const float* bias1, const float* bias2,
                              const float* bias_last_padded, int c_last, float* out_cf, void* out_cl, const float* seed_xyz,
                              const float* seed_cf, float* vote_xyz, int tiles_per_cta, sad_stream_t stream) {
  PwShape s;
  SAD_REQUIRE(pw_shape(kind, s), "pw_mlp: unknown kind %d", kind);
  SAD_REQUIRE(B >= 0 && n >= 128 && n % 128 == 0, "pw_mlp: rows per batch element must be a multiple of 128 (n=%d)", n);
  if (B == 0) return SAD_OK;
  SAD_REQUIRE(src_cl && w_image && bias1 && bias_last_padded && out_cf, "pw_mlp: null pointer");
  SAD_REQUIRE(s.nl == 2 || bias2, "pw_mlp: bias2 missing");
  SAD_REQUIRE((long long)B * n < 0x7FFFFFFFLL, "pw_mlp: too many rows");
  PwParams p;
  memset(&p, 0, sizeof(p));
  p.n = n; p.m = m;
  p.total_rows = (long long)B * n;
  p.num_tiles = (int)(p.total_rows / 128);
  p.src_cl = static_cast<const __nv_bfloat16*>(src_cl);
  p.w_img = static_cast<const uint8_t*>(w_image);
  p.bias1 = bias1; p.bias2 = bias2; p.bias_last = bias_last_padded;
  p.c_last = c_last;
  p.out_cf = out_cf; p.out_cl = static_cast<__nv_bfloat16*>(out_cl);
  if (kind == 0) {
    SAD_REQUIRE(known_cl && nn_idx && nn_w && m >= 3, "pw_mlp: FP stage needs the interpolation source, indices and weights");
    SAD_REQUIRE(c_last >= 8 && c_last <= 256, "pw_mlp: FP output width out of range");
    p.known_cl = static_cast<const __nv_bfloat16*>(known_cl);
    p.nn_idx = nn_idx; p.nn_w = nn_w;
    p.src_chunks = 4;
    p.last_relu = 1;
  } else {
    SAD_REQUIRE(seed_xyz && seed_cf && vote_xyz && c_last == 259, "pw_mlp: voting stage needs seeds and 3 + 256 outputs");
    p.src_chunks = 4;
    p.vote = 1;
    p.seed_xyz = seed_xyz; p.seed_cf = seed_cf; p.vote_xyz = vote_xyz;
    p.last_relu = 0;
  }
  int dev = 0, sms = 0;
  SAD_CUDA_OK(cudaGetDevice(&dev));
  SAD_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static thread_local int configured_dev = -1;
  if (configured_dev != dev) {
    SAD_CUDA_OK(cudaFuncSetAttribute(pw_mlp_kernel<8, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    SAD_CUDA_OK(cudaFuncSetAttribute(pw_mlp_kernel<4, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured_dev = dev;
  }
  const int tpc = tiles_per_cta < 1 ? 1 : tiles_per_cta;
  int grid = sad_ceil_div(p.num_tiles, tpc);
  if (grid > sms) grid = sms;
  if (kind == 0) pw_mlp_kernel<8, 2, 2><<<grid, kThreads, kSmem, (cudaStream_t)stream>>>(p);
  else pw_mlp_kernel<4, 3, 3><<<grid, kThreads, kSmem, (cudaStream_t)stream>>>(p);
  SAD_LAUNCH_CHECK("pw_mlp_kernel");
  return SAD_OK;
}
