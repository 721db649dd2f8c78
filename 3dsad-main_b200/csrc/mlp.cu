// a6  shared point-wise MLP (+ max-pool over nsample), fused with the neighbourhood gather
// -- SURVEY.md section 8(a) rows a5/a6, section 8(f) rank 1, hard parts H4/H5.
// (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
// One launch per SA / FP / voting stage.  For a tile of 128 rows (row = (query point,
// sample)) the kernel
//   1. GATHERS the layer-1 operand straight into shared memory in the tcgen05 canonical
//      K-major SWIZZLE_128B layout: channel-last bf16 feature rows are fetched by `idx` with
//      16-byte cp.async (the grouped tensor is never materialised in HBM); the relative,
//      radius-normalised xyz (+ optional fp32 scalar features) form one extra 16-wide K step;
//   2. runs the 2-3 layer MLP on the 5th-gen tensor cores: tcgen05.mma (bf16 in, fp32
//      accumulate in TMEM), weights streamed as pre-swizzled images by the TMA engine
//      (cp.async.bulk) through an mbarrier ring; hidden activations go TMEM -> registers
//      (bias + ReLU, fp32) -> bf16 -> shared memory and are the next layer's operand;
//   3. evaluates the LAST layer transposed (D^T = W . H^T), so a TMEM lane is an output
//      channel and the nsample rows of a query point are consecutive TMEM columns: the
//      max-pool is an in-register reduction in the epilogue (no shuffles, no extra pass).
// Warp roles: warps 0-3 gather + epilogue (thread == TMEM lane), warp 4 weight TMA, warp 5
// MMA issue + TMEM allocation.  Persistent CTAs, static tile round-robin.
#include <cuda_bf16.h>
#include <string.h>

#include "sad_common.cuh"

namespace {

using namespace sad;

constexpr int kWorkers = 128;
constexpr int kThreads = 192;
constexpr int kChunkBytes = 128 * 128;   // one 128-row x 64-bf16 K chunk (A operand / activations)
constexpr int kMaxLayers = 3;
constexpr int kMaxSlots = 8;

struct MlpParams {
  int B, N, P, S;
  long long total_rows;
  int num_tiles;
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
  int slot_bytes, nst, rw, act_chunks, steps_per_tile;
};

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of (row, 16-byte unit) inside a 128-row x 128-byte SWIZZLE_128B chunk
__device__ __forceinline__ uint32_t swz(int row, int unit) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((unit ^ (row & 7)) << 4));
}

// emit the pooled / plain outputs held by one thread (= output channel `ch`) for the 32
// consecutive rows starting at global row R0 (all rows of one tile share S).
template <int S>
__device__ __forceinline__ void emit_group(const MlpParams& p, const uint32_t (&v)[32], float& run, int ch,
                                           int c_last, float bias, long long R0, bool ch_ok) {
  if (S >= 32) {
    float m = __uint_as_float(v[0]);
#pragma unroll
    for (int i = 1; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
    const bool first = ((R0 % S) == 0);
    run = first ? m : fmaxf(run, m);
    if (((R0 + 32) % S) == 0 && ch_ok && R0 < p.total_rows) {
      const long long pt = R0 / S;
      float y = run + bias;
      if (p.last_relu) y = fmaxf(y, 0.f);
      if (p.out_cf) {
        const long long b = pt / p.P, j = pt % p.P;
        p.out_cf[(b * c_last + ch) * p.P + j] = y;
      }
      if (p.out_cl) p.out_cl[pt * c_last + ch] = __float2bfloat16_rn(y);
    }
  } else {
#pragma unroll
    for (int g = 0; g < 32 / S; ++g) {
      float m = __uint_as_float(v[g * S]);
#pragma unroll
      for (int i = 1; i < S; ++i) m = fmaxf(m, __uint_as_float(v[g * S + i]));
      const long long R = R0 + (long long)g * S;
      if (ch_ok && R < p.total_rows) {
        const long long pt = R / S;
        float y = m + bias;
        if (p.last_relu) y = fmaxf(y, 0.f);
        if (p.out_cf) {
          const long long b = pt / p.P, j = pt % p.P;
          p.out_cf[(b * c_last + ch) * p.P + j] = y;
        }
        if (p.out_cl) p.out_cl[pt * c_last + ch] = __float2bfloat16_rn(y);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) fused_mlp_kernel(const MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B atoms need 1024-B alignment
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  // carve-up (all offsets multiples of 1024)
  const uint32_t a_stage = base;                                     // 2 x 16 KB gather stages
  const uint32_t w_ring = a_stage + 2 * kChunkBytes;                 // nst x slot_bytes weight ring
  const uint32_t act = w_ring + (uint32_t)p.nst * p.slot_bytes;      // act_chunks x 16 KB activations
  uint8_t* misc = gbase + (act - base) + (size_t)p.act_chunks * kChunkBytes;
  uint64_t* wfull = reinterpret_cast<uint64_t*>(misc);               // [kMaxSlots]
  uint64_t* wfree = wfull + kMaxSlots;                               // [kMaxSlots]
  uint64_t* afull = wfree + kMaxSlots;                               // [2]
  uint64_t* afree = afull + 2;                                       // [2]
  uint64_t* dfull = afree + 2;                                       // [1]
  uint64_t* actfull = dfull + 1;                                     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(actfull + 1);
  long long* s_src = reinterpret_cast<long long*>(tmem_slot + 2);    // [128] source row (b*N + id) or -1

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nl = p.n_layers;
  const int c_last = p.c[nl - 1];
  const int nblk = (c_last + 127) / 128;
  const int chunks0 = p.kpad[0] / 64;
  const bool resident = p.steps_per_tile <= p.nst;

  if (tid == 0) {
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wfree[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&afull[i], kWorkers);
      mbar_init(&afree[i], 1);
    }
    mbar_init(dfull, 1);
    mbar_init(actfull, kWorkers);
    mbar_fence_init();
  }
  if (warp == 5) {   // TMEM allocation: one warp, power-of-two columns
    const uint32_t ncols = (uint32_t)(2 * p.rw);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 4) {
    // ================================================================== weight TMA producer
    if (lane == 0) {
      long long gs = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++iter) {
        if (resident && iter > 0) break;
        for (int li = 0; li < nl; ++li) {
          const bool last = (li == nl - 1);
          const int pieces = last ? nblk * (p.kpad[li] / 64) : (p.kpad[li] / 64);
          const uint32_t bytes = last ? (uint32_t)kChunkBytes : (uint32_t)p.c[li] * 128u;
          for (int pc = 0; pc < pieces; ++pc, ++gs) {
            const int slot = (int)(gs % p.nst);
            const long long use = gs / p.nst;
            if (use > 0) mbar_wait(&wfree[slot], (uint32_t)((use - 1) & 1));
            mbar_arrive_expect_tx(&wfull[slot], bytes);
            tma_bulk_g2s(gbase + (w_ring - base) + (size_t)slot * p.slot_bytes, p.w_img[li] + (size_t)pc * bytes, bytes,
                         &wfull[slot]);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      long long gs = 0, ga = 0;
      uint32_t acount = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++iter) {
        const bool wsync = !(resident && iter > 0);
        if (resident) gs = 0;
        for (int li = 0; li < nl; ++li) {
          const bool last = (li == nl - 1);
          const int chunks = p.kpad[li] / 64;
          if (li > 0) {   // previous epilogue has written ACT and drained its TMEM region
            mbar_wait(actfull, acount & 1);
            ++acount;
            tc_fence_after();
          }
          if (!last) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((li & 1) * p.rw);
            const uint32_t idesc = umma_idesc(128, p.c[li]);
            for (int kc = 0; kc < chunks; ++kc, ++gs) {
              const int slot = (int)(gs % p.nst);
              uint32_t a_addr;
              int ksteps = 4;
              if (li == 0) {
                const int stage = (int)(ga & 1);
                mbar_wait(&afull[stage], (uint32_t)((ga >> 1) & 1));
                a_addr = a_stage + stage * kChunkBytes;
                if (p.has_special && kc == chunks - 1) ksteps = 1;
              } else {
                a_addr = act + kc * kChunkBytes;
              }
              if (wsync) mbar_wait(&wfull[slot], (uint32_t)((gs / p.nst) & 1));
              tc_fence_after();
              const uint32_t b_addr = w_ring + slot * p.slot_bytes;
              for (int k = 0; k < ksteps; ++k)
                umma_bf16(d_tmem, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc,
                          (kc > 0 || k > 0) ? 1u : 0u);
              if (li == 0) {
                umma_commit(&afree[ga & 1]);
                ++ga;
              }
              if (wsync && !resident) umma_commit(&wfree[slot]);
            }
            umma_commit(dfull);
          } else {
            const uint32_t idesc = umma_idesc(128, 128);
            for (int blk = 0; blk < nblk; ++blk) {
              if (blk > 0) {   // v1: one output block in flight
                mbar_wait(actfull, acount & 1);
                ++acount;
                tc_fence_after();
              }
              const uint32_t d_tmem = tmem_base + (uint32_t)((blk & 1) * p.rw);
              for (int kc = 0; kc < chunks; ++kc, ++gs) {
                const int slot = (int)(gs % p.nst);
                if (wsync) mbar_wait(&wfull[slot], (uint32_t)((gs / p.nst) & 1));
                tc_fence_after();
                const uint32_t a_addr = w_ring + slot * p.slot_bytes;     // W_last block rows = M
                const uint32_t b_addr = act + kc * kChunkBytes;           // activations rows = N
                for (int k = 0; k < 4; ++k)
                  umma_bf16(d_tmem, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc,
                            (kc > 0 || k > 0) ? 1u : 0u);
                if (wsync && !resident) umma_commit(&wfree[slot]);
              }
              umma_commit(dfull);
            }
          }
        }
        // last block's epilogue must drain TMEM / ACT before the next tile's layer 1 reuses them
        mbar_wait(actfull, acount & 1);
        ++acount;
        tc_fence_after();
      }
    }
  } else {
    // ================================================================== gather + epilogue workers
    long long ga = 0;
    uint32_t dcount = 0;
    const int unit = tid & 7;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const long long R0 = (long long)tile * 128;
      // ---- per-row source bookkeeping (thread t <-> row t)
      {
        const long long R = R0 + tid;
        long long src = -1;
        if (R < p.total_rows) {
          const long long pt = R / p.S;
          const long long b = pt / p.P;
          const long long id = p.idx ? (long long)__ldg(p.idx + R) : (pt % p.P);
          src = b * p.N + id;
        }
        s_src[tid] = src;
      }
      named_bar_sync(1, kWorkers);

      // ---- layer-1 operand chunks
      for (int kc = 0; kc < chunks0; ++kc, ++ga) {
        const int stage = (int)(ga & 1);
        if (ga >= 2) mbar_wait(&afree[stage], (uint32_t)(((ga >> 1) - 1) & 1));
        const uint32_t dst = a_stage + stage * kChunkBytes;
        const int nf0 = p.C0 / 64, nf1 = p.C1in / 64;
        if (kc < nf0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = (tid >> 3) + 16 * i;
            const long long src = s_src[r];
            const __nv_bfloat16* g = p.feat_cl + (src < 0 ? 0 : src) * p.C0 + kc * 64 + unit * 8;
            cp_async16(dst + swz(r, unit), g, src < 0 ? 0u : 16u);
          }
        } else if (kc < nf0 + nf1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = (tid >> 3) + 16 * i;
            const long long R = R0 + r;
            const bool ok = R < p.total_rows;
            const __nv_bfloat16* g = p.feat2_cl + (ok ? R : 0) * p.C1in + (kc - nf0) * 64 + unit * 8;
            cp_async16(dst + swz(r, unit), g, ok ? 16u : 0u);
          }
        } else {
          // special 16-wide K step: [dx, dy, dz, extras..., 0]
          float vals[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) vals[i] = 0.f;
          const long long src = s_src[tid];
          if (src >= 0) {
            const long long pt = (R0 + tid) / p.S;
            if (p.xyz) {
              const float* a = p.xyz + src * 3;
              const float* q = p.new_xyz + pt * 3;
              float dx = __fsub_rn(__ldg(a), __ldg(q)), dy = __fsub_rn(__ldg(a + 1), __ldg(q + 1)),
                    dz = __fsub_rn(__ldg(a + 2), __ldg(q + 2));
              if (p.normalize) {
                const float r = p.radius_t ? __ldg(p.radius_t + pt) : p.radius;
                dx = __fdiv_rn(dx, r);
                dy = __fdiv_rn(dy, r);
                dz = __fdiv_rn(dz, r);
              }
              vals[0] = dx;
              vals[1] = dy;
              vals[2] = dz;
            }
#pragma unroll
            for (int e = 0; e < 13; ++e)
              if (e < p.E) vals[3 + e] = __ldg(p.extra + src * p.E + e);
          }
          st_shared_v4(dst + swz(tid, 0), pack_bf16(vals[0], vals[1]), pack_bf16(vals[2], vals[3]),
                       pack_bf16(vals[4], vals[5]), pack_bf16(vals[6], vals[7]));
          st_shared_v4(dst + swz(tid, 1), pack_bf16(vals[8], vals[9]), pack_bf16(vals[10], vals[11]),
                       pack_bf16(vals[12], vals[13]), pack_bf16(vals[14], vals[15]));
        }
        cp_async_wait_all();
        fence_proxy_async();
        mbar_arrive(&afull[stage]);
      }

      // ---- hidden-layer epilogues: TMEM -> +bias, ReLU -> bf16 -> ACT (thread == row)
      for (int li = 0; li < nl - 1; ++li) {
        mbar_wait(dfull, dcount & 1);
        ++dcount;
        tc_fence_after();
        const uint32_t d_tmem = lane_taddr + (uint32_t)((li & 1) * p.rw);
        const float* bias = p.bias[li];
        for (int c0 = 0; c0 < p.c[li]; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(d_tmem + c0, v);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float lo = fmaxf(__uint_as_float(v[2 * i]) + __ldg(bias + c0 + 2 * i), 0.f);
            const float hi = fmaxf(__uint_as_float(v[2 * i + 1]) + __ldg(bias + c0 + 2 * i + 1), 0.f);
            pk[i] = pack_bf16(lo, hi);
          }
          const uint32_t chunk = act + (c0 >> 6) * kChunkBytes;
          const int u0 = (c0 & 63) >> 3;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            st_shared_v4(chunk + swz(tid, u0 + u), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(actfull);
      }

      // ---- last layer (transposed): thread == output channel, columns == rows; pool over S
      for (int blk = 0; blk < nblk; ++blk) {
        mbar_wait(dfull, dcount & 1);
        ++dcount;
        tc_fence_after();
        const uint32_t d_tmem = lane_taddr + (uint32_t)((blk & 1) * p.rw);
        const int ch = blk * 128 + tid;
        const bool ch_ok = ch < c_last;
        const float bias = ch_ok ? __ldg(p.bias[nl - 1] + ch) : 0.f;
        float run = 0.f;
        for (int g = 0; g < 4; ++g) {
          uint32_t v[32];
          tmem_ld32(d_tmem + g * 32, v);
          const long long Rg = R0 + g * 32;
          switch (p.S) {
            case 1: emit_group<1>(p, v, run, ch, c_last, bias, Rg, ch_ok); break;
            case 2: emit_group<2>(p, v, run, ch, c_last, bias, Rg, ch_ok); break;
            case 4: emit_group<4>(p, v, run, ch, c_last, bias, Rg, ch_ok); break;
            case 8: emit_group<8>(p, v, run, ch, c_last, bias, Rg, ch_ok); break;
            case 16: emit_group<16>(p, v, run, ch, c_last, bias, Rg, ch_ok); break;
            case 32: emit_group<32>(p, v, run, ch, c_last, bias, Rg, ch_ok); break;
            case 64: emit_group<64>(p, v, run, ch, c_last, bias, Rg, ch_ok); break;
            default: emit_group<128>(p, v, run, ch, c_last, bias, Rg, ch_ok); break;
          }
        }
        tc_fence_before();
        mbar_arrive(actfull);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    const uint32_t ncols = (uint32_t)(2 * p.rw);
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
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
extern "C" long long sad_mlp_weight_image_bytes(int cout, int kpad, int is_last) {
  if (cout < 1 || kpad < 64 || (kpad % 64)) return -1;
  const long long chunks = kpad / 64;
  if (is_last) return (long long)((cout + 127) / 128) * chunks * kChunkBytes;
  return chunks * (long long)cout * 128;
}

// Host-side packer: fp32 W (cout x cin, row-major) -> bf16 image in the exact shared-memory byte
// layout the kernel consumes (K-major, SWIZZLE_128B, one piece per 64-wide K chunk; the last
// layer additionally blocked into 128-row pieces, zero padded).  perm[k] = source column of
// packed K index k, or -1 for a zero column.
extern "C" int sad_mlp_pack_weights(const float* W, int cout, int cin, const int32_t* perm, int kpad, int is_last,
                                    void* out_image) {
  SAD_REQUIRE(W && perm && out_image, "mlp_pack_weights: null pointer");
  SAD_REQUIRE(cout >= 1 && cin >= 1 && kpad >= 64 && kpad % 64 == 0, "mlp_pack_weights: bad sizes");
  SAD_REQUIRE(is_last || (cout % 16 == 0 && cout <= 256), "mlp_pack_weights: hidden width must be a multiple of 16, <= 256");
  const int chunks = kpad / 64;
  const int rows_per_piece = is_last ? 128 : cout;
  const int nblk = is_last ? (cout + 127) / 128 : 1;
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

extern "C" int sad_shared_mlp_fwd(int B, int N, int P, int S, const void* feat_cl, int C0, const void* feat2_cl,
                                  int C1in, const float* xyz, const float* new_xyz, const int32_t* idx, float radius,
                                  const float* radius_t, int normalize_xyz, const float* extra, int E, int n_layers,
                                  const void* const* w_img, const float* const* bias, const int* c_out, int last_relu,
                                  void* out_cl_bf16, float* out_cf_f32, sad_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
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
  if (B == 0 || P == 0) return SAD_OK;

  MlpParams p = {};
  p.B = B; p.N = N; p.P = P; p.S = S;
  p.total_rows = (long long)B * P * S;
  p.num_tiles = (int)((p.total_rows + 127) / 128);
  p.feat_cl = static_cast<const __nv_bfloat16*>(feat_cl); p.C0 = C0;
  p.feat2_cl = static_cast<const __nv_bfloat16*>(feat2_cl); p.C1in = C1in;
  p.xyz = xyz; p.new_xyz = new_xyz; p.idx = idx; p.radius_t = radius_t; p.radius = radius;
  p.normalize = normalize_xyz; p.extra = extra; p.E = E; p.has_special = has_special;
  p.n_layers = n_layers; p.last_relu = last_relu;
  p.out_cl = static_cast<__nv_bfloat16*>(out_cl_bf16); p.out_cf = out_cf_f32;
  int hidden_max = 0, steps = 0;
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
      steps += p.kpad[li] / 64;
    } else {
      steps += ((c_out[li] + 127) / 128) * (p.kpad[li] / 64);
    }
  }
  p.steps_per_tile = steps;
  p.rw = hidden_max > 128 ? 256 : 128;
  p.act_chunks = hidden_max / 64;
  p.slot_bytes = hidden_max > 128 ? 2 * kChunkBytes : kChunkBytes;
  const int misc = 1024;
  const int fixed = 2 * kChunkBytes + p.act_chunks * kChunkBytes + misc + 1024 /*alignment slack*/ + 128 * 8;
  int nst = steps <= kMaxSlots ? steps : 4;                     // everything resident when it fits in the ring
  while (nst > 2 && fixed + nst * p.slot_bytes > 227 * 1024) --nst;
  if (steps <= kMaxSlots && nst < steps) nst = nst < 4 ? nst : 4;
  SAD_REQUIRE(fixed + nst * p.slot_bytes <= 227 * 1024, "shared_mlp: shared-memory budget exceeded");
  p.nst = nst;
  const size_t smem = (size_t)fixed + (size_t)nst * p.slot_bytes;

  int dev = 0, sms = 0;
  SAD_CUDA_OK(cudaGetDevice(&dev));
  SAD_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static thread_local int configured_dev = -1;
  if (configured_dev != dev) {
    SAD_CUDA_OK(cudaFuncSetAttribute(fused_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured_dev = dev;
  }
  const int per_sm = (2 * p.rw <= 256 && smem <= 113 * 1024) ? 2 : 1;
  const int grid = p.num_tiles < sms * per_sm ? p.num_tiles : sms * per_sm;
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
