// Shared helpers for libsad_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sad_ops.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsad_b200 is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------- host side
void sad_set_error(const char* fmt, ...);
void sad_count_launch(int n);

#define SAD_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      sad_set_error(__VA_ARGS__);       \
      return SAD_EINVAL;                \
    }                                   \
  } while (0)

#define SAD_CUDA_OK(expr)                                                          \
  do {                                                                             \
    cudaError_t e_ = (expr);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      sad_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                    __LINE__);                                                     \
      return SAD_ECUDA;                                                            \
    }                                                                              \
  } while (0)

#define SAD_LAUNCH_CHECK(name)                                                   \
  do {                                                                           \
    cudaError_t e_ = cudaGetLastError();                                         \
    if (e_ != cudaSuccess) {                                                     \
      sad_set_error("launch of %s failed: %s", name, cudaGetErrorString(e_));    \
      return SAD_ECUDA;                                                          \
    }                                                                            \
    sad_count_launch(1);                                                         \
  } while (0)

static inline int sad_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Tuning / A-B hooks read the environment only in tools builds (tools/build_variant.py <name> -DSAD_TOOLS); the product
// library never looks at the environment.
#ifdef SAD_TOOLS
#include <stdlib.h>
static inline const char* sad_tool_env(const char* name) { return getenv(name); }
#else
static inline const char* sad_tool_env(const char*) { return nullptr; }
#endif

#ifdef SAD_TOOLS_ABLATE
// tools/build_variant.py builds only: SAD_ABLATE bit 0 = skip the scene-grid FPS, bit 1 = skip the fused MLP launches,
// bit 2 = skip the register-resident FPS (pipeline cost attribution; results are garbage by design).
#include <stdlib.h>
static inline int sad_ablate_mask() {
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("SAD_ABLATE");
    m = e ? atoi(e) : 0;
  }
  return m;
}
#endif

// -------------------------------------------------------------- device side
#ifdef __CUDACC__
namespace sad {

constexpr uint32_t FULL = 0xFFFFFFFFu;

// Squared distance under the arithmetic contract H1: one rounding per op, never
// contracted to FMA.  p = scanned candidate, q = query / last selected point.
__device__ __forceinline__ float sqdist(float px, float py, float pz, float qx, float qy, float qz) {
  const float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier (shared::cta) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA 1-D bulk copy global -> shared::cta, completion on an mbarrier ----
// (SASS: UBLKCP).  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- thread-block cluster ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive_release();
  cluster_wait_acquire();
}
// Map a local shared::cta address to the same offset in CTA `rank` of the cluster.
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}
__device__ __forceinline__ void st_cluster_b32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}

// ---- named barriers ----
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace sad
#endif  // __CUDACC__
