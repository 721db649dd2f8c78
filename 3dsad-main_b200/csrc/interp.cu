// a9 three_interpolate (forward + backward) -- SURVEY.md section 8(a) row a9.
// (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
// HBM-bound: out[b,c,i] = ((w0*f[i0]) + (w1*f[i1])) + (w2*f[i2]), evaluated in exactly
// that order with no FMA contraction so the fp32 result is bit-identical to the
// oracle.  One thread owns four consecutive unknown points: 12 indices + 12 weights are
// loaded once with 128-bit loads and reused over a chunk of channels; per channel the
// thread issues 12 independent 4-byte gathers (source rows are L1/L2 resident) and one
// 128-bit streaming store (a warp writes 512 contiguous bytes).
#include "sad_common.cuh"

namespace {

constexpr int TI_T = 128;
constexpr int TI_CCH = 16;

__device__ __forceinline__ float interp3(float w0, float f0, float w1, float f1, float w2, float f2) {
  return __fadd_rn(__fadd_rn(__fmul_rn(w0, f0), __fmul_rn(w1, f1)), __fmul_rn(w2, f2));
}

template <bool VEC>
__global__ void __launch_bounds__(TI_T)
interp_fwd_kernel(int C, int m, int n, const float* __restrict__ features, const int32_t* __restrict__ idx,
                  const float* __restrict__ weight, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * TI_CCH;
  const int cn = min(TI_CCH, C - c0);
  const int t = blockIdx.x * TI_T + threadIdx.x;
  const float* f = features + ((size_t)b * C + c0) * m;
  if (VEC) {
    if (t * 4 >= n) return;
    int id[12];
    float w[12];
    const int4* ip = reinterpret_cast<const int4*>(idx + (size_t)b * n * 3) + (size_t)t * 3;
    const float4* wp = reinterpret_cast<const float4*>(weight + (size_t)b * n * 3) + (size_t)t * 3;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int4 a = __ldg(ip + u);
      const float4 ww = __ldg(wp + u);
      id[4 * u] = a.x; id[4 * u + 1] = a.y; id[4 * u + 2] = a.z; id[4 * u + 3] = a.w;
      w[4 * u] = ww.x; w[4 * u + 1] = ww.y; w[4 * u + 2] = ww.z; w[4 * u + 3] = ww.w;
    }
    float* o = out + ((size_t)b * C + c0) * n + (size_t)t * 4;
#pragma unroll 2
    for (int c = 0; c < cn; ++c) {
      const float* fc = f + (size_t)c * m;
      float g[12];
#pragma unroll
      for (int u = 0; u < 12; ++u) g[u] = __ldg(fc + id[u]);
      float4 v;
      v.x = interp3(w[0], g[0], w[1], g[1], w[2], g[2]);
      v.y = interp3(w[3], g[3], w[4], g[4], w[5], g[5]);
      v.z = interp3(w[6], g[6], w[7], g[7], w[8], g[8]);
      v.w = interp3(w[9], g[9], w[10], g[10], w[11], g[11]);
      __stcs(reinterpret_cast<float4*>(o + (size_t)c * n), v);
    }
  } else {
    if (t >= n) return;
    const int32_t* ip = idx + ((size_t)b * n + t) * 3;
    const float* wp = weight + ((size_t)b * n + t) * 3;
    const int i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
    const float w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
    float* o = out + ((size_t)b * C + c0) * n + t;
#pragma unroll 4
    for (int c = 0; c < cn; ++c) {
      const float* fc = f + (size_t)c * m;
      __stcs(o + (size_t)c * n, interp3(w0, __ldg(fc + i0), w1, __ldg(fc + i1), w2, __ldg(fc + i2)));
    }
  }
}

// Row-staged variant: the chunk f[b, c0:c0+cn, :] (contiguous in the channel-first tensor) is brought into shared
// memory by one TMA bulk copy; the 12 gathers per thread and channel then hit shared-memory banks (~3 wavefronts per
// warp-wide gather instead of up to 32 L1 sector lookups).  Measured 52-58 % of the HBM peak at B >= 64; the bound
// is the shared-memory gather rate (12 four-byte gathers per 16 output bytes); a transposed [i][c] staging with
// 8-byte gathers was tried and lost more in the staging stores than it gained.
constexpr int TIS_T = 256;
__global__ void __launch_bounds__(TIS_T)
interp_fwd_staged_kernel(int C, int m, int quads, int cch, const float* __restrict__ features,
                         const int32_t* __restrict__ idx, const float* __restrict__ weight, float* __restrict__ out) {
  extern __shared__ __align__(128) float s_rows[];          // [cn][m]
  __shared__ __align__(8) uint64_t s_bar;
  using namespace sad;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * cch;
  const int cn = min(cch, C - c0);
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    const uint32_t bytes = (uint32_t)cn * (uint32_t)m * 4u;
    mbar_arrive_expect_tx(&s_bar, bytes);
    tma_bulk_g2s(s_rows, features + ((size_t)b * C + c0) * m, bytes, &s_bar);
  }
  __syncthreads();
  const int per = (quads + gridDim.x - 1) / gridDim.x;
  const int q0 = blockIdx.x * per, q1 = min(quads, q0 + per);
  const int4* ipb = reinterpret_cast<const int4*>(idx) + (size_t)b * quads * 3;
  const float4* wpb = reinterpret_cast<const float4*>(weight) + (size_t)b * quads * 3;
  float4* op = reinterpret_cast<float4*>(out) + ((size_t)b * C + c0) * quads;
  bool waited = false;
  for (int q = q0 + threadIdx.x; q < q1; q += TIS_T) {
    int id[12];
    float w[12];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int4 a = __ldg(ipb + (size_t)q * 3 + u);
      const float4 ww = __ldg(wpb + (size_t)q * 3 + u);
      id[4 * u] = a.x; id[4 * u + 1] = a.y; id[4 * u + 2] = a.z; id[4 * u + 3] = a.w;
      w[4 * u] = ww.x; w[4 * u + 1] = ww.y; w[4 * u + 2] = ww.z; w[4 * u + 3] = ww.w;
    }
    if (!waited) {
      mbar_wait(&s_bar, 0);
      waited = true;
    }
#pragma unroll 2
    for (int c = 0; c < cn; ++c) {
      const float* r = s_rows + (size_t)c * m;
      float4 v;
      v.x = interp3(w[0], r[id[0]], w[1], r[id[1]], w[2], r[id[2]]);
      v.y = interp3(w[3], r[id[3]], w[4], r[id[4]], w[5], r[id[5]]);
      v.z = interp3(w[6], r[id[6]], w[7], r[id[7]], w[8], r[id[8]]);
      v.w = interp3(w[9], r[id[9]], w[10], r[id[10]], w[11], r[id[11]]);
      __stcs(op + (size_t)c * quads + q, v);
    }
  }
  if (!waited) mbar_wait(&s_bar, 0);      // never leave with the bulk copy still in flight
}

__global__ void __launch_bounds__(TI_T)
interp_bwd_kernel(int C, int n, int m, const float* __restrict__ grad_out, const int32_t* __restrict__ idx,
                  const float* __restrict__ weight, float* __restrict__ grad_features) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * TI_CCH;
  const int cn = min(TI_CCH, C - c0);
  const int t = blockIdx.x * TI_T + threadIdx.x;
  if (t >= n) return;
  const int32_t* ip = idx + ((size_t)b * n + t) * 3;
  const float* wp = weight + ((size_t)b * n + t) * 3;
  const int i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
  const float w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
  const float* go = grad_out + ((size_t)b * C + c0) * n + t;
  float* g = grad_features + ((size_t)b * C + c0) * m;
#pragma unroll 4
  for (int c = 0; c < cn; ++c) {
    const float v = __ldcs(go + (size_t)c * n);
    float* gc = g + (size_t)c * m;
    atomicAdd(gc + i0, __fmul_rn(v, w0));
    atomicAdd(gc + i1, __fmul_rn(v, w1));
    atomicAdd(gc + i2, __fmul_rn(v, w2));
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" int sad_three_interpolate_fwd(int B, int C, int m, int n, const float* features, const int32_t* idx,
                                         const float* weight, float* out, sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && C >= 0 && m >= 1 && n >= 0, "three_interpolate: bad sizes B=%d C=%d m=%d n=%d", B, C, m, n);
  if (B == 0 || C == 0 || n == 0) return SAD_OK;
  SAD_REQUIRE(features && idx && weight && out, "three_interpolate: null pointer");
  SAD_REQUIRE(B <= 65535 && sad_ceil_div(C, TI_CCH) <= 65535, "three_interpolate: B/C exceed grid limits");
  const bool vec = (n % 4 == 0) && aligned16(idx) && aligned16(weight) && aligned16(out);
  const int cch = (m % 4 == 0 && m <= 4608) ? (18432 / m < 16 ? 18432 / m : 16) : 0;
  if (vec && cch >= 4 && 2LL * n >= m && aligned16(features)) {
    const int quads = n / 4;
    const int ychunks = sad_ceil_div(C, cch);
    long long x = sad_ceil_div(444, (long long)ychunks * B);
    const long long xmax = quads / TIS_T > 1 ? quads / TIS_T : 1;       // at least one full pass of the CTA
    if (x > xmax) x = xmax;
    if (x < 1) x = 1;
    SAD_REQUIRE(ychunks <= 65535, "three_interpolate: C exceeds grid limits");
    static thread_local int configured_dev = -1;
    int dev = 0;
    SAD_CUDA_OK(cudaGetDevice(&dev));
    if (configured_dev != dev) {
      SAD_CUDA_OK(cudaFuncSetAttribute(interp_fwd_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 73728));
      configured_dev = dev;
    }
    dim3 g2((unsigned)x, (unsigned)ychunks, (unsigned)B);
    interp_fwd_staged_kernel<<<g2, TIS_T, (size_t)cch * m * sizeof(float), (cudaStream_t)stream>>>(
        C, m, quads, cch, features, idx, weight, out);
    SAD_LAUNCH_CHECK("three_interpolate");
    return SAD_OK;
  }
  dim3 grid((unsigned)sad_ceil_div(vec ? n / 4 : n, TI_T), (unsigned)sad_ceil_div(C, TI_CCH), (unsigned)B);
  if (vec)
    interp_fwd_kernel<true><<<grid, TI_T, 0, (cudaStream_t)stream>>>(C, m, n, features, idx, weight, out);
  else
    interp_fwd_kernel<false><<<grid, TI_T, 0, (cudaStream_t)stream>>>(C, m, n, features, idx, weight, out);
  SAD_LAUNCH_CHECK("three_interpolate");
  return SAD_OK;
}

extern "C" int sad_three_interpolate_bwd(int B, int C, int n, int m, const float* grad_out, const int32_t* idx,
                                         const float* weight, float* grad_features, sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && C >= 0 && m >= 1 && n >= 0, "three_interpolate_bwd: bad sizes B=%d C=%d n=%d m=%d", B, C, n, m);
  if (B == 0 || C == 0) return SAD_OK;
  SAD_REQUIRE(grad_features, "three_interpolate_bwd: null pointer");
  SAD_CUDA_OK(cudaMemsetAsync(grad_features, 0, (size_t)B * C * m * sizeof(float), (cudaStream_t)stream));
  if (n == 0) return SAD_OK;
  SAD_REQUIRE(grad_out && idx && weight, "three_interpolate_bwd: null pointer");
  SAD_REQUIRE(B <= 65535 && sad_ceil_div(C, TI_CCH) <= 65535, "three_interpolate_bwd: B/C exceed grid limits");
  dim3 grid((unsigned)sad_ceil_div(n, TI_T), (unsigned)sad_ceil_div(C, TI_CCH), (unsigned)B);
  interp_bwd_kernel<<<grid, TI_T, 0, (cudaStream_t)stream>>>(C, n, m, grad_out, idx, weight, grad_features);
  SAD_LAUNCH_CHECK("three_interpolate_bwd");
  return SAD_OK;
}
