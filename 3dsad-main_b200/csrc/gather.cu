// a2 gather_operation, a5 grouping_operation (forward + backward) -- SURVEY.md 8(a).
// (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
// HBM-bound copies.  Channel-first public layout: out[b,c,j,s] = f[b,c,idx[b,j,s]].
// One thread owns FOUR consecutive (j,s) positions: it loads their indices once
// (one 128-bit load), then walks a chunk of channels issuing 4 independent 4-byte
// gathers (the source row f[b,c,:] is L1/L2 resident) and ONE 128-bit streaming store
// per channel, so every store instruction of a warp writes 512 contiguous bytes.
// gather_operation is grouping_operation with nsample == 1.
//
// Row-staged variant (group_fwd_staged_kernel), used when a chunk of source rows fits shared memory
// and every row is read many times (P*S >= 2N: the SA2..SA4 / vote-aggregation shapes): the chunk
// f[b, c0:c0+cn, :] is ONE contiguous block of the channel-first tensor, so a single TMA bulk copy
// brings it on chip; the 4-byte gathers then hit shared-memory banks instead of L1 sectors (a
// random warp-wide gather costs ~3 bank wavefronts instead of up to 32 sector lookups), and HBM sees
// each source row once plus the streaming 128-bit stores.
#include "sad_common.cuh"

namespace {

constexpr int GR_T = 256;
// Indices are untrusted input: forward gathers clamp them into [0, N) (one unsigned min; a negative index becomes N - 1),
// backward scatters skip anything outside -- no out-of-bounds access either way.
__device__ __forceinline__ int clamp_idx(int i, int N) { return (int)min((unsigned)i, (unsigned)(N - 1)); }
__device__ __forceinline__ int4 clamp_idx4(int4 i, int N) {
  return make_int4(clamp_idx(i.x, N), clamp_idx(i.y, N), clamp_idx(i.z, N), clamp_idx(i.w, N));
}
__device__ __forceinline__ bool idx_ok(int i, int N) { return (unsigned)i < (unsigned)N; }

constexpr int GR_CCH = 16;   // channels per thread (grid.y splits the rest)

template <bool VEC>
__global__ void __launch_bounds__(GR_T)
group_fwd_kernel(int C, int N, long long PS, const float* __restrict__ features,
                 const int32_t* __restrict__ idx, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * GR_CCH;
  const int cn = min(GR_CCH, C - c0);
  const long long t = (long long)blockIdx.x * GR_T + threadIdx.x;
  const float* f = features + ((size_t)b * C + c0) * N;
  if (VEC) {
    if (t * 4 >= PS) return;
    const int4 id = clamp_idx4(__ldg(reinterpret_cast<const int4*>(idx + (size_t)b * PS) + t), N);
    float* o = out + ((size_t)b * C + c0) * PS + t * 4;
    int c = 0;
    for (; c + 4 <= cn; c += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* fu = f + (size_t)(c + u) * N;
        v[u] = make_float4(__ldg(fu + id.x), __ldg(fu + id.y), __ldg(fu + id.z), __ldg(fu + id.w));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) __stcs(reinterpret_cast<float4*>(o + (size_t)(c + u) * PS), v[u]);
    }
    for (; c < cn; ++c) {
      const float* fu = f + (size_t)c * N;
      __stcs(reinterpret_cast<float4*>(o + (size_t)c * PS),
             make_float4(__ldg(fu + id.x), __ldg(fu + id.y), __ldg(fu + id.z), __ldg(fu + id.w)));
    }
  } else {
    if (t >= PS) return;
    const int id = clamp_idx(__ldg(idx + (size_t)b * PS + t), N);
    float* o = out + ((size_t)b * C + c0) * PS + t;
#pragma unroll 4
    for (int c = 0; c < cn; ++c) __stcs(o + (size_t)c * PS, __ldg(f + (size_t)c * N + id));
  }
}

__global__ void __launch_bounds__(GR_T)
group_fwd_staged_kernel(int C, int N, long long quads, int cch, const float* __restrict__ features,
                        const int32_t* __restrict__ idx, float* __restrict__ out) {
  extern __shared__ __align__(128) float s_rows[];          // [cn][N]
  __shared__ __align__(8) uint64_t s_bar;
  using namespace sad;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * cch;
  const int cn = min(cch, C - c0);
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    const uint32_t bytes = (uint32_t)cn * (uint32_t)N * 4u;
    mbar_arrive_expect_tx(&s_bar, bytes);
    tma_bulk_g2s(s_rows, features + ((size_t)b * C + c0) * N, bytes, &s_bar);
  }
  __syncthreads();
  // this CTA's share of the position quads
  const long long per = (quads + gridDim.x - 1) / gridDim.x;
  const long long q0 = (long long)blockIdx.x * per, q1 = min(quads, q0 + per);
  const int4* ip = reinterpret_cast<const int4*>(idx) + (size_t)b * quads;
  float4* op = reinterpret_cast<float4*>(out) + ((size_t)b * C + c0) * quads;
  long long q = q0 + threadIdx.x;
  int4 id = (q < q1) ? clamp_idx4(__ldg(ip + q), N) : make_int4(0, 0, 0, 0);
  mbar_wait(&s_bar, 0);
  for (; q < q1; q += GR_T) {
    const long long qn = q + GR_T;
    const int4 nxt = (qn < q1) ? clamp_idx4(__ldg(ip + qn), N) : make_int4(0, 0, 0, 0);      // next indices while this quad gathers
#pragma unroll 4
    for (int c = 0; c < cn; ++c) {
      const float* r = s_rows + (size_t)c * N;
      __stcs(op + (size_t)c * quads + q, make_float4(r[id.x], r[id.y], r[id.z], r[id.w]));
    }
    id = nxt;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(GR_T)
group_bwd_kernel(int C, int N, long long PS, const float* __restrict__ grad_out,
                 const int32_t* __restrict__ idx, float* __restrict__ grad_features) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * GR_CCH;
  const int cn = min(GR_CCH, C - c0);
  const long long t = (long long)blockIdx.x * GR_T + threadIdx.x;
  float* g = grad_features + ((size_t)b * C + c0) * N;
  if (VEC) {
    if (t * 4 >= PS) return;
    const int4 id = __ldg(reinterpret_cast<const int4*>(idx + (size_t)b * PS) + t);
    const float* go = grad_out + ((size_t)b * C + c0) * PS + t * 4;
#pragma unroll 4
    for (int c = 0; c < cn; ++c) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(go + (size_t)c * PS));
      float* gc = g + (size_t)c * N;
      if (idx_ok(id.x, N)) atomicAdd(gc + id.x, v.x);
      if (idx_ok(id.y, N)) atomicAdd(gc + id.y, v.y);
      if (idx_ok(id.z, N)) atomicAdd(gc + id.z, v.z);
      if (idx_ok(id.w, N)) atomicAdd(gc + id.w, v.w);
    }
  } else {
    if (t >= PS) return;
    const int id = __ldg(idx + (size_t)b * PS + t);
    if (!idx_ok(id, N)) return;
    const float* go = grad_out + ((size_t)b * C + c0) * PS + t;
#pragma unroll 4
    for (int c = 0; c < cn; ++c) atomicAdd(g + (size_t)c * N + id, __ldcs(go + (size_t)c * PS));
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int group_fwd(const char* name, int B, int C, int N, long long PS, const float* features, const int32_t* idx,
              float* out, cudaStream_t stream) {
  SAD_REQUIRE(B >= 0 && C >= 0 && N >= 1 && PS >= 0, "%s: bad sizes B=%d C=%d N=%d PS=%lld", name, B, C, N, PS);
  if (B == 0 || C == 0 || PS == 0) return SAD_OK;
  SAD_REQUIRE(features && idx && out, "%s: null pointer", name);
  SAD_REQUIRE(B <= 65535 && sad_ceil_div(C, GR_CCH) <= 65535, "%s: B/C exceed grid limits", name);
  const bool vec = (PS % 4 == 0) && aligned16(idx) && aligned16(out);
  // row-staged path: rows fit (<= 72 KB per CTA, so three CTAs share an SM), are 16-byte granular, and are re-read
  // often enough (P*S >= 2N) to pay for staging
  const int cch = (N % 4 == 0 && N <= 4608) ? (int)(18432 / N < 16 ? 18432 / N : 16) : 0;
  if (vec && cch >= 4 && PS >= 2LL * N && aligned16(features)) {
    const long long quads = PS / 4;
    const int ychunks = sad_ceil_div(C, cch);
    long long x = sad_ceil_div(444, (long long)ychunks * B);            // ~3 CTAs per SM in flight ...
    const long long xmax = quads / N > 1 ? quads / N : 1;               // ... but >= N quads each (4x the staged bytes)
    if (x > xmax) x = xmax;
    if (x < 1) x = 1;
    SAD_REQUIRE(ychunks <= 65535, "%s: C exceeds grid limits", name);
    const size_t smem = (size_t)cch * N * sizeof(float);
    static thread_local int configured_dev = -1;
    int dev = 0;
    SAD_CUDA_OK(cudaGetDevice(&dev));
    if (configured_dev != dev) {
      SAD_CUDA_OK(cudaFuncSetAttribute(group_fwd_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 73728));
      configured_dev = dev;
    }
    dim3 g2((unsigned)x, (unsigned)ychunks, (unsigned)B);
    group_fwd_staged_kernel<<<g2, GR_T, smem, stream>>>(C, N, quads, cch, features, idx, out);
    SAD_LAUNCH_CHECK(name);
    return SAD_OK;
  }
  dim3 grid((unsigned)sad_ceil_div(vec ? PS / 4 : PS, GR_T), (unsigned)sad_ceil_div(C, GR_CCH), (unsigned)B);
  if (vec)
    group_fwd_kernel<true><<<grid, GR_T, 0, stream>>>(C, N, PS, features, idx, out);
  else
    group_fwd_kernel<false><<<grid, GR_T, 0, stream>>>(C, N, PS, features, idx, out);
  SAD_LAUNCH_CHECK(name);
  return SAD_OK;
}

int group_bwd(const char* name, int B, int C, int N, long long PS, const float* grad_out, const int32_t* idx,
              float* grad_features, cudaStream_t stream) {
  SAD_REQUIRE(B >= 0 && C >= 0 && N >= 1 && PS >= 0, "%s: bad sizes B=%d C=%d N=%d PS=%lld", name, B, C, N, PS);
  if (B == 0 || C == 0) return SAD_OK;
  SAD_REQUIRE(grad_features, "%s: null pointer", name);
  SAD_CUDA_OK(cudaMemsetAsync(grad_features, 0, (size_t)B * C * N * sizeof(float), stream));
  if (PS == 0) return SAD_OK;
  SAD_REQUIRE(grad_out && idx, "%s: null pointer", name);
  SAD_REQUIRE(B <= 65535 && sad_ceil_div(C, GR_CCH) <= 65535, "%s: B/C exceed grid limits", name);
  const bool vec = (PS % 4 == 0) && aligned16(idx) && aligned16(grad_out);
  dim3 grid((unsigned)sad_ceil_div(vec ? PS / 4 : PS, GR_T), (unsigned)sad_ceil_div(C, GR_CCH), (unsigned)B);
  if (vec)
    group_bwd_kernel<true><<<grid, GR_T, 0, stream>>>(C, N, PS, grad_out, idx, grad_features);
  else
    group_bwd_kernel<false><<<grid, GR_T, 0, stream>>>(C, N, PS, grad_out, idx, grad_features);
  SAD_LAUNCH_CHECK(name);
  return SAD_OK;
}

}  // namespace

extern "C" int sad_grouping_operation_fwd(int B, int C, int N, int npoint, int nsample, const float* features,
                                          const int32_t* idx, float* out, sad_stream_t stream) {
  SAD_REQUIRE(npoint >= 0 && nsample >= 0, "grouping_operation: bad npoint/nsample");
  return group_fwd("grouping_operation", B, C, N, (long long)npoint * nsample, features, idx, out,
                   (cudaStream_t)stream);
}

extern "C" int sad_grouping_operation_bwd(int B, int C, int N, int npoint, int nsample, const float* grad_out,
                                          const int32_t* idx, float* grad_features, sad_stream_t stream) {
  SAD_REQUIRE(npoint >= 0 && nsample >= 0, "grouping_operation_bwd: bad npoint/nsample");
  return group_bwd("grouping_operation_bwd", B, C, N, (long long)npoint * nsample, grad_out, idx, grad_features,
                   (cudaStream_t)stream);
}

// new_xyz (B,P,3) = xyz[b, inds[b,j]] straight from the (B,N,3) layout (the lineage idiom transposes twice around
// gather_operation: three launches); optionally also the padded float4 rows the fused SA kernel gathers from.
__global__ void __launch_bounds__(256)
gather_points_kernel(long long rows, int N, int P, const float* __restrict__ xyz, const int32_t* __restrict__ inds,
                     float* __restrict__ out, float4* __restrict__ out_xyzw) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows) return;
  const long long b = i / P;
  int k = __ldg(inds + i);
  k = (unsigned)k < (unsigned)N ? k : 0;
  const float* p = xyz + ((size_t)b * N + k) * 3;
  const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
  out[3 * i] = x;
  out[3 * i + 1] = y;
  out[3 * i + 2] = z;
  if (out_xyzw) out_xyzw[i] = make_float4(x, y, z, 0.f);
}

extern "C" int sad_gather_points_fwd(int B, int N, int npoint, const float* xyz, const int32_t* inds, float* new_xyz,
                                     void* new_xyzw, sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && N >= 1 && npoint >= 0, "gather_points: bad sizes");
  if (B == 0 || npoint == 0) return SAD_OK;
  SAD_REQUIRE(xyz && inds && new_xyz, "gather_points: null pointer");
  const long long rows = (long long)B * npoint;
  gather_points_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, N, npoint, xyz, inds, new_xyz,
                                                                                        static_cast<float4*>(new_xyzw));
  SAD_LAUNCH_CHECK("gather_points_kernel");
  return SAD_OK;
}

extern "C" int sad_gather_operation_fwd(int B, int C, int N, int npoint, const float* features,
                                        const int32_t* idx, float* out, sad_stream_t stream) {
  SAD_REQUIRE(npoint >= 0, "gather_operation: bad npoint");
  return group_fwd("gather_operation", B, C, N, (long long)npoint, features, idx, out, (cudaStream_t)stream);
}

extern "C" int sad_gather_operation_bwd(int B, int C, int N, int npoint, const float* grad_out,
                                        const int32_t* idx, float* grad_features, sad_stream_t stream) {
  SAD_REQUIRE(npoint >= 0, "gather_operation_bwd: bad npoint");
  return group_bwd("gather_operation_bwd", B, C, N, (long long)npoint, grad_out, idx, grad_features,
                   (cudaStream_t)stream);
}
