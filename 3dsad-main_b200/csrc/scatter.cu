// a2 / a5 / a9 backward, DETERMINISTIC mode -- SURVEY.md section 7 H6, section 8(f) rank 4; VERDICT r1 item 7(c).
// (No reference file exists to cite: /root/reference is README.md:1-2 only.)
//
// The default backward kernels scatter with red.global.add.f32: the order in which the addends of one destination
// arrive changes from run to run, and fp32 addition is not associative.  This mode turns the scatter into a gather
// with a fixed summation order:
//   1. sad_scatter_plan_build: per scene a stable counting sort of the source positions by destination --
//      offsets (N + 1) and order (PS), positions of one destination in ASCENDING order.  One CTA per scene: histogram
//      (shared-memory atomics on integers: deterministic totals) -> block scan -> one warp walks the positions from
//      the last to the first, 32 at a time, ranks equal destinations with match.any and takes slots from the end of
//      each segment (the cursor array ends as the exclusive scan, i.e. as `offsets`).
//   2. sad_*_bwd_det: one thread per destination adds its segment sequentially, ((0 + v1) + v2) + ..., one rounding per
//      add, no atomics, every output written exactly once (no memset).
// Ascending source position is the order in which the oracle's np.add.at accumulates, so the result is not only
// reproducible bit for bit from run to run but bit-identical to oracle.*_grad (tests/test_ops_gpu.py).
// three_interpolate: the oracle adds neighbour slot t = 0 for every point, then t = 1, then t = 2; the plan's position
// key is therefore p = t * n + i for the source element idx[b, i, t] (`slots` = 3).
#include "sad_common.cuh"

namespace {

constexpr int SP_T = 1024;

// destination of plan position p: plain (slots == 1): idx[p]; interleaved source (slots == 3): idx[(p % n) * 3 + p / n]
__device__ __forceinline__ int plan_dest(const int32_t* idx, long long p, int n_inner, int slots) {
  if (slots == 1) return __ldg(idx + p);
  const int t = (int)(p / n_inner);
  const int i = (int)(p - (long long)t * n_inner);
  return __ldg(idx + (size_t)i * slots + t);
}

__global__ void __launch_bounds__(SP_T)
scatter_plan_kernel(int N, long long PS, int n_inner, int slots, int cur_in_smem, const int32_t* __restrict__ idx,
                    int32_t* __restrict__ order, int32_t* __restrict__ offsets) {
  extern __shared__ int s_cur[];
  __shared__ int s_part[SP_T];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int32_t* ib = idx + (size_t)b * PS;
  int32_t* ob = order + (size_t)b * PS;
  int32_t* fb = offsets + (size_t)b * (N + 1);
  int* cur = cur_in_smem ? s_cur : fb;
  for (int n = tid; n < N; n += SP_T) cur[n] = 0;
  __syncthreads();
  for (long long p = tid; p < PS; p += SP_T) {
    const int d = plan_dest(ib, p, n_inner, slots);
    if (d >= 0 && d < N) atomicAdd(cur + d, 1);
  }
  __syncthreads();
  // inclusive scan of cur[0..N): thread = contiguous run, block scan of the run totals
  const int per = (N + SP_T - 1) / SP_T;
  const int n0 = min(N, tid * per), n1 = min(N, n0 + per);
  int sum = 0;
  for (int n = n0; n < n1; ++n) sum += cur[n];
  s_part[tid] = sum;
  __syncthreads();
  for (int off = 1; off < SP_T; off <<= 1) {
    const int v = tid >= off ? s_part[tid - off] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int run = s_part[tid] - sum;
  for (int n = n0; n < n1; ++n) {
    run += cur[n];
    cur[n] = run;                 // = offsets[n + 1]
  }
  const int total = s_part[SP_T - 1];
  __syncthreads();
  // stable placement, last position first: slots are taken from the END of each destination's segment
  if (tid < 32) {
    const int lane = tid;
    const unsigned lt_gt = lane == 31 ? 0u : (0xFFFFFFFFu << (lane + 1));      // lanes above this one
    const long long chunks = (PS + 31) / 32;
    for (long long c = chunks - 1; c >= 0; --c) {
      const long long p = c * 32 + lane;
      int d = -1;
      if (p < PS) d = plan_dest(ib, p, n_inner, slots);
      const bool ok = d >= 0 && d < N;
      const unsigned live = __ballot_sync(0xFFFFFFFFu, ok);
      if (ok) {
        const unsigned peers = __match_any_sync(live, d);
        const int above = __popc(peers & lt_gt);                 // peers at higher positions take the higher slots
        const int base = cur[d];
        __syncwarp(live);
        ob[base - 1 - above] = (int32_t)p;
        if ((peers & lt_gt) == 0u) cur[d] = base - __popc(peers);   // the highest peer updates the cursor
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (cur_in_smem)
    for (int n = tid; n < N; n += SP_T) fb[n] = cur[n];
  if (tid == 0) fb[N] = total;
}

constexpr int SR_T = 128, SR_CCH = 16;
// MODE 0: grad_features[b,c,n] = sum over the segment of grad_out[b,c,p]              (grouping / gather)
// MODE 1: ... of grad_out[b,c,i] * weight[b,i,t], p = t * n_inner + i                 (three_interpolate)
template <int MODE>
__global__ void __launch_bounds__(SR_T)
segment_sum_kernel(int C, int N, long long PS, int n_inner, const float* __restrict__ grad_out,
                   const float* __restrict__ weight, const int32_t* __restrict__ order,
                   const int32_t* __restrict__ offsets, float* __restrict__ grad_features) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * SR_CCH;
  const int cn = min(SR_CCH, C - c0);
  const int n = blockIdx.x * SR_T + threadIdx.x;
  if (n >= N) return;
  const int32_t* fb = offsets + (size_t)b * (N + 1);
  const int32_t* ob = order + (size_t)b * PS;
  const int k0 = __ldg(fb + n), k1 = __ldg(fb + n + 1);
  const long long src_len = MODE == 0 ? PS : (long long)n_inner;
  const float* go = grad_out + ((size_t)b * C + c0) * src_len;
  float acc[SR_CCH];
#pragma unroll
  for (int c = 0; c < SR_CCH; ++c) acc[c] = 0.f;
  for (int k = k0; k < k1; ++k) {
    const int p = __ldg(ob + k);
    long long src = p;
    float w = 1.f;
    if (MODE == 1) {
      const int t = p / n_inner;
      const int i = p - t * n_inner;
      src = i;
      w = __ldg(weight + ((size_t)b * n_inner + i) * 3 + t);
    }
#pragma unroll
    for (int c = 0; c < SR_CCH; ++c) {
      if (c < cn) {
        float v = __ldg(go + (size_t)c * src_len + src);
        if (MODE == 1) v = __fmul_rn(v, w);
        acc[c] = __fadd_rn(acc[c], v);
      }
    }
  }
  float* g = grad_features + ((size_t)b * C + c0) * N + n;
#pragma unroll
  for (int c = 0; c < SR_CCH; ++c)
    if (c < cn) g[(size_t)c * N] = acc[c];
}

int plan_build(const char* name, int B, int N, long long PS, int n_inner, int slots, const int32_t* idx, int32_t* order,
               int32_t* offsets, cudaStream_t stream) {
  SAD_REQUIRE(B >= 0 && N >= 1 && PS >= 0 && PS < (1LL << 31), "%s: bad sizes B=%d N=%d PS=%lld", name, B, N, PS);
  if (B == 0) return SAD_OK;
  SAD_REQUIRE(offsets && (PS == 0 || (idx && order)), "%s: null pointer", name);
  SAD_REQUIRE(B <= 65535 * 32768, "%s: B exceeds grid limits", name);
  const size_t need = (size_t)N * sizeof(int);
  const int in_smem = need <= 200 * 1024;
  static thread_local int configured_dev = -1;
  int dev = 0;
  SAD_CUDA_OK(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    SAD_CUDA_OK(cudaFuncSetAttribute(scatter_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured_dev = dev;
  }
  scatter_plan_kernel<<<B, SP_T, in_smem ? need : 0, stream>>>(N, PS, n_inner, slots, in_smem, idx, order, offsets);
  SAD_LAUNCH_CHECK(name);
  return SAD_OK;
}

}  // namespace

extern "C" int sad_scatter_plan_build(int B, int N, long long PS, const int32_t* idx, int32_t* order, int32_t* offsets,
                                      sad_stream_t stream) {
  return plan_build("scatter_plan_build", B, N, PS, 1, 1, idx, order, offsets, (cudaStream_t)stream);
}

extern "C" int sad_interp_plan_build(int B, int n, int m, const int32_t* idx, int32_t* order, int32_t* offsets,
                                     sad_stream_t stream) {
  SAD_REQUIRE(n >= 0, "interp_plan_build: bad n");
  return plan_build("interp_plan_build", B, m, 3LL * n, n > 0 ? n : 1, 3, idx, order, offsets, (cudaStream_t)stream);
}

extern "C" int sad_scatter_add_det(int B, int C, int N, long long PS, const float* grad_out, const int32_t* order,
                                   const int32_t* offsets, float* grad_features, sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && C >= 0 && N >= 1 && PS >= 0, "scatter_add_det: bad sizes");
  if (B == 0 || C == 0) return SAD_OK;
  SAD_REQUIRE(offsets && grad_features && (PS == 0 || (grad_out && order)), "scatter_add_det: null pointer");
  SAD_REQUIRE(B <= 65535 && sad_ceil_div(C, SR_CCH) <= 65535, "scatter_add_det: B/C exceed grid limits");
  dim3 grid((unsigned)sad_ceil_div(N, SR_T), (unsigned)sad_ceil_div(C, SR_CCH), (unsigned)B);
  segment_sum_kernel<0><<<grid, SR_T, 0, (cudaStream_t)stream>>>(C, N, PS, 1, grad_out, nullptr, order, offsets, grad_features);
  SAD_LAUNCH_CHECK("scatter_add_det");
  return SAD_OK;
}

extern "C" int sad_three_interpolate_bwd_det(int B, int C, int n, int m, const float* grad_out, const float* weight,
                                             const int32_t* order, const int32_t* offsets, float* grad_features,
                                             sad_stream_t stream) {
  SAD_REQUIRE(B >= 0 && C >= 0 && m >= 1 && n >= 0, "three_interpolate_bwd_det: bad sizes");
  if (B == 0 || C == 0) return SAD_OK;
  SAD_REQUIRE(offsets && grad_features && (n == 0 || (grad_out && order && weight)), "three_interpolate_bwd_det: null pointer");
  SAD_REQUIRE(B <= 65535 && sad_ceil_div(C, SR_CCH) <= 65535, "three_interpolate_bwd_det: B/C exceed grid limits");
  dim3 grid((unsigned)sad_ceil_div(m, SR_T), (unsigned)sad_ceil_div(C, SR_CCH), (unsigned)B);
  segment_sum_kernel<1><<<grid, SR_T, 0, (cudaStream_t)stream>>>(C, m, 3LL * n, n > 0 ? n : 1, grad_out, weight, order, offsets,
                                                                 grad_features);
  SAD_LAUNCH_CHECK("three_interpolate_bwd_det");
  return SAD_OK;
}
