// Dependent-chain latencies of the warp primitives the FPS round is made of (one warp, one SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_bench tools/lat_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
constexpr int IT = 4096;

template <int MODE>
__global__ void k(unsigned* out, long long* cyc, unsigned seed) {
  __shared__ unsigned sm[64];
  __shared__ __align__(8) unsigned long long bar;
  sm[threadIdx.x & 63] = threadIdx.x;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
  __syncthreads();
  unsigned v = seed + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < IT; ++i) {
    if (MODE == 0) v = __reduce_max_sync(FULL, v) + threadIdx.x;
    if (MODE == 1) v = __ballot_sync(FULL, v & 1) + threadIdx.x;
    if (MODE == 2) v = __shfl_sync(FULL, v, (v + 1) & 31) + 1;
    if (MODE == 3) v = sm[v & 63] + 1;
    if (MODE == 4) v = __ffs(v | 1) + v;
    if (MODE == 5) { __syncthreads(); v += 1; }
    if (MODE == 6) { __syncwarp(); v += 1; }
    if (MODE == 7) v = __float_as_uint(fminf(__uint_as_float(v), 3.0f) + 1.0f);
    if (MODE == 8) {   // local mbarrier arrive + try_wait round trip (phase flips every iteration)
      unsigned a = (unsigned)__cvta_generic_to_shared(&bar);
      if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
      unsigned ok = 0;
      while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0,1,0,p;}" : "=r"(ok) : "r"(a), "r"(i & 1) : "memory");
      v += ok;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  unsigned* out; long long* cyc; long long h;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  k<MODE><<<1, threads>>>(out, cyc, 1);
  k<MODE><<<1, threads>>>(out, cyc, 1);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s threads=%4d  %7.1f cycles/iter  (%s)\n", name, threads, (double)h / IT, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("redux.max + add", 32);
  run<1>("ballot + add", 32);
  run<2>("shfl + add", 32);
  run<3>("lds + add", 32);
  run<4>("ffs + add", 32);
  run<5>("__syncthreads (4 warps)", 128);
  run<5>("__syncthreads (16 warps)", 512);
  run<5>("__syncthreads (32 warps)", 1024);
  run<0>("redux.max + add (32 warps)", 1024);
  run<3>("lds + add (32 warps)", 1024);
  run<6>("__syncwarp", 32);
  run<7>("fmin + fadd", 32);
  run<8>("mbarrier arrive+try_wait (1 warp)", 32);
  run<8>("mbarrier arrive+try_wait (4 warps)", 128);
  return 0;
}
