// Micro-benchmark: issue throughput of the fp32 instructions the search / FPS kernels are
// made of (scalar vs packed f32x2), to decide whether packed math is worth using.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes tools/fp32_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
  float r[ILP * 2];
#pragma unroll
  for (int i = 0; i < ILP * 2; ++i) r[i] = a + threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) {  // FADD
        r[2 * i] = __fadd_rn(r[2 * i], b);
        r[2 * i + 1] = __fadd_rn(r[2 * i + 1], b);
      } else if (MODE == 1) {  // FMUL
        r[2 * i] = __fmul_rn(r[2 * i], b);
        r[2 * i + 1] = __fmul_rn(r[2 * i + 1], b);
      } else if (MODE == 2) {  // FFMA
        r[2 * i] = __fmaf_rn(r[2 * i], b, a);
        r[2 * i + 1] = __fmaf_rn(r[2 * i + 1], b, a);
      } else if (MODE == 3) {  // add.f32x2
        unsigned long long v, w;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(r[2 * i]), "f"(r[2 * i + 1]));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(w) : "f"(b));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(w));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(r[2 * i]), "=f"(r[2 * i + 1]) : "l"(v));
      } else if (MODE == 4) {  // mul.f32x2
        unsigned long long v, w;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(r[2 * i]), "f"(r[2 * i + 1]));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(w) : "f"(b));
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(w));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(r[2 * i]), "=f"(r[2 * i + 1]) : "l"(v));
      } else if (MODE == 5) {  // fma.f32x2
        unsigned long long v, w, u;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(r[2 * i]), "f"(r[2 * i + 1]));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(w) : "f"(b));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(u) : "f"(a));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(w), "l"(u));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(r[2 * i]), "=f"(r[2 * i + 1]) : "l"(v));
      } else if (MODE == 6) {  // FMNMX
        r[2 * i] = fminf(r[2 * i], b + i);
        r[2 * i + 1] = fminf(r[2 * i + 1], a + i);
      } else if (MODE == 7) {  // the scalar distance test: 3 sub, 3 mul, 2 add
        const float dx = __fsub_rn(r[2 * i], a), dy = __fsub_rn(r[2 * i + 1], b), dz = __fsub_rn(r[2 * i], b);
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        r[2 * i] = fminf(r[2 * i], d);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP * 2; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_iter) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<MODE><<<148 * 8, 256>>>(out, 1.0f, 1.000001f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE><<<148 * 8, 256>>>(out, 1.0f, 1.000001f);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double thread_instr = (double)148 * 8 * 256 * ITERS * ops_per_iter;
  printf("%-28s %8.3f ms  %8.1f G lane-results/s  (%.1f results/clk/SM @1.9GHz)  err=%s\n", name, ms,
         thread_instr / ms / 1e6, thread_instr / ms / 1e6 / 148 / 1.9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("FADD scalar", ILP * 2);
  run<1>("FMUL scalar", ILP * 2);
  run<2>("FFMA scalar", ILP * 2);
  run<3>("add.f32x2 (results)", ILP * 2);
  run<4>("mul.f32x2 (results)", ILP * 2);
  run<5>("fma.f32x2 (results)", ILP * 2);
  run<6>("FMNMX", ILP * 2);
  run<7>("dist test (pairs)", ILP);
  return 0;
}
