// Microbenchmark: throughput of a compare-exchange (CE) on sm_100a, three codings:
//   0 plain    : lo = FMNMX(a,b); hi = FMNMX(a,b)                      (2 ALU-pipe ops)
//   1 imad     : lo = FMNMX(a,b); hi = a + b - lo in integer arithmetic via two IMADs with a runtime
//                multiplier (1, -1) so that ptxas cannot fold them into IADD3   (1 ALU + 2 FMA-pipe ops)
//   2 mixed    : every third CE plain, the others imad
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ce profiles/microbench_ce.cu && /tmp/ce
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void ce_plain(float& a, float& b) { float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }
__device__ __forceinline__ void ce_imad(float& a, float& b, unsigned one, unsigned mone) {
  float lo = fminf(a, b);
  unsigned s, h;
  asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s) : "r"(__float_as_uint(a)), "r"(one), "r"(__float_as_uint(b)));
  asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(h) : "r"(__float_as_uint(lo)), "r"(mone), "r"(s));
  a = lo; b = __uint_as_float(h);
}
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* p, int reps, unsigned one, unsigned mone) {
  float r[32];
#pragma unroll
  for (int i = 0; i < 32; i++) r[i] = p[threadIdx.x + i * blockDim.x + blockIdx.x * 32 * blockDim.x];
  for (int rep = 0; rep < reps; rep++) {
#pragma unroll
    for (int j = 16; j > 0; j >>= 1)
#pragma unroll
      for (int i = 0; i < 32; i++)
        if (!(i & j)) {
          if (MODE == 0 || (MODE == 2 && (i % 3 == 0))) ce_plain(r[i], r[i | j]);
          else ce_imad(r[i], r[i | j], one, mone);
        }
  }
#pragma unroll
  for (int i = 0; i < 32; i++) p[threadIdx.x + i * blockDim.x + blockIdx.x * 32 * blockDim.x] = r[i];
}
template <int MODE> float run(float* p, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, 1024>>>(p, 10, 1u, 0xffffffffu);
  cudaEventRecord(e0);
  k<MODE><<<148, 1024>>>(p, reps, 1u, 0xffffffffu);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* p; cudaMalloc(&p, 148 * 1024 * 32 * 4); cudaMemset(p, 0, 148 * 1024 * 32 * 4);
  const int reps = 2000;
  const double ces = 148.0 * 1024 * reps * 80;  // 5 stages x 16 CE per thread per rep
  float t0 = run<0>(p, reps), t1 = run<1>(p, reps), t2 = run<2>(p, reps);
  printf("{\"plain_ms\": %.3f, \"imad_ms\": %.3f, \"mixed_ms\": %.3f, \"plain_GCEps\": %.1f, \"imad_GCEps\": %.1f, \"mixed_GCEps\": %.1f}\n",
         t0, t1, t2, ces / t0 / 1e6, ces / t1 / 1e6, ces / t2 / 1e6);
  return 0;
}
