// Microbenchmarks behind two round-2 design decisions ("what bounds the train kernel"):
//   (1) shared-memory atomics vs plain read-modify-write for a per-column bucket histogram
//       ([bucket][lane] counters: bank = lane, so a warp never has a bank conflict), to decide whether a
//       distribution (bucket) sort can replace the compare-exchange network;
//   (2) 3-input min/max (sm_100 `min.f32 d, a, b, c`) vs the 2-input form as a sorting-network primitive.
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb profiles/microbench_smem.cu && /tmp/mb
#include <cuda_runtime.h>
#include <cstdio>

constexpr int kBuckets = 512;

__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// MODE 0: atomicAdd, result unused (RED); 1: atomicAdd, result used; 2: plain LDS + IADD + STS (private-safe only)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) hist_kernel(unsigned* out, int reps) {
  extern __shared__ unsigned h[];  // [kBuckets][32]
  for (int i = threadIdx.x; i < kBuckets * 32; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  unsigned s = threadIdx.x * 2654435761u + blockIdx.x, acc = 0;
  for (int r = 0; r < reps; ++r) {
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const unsigned b = lcg(s) & (kBuckets - 1);
      unsigned* p = h + b * 32 + lane;
      if (MODE == 0) atomicAdd(p, 1u);
      else if (MODE == 1) acc += atomicAdd(p, 1u);
      else { *(volatile unsigned*)p = *(volatile unsigned*)p + 1u; }
    }
  }
  __syncthreads();
  unsigned t = acc;
  for (int i = threadIdx.x; i < kBuckets * 32; i += blockDim.x) t += h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE> float run_hist(unsigned* out, int reps) {
  cudaFuncSetAttribute(hist_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBuckets * 32 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  hist_kernel<MODE><<<148, 1024, kBuckets * 32 * 4>>>(out, 2);
  cudaEventRecord(e0);
  hist_kernel<MODE><<<148, 1024, kBuckets * 32 * 4>>>(out, reps);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

// 3-sorter codings on 33 registers (11 triples per stage-like sweep)
__device__ __forceinline__ float min3(float a, float b, float c) {
  float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
template <int MODE>
__global__ void __launch_bounds__(1024, 1) sort3_kernel(float* p, int reps, unsigned one, unsigned mone) {
  float r[33];
#pragma unroll
  for (int i = 0; i < 33; i++) r[i] = p[threadIdx.x + i * blockDim.x];
  for (int rep = 0; rep < reps; rep++) {
#pragma unroll
    for (int sh = 0; sh < 3; ++sh)
#pragma unroll
      for (int i = 0; i < 30; i += 3) {
        float &a = r[i + sh], &b = r[i + sh + 1], &c = r[i + sh + 2];
        if (MODE == 0) {  // three 2-input compare-exchanges (6 FMNMX)
          float t;
          t = fminf(a, b); b = fmaxf(a, b); a = t;
          t = fminf(b, c); c = fmaxf(b, c); b = t;
          t = fminf(a, b); b = fmaxf(a, b); a = t;
        } else {          // min3 + max3 + median rebuilt in integer arithmetic (exact): a + b + c - lo - hi
          const float lo = min3(a, b, c), hi = max3(a, b, c);
          unsigned s1, s2, s3, s4;
          asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s1) : "r"(__float_as_uint(a)), "r"(one), "r"(__float_as_uint(b)));
          asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s2) : "r"(__float_as_uint(c)), "r"(one), "r"(s1));
          asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s3) : "r"(__float_as_uint(lo)), "r"(mone), "r"(s2));
          asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s4) : "r"(__float_as_uint(hi)), "r"(mone), "r"(s3));
          a = lo; b = __uint_as_float(s4); c = hi;
        }
      }
  }
#pragma unroll
  for (int i = 0; i < 33; i++) p[threadIdx.x + i * blockDim.x] = r[i];
}
template <int MODE> float run_sort3(float* p, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  sort3_kernel<MODE><<<148, 1024>>>(p, 4, 1u, 0xffffffffu);
  cudaEventRecord(e0);
  sort3_kernel<MODE><<<148, 1024>>>(p, reps, 1u, 0xffffffffu);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  unsigned* out; cudaMalloc(&out, 148 * 1024 * 4);
  float* p; cudaMalloc(&p, 1024 * 33 * 4); cudaMemset(p, 0, 1024 * 33 * 4);
  const int reps = 400;
  const double ops = 148.0 * 1024 * 32 * reps;  // lane-updates
  const float t0 = run_hist<0>(out, reps), t1 = run_hist<1>(out, reps), t2 = run_hist<2>(out, reps);
  const double clk = 1.965e9;
  printf("{\"hist_lane_updates_per_clk_per_sm\": {\"atomic_noret\": %.2f, \"atomic_ret\": %.2f, \"lds_add_sts\": %.2f},\n",
         ops / (t0 * 1e-3) / 148 / clk, ops / (t1 * 1e-3) / 148 / clk, ops / (t2 * 1e-3) / 148 / clk);
  const int reps3 = 2000;
  const double s3 = 148.0 * 1024 * 30.0 * reps3;  // 3-sorters
  const float u0 = run_sort3<0>(p, reps3), u1 = run_sort3<1>(p, reps3);
  printf(" \"sort3_per_clk_per_sm\": {\"three_ce\": %.2f, \"min3_max3_imad\": %.2f}, \"err\": \"%s\"}\n",
         s3 / (u0 * 1e-3) / 148 / clk, s3 / (u1 * 1e-3) / 148 / clk, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
