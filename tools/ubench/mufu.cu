// MUFU.EX2 / F2FP throughput per SM on sm_100a (sizing the softmax of the attention kernels).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mufu tools/ubench/mufu.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // ex2 only
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 1) {  // fma + ex2 + add (softmax inner loop without the pack)
        float t;
        asm volatile("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(t) : "f"(a[i]), "f"(0.5f), "f"(-1.0f));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(t));
        a[i] = a[i] * 0.25f + t;
      } else if (MODE == 2) {  // bf16x2 pack only
        unsigned r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        acc ^= r;
        a[i] += 1.0f;
      } else {  // full inner loop: fma, ex2, add, pack every second element
        float t;
        asm volatile("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(t) : "f"(a[i]), "f"(0.5f), "f"(-1.0f));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(t));
        a[i] = a[i] * 0.25f + t;
        if (i & 1) {
          unsigned r;
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[i - 1]));
          acc ^= r;
        }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc & 1);
}

template <int MODE>
void run(const char* name, int warps_per_sm, int sms, float* out) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<sms, warps_per_sm * 32>>>(out, 16, 0.1f);
  cudaEventRecord(e0);
  k<MODE><<<sms, warps_per_sm * 32>>>(out, iters, 0.1f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double ops = double(iters) * 8 * warps_per_sm * 32;  // per SM
  const double clk = ms * 1e-3 * clk_khz * 1e3;
  printf("%-28s warps/SM %2d: %.2f loop-elements/clk/SM (%.3f ms, nominal clock %d MHz)\n", name, warps_per_sm, ops / clk, ms,
         clk_khz / 1000);
}

int main() {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sms * 1024 * sizeof(float));
  for (int w : {4, 8, 16, 32}) {
    run<0>("ex2 only", w, sms, out);
    run<1>("fma+ex2+fma", w, sms, out);
    run<2>("cvt.bf16x2 (+add)", w, sms, out);
    run<3>("fma+ex2+fma+cvt/2", w, sms, out);
  }
  return 0;
}
