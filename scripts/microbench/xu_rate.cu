// Microbenchmark: issue rate of the XU-pipe instructions the softmax uses, one warp per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_rate xu_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

// MODE 0: ex2.approx.ftz.f32   1: ex2.approx.f16x2   2: ex2.approx.ftz.bf16x2
// MODE 3: cvt.rn.bf16x2.f32 (F2FP)   4: ex2 f32 + F2FP bf16 pack per 2   5: FFMA (fp32 pipe reference)
// MODE 6: fma.rn.f32x2
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(const float* in, uint32_t* out, int iters, long long* cycles) {
  constexpr int N = 32;
  float x[N];
  uint32_t h[N];
#pragma unroll
  for (int t = 0; t < N; ++t) { x[t] = in[(threadIdx.x + t * 7) & 1023]; h[t] = __float_as_uint(x[t]); }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < N; ++t) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[t]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[t]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[t]));
      if (MODE == 3) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[t]) : "f"(x[t]), "f"(x[(t + 1) % N]));
      if (MODE == 4) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[t]));
        if (t & 1) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[t]) : "f"(x[t]), "f"(x[t - 1]));
      }
      if (MODE == 5) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[t]) : "f"(x[(t + 1) % N]));
      if (MODE == 6 && (t & 1)) {
        uint64_t a, b;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x[t - 1]), "f"(x[t]));
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(a));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[t - 1]), "=f"(x[t]) : "l"(a));
      }
    }
  }
  const long long t1 = clock64();
  uint32_t sink = 0;
#pragma unroll
  for (int t = 0; t < N; ++t) sink ^= h[t] ^ __float_as_uint(x[t]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, const float* in, uint32_t* out, long long* dc, double per) {
  const int iters = 500;
  for (int rep = 0; rep < 2; ++rep) k<MODE><<<148, 128>>>(in, out, iters, dc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s error %s\n", name, cudaGetErrorString(e)); return; }
  long long hc[148]; cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += hc[i]; avg /= 148;
  printf("%-52s %6.2f cycles per warp instruction group (%.0f per loop of 32)\n", name, avg / iters / per, avg / iters);
}

int main() {
  float* in; uint32_t* out; long long* dc;
  cudaMalloc(&in, 4096);
  { float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = -0.001f * i; cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice); }
  cudaMalloc(&out, 148 * 128 * 4); cudaMalloc(&dc, 148 * 8);
  run<0>("ex2.approx.ftz.f32", in, out, dc, 32);
  run<1>("ex2.approx.f16x2 (2 values)", in, out, dc, 32);
  run<2>("ex2.approx.ftz.bf16x2 (2 values)", in, out, dc, 32);
  run<3>("cvt.rn.bf16x2.f32 (F2FP)", in, out, dc, 32);
  run<4>("2 x ex2 f32 + 1 x F2FP (per pair)", in, out, dc, 16);
  run<5>("fma.rn.f32", in, out, dc, 32);
  run<6>("fma.rn.f32x2 (per pair)", in, out, dc, 16);
  return 0;
}
