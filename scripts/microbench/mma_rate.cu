// Microbenchmark: tcgen05.mma issue/execute rate from ONE issuing thread, M=128, K=16, bf16,
// as a function of N, with (a) descriptors rebuilt per MMA like our kernels do, (b) hoisted.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu -I ../../omni_video_b200/csrc
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ovb;

__global__ void __launch_bounds__(128, 1) k(int N, int iters, int per_commit, int mode, long long* cycles) {
  const int hoist = 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 8); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(mode == 2 ? 64 : 128, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = sa + 32768;
    const uint64_t da0 = umma_desc_sw128(sa, 16, 1024), db0 = umma_desc_sw128(sb, 16, 1024);
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      for (int j = 0; j < per_commit; ++j) {
        if (mode == 1) umma_ts(tmem, tmem + 256 + (j & 3) * 8, db0 + ((j & 3) * 2), idesc, 1);
        else if (hoist) umma_ss(tmem, da0 + ((j & 3) * 2), db0 + ((j & 3) * 2), idesc, 1);
        else umma_ss(tmem, umma_desc_sw128(sa + (j & 3) * 32, 16, 1024), umma_desc_sw128(sb + (j & 3) * 32, 16, 1024), idesc, (i | j) != 0);
      }
      umma_commit(&bar);
      if ((i & 7) == 7) {  // keep a bounded number of commits outstanding
        mbar_wait(&bar, phase); phase ^= 1;
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

int main() {
  long long* dc; cudaMalloc(&dc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  printf("mode 0=SS M128, 1=TS M128 (A in TMEM), 2=SS M64\n   N per_commit mode | cycles/MMA  (tensor floor N/2)  TFLOP/s chip @1.9GHz\n");
  for (int hoist : {0, 1, 2}) for (int N : {32, 64, 128, 192, 256}) for (int pc : {8, 16}) {
    const int iters = 2048;
    for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 100 * 1024>>>(N, iters, pc, hoist, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long hc[148]; cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += hc[i]; avg /= 148;
    double cpm = avg / iters / pc;
    printf("%4d %6d %6d | %9.1f %9.1f %12.0f\n", N, pc, hoist, cpm, N / 2.0, 2.0 * 128 * N * 16 / cpm * 148 * 1.9e9 / 1e12);
  }
  return 0;
}
