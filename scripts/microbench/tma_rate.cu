// Microbenchmark: how fast can one SM's TMA unit fill shared memory, as a function of the box
// row width (32/64/128 bytes), for 2-D boxes read from an L2-resident tensor?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu -I ../../omni_video_b200/csrc
// Every CTA (one per SM) issues `iters` box loads of `rows` x `width_bytes` into a ring of
// `stages` buffers; a consumer warp only waits on the full barrier and releases the slot.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ovb;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ CUtensorMap tm, int rows, int width_elems,
                                           int stages, int iters, int n_row_tiles, long long* cycles, int P, int nomod, int nload) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = (nload * rows * width_elems * 2 + 1023) / 1024 * 1024;
  uint64_t* full_all = (uint64_t*)(smem + P * stages * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 64; ++s) { mbar_init(&full_all[s], 1); mbar_init(&full_all[64 + s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  long long t0 = clock64();
  if (warp < P && lane == 0) {
    uint64_t* full = full_all + warp * 8; uint64_t* empty = full_all + 64 + warp * 8;
    uint8_t* base = smem + warp * stages * stage_bytes;
    int stage = 0; uint32_t phase = 0;
    int tile = (blockIdx.x * 7 + warp * 31) % n_row_tiles;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_expect_tx(&full[stage], nload * rows * width_elems * 2);
      if (nomod) { tile += 1; if (tile >= n_row_tiles) tile = 0; } else tile = (blockIdx.x * 7 + i + warp * 31) % n_row_tiles;
      for (int l = 0; l < nload; ++l)
        tma_load_2d(base + stage * stage_bytes + l * rows * width_elems * 2, &tm, &full[stage], ((i + l) & 3) * width_elems, tile * rows);
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 4 && warp < 4 + P && lane == 0) {
    uint64_t* full = full_all + (warp - 4) * 8; uint64_t* empty = full_all + 64 + (warp - 4) * 8;
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&full[stage], phase);
      mbar_arrive(&empty[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const int K = 512, M = 32768;  // 32 MB bf16 tensor: L2 resident
  void* d; cudaMalloc(&d, (size_t)K * M * 2); cudaMemset(d, 1, (size_t)K * M * 2);
  long long* dc; cudaMalloc(&dc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  printf("rows width_B stages | cycles/box  cycles/row  B/cycle/SM  chip TB/s@1.9GHz\n");
  for (int width_b : {128}) for (int rows : {32, 128}) for (int stages : {3}) for (int P : {1, 2}) for (int nload : {1, 2, 4, 8}) { int nomod = 1;
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)(width_b / 2), (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
    CUtensorMapSwizzle sw = width_b == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : width_b == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    const int stage_bytes = (nload * rows * width_b + 1023) / 1024 * 1024;
    if (P * stages * stage_bytes > 200 * 1024) continue;
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep)
      k<<<148, 256, P * stages * stage_bytes + 4096>>>(tm, rows, width_b / 2, stages, iters, M / rows, dc, P, nomod, nload);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long hc[148]; cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += hc[i]; avg /= 148;
    double cpb = avg / iters;
    printf("%4d %5d %4d P=%d nload=%d | %9.1f cycles/iteration/producer %9.1f cycles/box(SM) %9.1f B/cycle/SM\n", rows, width_b, stages, P, nload, cpb, cpb / P / nload,
           P * nload * rows * width_b / cpb);
  }
  return 0;
}
