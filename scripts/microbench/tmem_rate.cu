// Microbenchmark: what tcgen05.st / tcgen05.ld cost the ISSUING warp, alone and inside the attention
// softmax pass (S fp32 from TMEM -> exp2 -> P bf16 back to TMEM).  4 warps, one per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_rate tmem_rate.cu -I ../../omni_video_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ovb;

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// MODE 0: back-to-back st.x16            MODE 1: back-to-back ld.x32 (+wait)
// MODE 2: st.x16 + 64 dependent FFMAs    MODE 3: 64 dependent FFMAs only
// MODE 4: softmax pass, P -> TMEM x16 in the loop (the kernel's form)
// MODE 5: softmax pass, P not stored (xor sink)
// MODE 6: softmax pass, P -> TMEM as 8 x8 stores
// MODE 7: softmax pass, P -> TMEM as 2 x32 stores after the loop
// MODE 8: softmax pass, P -> shared memory (16-byte st.shared, swizzle-free row-per-thread)
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(const float* in, uint32_t* out, int iters, long long* cycles, int active_warps) {
  __shared__ uint32_t tmem_slot;
  extern __shared__ __align__(1024) uint8_t dyn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot + (static_cast<uint32_t>(warp * 32) << 16);
  // fill S columns [0,128) with scores
  {
    uint32_t r[32];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int t = 0; t < 32; ++t) r[t] = __float_as_uint(in[(threadIdx.x * 131 + (c * 32 + t) * 7) & 1023]);
      tmem_st32(tbase + c * 32, r);
    }
    tmem_wait_st();
  }
  uint32_t sink = 0;
  float chain = in[lane];
  __syncthreads();
  long long t0 = 0, t1 = 0;
  if (warp < active_warps) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0) {
        uint32_t pk[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) pk[t] = sink + t;
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st16(tbase + 256 + c * 16, pk);
        sink += 1;
      } else if (MODE == 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(tbase + c * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int t = 0; t < 32; ++t) sink ^= r[t];
        }
      } else if (MODE == 2 || MODE == 3) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (MODE == 2) {
            uint32_t pk[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) pk[t] = __float_as_uint(chain) + t;
            tmem_st16(tbase + 256 + c * 16, pk);
          }
#pragma unroll
          for (int t = 0; t < 64; ++t) chain = fmaf(chain, 1.0001f, 0.5f);
        }
      } else {
        float s[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(tbase + c * 32, r);
#pragma unroll
          for (int t = 0; t < 32; ++t) s[c * 32 + t] = __uint_as_float(r[t]);
        }
        tmem_wait_ld();
        const float scale = 0.127f, neg_m = -0.73f + 1e-6f * it;
        const uint64_t scale2 = pack_f32x2(scale, scale), negm2 = pack_f32x2(neg_m, neg_m);
        uint64_t acc_a = 0ull, acc_b = 0ull;
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        uint32_t pk_all[64];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t* pk = pk_all + c * 16;
#pragma unroll
          for (int t = 0; t < 32; t += 4) {
            const float s0 = s[c * 32 + t], s1 = s[c * 32 + t + 1], s2 = s[c * 32 + t + 2], s3 = s[c * 32 + t + 3];
            m0 = fmaxf(m0, s0); m1 = fmaxf(m1, s1); m2 = fmaxf(m2, s2); m3 = fmaxf(m3, s3);
            float x0, x1, x2, x3;
            unpack_f32x2(ffma2(pack_f32x2(s0, s1), scale2, negm2), x0, x1);
            unpack_f32x2(ffma2(pack_f32x2(s2, s3), scale2, negm2), x2, x3);
            const float e0 = fast_exp2(x0), e1 = fast_exp2(x1), e2 = fast_exp2(x2), e3 = fast_exp2(x3);
            acc_a = fadd2(acc_a, pack_f32x2(e0, e1));
            acc_b = fadd2(acc_b, pack_f32x2(e2, e3));
            pk[t / 2] = pack_bf16x2(e0, e1);
            pk[t / 2 + 1] = pack_bf16x2(e2, e3);
          }
          if (MODE == 4) tmem_st16(tbase + 256 + c * 16, pk);
          if (MODE == 6) { tmem_st8(tbase + 256 + c * 16, pk); tmem_st8(tbase + 256 + c * 16 + 8, pk + 8); }
          if (MODE == 8) {
            uint4* row = reinterpret_cast<uint4*>(dyn + threadIdx.x * 256 + c * 64);
#pragma unroll
            for (int q = 0; q < 4; ++q) row[q ^ (lane & 3)] = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
          }
        }
        if (MODE == 5) {
#pragma unroll
          for (int t = 0; t < 64; ++t) sink ^= pk_all[t];
        }
        if (MODE == 7) {
          tmem_st32(tbase + 256, reinterpret_cast<const uint32_t(&)[32]>(pk_all[0]));
          tmem_st32(tbase + 288, reinterpret_cast<const uint32_t(&)[32]>(pk_all[32]));
        }
        float a0, a1, a2, a3;
        unpack_f32x2(acc_a, a0, a1); unpack_f32x2(acc_b, a2, a3);
        chain += (a0 + a1) + (a2 + a3) + fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        if (MODE == 4 || MODE == 6 || MODE == 7) tmem_wait_st();
        if (MODE == 8) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      }
    }
    if (MODE == 0 || MODE == 2) tmem_wait_st();
    t1 = clock64();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink ^ __float_as_uint(chain);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_slot);
}


// spread of the polynomial pairs inside a group of 8 pairs (so they interleave with the MUFU pairs)
__device__ constexpr int kPolyOrder[8] = {0, 4, 2, 6, 1, 5, 3, 7};

// Kernel-like environment around the M4 pass: 384 threads, setmaxnreg 72/216, warps 4-7 (and 8-11 when
// BOTH) run the pass on their own TMEM quadrant/columns; POLL: lane 0 of warps 0 and 1 sits in a
// try_wait (suspend hint) on an mbarrier that completes only at the end, like the TMA/MMA threads.
template <int BOTH, int POLL, int POLY8 = 0>
__global__ void __launch_bounds__(384, 1) kenv(const float* in, uint32_t* out, int iters, long long* cycles) {
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t done_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (threadIdx.x == 32) { mbar_init(&done_bar, BOTH ? 8 : 4); fence_barrier_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;\n");
    if (POLL && warp < 2 && lane == 0) mbar_wait(&done_bar, 0);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;\n");
    const int wg = (warp - 4) >> 2;
    if (wg == 0 || BOTH) {
      const uint32_t tbase = tmem_slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + wg * 128;
      {
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int t = 0; t < 32; ++t) r[t] = __float_as_uint(in[(threadIdx.x * 131 + (c * 32 + t) * 7) & 1023]);
          tmem_st32(tbase + c * 32, r);
        }
        tmem_wait_st();
      }
      float chain = in[lane];
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        float s[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(tbase + c * 32, r);
#pragma unroll
          for (int t = 0; t < 32; ++t) s[c * 32 + t] = __uint_as_float(r[t]);
        }
        tmem_wait_ld();
        const float scale = 0.127f, neg_m = -0.73f + 1e-6f * it;
        const uint64_t scale2 = pack_f32x2(scale, scale), negm2 = pack_f32x2(neg_m, neg_m);
        uint64_t acc_a = 0ull, acc_b = 0ull;
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int t = 0; t < 32; t += 4) {
            const float s0 = s[c * 32 + t], s1 = s[c * 32 + t + 1], s2 = s[c * 32 + t + 2], s3 = s[c * 32 + t + 3];
            m0 = fmaxf(m0, s0); m1 = fmaxf(m1, s1); m2 = fmaxf(m2, s2); m3 = fmaxf(m3, s3);
            float x0, x1, x2, x3;
            unpack_f32x2(ffma2(pack_f32x2(s0, s1), scale2, negm2), x0, x1);
            unpack_f32x2(ffma2(pack_f32x2(s2, s3), scale2, negm2), x2, x3);
            // pair index within a group of 8 pairs (16 elements): pairs < POLY8 go to the FMA pipe
            const int pa = (t / 2) & 7, pb = (t / 2 + 1) & 7;
            float e0, e1, e2, e3;
            if (kPolyOrder[pa] < POLY8) unpack_f32x2(poly_exp2_x2(pack_f32x2(fmaxf(x0, -120.f), fmaxf(x1, -120.f))), e0, e1);
            else { e0 = fast_exp2(x0); e1 = fast_exp2(x1); }
            if (kPolyOrder[pb] < POLY8) unpack_f32x2(poly_exp2_x2(pack_f32x2(fmaxf(x2, -120.f), fmaxf(x3, -120.f))), e2, e3);
            else { e2 = fast_exp2(x2); e3 = fast_exp2(x3); }
            acc_a = fadd2(acc_a, pack_f32x2(e0, e1));
            acc_b = fadd2(acc_b, pack_f32x2(e2, e3));
            pk[t / 2] = pack_bf16x2(e0, e1);
            pk[t / 2 + 1] = pack_bf16x2(e2, e3);
          }
          tmem_st16(tbase + 256 + c * 16, pk);
        }
        float a0, a1, a2, a3;
        unpack_f32x2(acc_a, a0, a1); unpack_f32x2(acc_b, a2, a3);
        chain += (a0 + a1) + (a2 + a3) + fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        tmem_wait_st();
      }
      const long long t1 = clock64();
      out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(chain);
      if (threadIdx.x == 128) cycles[blockIdx.x] = t1 - t0;
      __syncwarp();
      if (lane == 0) mbar_arrive(&done_bar);
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_slot);
}

template <int BOTH, int POLL, int POLY8 = 0>
void run_env(const char* name, const float* in, uint32_t* out, long long* dc) {
  const int iters = 200;
  for (int rep = 0; rep < 2; ++rep) kenv<BOTH, POLL, POLY8><<<148, 384>>>(in, out, iters, dc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s error %s\n", name, cudaGetErrorString(e)); return; }
  long long hc[148]; cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += hc[i]; avg /= 148;
  printf("%-64s          %7.0f cycles per iteration\n", name, avg / iters);
}

template <int MODE>
void run(const char* name, const float* in, uint32_t* out, long long* dc, int active) {
  const int iters = 200;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int rep = 0; rep < 2; ++rep) k<MODE><<<148, 128, 65536>>>(in, out, iters, dc, active);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s error %s\n", name, cudaGetErrorString(e)); return; }
  long long hc[148]; cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += hc[i]; avg /= 148;
  printf("%-64s warps=%d  %7.0f cycles per iteration\n", name, active, avg / iters);
}

int main() {
  float* in; uint32_t* out; long long* dc;
  cudaMalloc(&in, 4096);
  { float h[1024]; unsigned x = 12345u; for (int i = 0; i < 1024; ++i) { x = x * 1664525u + 1013904223u; h[i] = -80.0f * (float)(x >> 8) / 16777216.0f; }
    cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice); }
  cudaMalloc(&out, 148 * 384 * 4); cudaMalloc(&dc, 148 * 8);
  run_env<0, 0>("ENV 384 thr, setmaxnreg, one WG, no pollers (M4 pass)", in, out, dc);
  run_env<0, 1>("ENV 384 thr, setmaxnreg, one WG, 2 pollers", in, out, dc);
  run_env<1, 0>("ENV 384 thr, setmaxnreg, both WGs, no pollers", in, out, dc);
  run_env<1, 1>("ENV 384 thr, setmaxnreg, both WGs, 2 pollers", in, out, dc);
  run_env<0, 1, 1>("ENV one WG,   12.5% of pairs by polynomial", in, out, dc);
  run_env<0, 1, 2>("ENV one WG,   25%   of pairs by polynomial", in, out, dc);
  run_env<0, 1, 3>("ENV one WG,   37.5% of pairs by polynomial", in, out, dc);
  run_env<0, 1, 4>("ENV one WG,   50%   of pairs by polynomial", in, out, dc);
  run_env<1, 1, 1>("ENV both WGs, 12.5% of pairs by polynomial", in, out, dc);
  run_env<1, 1, 2>("ENV both WGs, 25%   of pairs by polynomial", in, out, dc);
  run_env<1, 1, 3>("ENV both WGs, 37.5% of pairs by polynomial", in, out, dc);
  run_env<1, 1, 4>("ENV both WGs, 50%   of pairs by polynomial", in, out, dc);
  run_env<1, 1, 5>("ENV both WGs, 62.5% of pairs by polynomial", in, out, dc);
  for (int active = 4; active >= 1; active -= 3) {
    run<0>("M0 4 x tcgen05.st.32x32b.x16 back to back", in, out, dc, active);
    run<1>("M1 4 x (tcgen05.ld.32x32b.x32 + wait)", in, out, dc, active);
    run<3>("M3 4 x 64 dependent FFMA", in, out, dc, active);
    run<2>("M2 4 x (st.x16 + 64 dependent FFMA)", in, out, dc, active);
    run<5>("M5 softmax pass, P not stored", in, out, dc, active);
    run<4>("M4 softmax pass, P -> TMEM st.x16 in loop (kernel form)", in, out, dc, active);
    run<6>("M6 softmax pass, P -> TMEM 8 x st.x8", in, out, dc, active);
    run<7>("M7 softmax pass, P -> TMEM 2 x st.x32 after loop", in, out, dc, active);
    run<8>("M8 softmax pass, P -> shared memory", in, out, dc, active);
  }
  return 0;
}
