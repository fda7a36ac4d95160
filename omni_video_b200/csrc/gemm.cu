// ovb_gemm_bf16: persistent, warp-specialised bf16 GEMM for sm_100a.
//
//   out = epilogue(A[M,K] * W[N,K]^T + bias)
//
// One CTA per SM (persistent over 128x256 output tiles, N fastest so that the 148 resident CTAs
// share a handful of A row-panels and the whole weight matrix stays in the 126 MB L2).
//   warp 0      TMA producer: 64-wide K slices of A (128 rows) and W (256 rows), 128B-swizzled,
//               into a 4-stage shared-memory ring (48 KB / stage)
//   warp 1      single elected thread issues tcgen05.mma (M=128, N=256, K=16) x4 per stage;
//               tcgen05.commit releases the stage and, on the last K slice, publishes the tile
//   warp 2      TMEM allocator (all 512 columns = two 128x256 fp32 accumulators, so the epilogue
//               of tile i overlaps the main loop of tile i+1)
//   warps 4-7   epilogue: tcgen05.ld 32 columns at a time (one accumulator row per thread),
//               bias / GELU / gate / fp32-residual math in registers, 16-byte global stores
//
// Algorithmic bytes per tile are tiny next to the FLOPs (arithmetic intensity >= 700 FLOP/B for
// the smallest DiT shape), so the bounding roofline is the tensor pipe; see DESIGN.md §4.
#include "host_util.h"
#include "ptx.cuh"

namespace ovb {

constexpr int kBM = 128;
constexpr int kBN = 256;
constexpr int kBK = 64;
constexpr int kStages = 4;
constexpr int kABytes = kBM * kBK * 2;  // 16 KB
constexpr int kBBytes = kBN * kBK * 2;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kGemmThreads = 256;
constexpr int kGemmSmem = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmParams {
  int M, N, K;
  int k_per_split;  // K / k_split
  const __nv_bfloat16* bias;
  void* out;
  long long ldo;
  int epilogue;
  const void* resid;
  long long ldr;
  const float* gate;
  long long gate_row_stride, gate_sample_stride;
  int rows_per_sample;
  int n_per_split;  // N / n_split (== N when off)
  long long out_split_stride;
  int num_m_tiles, num_n_tiles;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;  // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int kb_per_split = (p.k_per_split + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // elect.sync (not `lane == 0`): the compiler then knows exactly one lane runs the loop and
    // feeds UTMALDG / UTCHMMA from uniform registers without a per-instruction waterfall loop
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.num_n_tiles;
        const int m_tile = tile / p.num_n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(&full_bar[stage], kStageBytes);
          const int split = kb / kb_per_split;
          const int kcol = (kb - split * kb_per_split) * kBK;
          tma_load_3d(sa, &tmA, &full_bar[stage], kcol, m_tile * kBM, split);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * kBK, n_tile * kBN);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(kBM, kBN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = umma_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = umma_desc_sw128(sb + k * 32, 16, 1024);
            umma_ss(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.num_n_tiles;
      const int m_tile = tile / p.num_n_tiles;
      const int row = m_tile * kBM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const float* gate_row = nullptr;
      if (p.epilogue == OVB_EPI_GATE_RESID_F32 && row_ok) {
        const int sample = row / p.rows_per_sample;
        const int r_in = row - sample * p.rows_per_sample;
        gate_row = p.gate + sample * p.gate_sample_stride + r_in * p.gate_row_stride;
      }
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        const int col0 = n_tile * kBN + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kBN + c * 32, r);
        tmem_wait_ld();
        if (row_ok) {
          const int seg = col0 / p.n_per_split;
          const int col_in = col0 - seg * p.n_per_split;
          const long long obase = seg * p.out_split_stride + static_cast<long long>(row) * p.ldo + col_in;
#pragma unroll
          for (int g = 0; g < 4; ++g) {  // 8 columns per group
            const int col = col0 + g * 8;
            if (col >= p.N) break;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
            if (p.bias != nullptr) {
              const uint4 bw = *reinterpret_cast<const uint4*>(p.bias + col);
              const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&bw);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(b2[j]);
                v[2 * j] += f.x;
                v[2 * j + 1] += f.y;
              }
            }
            if (p.epilogue == OVB_EPI_BF16 || p.epilogue == OVB_EPI_GELU_BF16 ||
                p.epilogue == OVB_EPI_RESID_BF16) {
              if (p.epilogue == OVB_EPI_GELU_BF16) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = gelu_tanh(bf16_round(v[j]));
              } else if (p.epilogue == OVB_EPI_RESID_BF16) {
                const uint4 rw = *reinterpret_cast<const uint4*>(
                    reinterpret_cast<const __nv_bfloat16*>(p.resid) + static_cast<long long>(row) * p.ldr + col);
                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rw);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __bfloat1622float2(r2[j]);
                  v[2 * j] = bf16_round(v[2 * j]) + f.x;
                  v[2 * j + 1] = bf16_round(v[2 * j + 1]) + f.y;
                }
              }
              uint4 o;
              o.x = pack_bf16x2(v[0], v[1]);
              o.y = pack_bf16x2(v[2], v[3]);
              o.z = pack_bf16x2(v[4], v[5]);
              o.w = pack_bf16x2(v[6], v[7]);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + obase + g * 8) = o;
            } else {
              float* optr = reinterpret_cast<float*>(p.out) + obase + g * 8;
              if (p.epilogue != OVB_EPI_F32) {
                const float* rptr = reinterpret_cast<const float*>(p.resid) + static_cast<long long>(row) * p.ldr + col;
                const float4 x0 = *reinterpret_cast<const float4*>(rptr);
                const float4 x1 = *reinterpret_cast<const float4*>(rptr + 4);
                float gt[8];
                if (p.epilogue == OVB_EPI_GATE_RESID_F32) {
                  const float4 g0 = *reinterpret_cast<const float4*>(gate_row + col);
                  const float4 g1 = *reinterpret_cast<const float4*>(gate_row + col + 4);
                  gt[0] = g0.x; gt[1] = g0.y; gt[2] = g0.z; gt[3] = g0.w;
                  gt[4] = g1.x; gt[5] = g1.y; gt[6] = g1.z; gt[7] = g1.w;
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) gt[j] = 1.0f;
                }
                const float xr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = xr[j] + bf16_round(v[j]) * gt[j];
              }
              *reinterpret_cast<float4*>(optr) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(optr + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace ovb

extern "C" int ovb_gemm_bf16(const ovb_gemm_args* a, void* stream) {
  using namespace ovb;
  OVB_CHECK_ARG(a != nullptr, "ovb_gemm_bf16: null args");
  OVB_CHECK_ARG(a->A && a->W && a->out, "ovb_gemm_bf16: null A/W/out");
  OVB_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "ovb_gemm_bf16: bad shape M=%d N=%d K=%d", a->M,
                a->N, a->K);
  OVB_CHECK_ARG(a->N % 8 == 0 && a->K % 8 == 0, "ovb_gemm_bf16: N and K must be multiples of 8");
  OVB_CHECK_ARG(a->lda % 8 == 0 && a->ldw % 8 == 0 && a->ldo % 8 == 0,
                "ovb_gemm_bf16: leading dims must be multiples of 8 elements");
  OVB_CHECK_ARG(a->epilogue >= OVB_EPI_BF16 && a->epilogue <= OVB_EPI_RESID_BF16,
                "ovb_gemm_bf16: unknown epilogue %d", a->epilogue);
  const bool needs_resid =
      a->epilogue == OVB_EPI_RESID_F32 || a->epilogue == OVB_EPI_GATE_RESID_F32;
  OVB_CHECK_ARG(!needs_resid || (a->resid != nullptr && a->ldr % 4 == 0),
                "ovb_gemm_bf16: residual epilogue needs resid with ldr %% 4 == 0");
  OVB_CHECK_ARG(a->epilogue != OVB_EPI_RESID_BF16 ||
                    (a->resid != nullptr && a->ldr % 8 == 0 && a->n_split <= 1 &&
                     reinterpret_cast<uintptr_t>(a->resid) % 16 == 0),
                "ovb_gemm_bf16: bf16 residual epilogue needs a 16-byte aligned resid, ldr %% 8 == 0, no n_split");
  OVB_CHECK_ARG(a->epilogue != OVB_EPI_GATE_RESID_F32 ||
                    (a->gate != nullptr && a->rows_per_sample > 0 && a->gate_row_stride % 4 == 0 &&
                     a->gate_sample_stride % 4 == 0),
                "ovb_gemm_bf16: gate epilogue needs gate, rows_per_sample > 0, strides %% 4 == 0");
  const int n_split = a->n_split > 1 ? a->n_split : 1;
  const int k_split = a->k_split > 1 ? a->k_split : 1;
  OVB_CHECK_ARG(a->N % n_split == 0 && (n_split == 1 || (a->N / n_split) % 32 == 0),
                "ovb_gemm_bf16: N/n_split must be a multiple of 32");
  OVB_CHECK_ARG(a->K % k_split == 0 && (k_split == 1 || (a->K / k_split) % kBK == 0),
                "ovb_gemm_bf16: K/k_split must be a multiple of 64");
  OVB_CHECK_ARG((reinterpret_cast<uintptr_t>(a->A) | reinterpret_cast<uintptr_t>(a->W) |
                 reinterpret_cast<uintptr_t>(a->out)) % 16 == 0,
                "ovb_gemm_bf16: pointers must be 16-byte aligned");

  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(a->K / k_split), static_cast<uint64_t>(a->M),
                              static_cast<uint64_t>(k_split)};
    const uint64_t strides[2] = {static_cast<uint64_t>(a->lda) * 2,
                                 k_split > 1 ? static_cast<uint64_t>(a->a_split_stride) * 2
                                             : static_cast<uint64_t>(a->lda) * 2 * a->M};
    const uint32_t box[3] = {kBK, kBM, 1};
    int rc = encode_tmap_bf16(&tmA, a->A, 3, dims, strides, box);
    if (rc != OVB_OK) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->ldw) * 2};
    const uint32_t box[2] = {kBK, kBN};
    int rc = encode_tmap_bf16(&tmB, a->W, 2, dims, strides, box);
    if (rc != OVB_OK) return rc;
  }

  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.k_per_split = a->K / k_split;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(a->bias);
  p.out = a->out;
  p.ldo = a->ldo;
  p.epilogue = a->epilogue;
  p.resid = a->resid;
  p.ldr = a->ldr;
  p.gate = reinterpret_cast<const float*>(a->gate);
  p.gate_row_stride = a->gate_row_stride;
  p.gate_sample_stride = a->gate_sample_stride;
  p.rows_per_sample = a->rows_per_sample > 0 ? a->rows_per_sample : a->M;
  p.n_per_split = a->N / n_split;
  p.out_split_stride = n_split > 1 ? a->out_split_stride : 0;
  p.num_m_tiles = (a->M + kBM - 1) / kBM;
  p.num_n_tiles = (a->N + kBN - 1) / kBN;

  static bool attr_set = false;
  if (!attr_set) {
    OVB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    attr_set = true;
  }
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int sms = device_sm_count();
  const int grid = num_tiles < sms ? num_tiles : sms;
  gemm_bf16_kernel<<<grid, kGemmThreads, kGemmSmem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}
