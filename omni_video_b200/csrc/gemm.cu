// ovb_gemm_bf16: persistent, warp-specialised bf16 GEMM for sm_100a on CTA PAIRS.
//
//   out = epilogue(A[M,K] * W[N,K]^T + bias)
//
// Two CTAs on the two SMs of a TPC form a cluster and compute one 256x256 output tile with
// tcgen05.mma.cta_group::2 (M = 256): CTA r of the pair owns rows [128r, 128r+128) of the tile (its
// 128 TMEM lanes) and stages only ITS half of the operands — 128 rows of A and 128 of the 256 rows
// of W — so each SM reads 8 KB of shared memory per MMA instead of 12 KB.  Measured on B200
// (scripts/microbench/mma_rate.cu): a 128x256x16 MMA fed from one SM's shared memory takes ~163
// cycles against the 128-cycle tensor floor because the B operand streams at ~70 B/cycle; halving
// B per SM removes that bound.
//   warp 0      TMA producer (each CTA loads its own halves; the transaction bytes of both CTAs
//               are signalled on the LEADER's full barrier), 6-stage ring, 32 KB / stage / CTA
//   warp 1      leader CTA only: one elected thread issues the MMAs; tcgen05.commit multicasts the
//               "stage free" / "accumulator ready" arrivals to both CTAs
//   warp 2      TMEM allocator (cta_group::2, all 512 columns = two 256-column accumulators, so the
//               epilogue of tile i overlaps the main loop of tile i+1)
//   warps 4-7   epilogue: tcgen05.ld 32 columns -> shared-memory transpose (XOR-swizzled 16-byte
//               chunks, conflict free) -> bias / GELU / gate / residual math on ROW-CONTIGUOUS
//               16-byte vectors, so every global access of a warp covers whole 64/128-byte row
//               segments instead of 32 scattered rows.
//
// Algorithmic bytes per tile are tiny next to the FLOPs (arithmetic intensity >= 700 FLOP/B for
// the smallest DiT shape), so the bounding roofline is the tensor pipe; see DESIGN.md §3.
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace ovb {

constexpr int kBM = 128;          // rows per CTA (256 per pair)
constexpr int kBN = 256;          // tile width; each CTA stages kBN/2 rows of W
constexpr int kBK = 64;
constexpr int kStages = 6;
constexpr int kABytes = kBM * kBK * 2;        // 16 KB
constexpr int kBBytes = (kBN / 2) * kBK * 2;  // 16 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kGemmThreads = 256;   // gemm_resid_tma_kernel: 4 service warps + 4 epilogue warps
// gemm_bf16_kernel runs EIGHT epilogue warps, two per scheduler and TMEM lane quadrant (warps w and
// w + 4 may both access lanes 32 (w % 4) .. + 31), each taking every other 32-column chunk.  One warp
// per scheduler executes the epilogue at 5-7 cycles per instruction (dependent chains, MUFU latency,
// nobody to switch to): with the GELU — or at K = 1536 any — epilogue that outlasted the main loop
// (FFN1 [65520, 8960, 1536]: 1.53 ms with GELU against 1.14 ms without).
constexpr int kEpiWarps = 8;
constexpr int kGemmThreadsEpi8 = 128 + 32 * kEpiWarps;
constexpr int kStagingBytes = kEpiWarps * 32 * 32 * 4;  // one 32x32 fp32 transpose buffer per epilogue warp
constexpr int kGemmSmem = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;

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
  int num_m_tiles, num_n_tiles;  // 256 x 256 pair tiles
};

template <int kEpi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreadsEpi8, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  float* staging = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kStagingBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;  // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained (leader's copy is used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int kb_per_split = (p.k_per_split + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's producer: arrive.expect_tx for both CTAs' bytes
      mbar_init(&empty_bar[s], 1);  // one multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * kEpiWarps);  // epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // peer barriers initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    // elect.sync (not `lane == 0`): the compiler then knows exactly one lane runs the loop and
    // feeds UTMALDG / UTCHMMA from uniform registers without a per-instruction waterfall loop
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int n_tile = tile % p.num_n_tiles;
        const int m_tile = tile / p.num_n_tiles;
        const int row0 = m_tile * 2 * kBM + rank * kBM;
        const int wrow0 = n_tile * kBN + rank * (kBN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes);
          const int split = kb / kb_per_split;
          const int kcol = (kb - split * kb_per_split) * kBK;
          tma_load_3d_pair(sa, &tmA, leader_full, kcol, row0, split);
          tma_load_2d_pair(sb, &tmB, leader_full, kb * kBK, wrow0);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(2 * kBM, kBN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
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
            umma_ss_pair(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit_pair(&empty_bar[stage], 3);
          if (kb == num_kb - 1) umma_commit_pair(&tfull_bar[acc], 3);
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
    // ------------------------------------------------------------ epilogue (both CTAs)
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int chunk0 = (warp - 4) >> 2;  // this warp's first 32-column chunk; it takes every other one
    float* stg = staging + (warp - 4) * (32 * 32);
    const uint32_t stg_addr = smem_u32(stg);
    const uint32_t leader_tempty0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    const int rr_off = lane >> 3;  // row inside a group of 4 in the read phase
    const int cj = lane & 7;       // 16-byte chunk (4 fp32 columns) in the read phase
    constexpr bool f32_resid = kEpi == OVB_EPI_RESID_F32 || kEpi == OVB_EPI_GATE_RESID_F32;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int n_tile = tile % p.num_n_tiles;
      const int m_tile = tile / p.num_n_tiles;
      const int row_base = m_tile * 2 * kBM + rank * kBM + quad * 32;
      // bias of this lane's 4 columns, one chunk ahead: the first fetch is issued before the
      // accumulator is ready so its latency hides behind the main loop
      uint2 bw_next = make_uint2(0u, 0u);
      {
        const int col = n_tile * kBN + chunk0 * 32 + cj * 4;
        if (p.bias != nullptr && col < p.N) bw_next = *reinterpret_cast<const uint2*>(p.bias + col);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = chunk0; c < kBN / 32; c += 2) {
        const int col0 = n_tile * kBN + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kBN + c * 32, r);
        const int col = col0 + cj * 4;
        const bool col_ok = col < p.N;
        const uint2 bw = bw_next;
        if (p.bias != nullptr && col + 64 < p.N && c + 2 < kBN / 32)
          bw_next = *reinterpret_cast<const uint2*>(p.bias + col + 64);
        // operands of the residual epilogues: all 8 row groups are requested up front, before
        // any store of this chunk is issued (a store may alias them as far as the compiler knows)
        float4 xr[8], gt[8];
        uint2 xb[8];
        if constexpr (f32_resid) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = row_base + 4 * i + rr_off;
            xr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            gt[i] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (row < p.M && col_ok) {
              xr[i] = *reinterpret_cast<const float4*>(
                  reinterpret_cast<const float*>(p.resid) + static_cast<long long>(row) * p.ldr + col);
              if constexpr (kEpi == OVB_EPI_GATE_RESID_F32) {
                const int sample = row / p.rows_per_sample;
                const int r_in = row - sample * p.rows_per_sample;
                gt[i] = *reinterpret_cast<const float4*>(p.gate + sample * p.gate_sample_stride +
                                                         r_in * p.gate_row_stride + col);
              }
            }
          }
        } else if constexpr (kEpi == OVB_EPI_RESID_BF16) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = row_base + 4 * i + rr_off;
            xb[i] = make_uint2(0u, 0u);
            if (row < p.M && col_ok)
              xb[i] = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.resid) +
                                                      static_cast<long long>(row) * p.ldr + col);
          }
        }
        tmem_wait_ld();
        // transpose through shared memory: thread = row writes 8 chunks, chunk j at j ^ (row & 7)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t a = stg_addr + lane * 128 + ((j ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "r"(r[4 * j]),
                       "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                       : "memory");
        }
        __syncwarp();
        float b4[4];
        {
          const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw.x));
          const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw.y));
          b4[0] = f0.x; b4[1] = f0.y; b4[2] = f1.x; b4[3] = f1.y;
        }
        const int seg = col0 / p.n_per_split;
        const long long ocol = seg * p.out_split_stride + (col - seg * p.n_per_split);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + rr_off;
          const int row = row_base + rr;
          float v[4];
          {
            const uint32_t a = stg_addr + rr * 128 + ((cj ^ (rr & 7)) << 4);
            uint32_t u0, u1, u2, u3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n"
                         : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3)
                         : "r"(a)
                         : "memory");
            v[0] = __uint_as_float(u0) + b4[0];
            v[1] = __uint_as_float(u1) + b4[1];
            v[2] = __uint_as_float(u2) + b4[2];
            v[3] = __uint_as_float(u3) + b4[3];
          }
          if (row < p.M && col_ok) {
            const long long o = static_cast<long long>(row) * p.ldo + ocol;
            if constexpr (f32_resid) {
              v[0] = xr[i].x + bf16_round(v[0]) * gt[i].x;
              v[1] = xr[i].y + bf16_round(v[1]) * gt[i].y;
              v[2] = xr[i].z + bf16_round(v[2]) * gt[i].z;
              v[3] = xr[i].w + bf16_round(v[3]) * gt[i].w;
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o) =
                  make_float4(v[0], v[1], v[2], v[3]);
            } else if constexpr (kEpi == OVB_EPI_F32) {
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o) =
                  make_float4(v[0], v[1], v[2], v[3]);
            } else {
              if constexpr (kEpi == OVB_EPI_GELU_BF16) {
                // the Linear's bf16 result (model.py:211 under autocast), then GELU in fp32, two values
                // per instruction: at K = 1536 this epilogue used to outlast the main loop (FFN1 ran at
                // 1176 TF/s against 1561 with the plain epilogue)
                uint32_t p01 = pack_bf16x2(v[0], v[1]), p23 = pack_bf16x2(v[2], v[3]);
                const uint64_t g01 = gelu_tanh_x2(pack_f32x2(__uint_as_float(p01 << 16), __uint_as_float(p01 & 0xffff0000u)));
                const uint64_t g23 = gelu_tanh_x2(pack_f32x2(__uint_as_float(p23 << 16), __uint_as_float(p23 & 0xffff0000u)));
                unpack_f32x2(g01, v[0], v[1]);
                unpack_f32x2(g23, v[2], v[3]);
              } else if constexpr (kEpi == OVB_EPI_RESID_BF16) {
                const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xb[i].x));
                const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xb[i].y));
                v[0] = bf16_round(v[0]) + f0.x; v[1] = bf16_round(v[1]) + f0.y;
                v[2] = bf16_round(v[2]) + f1.x; v[3] = bf16_round(v[3]) + f1.y;
              }
              uint2 ov;
              ov.x = pack_bf16x2(v[0], v[1]);
              ov.y = pack_bf16x2(v[2], v[3]);
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = ov;
            }
          }
        }
        __syncwarp();  // staging buffer is rewritten by the next chunk
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tempty0 + acc * 8);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // both CTAs are done with TMEM and with each other's barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------
// fp32-residual epilogues with TMA on both sides of the residual stream:
//     out(fp32) = resid(fp32) + bf16(acc + bias) * gate            (reference model.py:246, :250, :254)
// The residual stream costs 8 bytes per output element (403 MB per [32760,1536] GEMM), all of it
// latency-bound when four epilogue warps fetch it with ordinary loads (16 KB in flight per SM).
// Here one epilogue thread runs a 4-deep ring of [128 rows x 32 columns] fp32 boxes: TMA loads the
// residual box two chunks ahead, each thread reads ITS row from the box (TMA's 128-byte swizzle is
// conflict free for row-per-thread 16-byte accesses), writes the result back in place, and the box
// goes to `out` with one TMA store.  Main loop as above with a 4-stage operand ring.
// 5 operand stages + 4 residual boxes fill the 227 KB (4 stages: FFN2 [65520,1536,8960] 1.30 ms, 5: 1.17 ms =
// the plain bf16 epilogue's time; O-projection 0.271 -> 0.252 ms).  Pulling the next tile's residual columns into
// L2 ahead of time (cp.async.bulk.prefetch.tensor) was measured too and made both shapes 2-8 % slower.
constexpr int kRStages = 5;                       // operand ring
constexpr int kRing = 4;                          // residual boxes
constexpr int kRBoxBytes = kBM * 32 * 4;          // 16 KB
constexpr int kGemmResidSmem = kRStages * kStageBytes + kRing * kRBoxBytes + 1024 + 256;

template <bool kGate>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_resid_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                      const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* ring = smem + kRStages * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + kRing * kRBoxBytes);
  uint64_t* empty_bar = full_bar + kRStages;
  uint64_t* tfull_bar = empty_bar + kRStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rfull_bar = tempty_bar + 2;  // [kRing] residual box landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull_bar + kRing);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int kb_per_split = (p.k_per_split + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kRStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 8);
    }
    for (int b = 0; b < kRing; ++b) mbar_init(&rfull_bar[b], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int n_tile = tile % p.num_n_tiles;
        const int m_tile = tile / p.num_n_tiles;
        const int row0 = m_tile * 2 * kBM + rank * kBM;
        const int wrow0 = n_tile * kBN + rank * (kBN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes);
          const int split = kb / kb_per_split;
          const int kcol = (kb - split * kb_per_split) * kBK;
          tma_load_3d_pair(sa, &tmA, leader_full, kcol, row0, split);
          tma_load_2d_pair(sb, &tmB, leader_full, kb * kBK, wrow0);
          if (++stage == kRStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(2 * kBM, kBN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
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
            umma_ss_pair(d_tmem, umma_desc_sw128(sa + k * 32, 16, 1024),
                         umma_desc_sw128(sb + k * 32, 16, 1024), idesc, (kb | k) != 0);
          }
          umma_commit_pair(&empty_bar[stage], 3);
          if (kb == num_kb - 1) umma_commit_pair(&tfull_bar[acc], 3);
          if (++stage == kRStages) {
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
    // ------------------------------------------------------------ epilogue (both CTAs)
    const int quad = warp & 3;
    const int r_in_cta = quad * 32 + lane;             // this thread's row of the CTA's 128
    const bool leader_thread = warp == 4 && lane == 0;  // runs the residual ring
    const uint32_t leader_tempty0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    constexpr int kChunks = kBN / 32;
    constexpr int kAhead = 2;  // residual boxes requested ahead of the one being processed
    // number of 32-column chunks this CTA will process, in order (tiles x chunks that start < N)
    const int chunks_per_tile_full = kChunks;
    auto tile_of = [&](int it) { return pair + it * num_pairs; };
    auto chunk_coords = [&](long long g, int& row0, int& col0) -> bool {
      const int it = static_cast<int>(g / chunks_per_tile_full);
      const int c = static_cast<int>(g - static_cast<long long>(it) * chunks_per_tile_full);
      const int tile = tile_of(it);
      if (tile >= num_tiles) return false;
      const int n_tile = tile % p.num_n_tiles;
      const int m_tile = tile / p.num_n_tiles;
      row0 = m_tile * 2 * kBM + rank * kBM;
      col0 = n_tile * kBN + c * 32;
      return true;
    };
    auto request = [&](long long g) {  // leader thread: TMA-load the residual box of chunk g
      int row0, col0;
      if (!chunk_coords(g, row0, col0)) return;
      const int b = static_cast<int>(g % kRing);
      mbar_expect_tx(&rfull_bar[b], kRBoxBytes);
      tma_load_2d(ring + b * kRBoxBytes, &tmR, &rfull_bar[b], col0, row0);
    };
    if (leader_thread) {
      for (int g = 0; g < kAhead; ++g) request(g);
    }
    long long g = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int n_tile = tile % p.num_n_tiles;
      const int m_tile = tile / p.num_n_tiles;
      const int row = m_tile * 2 * kBM + rank * kBM + r_in_cta;
      const float* gate_row = nullptr;
      if (kGate) {
        const int rclamp = row < p.M ? row : p.M - 1;
        const int sample = rclamp / p.rows_per_sample;
        const int r_in = rclamp - sample * p.rows_per_sample;
        gate_row = p.gate + sample * p.gate_sample_stride + r_in * p.gate_row_stride;
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c, ++g) {
        const int col0 = n_tile * kBN + c * 32;
        const int b = static_cast<int>(g % kRing);
        uint8_t* box = ring + b * kRBoxBytes;
        if (leader_thread) {
          // box (g + kAhead) % kRing was last used by chunk g + kAhead - kRing = g - 2: its store
          // must have finished READING shared memory; stores are committed one group per chunk
          bulk_wait_group_read<kRing - kAhead - 1>();
          request(g + kAhead);
        }
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kBN + c * 32, v);
        // bias and gate for these 32 columns (same for every row of a sample: L1 broadcast), kept as fp32
        // pairs: the update below is packed (FADD2, one cvt per pair, FFMA2) — this warp is alone on its
        // scheduler and the instruction count per chunk decides whether the epilogue hides behind a
        // K = 1536 main loop
        uint64_t bs2[16], gt2[16];
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (p.bias != nullptr && col0 + j < p.N) {
            const uint4 bw = *reinterpret_cast<const uint4*>(p.bias + col0 + j);
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&bw);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = __bfloat1622float2(b2[q]);
              bs2[j / 2 + q] = pack_f32x2(f.x, f.y);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) bs2[j / 2 + q] = 0;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 gq = make_float4(1.f, 1.f, 1.f, 1.f);
          if (kGate && col0 + j < p.N) gq = *reinterpret_cast<const float4*>(gate_row + col0 + j);
          gt2[j / 2] = pack_f32x2(gq.x, gq.y);
          gt2[j / 2 + 1] = pack_f32x2(gq.z, gq.w);
        }
        mbar_wait(&rfull_bar[b], static_cast<uint32_t>((g / kRing) & 1));
        tmem_wait_ld();
        const uint32_t row_addr = smem_u32(box) + r_in_cta * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t a = row_addr + ((j ^ (r_in_cta & 7)) << 4);
          uint64_t x01, x23;
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];\n" : "=l"(x01), "=l"(x23) : "r"(a) : "memory");
          // x += bf16(acc + bias) * gate, two columns per instruction
          float s0, s1, s2, s3;
          unpack_f32x2(fadd2(pack_f32x2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), bs2[2 * j]), s0, s1);
          unpack_f32x2(fadd2(pack_f32x2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), bs2[2 * j + 1]), s2, s3);
          const uint32_t c01 = pack_bf16x2(s0, s1), c23 = pack_bf16x2(s2, s3);
          x01 = ffma2(pack_f32x2(__uint_as_float(c01 << 16), __uint_as_float(c01 & 0xffff0000u)), gt2[2 * j], x01);
          x23 = ffma2(pack_f32x2(__uint_as_float(c23 << 16), __uint_as_float(c23 & 0xffff0000u)), gt2[2 * j + 1], x23);
          asm volatile("st.shared.v2.b64 [%0], {%1, %2};\n" ::"r"(a), "l"(x01), "l"(x23) : "memory");
        }
        fence_proxy_async_smem();      // generic-proxy writes visible to the TMA store
        named_bar_sync(1, 128);        // all four epilogue warps have written their rows
        if (leader_thread) {
          int row0, cc0;
          chunk_coords(g, row0, cc0);
          if (cc0 < p.N) tma_store_2d(&tmO, box, cc0, row0);
          bulk_commit_group();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tempty0 + acc * 8);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (leader_thread) bulk_wait_group<0>();  // all stores complete before the CTA exits
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
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
    const uint32_t box[2] = {kBK, kBN / 2};
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
  p.num_m_tiles = (a->M + 2 * kBM - 1) / (2 * kBM);
  p.num_n_tiles = (a->N + kBN - 1) / kBN;

  static bool attr_set[64] = {};
  if (once_per_device(attr_set)) {
    OVB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<OVB_EPI_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    OVB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<OVB_EPI_GELU_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    OVB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<OVB_EPI_RESID_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    OVB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<OVB_EPI_GATE_RESID_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    OVB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<OVB_EPI_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    OVB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<OVB_EPI_RESID_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
  }
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int sms = device_sm_count();
  const int pairs = sms / 2;
  const int grid = 2 * (num_tiles < pairs ? num_tiles : pairs);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool resid_tma = needs_resid && n_split == 1 && a->ldo % 4 == 0 &&
                         reinterpret_cast<uintptr_t>(a->resid) % 16 == 0 &&
                         (a->epilogue == OVB_EPI_RESID_F32 || a->gate_row_stride == 0) &&
                         getenv("OVB_GEMM_NO_TMA_EPILOGUE") == nullptr;
  if (resid_tma) {
    CUtensorMap tmR, tmO;
    const uint64_t dims[2] = {static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->M)};
    const uint32_t box[2] = {32, kBM};
    const uint64_t rs[1] = {static_cast<uint64_t>(a->ldr) * 4};
    const uint64_t os[1] = {static_cast<uint64_t>(a->ldo) * 4};
    int rc = encode_tmap_f32(&tmR, a->resid, 2, dims, rs, box);
    if (rc != OVB_OK) return rc;
    rc = encode_tmap_f32(&tmO, a->out, 2, dims, os, box);
    if (rc != OVB_OK) return rc;
    static bool attr2[64] = {};
    if (once_per_device(attr2)) {
      OVB_CHECK_CUDA(cudaFuncSetAttribute(gemm_resid_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmResidSmem));
      OVB_CHECK_CUDA(cudaFuncSetAttribute(gemm_resid_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmResidSmem));
    }
    if (a->epilogue == OVB_EPI_GATE_RESID_F32)
      gemm_resid_tma_kernel<true><<<grid, kGemmThreads, kGemmResidSmem, st>>>(tmA, tmB, tmR, tmO, p);
    else
      gemm_resid_tma_kernel<false><<<grid, kGemmThreads, kGemmResidSmem, st>>>(tmA, tmB, tmR, tmO, p);
    OVB_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return OVB_OK;
  }
  switch (a->epilogue) {
    case OVB_EPI_BF16: gemm_bf16_kernel<OVB_EPI_BF16><<<grid, kGemmThreadsEpi8, kGemmSmem, st>>>(tmA, tmB, p); break;
    case OVB_EPI_GELU_BF16: gemm_bf16_kernel<OVB_EPI_GELU_BF16><<<grid, kGemmThreadsEpi8, kGemmSmem, st>>>(tmA, tmB, p); break;
    case OVB_EPI_RESID_F32: gemm_bf16_kernel<OVB_EPI_RESID_F32><<<grid, kGemmThreadsEpi8, kGemmSmem, st>>>(tmA, tmB, p); break;
    case OVB_EPI_GATE_RESID_F32: gemm_bf16_kernel<OVB_EPI_GATE_RESID_F32><<<grid, kGemmThreadsEpi8, kGemmSmem, st>>>(tmA, tmB, p); break;
    case OVB_EPI_F32: gemm_bf16_kernel<OVB_EPI_F32><<<grid, kGemmThreadsEpi8, kGemmSmem, st>>>(tmA, tmB, p); break;
    default: gemm_bf16_kernel<OVB_EPI_RESID_BF16><<<grid, kGemmThreadsEpi8, kGemmSmem, st>>>(tmA, tmB, p); break;
  }
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}
