// ovb_conv3d: implicit-GEMM convolution over channels-last video tensors for sm_100a.
//
//   out[t, h, w, n] = bias[n] + sum_{tap, c} x[t*st + dt(tap), h*sh + dh(tap), w*sw + dw(tap), c]
//                                            * W[n, tap*Cin + c]            (+ resid[t, h, w, n])
//
// This is the CausalConv3d / Conv2d workhorse of the Wan2.1 VAE (reference
// omnivideo/modules/vae2_1.py:16-35, :75-95).  GEMM view: M = output voxels, N = Cout,
// K = taps * Cin.  An M tile is a bw x bh box of 128 output voxels of one frame; for every tap the
// A operand of that tile is the SAME box shifted by (dt, dh, dw), so the producer warp fetches it
// with one 4-D TMA box load per (tap, channel chunk) straight from the activation tensor — no
// im2col buffer, and TMA's zero fill outside the tensor IS the spatial zero padding and the
// causal (left-only) temporal padding.  The box lands in shared memory as 128 rows x KC channels,
// K-major with the hardware swizzle, i.e. exactly the tcgen05.mma A-operand layout.
//
// Same warp-specialised persistent structure as gemm.cu, on CTA PAIRS (tcgen05.mma.cta_group::2,
// M = 256): the two CTAs of a cluster take two consecutive M tiles (two voxel boxes) of the same N
// tile; each loads its own A box and only HALF of the weight rows, which halves the B-operand bytes
// an SM streams from shared memory per MMA (the measured limiter at these tile widths, see
// gemm.cu).  Warp 0 issues the A loads, warp 3 the B loads (TMA issue from one thread costs
// ~80 cycles per load plus ~250 per barrier round trip — scripts/microbench/tma_rate.cu — so two
// issuing threads and one stage per TAP (all channel chunks of a tap together, when they fit)
// keep the producer ahead of the tensor pipe); warp 1 of the leader issues the MMAs, warp 2
// owns TMEM (two accumulators, epilogue of tile i overlaps the main loop of tile i+1), warps 4-7
// run the epilogue.  Tile width BN, channel chunk KC (64/32/16 -> 128/64/32-byte swizzle), chunks
// per stage and ring depth are runtime parameters so one kernel serves every layer of the VAE
// (Cin, Cout in {16, 32, 96, 192, 384, 768} and the 3-channel head).
//
// The output address of voxel (t, h, w), channel n is
//     out_base + t*os_t + h*os_h + w*os_w + (n / n_group)*group_stride + (n % n_group)
// which lets the same kernel (a) write a 2x nearest-upsampled phase (os_h, os_w doubled, one launch
// per phase with phase-specific 2x2 taps) and (b) interleave the two channel halves of
// upsample3d's time_conv as consecutive frames (vae2_1.py:133-136).
#include <stdlib.h>
#include <string.h>

#include "host_util.h"
#include "ptx.cuh"

namespace ovb {

constexpr int kConvThreads = 256;
constexpr int kConvMaxStages = 8;
constexpr int kConvMaxTaps = 27;
constexpr int kConvSmemBudget = 225 * 1024;
constexpr int kConvStagingBytes = 4 * 32 * 32 * 4;

struct ConvParams {
  int To, Ho, Wo;
  int bw_log2, bh;   // box: (1 << bw_log2) x bh == 128 voxels
  int n_wt, n_ht;
  int st, sh, sw;
  int Cin, KC, n_kc;
  int kc_per_stage, a_chunk_bytes, b_chunk_bytes;
  int n_taps;
  int Cout, BN, n_nt;
  int stages, stage_bytes;
  const float* bias;
  void* out;
  const __nv_bfloat16* resid;
  long long os_t, os_h, os_w, out_base;
  int n_group;
  long long group_stride;
  int out_mode;
  long long plane_stride;
  // fused RMS_norm(+SiLU) over the channels of every output voxel (needs the whole channel range in
  // one N tile): norm_out (or, when null, out itself) receives act(v / max(|v|_2, 1e-12) * sqrt(C) * gamma)
  const float* norm_gamma;
  __nv_bfloat16* norm_out;
  int norm_silu;
  signed char dt[kConvMaxTaps], dh[kConvMaxTaps], dw[kConvMaxTaps];
  // Brick mode (stride-1 convs whose taps come in (dh0, dh0+1, dh0+2) triples sharing (dt, dw)): one
  // A load per triple fetches bh+2 input rows, and the three taps' A tiles are that brick read at row
  // offsets 0, bw, 2*bw — atom-aligned (bw % 8 == 0), so plain descriptor start-address offsets do it.
  int hb;  // tile order: bands of hb tile rows, all frames of a band before the next band (see decode_tile)
  int brick, n_groups, a_rows;
  signed char g_dt[9], g_dw[9], g_dh0[9];
  unsigned char g_tap[9][3];
};

// M tile index -> (frame, tile row, tile column).  Order: bands of `hb` tile rows; inside a band frame by
// frame, inside a frame row by row.  The three input frames a causal 3x3x3 tap set reads for one band then
// stay in L2 while the band is walked through time: every input frame comes from DRAM once instead of
// three times (the whole frame triple of a full-resolution layer is 230 MB, the L2 126 MB).
__device__ __forceinline__ void decode_tile(const ConvParams& p, int m, int& t, int& ht, int& wt) {
  const int per_band = p.hb * p.n_wt * p.To;
  const int band = m / per_band;
  int r = m - band * per_band;
  const int rows = min(p.hb, p.n_ht - band * p.hb);  // the last band may be shorter
  const int bt = rows * p.n_wt;
  t = r / bt;
  r -= t * bt;
  const int hr = r / p.n_wt;
  ht = band * p.hb + hr;
  wt = r - hr * p.n_wt;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv3d_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const ConvParams p) {  // tmA: box of 128 voxels, or (brick mode) of bw x (bh+2) voxels
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  float* staging = reinterpret_cast<float*>(smem + p.stages * p.stage_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes + kConvStagingBytes);
  uint64_t* empty_bar = full_bar + kConvMaxStages;
  uint64_t* tfull_bar = empty_bar + kConvMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int m_tiles = p.To * p.n_ht * p.n_wt;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int num_tiles = m_pairs * p.n_nt;  // pair tiles (256 voxels x BN)
  const int stages_per_tap = p.n_kc / p.kc_per_stage;
  const int num_ks = p.brick ? p.n_groups : p.n_taps * stages_per_tap;
  const int bw = 1 << p.bw_log2;
  const int a_stage_bytes = (p.brick ? p.n_kc : p.kc_per_stage) * p.a_chunk_bytes;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 2);  // leader's A producer + leader's B producer (expect_tx each)
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 8);  // 4 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ A producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = static_cast<uint32_t>(2 * p.kc_per_stage * 128 * p.KC * 2);
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m = (tile / p.n_nt) * 2 + static_cast<int>(rank);
        const bool real = m < m_tiles;
        int t = p.To + 4, ht = 0, wt = 0;  // odd tail: a box fully outside -> zero fill
        if (real) decode_tile(p, m, t, ht, wt);
        const int w0 = (wt << p.bw_log2) * p.sw, h0 = ht * p.bh * p.sh, t0 = t * p.st;
        if (p.brick) {
          const uint32_t txb = static_cast<uint32_t>(2 * p.n_kc * p.a_rows * p.KC * 2);
          for (int g = 0; g < p.n_groups; ++g) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * p.stage_bytes;
            const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (rank == 0) mbar_expect_tx(&full_bar[stage], txb);
            for (int q = 0; q < p.n_kc; ++q)
              tma_load_4d_pair(sa + q * p.a_chunk_bytes, &tmA, leader_full, q * p.KC, w0 + p.g_dw[g],
                               h0 + p.g_dh0[g], t0 + p.g_dt[g]);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          continue;
        }
        for (int tap = 0; tap < p.n_taps; ++tap) {
          const int ct = t0 + p.dt[tap], ch = h0 + p.dh[tap], cw = w0 + p.dw[tap];
          for (int g = 0; g < stages_per_tap; ++g) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * p.stage_bytes;
            const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (rank == 0) mbar_expect_tx(&full_bar[stage], tx);
            for (int q = 0; q < p.kc_per_stage; ++q)
              tma_load_4d_pair(sa + q * p.a_chunk_bytes, &tmA, leader_full,
                               (g * p.kc_per_stage + q) * p.KC, cw, ch, ct);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ B producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int half = p.BN >> 1;
      const uint32_t tx = static_cast<uint32_t>(2 * p.kc_per_stage * half * p.KC * 2);
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int nt = tile % p.n_nt;
        const int wrow0 = nt * p.BN + static_cast<int>(rank) * half;
        if (p.brick) {
          const uint32_t txb = static_cast<uint32_t>(2 * 3 * p.n_kc * half * p.KC * 2);
          for (int g = 0; g < p.n_groups; ++g) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sb = smem + stage * p.stage_bytes + a_stage_bytes;
            const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (rank == 0) mbar_expect_tx(&full_bar[stage], txb);
            for (int i = 0; i < 3; ++i)
              for (int q = 0; q < p.n_kc; ++q)
                tma_load_2d_pair(sb + (i * p.n_kc + q) * p.b_chunk_bytes, &tmB, leader_full,
                                 p.g_tap[g][i] * p.Cin + q * p.KC, wrow0);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          continue;
        }
        for (int tap = 0; tap < p.n_taps; ++tap) {
          for (int g = 0; g < stages_per_tap; ++g) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sb = smem + stage * p.stage_bytes + a_stage_bytes;
            const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (rank == 0) mbar_expect_tx(&full_bar[stage], tx);
            for (int q = 0; q < p.kc_per_stage; ++q)
              tma_load_2d_pair(sb + q * p.b_chunk_bytes, &tmB, leader_full,
                               tap * p.Cin + (g * p.kc_per_stage + q) * p.KC, wrow0);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(256, p.BN, 0, 0);
      const uint32_t row_bytes = p.KC * 2;
      const int k_steps = p.KC / 16;
      // Shared-memory descriptors: everything but the start address is the same for every MMA of the
      // launch, and all offsets are multiples of 16 bytes, so one MMA costs one 32-bit add on the
      // descriptor's low word.  (Rebuilding the descriptor from runtime parameters inside the triple
      // loop made the ISSUE the bottleneck: ~140 cycles per MMA against a 48-cycle tensor floor at
      // N = 96.)
      const uint64_t desc_tmpl = umma_desc_kmajor(0, row_bytes);
      const uint32_t a_chunk16 = static_cast<uint32_t>(p.a_chunk_bytes) >> 4;
      const uint32_t b_chunk16 = static_cast<uint32_t>(p.b_chunk_bytes) >> 4;
      const uint32_t tap_rows16 = (row_bytes << p.bw_log2) >> 4;  // brick mode: bw rows down the brick
      const int n_kc = p.n_kc, kc_per_stage = p.kc_per_stage, brick = p.brick, n_stages = p.stages;
      const uint32_t stage16 = static_cast<uint32_t>(p.stage_bytes) >> 4;
      const uint32_t a_stage16 = static_cast<uint32_t>(a_stage_bytes) >> 4;
      const uint32_t smem16 = (smem_u32(smem) & 0x3FFFF) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int ks = 0; ks < num_ks; ++ks) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa16 = smem16 + stage * stage16;
          const uint32_t sb16 = sa16 + a_stage16;
          uint32_t first = ks == 0 ? 0u : 1u;   // accumulate flag of the next MMA
          if (brick) {
            for (int i = 0; i < 3; ++i) {
              uint32_t a16 = sa16 + i * tap_rows16;
              uint32_t b16 = sb16 + i * n_kc * b_chunk16;
              for (int q = 0; q < n_kc; ++q) {
                for (int k = 0; k < k_steps; ++k) {
                  umma_ss_pair(d_tmem, desc_tmpl | ((a16 + 2 * k) & 0x3FFFu), desc_tmpl | ((b16 + 2 * k) & 0x3FFFu), idesc,
                               first);
                  first = 1u;
                }
                a16 += a_chunk16;
                b16 += b_chunk16;
              }
            }
          } else {
            uint32_t a16 = sa16, b16 = sb16;
            for (int q = 0; q < kc_per_stage; ++q) {
              for (int k = 0; k < k_steps; ++k) {
                umma_ss_pair(d_tmem, desc_tmpl | ((a16 + 2 * k) & 0x3FFFu), desc_tmpl | ((b16 + 2 * k) & 0x3FFFu), idesc,
                               first);
                first = 1u;
              }
              a16 += a_chunk16;
              b16 += b_chunk16;
            }
          }
          umma_commit_pair(&empty_bar[stage], 3);
          if (ks == num_ks - 1) umma_commit_pair(&tfull_bar[acc], 3);
          if (++stage == n_stages) {
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
    const int r = quad * 32 + lane;  // accumulator row = voxel inside the box
    const uint32_t stg_addr = smem_u32(staging + quad * (32 * 32));
    const uint32_t leader_tempty0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    const int rr_off = lane >> 3, cj = lane & 7;
    const bool planes = p.out_mode == OVB_CONV_OUT_BF16 && p.n_group < 16;  // narrow channel groups
    const bool transposed = p.out_mode == OVB_CONV_OUT_BF16 && (p.BN & 31) == 0 && !planes;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int nt = tile % p.n_nt;
      const int m = (tile / p.n_nt) * 2 + static_cast<int>(rank);
      const bool real = m < m_tiles;
      int t = 0, ht = 0, wt = 0;
      if (real) decode_tile(p, m, t, ht, wt);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#ifndef OVB_CONV_DBG
#define OVB_CONV_DBG 0
#endif
      if (OVB_CONV_DBG == 2) {
        // experiment: no epilogue at all (results are garbage): the main loop alone
      } else
      if (transposed) {
        // 32-column chunks through a swizzled shared-memory transpose (as in gemm.cu): in the read
        // phase 8 lanes cover 64 contiguous output bytes of one voxel and a warp instruction 4
        // voxels, instead of 32 voxels x 32 bytes.
        // With a fused RMS_norm the tile is walked twice: pass 0 only accumulates every voxel's sum
        // of squares (of the bf16-rounded values the unfused formulation would have stored), pass 1
        // recomputes the values from TMEM and writes them normalised.
        // One epilogue warp per scheduler executes this alone, at ~3-4 cycles per dependent
        // instruction: the instruction COUNT is what has to stay below the main loop's duration, so
        // everything that does not depend on the chunk is computed once per tile.
        const int n_chunks = p.BN / 32;
        const bool do_norm = p.norm_gamma != nullptr;
        const bool has_resid = p.resid != nullptr;
        __nv_bfloat16* const out_raw = (!do_norm || p.norm_out != nullptr) ? reinterpret_cast<__nv_bfloat16*>(p.out) : nullptr;
        __nv_bfloat16* const out_nrm = do_norm ? (p.norm_out != nullptr ? p.norm_out : reinterpret_cast<__nv_bfloat16*>(p.out)) : nullptr;
        // this lane's 8 voxels of the read phase: offset of channel 0, or -1
        long long vbase[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = quad * 32 + 4 * i + rr_off;
          const int h = ht * p.bh + (rr >> p.bw_log2), w = (wt << p.bw_log2) + (rr & (bw - 1));
          vbase[i] = (real && h < p.Ho && w < p.Wo) ? p.out_base + t * p.os_t + h * p.os_h + w * p.os_w : -1;
        }
        // chunk c of the accumulator -> the bf16 values of this lane's 8 voxels x 4 columns (bias and
        // residual applied, packed as the raw tensor holds them)
        auto chunk_values = [&](const int c, uint2 (&o8)[8], bool& col_ok, long long& coff) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256 + c * 32, v);
          const int col = nt * p.BN + c * 32 + cj * 4;
          col_ok = col < p.Cout;
          const int grp = col / p.n_group;
          coff = grp * p.group_stride + (col - grp * p.n_group);
          uint2 xb[8];  // residual values, requested up front.  The loads are UNCONDITIONAL (voxels or columns
          // outside the tensor read element 0; their results are never stored): a predicated load needs a
          // select behind it, and ptxas then waits for each load before it issues the next — eight serialised
          // DRAM latencies per chunk, which made the residual forms epilogue-bound (ncu source page, round 2).
          if (has_resid) {
            const long long csafe = col_ok ? coff : 0;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              xb[i] = *reinterpret_cast<const uint2*>(p.resid + (vbase[i] >= 0 ? vbase[i] : 0) + csafe);
          }
          uint64_t b01 = 0, b23 = 0;  // bias of this lane's 4 columns as two fp32 pairs
          if (p.bias != nullptr && col_ok) {
            const float4 bv = *reinterpret_cast<const float4*>(p.bias + col);
            b01 = pack_f32x2(bv.x, bv.y);
            b23 = pack_f32x2(bv.z, bv.w);
          }
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = stg_addr + lane * 128 + ((j ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "r"(v[4 * j]),
                         "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                         : "memory");
          }
          __syncwarp();
          // Packed arithmetic throughout (one epilogue warp per scheduler: the instruction count is the
          // cost): fp32x2 bias add, one cvt per pair, and the residual added as bf16x2 — the exact sum of
          // the bf16-rounded conv result and the bf16 residual, rounded once to bf16.
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + rr_off;
            const uint32_t a = stg_addr + rr * 128 + ((cj ^ (rr & 7)) << 4);
            uint64_t u01, u23;
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];\n" : "=l"(u01), "=l"(u23) : "r"(a) : "memory");
            float f0, f1, f2, f3;
            unpack_f32x2(fadd2(u01, b01), f0, f1);
            unpack_f32x2(fadd2(u23, b23), f2, f3);
            uint32_t c01 = pack_bf16x2(f0, f1), c23 = pack_bf16x2(f2, f3);
            if (has_resid) {
              // conv result is stored as bf16 before the add in the unfused formulation
              const __nv_bfloat162 s01 = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&c01),
                                                 *reinterpret_cast<const __nv_bfloat162*>(&xb[i].x));
              const __nv_bfloat162 s23 = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&c23),
                                                 *reinterpret_cast<const __nv_bfloat162*>(&xb[i].y));
              c01 = *reinterpret_cast<const uint32_t*>(&s01);
              c23 = *reinterpret_cast<const uint32_t*>(&s23);
            }
            o8[i].x = c01;
            o8[i].y = c23;
          }
          __syncwarp();
        };
        // sum of squares of this lane's 4 values of one voxel (the other 7 lanes that share the voxel,
        // cj = lane & 7, are added once all chunks are in: reduce8)
        auto sumsq4 = [&](const uint2 o, const bool col_ok) {
          const float2 r01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o.x));
          const float2 r23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o.y));
          return col_ok ? (r01.x * r01.x + r01.y * r01.y) + (r23.x * r23.x + r23.y * r23.y) : 0.f;
        };
        auto reduce8 = [&](float sq) {
          sq += __shfl_xor_sync(0xffffffffu, sq, 1);
          sq += __shfl_xor_sync(0xffffffffu, sq, 2);
          sq += __shfl_xor_sync(0xffffffffu, sq, 4);
          return sq;
        };
        // g01/g23: gamma of this lane's 4 columns, halved when SiLU follows: SiLU(y) = h + h tanh(h), h = y/2
        // (as rmsnorm_silu_cl_kernel; the factor 1/2 is exact, so (r*inv)*(g/2) == ((r*inv)*g)/2 bit for bit)
        auto normalised = [&](const uint2 o, const float inv, const uint64_t g01, const uint64_t g23) {
          const float2 r01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o.x));
          const float2 r23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o.y));
          const uint64_t inv2 = pack_f32x2(inv, inv);
          uint64_t h01 = fmul2(fmul2(pack_f32x2(r01.x, r01.y), inv2), g01);
          uint64_t h23 = fmul2(fmul2(pack_f32x2(r23.x, r23.y), inv2), g23);
          if (p.norm_silu) {
            float h0, h1, h2, h3, t0, t1, t2, t3;
            unpack_f32x2(h01, h0, h1);
            unpack_f32x2(h23, h2, h3);
            asm("tanh.approx.f32 %0, %1;\n" : "=f"(t0) : "f"(h0));
            asm("tanh.approx.f32 %0, %1;\n" : "=f"(t1) : "f"(h1));
            asm("tanh.approx.f32 %0, %1;\n" : "=f"(t2) : "f"(h2));
            asm("tanh.approx.f32 %0, %1;\n" : "=f"(t3) : "f"(h3));
            h01 = ffma2(h01, pack_f32x2(t0, t1), h01);
            h23 = ffma2(h23, pack_f32x2(t2, t3), h23);
          }
          float y0, y1, y2, y3;
          unpack_f32x2(h01, y0, y1);
          unpack_f32x2(h23, y2, y3);
          uint2 r;
          r.x = pack_bf16x2(y0, y1);
          r.y = pack_bf16x2(y2, y3);
          return r;
        };
        auto load_gamma = [&](const int c, uint64_t& g01, uint64_t& g23) {
          const int col = nt * p.BN + c * 32 + cj * 4;
          g01 = g23 = 0;
          if (col < p.Cout) {
            const float4 gv = *reinterpret_cast<const float4*>(p.norm_gamma + col);
            const float k = p.norm_silu ? 0.5f : 1.0f;
            g01 = pack_f32x2(k * gv.x, k * gv.y);
            g23 = pack_f32x2(k * gv.z, k * gv.w);
          }
        };
        const float sqrt_c = sqrtf(static_cast<float>(p.Cout));
        float ss[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (!do_norm) {
#pragma unroll 1
          for (int c = 0; c < n_chunks; ++c) {
            if (nt * p.BN + c * 32 >= p.Cout) break;  // warp-uniform
            uint2 o8[8];
            bool col_ok;
            long long coff;
            chunk_values(c, o8, col_ok, coff);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (col_ok && vbase[i] >= 0) *reinterpret_cast<uint2*>(out_raw + vbase[i] + coff) = o8[i];
          }
        } else if (n_chunks <= 3) {
          // <= 96 channels: ONE walk over the accumulator; the bf16 values stay in registers (two per
          // voxel and chunk) until the voxel's sum of squares is complete
          uint2 keep[3][8];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if (c < n_chunks) {
              bool col_ok;
              long long coff;
              chunk_values(c, keep[c], col_ok, coff);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                ss[i] += sumsq4(keep[c][i], col_ok);
                if (out_raw != nullptr && col_ok && vbase[i] >= 0)
                  *reinterpret_cast<uint2*>(out_raw + vbase[i] + coff) = keep[c][i];
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) ss[i] = sqrt_c * rsqrtf(fmaxf(reduce8(ss[i]), 1e-24f));  // sqrt(C) / max(|v|, 1e-12)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if (c < n_chunks) {
              const int col = nt * p.BN + c * 32 + cj * 4;
              const int grp = col / p.n_group;
              const long long coff = grp * p.group_stride + (col - grp * p.n_group);
              uint64_t g01, g23;
              load_gamma(c, g01, g23);
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (col < p.Cout && vbase[i] >= 0)
                  *reinterpret_cast<uint2*>(out_nrm + vbase[i] + coff) = normalised(keep[c][i], ss[i], g01, g23);
            }
          }
        } else {
          // more channels than the registers can keep: the tile is walked twice
#pragma unroll 1
          for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) ss[i] = sqrt_c * rsqrtf(fmaxf(reduce8(ss[i]), 1e-24f));
            }
#pragma unroll 1
            for (int c = 0; c < n_chunks; ++c) {
              if (nt * p.BN + c * 32 >= p.Cout) break;  // warp-uniform
              uint2 o8[8];
              bool col_ok;
              long long coff;
              chunk_values(c, o8, col_ok, coff);
              if (pass == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) ss[i] += sumsq4(o8[i], col_ok);
              } else {
                uint64_t g01, g23;
                load_gamma(c, g01, g23);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  if (col_ok && vbase[i] >= 0) {
                    if (out_raw != nullptr) *reinterpret_cast<uint2*>(out_raw + vbase[i] + coff) = o8[i];
                    *reinterpret_cast<uint2*>(out_nrm + vbase[i] + coff) = normalised(o8[i], ss[i], g01, g23);
                  }
                }
              }
            }
          }
        }
      } else if (planes) {
        // Per-group planes such as [group][T][H][W][4] (ovb_tap_sum's input): voxels are contiguous inside a
        // plane, so the accumulator's own mapping (lane = voxel) already coalesces — a warp instruction
        // writes 8 bytes x 32 consecutive voxels of one plane; no transpose.  The epilogue is all there is to
        // such a conv (K is one tap), so its instruction count is what matters: the plane pointer advances
        // incrementally, the bias test is hoisted out of the loops.
        const int hl = r >> p.bw_log2, wl = r & (bw - 1);
        const int h = ht * p.bh + hl, w = (wt << p.bw_log2) + wl;
        const bool ok = real && h < p.Ho && w < p.Wo;
        const int n_first = nt * p.BN;
        const int g0 = n_first / p.n_group;
        int inner = n_first - g0 * p.n_group;  // column inside its group
        __nv_bfloat16* pp = reinterpret_cast<__nv_bfloat16*>(p.out) + p.out_base + t * p.os_t + h * p.os_h +
                            w * p.os_w + g0 * p.group_stride + inner;
        const long long wrap = p.group_stride - (p.n_group - 4);
        const bool has_bias = p.bias != nullptr;
        const int n_chunks = p.BN / 32;
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
          const int n0 = n_first + c * 32;
          if (n0 >= p.Cout) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256 + c * 32, v);
          tmem_wait_ld();
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.Cout) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(p.bias + n0 + j));
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {  // n_group % 4 == 0: four columns never straddle two groups
            if (ok && n0 + j < p.Cout) {
              uint2 o;
              o.x = pack_bf16x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
              o.y = pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              *reinterpret_cast<uint2*>(pp) = o;
            }
            inner += 4;
            if (inner == p.n_group) {
              inner = 0;
              pp += wrap;
            } else {
              pp += 4;
            }
          }
          __syncwarp();
        }
      } else {
        const int hl = r >> p.bw_log2, wl = r & (bw - 1);
        const int h = ht * p.bh + hl, w = (wt << p.bw_log2) + wl;
        const bool ok = real && h < p.Ho && w < p.Wo;
        const long long vox = p.out_base + t * p.os_t + h * p.os_h + w * p.os_w;
        const int n_chunks = p.BN / 16;
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
          const int n0 = nt * p.BN + c * 16;
          if (n0 >= p.Cout) break;  // warp-uniform
          uint32_t v[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256 + c * 16, v);
          tmem_wait_ld();
          if (ok) {
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (n0 + j < p.Cout) f[j] += __ldg(p.bias + n0 + j);
            }
            if (p.out_mode == OVB_CONV_OUT_BF16) {
              const int grp = n0 / p.n_group;
              const long long off = vox + grp * p.group_stride + (n0 - grp * p.n_group);
              if (p.resid != nullptr) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.resid + off);
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                  const uint4 rv = rp[g];
                  const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 a = __bfloat1622float2(r2[j]);
                    f[g * 8 + 2 * j] = bf16_round(f[g * 8 + 2 * j]) + a.x;
                    f[g * 8 + 2 * j + 1] = bf16_round(f[g * 8 + 2 * j + 1]) + a.y;
                  }
                }
              }
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                uint4 o;
                o.x = pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]);
                o.y = pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]);
                o.z = pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]);
                o.w = pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]);
                op[g] = o;
              }
            } else {  // OVB_CONV_OUT_F32_PLANAR(_CLAMP): out[n][t][h][w]
              float* op = reinterpret_cast<float*>(p.out) + vox;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (n0 + j < p.Cout) {
                  float y = f[j];
                  if (p.out_mode == OVB_CONV_OUT_F32_PLANAR_CLAMP) y = fminf(fmaxf(y, -1.0f), 1.0f);
                  op[(n0 + j) * p.plane_stride] = y;
                }
              }
            }
          }
          __syncwarp();
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
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

}  // namespace ovb

extern "C" int ovb_conv3d(const ovb_conv3d_args* a, void* stream) {
  using namespace ovb;
  OVB_CHECK_ARG(a != nullptr, "ovb_conv3d: null args");
  OVB_CHECK_ARG(a->x && a->w && a->out, "ovb_conv3d: null x/w/out");
  OVB_CHECK_ARG(a->Ti > 0 && a->Hi > 0 && a->Wi > 0 && a->To > 0 && a->Ho > 0 && a->Wo > 0,
                "ovb_conv3d: bad extents in [%d,%d,%d] out [%d,%d,%d]", a->Ti, a->Hi, a->Wi, a->To,
                a->Ho, a->Wo);
  OVB_CHECK_ARG(a->Cin > 0 && a->Cin % 16 == 0, "ovb_conv3d: Cin=%d must be a multiple of 16", a->Cin);
  OVB_CHECK_ARG(a->Cout > 0 && a->w_rows >= a->Cout && a->w_rows % 16 == 0,
                "ovb_conv3d: Cout=%d w_rows=%d (w_rows must be >= Cout and a multiple of 16)", a->Cout,
                a->w_rows);
  OVB_CHECK_ARG(a->n_taps >= 1 && a->n_taps <= kConvMaxTaps && a->tap_dt && a->tap_dh && a->tap_dw,
                "ovb_conv3d: n_taps=%d outside [1,%d] or null tap arrays", a->n_taps, kConvMaxTaps);
  OVB_CHECK_ARG(a->st >= 1 && a->sh >= 1 && a->sw >= 1 && a->st <= 2 && a->sh <= 2 && a->sw <= 2,
                "ovb_conv3d: strides must be 1 or 2");
  OVB_CHECK_ARG(a->out_mode >= OVB_CONV_OUT_BF16 && a->out_mode <= OVB_CONV_OUT_F32_PLANAR_CLAMP,
                "ovb_conv3d: unknown out_mode %d", a->out_mode);
  OVB_CHECK_ARG((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->w) |
                 reinterpret_cast<uintptr_t>(a->out) | reinterpret_cast<uintptr_t>(a->resid)) % 16 == 0,
                "ovb_conv3d: pointers must be 16-byte aligned");
  // Channel groups narrower than 16 (down to 4 channels = one 8-byte store per voxel and group) turn
  // the output into per-group planes, e.g. [group][T][H][W][4]: the per-tap partial sums ovb_tap_sum consumes.
  const bool narrow = a->out_mode == OVB_CONV_OUT_BF16 && a->n_group > 0 && a->n_group < 16;
  if (a->out_mode == OVB_CONV_OUT_BF16) {
    const int al = narrow ? 4 : 8, cal = narrow ? 4 : 16;
    OVB_CHECK_ARG(a->Cout % cal == 0, "ovb_conv3d: bf16 output needs Cout %% %d == 0 (got %d)", cal, a->Cout);
    OVB_CHECK_ARG(a->os_t % al == 0 && a->os_h % al == 0 && a->os_w % al == 0 && a->out_base % al == 0 &&
                      a->group_stride % al == 0,
                  "ovb_conv3d: bf16 output strides must be multiples of %d elements", al);
    OVB_CHECK_ARG(a->n_group == 0 || (a->n_group % cal == 0 && a->Cout % a->n_group == 0),
                  "ovb_conv3d: n_group=%d must be a multiple of 16 (or of 4, below 16) dividing Cout", a->n_group);
    OVB_CHECK_ARG(!narrow || (a->resid == nullptr && a->norm_gamma == nullptr && a->w_rows % 32 == 0 &&
                              a->w_rows <= 256),
                  "ovb_conv3d: channel groups below 16 need w_rows %% 32 == 0, w_rows <= 256, no resid, no norm");
  } else {
    OVB_CHECK_ARG(a->resid == nullptr && a->n_group == 0,
                  "ovb_conv3d: planar fp32 output supports neither resid nor channel groups");
  }

  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.To = a->To; p.Ho = a->Ho; p.Wo = a->Wo;
  p.st = a->st; p.sh = a->sh; p.sw = a->sw;
  // ---- M tile shape: bw x bh = 128 voxels with the least padding waste
  {
    long long best = -1;
    for (int l2 = 7; l2 >= 3; --l2) {
      const int bw = 1 << l2, bh = 128 >> l2;
      const long long cover = static_cast<long long>((a->Wo + bw - 1) / bw) * bw *
                              ((a->Ho + bh - 1) / bh) * bh;
      if (best < 0 || cover < best) {
        best = cover;
        p.bw_log2 = l2;
        p.bh = bh;
      }
    }
  }
  // ---- K chunk
  p.Cin = a->Cin;
  p.KC = a->Cin % 64 == 0 ? 64 : (a->Cin % 32 == 0 ? 32 : 16);
  p.n_kc = a->Cin / p.KC;
  p.n_taps = a->n_taps;
  for (int i = 0; i < a->n_taps; ++i) {
    p.dt[i] = a->tap_dt[i]; p.dh[i] = a->tap_dh[i]; p.dw[i] = a->tap_dw[i];
  }
  // ---- brick mode: do the taps come in (dh0, dh0+1, dh0+2) triples that share (dt, dw)?
  bool brick_taps = a->st == 1 && a->sh == 1 && a->sw == 1 && a->n_taps % 3 == 0 && a->n_taps <= 27 &&
                    getenv("OVB_CONV_NO_BRICK") == nullptr;
  if (brick_taps) {
    bool used[kConvMaxTaps] = {};
    int ng = 0;
    for (int i = 0; i < a->n_taps && brick_taps; ++i) {
      if (used[i]) continue;
      // i must be the top tap (smallest dh) of its triple
      int idx[3] = {i, -1, -1};
      for (int j = 0; j < a->n_taps; ++j) {
        if (used[j] || j == i || p.dt[j] != p.dt[i] || p.dw[j] != p.dw[i]) continue;
        if (p.dh[j] == p.dh[i] + 1) idx[1] = j;
        if (p.dh[j] == p.dh[i] + 2) idx[2] = j;
        if (p.dh[j] < p.dh[i]) idx[0] = -1;  // not the top tap: handled when its top tap comes up
      }
      if (idx[0] < 0) continue;
      if (idx[1] < 0 || idx[2] < 0 || ng >= 9) { brick_taps = false; break; }
      p.g_dt[ng] = p.dt[i]; p.g_dw[ng] = p.dw[i]; p.g_dh0[ng] = p.dh[i];
      for (int u = 0; u < 3; ++u) { p.g_tap[ng][u] = static_cast<unsigned char>(idx[u]); used[idx[u]] = true; }
      ++ng;
    }
    for (int i = 0; i < a->n_taps && brick_taps; ++i) brick_taps = used[i];
    p.n_groups = ng;
    if (ng * 3 != a->n_taps) brick_taps = false;
  }
  if (brick_taps) {
    // tallest box (least halo overhead (bh+2)/bh) among the shapes within 2 % of the least padding waste
    long long best_cover = -1;
    for (int l2 = 7; l2 >= 3; --l2) {
      const int bwc = 1 << l2, bhc = 128 >> l2;
      const long long cover = static_cast<long long>((a->Wo + bwc - 1) / bwc) * bwc * ((a->Ho + bhc - 1) / bhc) * bhc;
      if (best_cover < 0 || cover < best_cover) best_cover = cover;
    }
    const char* force = getenv("OVB_CONV_BOX_L2");  // experiments: log2 of the box width
    const int lo = force != nullptr ? atoi(force) : 3, hi = force != nullptr ? atoi(force) : 7;
    for (int l2 = lo; l2 <= hi; ++l2) {
      const int bwc = 1 << l2, bhc = 128 >> l2;
      const long long cover = static_cast<long long>((a->Wo + bwc - 1) / bwc) * bwc * ((a->Ho + bhc - 1) / bhc) * bhc;
      if (force != nullptr || cover * 100 <= best_cover * 102) { p.bw_log2 = l2; p.bh = bhc; break; }
    }
  }
  const int bw = 1 << p.bw_log2;
  p.n_wt = (a->Wo + bw - 1) / bw;
  p.n_ht = (a->Ho + p.bh - 1) / p.bh;
  {
    // band height: ~8 MB of input per frame and band, so that three frames of a band sit in L2 with room to spare
    const long long tile_row_bytes = static_cast<long long>(p.bh) * a->sh * a->Wi * a->Cin * 2;
    long long hb = (8ll << 20) / (tile_row_bytes > 0 ? tile_row_bytes : 1);
    if (hb < 1) hb = 1;
    if (hb > p.n_ht) hb = p.n_ht;
    p.hb = static_cast<int>(hb);
  }
  // ---- N tile
  p.Cout = a->Cout;
  {
    const int g = a->n_group >= 16 ? a->n_group : a->w_rows;
    int bn = 0;
    for (int c = 256; c >= 16; c -= 16)
      if (g % c == 0) { bn = c; break; }
    if (a->n_group == 0 && bn < 64 && a->w_rows > 256) bn = 256;  // awkward Cout: partial last tile
    if (narrow) bn = a->w_rows;  // one N tile, a multiple of 32 (the planes epilogue reads 32-column chunks)
    p.BN = bn;
    p.n_nt = (a->w_rows + bn - 1) / bn;
  }
  OVB_CHECK_ARG(p.BN >= 16 && p.BN % 16 == 0, "ovb_conv3d: no valid N tile for Cout=%d", a->Cout);
  // ---- shared-memory ring: one stage holds `kc_per_stage` channel chunks of one tap
  p.a_chunk_bytes = 128 * p.KC * 2;                                       // multiple of 4096
  p.b_chunk_bytes = ((p.BN / 2) * p.KC * 2 + 1023) / 1024 * 1024;
  const int chunk_bytes = p.a_chunk_bytes + p.b_chunk_bytes;
  const int ring_budget = kConvSmemBudget - kConvStagingBytes - 2048;
  p.kc_per_stage = 1;
  for (int q = p.n_kc; q >= 1; --q) {  // largest divisor of n_kc that still leaves >= 4 stages
    if (p.n_kc % q == 0 && q * chunk_bytes * 4 <= ring_budget) {
      p.kc_per_stage = q;
      break;
    }
  }
  p.stage_bytes = p.kc_per_stage * chunk_bytes;
  p.stages = ring_budget / p.stage_bytes;
  p.a_rows = 128;
  if (brick_taps) {
    // one stage = one tap triple, all channel chunks: A brick of (bh+2) x bw voxels + 3 taps of weights
    const int a_rows = (p.bh + 2) * bw;
    const int a_chunk = (a_rows * p.KC * 2 + 1023) / 1024 * 1024;
    const int stage = p.n_kc * a_chunk + 3 * p.n_kc * p.b_chunk_bytes;
    if (ring_budget / stage >= 3) {
      p.brick = 1;
      p.a_rows = a_rows;
      p.a_chunk_bytes = a_chunk;
      p.kc_per_stage = p.n_kc;
      p.stage_bytes = stage;
      p.stages = ring_budget / stage;
    }
  }
  if (!p.brick) {
    // (the box shape chosen for brick mode is still a valid 128-voxel box)
    p.n_groups = 0;
  }
  if (p.stages > kConvMaxStages) p.stages = kConvMaxStages;
  OVB_CHECK_ARG(p.stages >= 2, "ovb_conv3d: tile does not fit shared memory (BN=%d KC=%d)", p.BN, p.KC);
  p.bias = a->bias;
  p.out = a->out;
  p.resid = reinterpret_cast<const __nv_bfloat16*>(a->resid);
  p.os_t = a->os_t; p.os_h = a->os_h; p.os_w = a->os_w; p.out_base = a->out_base;
  p.n_group = a->n_group > 0 ? a->n_group : 0;
  p.group_stride = a->group_stride;
  if (p.n_group == 0) p.n_group = 1 << 30;  // n / n_group == 0 for every n
  p.out_mode = a->out_mode;
  p.plane_stride = a->plane_stride;
  p.norm_gamma = a->norm_gamma;
  p.norm_out = reinterpret_cast<__nv_bfloat16*>(a->norm_out);
  p.norm_silu = a->norm_silu;
  if (a->norm_gamma != nullptr) {
    OVB_CHECK_ARG(a->out_mode == OVB_CONV_OUT_BF16 && p.n_nt == 1 && p.BN % 32 == 0 && a->n_group == 0 &&
                      a->Cout == a->w_rows,
                  "ovb_conv3d: the fused RMS_norm needs bf16 output, no channel groups and all %d channels in "
                  "one N tile (BN=%d, tiles=%d)", a->Cout, p.BN, p.n_nt);
    OVB_CHECK_ARG(reinterpret_cast<uintptr_t>(a->norm_gamma) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(a->norm_out) % 16 == 0,
                  "ovb_conv3d: norm_gamma / norm_out must be 16-byte aligned");
  } else {
    OVB_CHECK_ARG(a->norm_out == nullptr, "ovb_conv3d: norm_out without norm_gamma");
  }

  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(a->Cin), static_cast<uint64_t>(a->Wi),
                              static_cast<uint64_t>(a->Hi), static_cast<uint64_t>(a->Ti)};
    const uint64_t strides[3] = {static_cast<uint64_t>(a->Cin) * 2,
                                 static_cast<uint64_t>(a->Cin) * 2 * a->Wi,
                                 static_cast<uint64_t>(a->Cin) * 2 * a->Wi * a->Hi};
    // with an element stride s the box spans s*n source elements and delivers n of them
    const uint32_t box[4] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(bw * a->sw),
                             static_cast<uint32_t>((p.brick ? p.bh + 2 : p.bh) * a->sh), 1};
    const uint32_t estr[4] = {1, static_cast<uint32_t>(a->sw), static_cast<uint32_t>(a->sh), 1};
    OVB_CHECK_ARG(box[1] <= 256 && box[2] <= 256, "ovb_conv3d: strided box too large");
    int rc = encode_tmap_bf16_ex(&tmA, a->x, 4, dims, strides, box, estr, p.KC * 2);
    if (rc != OVB_OK) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->n_taps) * a->Cin, static_cast<uint64_t>(a->w_rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->n_taps) * a->Cin * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(p.KC), static_cast<uint32_t>(p.BN / 2)};
    int rc = encode_tmap_bf16_ex(&tmB, a->w, 2, dims, strides, box, nullptr, p.KC * 2);
    if (rc != OVB_OK) return rc;
  }

  const int smem_bytes = p.stages * p.stage_bytes + kConvStagingBytes + 1024 + 512;
  static bool attr_set[64] = {};
  if (once_per_device(attr_set)) {
    OVB_CHECK_CUDA(cudaFuncSetAttribute(conv3d_igemm_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kConvSmemBudget + 2048));
  }
  const long long m_tiles = static_cast<long long>(p.To) * p.n_ht * p.n_wt;
  const long long num_tiles = ((m_tiles + 1) / 2) * p.n_nt;  // pair tiles
  OVB_CHECK_ARG(m_tiles < (1ll << 30) && num_tiles < (1ll << 30), "ovb_conv3d: too many tiles");
  const int pairs = device_sm_count() / 2;
  const int grid = 2 * (num_tiles < pairs ? static_cast<int>(num_tiles) : pairs);
  conv3d_igemm_kernel<<<grid, kConvThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}
