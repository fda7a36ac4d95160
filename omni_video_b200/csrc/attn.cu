// ovb_attn_fwd: flash-attention forward for sm_100a (head_dim 128, non-causal, key-length mask).
//
// One CTA = 256 query rows of one (batch, head): two 128-row Q tiles that ping-pong on the tensor
// core so that the softmax of one tile overlaps the MMAs of the other.
//   warp 0        TMA producer: Q (once), then K_j / V_j 128-row tiles into two 2-stage rings
//   warp 1        one thread issues tcgen05.mma:  S_i = Q_i K_j^T  (SS, both K-major)  and
//                 O_i += P_i V_j  (TS: P read from TMEM, V MN-major from shared memory)
//   warp 2        TMEM allocator (512 columns: S0 | S1 | O0 | O1, each 128 lanes x 128 fp32)
//   warps 4-7     softmax warpgroup for Q tile 0 : one thread per query row
//   warps 8-11    softmax warpgroup for Q tile 1
// Per KV tile a softmax thread reads its S row (tcgen05.ld), keeps a running max with LAZY
// rescaling (O and the row sum are only rescaled when the max grows by more than 2^8, which also
// makes a separate correction warpgroup unnecessary: "S_j ready" implies "P_{j-1} V_{j-1} done",
// so the owner of the row can rescale O in TMEM itself), exponentiates with ex2.approx in the
// log2 domain, and writes bf16 P over the first 64 columns of its own S region (tcgen05.st).
// Keys >= k_len are masked to -inf; K/V rows past the tensor end are zero-filled by TMA.
// P is published in two parts (keys [0,96), then the rest) so that PV starts before the row is done;
// the launch is a 1-D grid of (q block, head, sample) items whose last partial wave is cut along the
// keys (AttnParams, plan_split, attn_combine_kernel); in the scatter form the epilogue stores every
// row into the buffer of the sequence-parallel rank that owns it (out_row_ptr).
//
// FLOPs per CTA-step (one KV tile, both Q tiles): 4 * 128*128*128 MACs; bytes from L2: 64 KB.
// Co-limited by MUFU.EX2 (8 cycles per warp instruction: 2048 XU cycles per step against 2848
// tensor-pipe cycles) through the per-tile chain QK -> softmax -> PV -> QK; DESIGN.md §3.1.
#include <type_traits>

#include "host_util.h"
#include "ptx.cuh"

namespace ovb {

#ifndef OVB_TRACE_QUAD
#define OVB_TRACE_QUAD 0
#endif
#ifdef OVB_ATTN_TRACE
// Debug build only (scripts/attn_trace.py): clock64() stamps of one CTA's pipeline events.
__device__ long long* g_attn_trace = nullptr;
#define OVB_TRACE(cond, slot)                                                      \
  do {                                                                             \
    if ((cond) && blockIdx.x == 3 && blockIdx.y == 0 && g_attn_trace != nullptr)  \
      g_attn_trace[(slot)] = clock64();                                            \
  } while (0)
#else
#define OVB_TRACE(cond, slot) do { } while (0)
#endif

constexpr int kAttnThreads = 384;
constexpr int kTileBytes = 128 * 128 * 2;  // 32 KB : one 128x128 bf16 tile (two 64-col TMA boxes)
constexpr int kBoxBytes = 128 * 64 * 2;    // 16 KB
constexpr int kKVStages = 2;
constexpr int kPartSteps = 6;  // PV k-steps (16 keys each) covered by the first P publication
constexpr int kAttnSmem = 2 * kTileBytes + 2 * kKVStages * kTileBytes + 1024 + 256;

// TMEM column map
constexpr uint32_t kColS0 = 0, kColS1 = 128, kColO0 = 256, kColO1 = 384;

struct AttnParams {
  __nv_bfloat16* out;
  long long o_row_stride, o_batch_stride;
  int Lq, Lk, H;
  float scale_log2;  // softmax_scale * log2(e)
  int kl0, kl1, kl2, kl3, kl4, kl5, kl6, kl7;  // per-sample valid keys (B <= 8)
  // Work decomposition: item = (q block, head, sample), q block fastest (CTAs that run together
  // share a head's K/V in L2).  CTAs [0, n_whole) take one whole item each; the last, partial wave
  // of items is cut `parts` ways along the keys so that it fills the machine: CTA n_whole + r
  // computes keys slice r % parts of item n_whole + r / parts and leaves an UNNORMALISED partial
  // (O, running max, row sum) in the workspace for attn_combine_kernel.
  int nqb, n_whole, parts;
  float* ws_o;   // [n_tail_ctas][256][128] fp32
  float2* ws_ml; // [n_tail_ctas][256] (max in the log2 domain, row sum)
  // scatter form: row r of the (single) sample goes to outp[r / rows_per_out] (maybe peer memory)
  __nv_bfloat16* outp[8];
  int rows_per_out;  // 0: plain form
};

__device__ __forceinline__ __nv_bfloat16* out_row_ptr(const AttnParams& p, int batch, int q_row, int head) {
  if (p.rows_per_out > 0) {
    const int owner = q_row / p.rows_per_out;
    return p.outp[owner] + static_cast<long long>(q_row - owner * p.rows_per_out) * p.o_row_stride + head * 128;
  }
  return p.out + static_cast<long long>(batch) * p.o_batch_stride +
         static_cast<long long>(q_row) * p.o_row_stride + head * 128;
}

__device__ __forceinline__ void setmaxnreg_dec_72() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;\n");
}
__device__ __forceinline__ void setmaxnreg_inc_216() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 216;\n");
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                  // 2 tiles
  uint8_t* sK = sQ + 2 * kTileBytes;                   // kKVStages tiles
  uint8_t* sV = sK + kKVStages * kTileBytes;           // kKVStages tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKVStages * kTileBytes);
  uint64_t* q_full = bars;          // [1]
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [2] per Q tile
  uint64_t* p_full = bars + 11;     // [2] per Q tile
  uint64_t* o_full = bars + 13;     // [2] per Q tile
  uint64_t* p_part = bars + 15;     // [2] per Q tile: P columns of keys [0, 96) published
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int item = blockIdx.x, part = 0, nparts = 1, slot = -1;
  if (item >= p.n_whole) {
    slot = item - p.n_whole;
    item = p.n_whole + slot / p.parts;
    part = slot % p.parts;
    nparts = p.parts;
  }
  const int q_blk = item % p.nqb;
  const int head = (item / p.nqb) % p.H;
  const int batch = item / (p.nqb * p.H);
  const int k_len = batch == 0 ? p.kl0 : batch == 1 ? p.kl1 : batch == 2 ? p.kl2 : batch == 3 ? p.kl3
                  : batch == 4 ? p.kl4 : batch == 5 ? p.kl5 : batch == 6 ? p.kl6 : p.kl7;
  const int n_kv_all = (k_len + 127) / 128;
  const int kv0 = (n_kv_all * part) / nparts;               // first KV tile of this CTA
  const int n_kv = (n_kv_all * (part + 1)) / nparts - kv0;  // tiles of this CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);  // one arrival per softmax warp
      mbar_init(&p_part[i], 4);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec_72();
    if (warp == 0 && n_kv > 0 && elect_one()) {
      // -------------------------------------------------------- TMA producer
      const int col = head * 128;
      mbar_expect_tx(q_full, 2 * kTileBytes);
      for (int i = 0; i < 2; ++i) {
        const int row = q_blk * 256 + i * 128;
        tma_load_3d(sQ + i * kTileBytes, &tmQ, q_full, col, row, batch);
        tma_load_3d(sQ + i * kTileBytes + kBoxBytes, &tmQ, q_full, col + 64, row, batch);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kKVStages;
        const uint32_t ph = (j / kKVStages) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], kTileBytes);
        const int krow = (kv0 + j) * 128;
        tma_load_3d(sK + s * kTileBytes, &tmK, &k_full[s], col, krow, batch);
        tma_load_3d(sK + s * kTileBytes + kBoxBytes, &tmK, &k_full[s], col + 64, krow, batch);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], kTileBytes);
        tma_load_3d(sV + s * kTileBytes, &tmV, &v_full[s], col, krow, batch);
        tma_load_3d(sV + s * kTileBytes + kBoxBytes, &tmV, &v_full[s], col + 64, krow, batch);
      }
    } else if (warp == 1 && n_kv > 0 && elect_one()) {
      // -------------------------------------------------------- MMA issuer
      const uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(128, 128, 0, 1);  // V is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t k_addr = smem_u32(sK);
      const uint32_t v_addr = smem_u32(sV);

      auto issue_qk = [&](int i, int s) {
        const uint32_t qa = q_addr + i * kTileBytes;
        const uint32_t ka = k_addr + s * kTileBytes;
        const uint32_t d = tmem_base + (i == 0 ? kColS0 : kColS1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * kBoxBytes + (kk & 3) * 32;
          umma_ss(d, umma_desc_sw128(qa + off, 16, 1024), umma_desc_sw128(ka + off, 16, 1024),
                  idesc_qk, kk != 0);
        }
      };
      auto issue_pv = [&](int i, int s, bool first, int kk0, int kk1) {
        const uint32_t va = v_addr + s * kTileBytes;
        const uint32_t pt = tmem_base + (i == 0 ? kColS0 : kColS1);
        const uint32_t d = tmem_base + (i == 0 ? kColO0 : kColO1);
        for (int kk = kk0; kk < kk1; ++kk) {
          // 16 kv rows per MMA: 2048 B down the V tile; second 64 columns are kBoxBytes away (LBO)
          umma_ts(d, pt + kk * 8, umma_desc_sw128(va + kk * 2048, kBoxBytes, 1024), idesc_pv,
                  !(first && kk == 0));
        }
      };

      // O_i += P_i V: k-steps [0,6) as soon as keys [0,96) of P are published, the rest when the
      // softmax warpgroup has finished the row.
      auto pv_tile = [&](int i, int s, int jj) {
        mbar_wait(&p_part[i], jj & 1);
        tc_fence_after();
        OVB_TRACE(jj + i >= 64 && jj + i < 72, (jj + i - 64) * 32 + (i == 0 ? 4 : 6));   // MMA: P part seen
        issue_pv(i, s, jj == 0, 0, kPartSteps);
        mbar_wait(&p_full[i], jj & 1);
        tc_fence_after();
        issue_pv(i, s, jj == 0, kPartSteps, 8);
      };

      mbar_wait(q_full, 0);
      tc_fence_after();
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kKVStages;
        const uint32_t ph = (j / kKVStages) & 1;
        mbar_wait(&k_full[s], ph);
        tc_fence_after();
        OVB_TRACE(j >= 64 && j < 72, (j - 64) * 32 + 0);   // MMA: about to issue QK0(j)
        issue_qk(0, s);
        umma_commit(&s_full[0]);
        OVB_TRACE(j >= 64 && j < 72, (j - 64) * 32 + 1);   // MMA: QK0 issued
        if (j > 0) {
          const int sp = (j - 1) % kKVStages;
          pv_tile(1, sp, j - 1);
          OVB_TRACE(j >= 64 && j < 72, (j - 64) * 32 + 2);   // MMA: PV1(j-1) issued
          umma_commit(&v_empty[sp]);
        }
        issue_qk(1, s);
        umma_commit(&s_full[1]);
        umma_commit(&k_empty[s]);
        OVB_TRACE(j >= 64 && j < 72, (j - 64) * 32 + 3);   // MMA: QK1(j) issued
        mbar_wait(&v_full[s], ph);
        pv_tile(0, s, j);
        OVB_TRACE(j >= 64 && j < 72, (j - 64) * 32 + 5);   // MMA: PV0(j) issued
      }
      umma_commit(&o_full[0]);
      {
        const int sp = (n_kv - 1) % kKVStages;
        pv_tile(1, sp, n_kv - 1);
        umma_commit(&v_empty[sp]);
        umma_commit(&o_full[1]);
      }
    }
  } else {
    // ---------------------------------------------------------- softmax warpgroups
    setmaxnreg_inc_216();
    const int wg = (warp - 4) >> 2;  // Q tile index
    const int quad = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + (wg == 0 ? kColS0 : kColS1);
    const uint32_t tO = tmem_base + lane_base + (wg == 0 ? kColO0 : kColO1);
    const int q_row = q_blk * 256 + wg * 128 + quad * 32 + lane;
    const float scale = p.scale_log2;
    const uint32_t rt_zero = static_cast<uint32_t>(p.Lk) >> 31;  // 0, but not provably so

    float m_used = -INFINITY;  // reference max (raw score units) the exponentials are taken against
    float l_sum = 0.f;

    // One KV tile of this row.  kMask (only the LAST tile of a sequence whose length is not a
    // multiple of 128) is a compile-time flag on purpose: written as a run-time `if`, the compiler
    // if-converts the 128 compare+select pairs into EVERY tile's instruction stream.
    //
    // What paces this loop (scripts/microbench/{xu_rate,tmem_rate}.cu): MUFU.EX2 issues once every
    // 8 cycles per scheduler, so the 128 exponentials of a row cost >= 1024 cycles — and the two
    // warpgroups share the schedulers, so a pass that overlaps the other tile's pass sees half that
    // rate.  tcgen05.st costs ~28 cycles and F2FP ~1.3; neither matters.  The kernel's period is
    // the chain  QK_i(j) -> softmax_i(j) -> PV_i(j) -> QK_i(j+1), so everything here is about
    // SHORTENING THE CHAIN, not about instruction count:
    //   * the row is loaded in two waits (chunk 0 first) so the first exponentials start ~120
    //     cycles earlier;
    //   * the pass runs SPECULATIVELY against the reference max carried over from earlier tiles,
    //     the tile's own maximum being collected in the shadow of the MUFU-paced loop (a separate
    //     max pass cost 345 cycles on the chain);
    //   * P is published in two parts: keys [0,96) as soon as their stores have landed, so the
    //     tensor core starts 6 of the 8 PV k-steps while the last 32 exponentials are still being
    //     evaluated.  For that the WHOLE tile's maximum must be known at the 3/4 point (nothing
    //     published may have to be taken back), so chunk 3's maxima are folded into chunks 1-2.
    // The reference is validated at the 3/4 point; if some row outgrew it by more than 2^8 — always
    // on the first tile, rarely later — O and l move to the new reference and the pass is redone
    // from the registers before anything is published.
    auto process_tile = [&](int j, auto mask_tag) {
      constexpr bool kMask = decltype(mask_tag)::value;
      mbar_wait(&s_full[wg], j & 1);
      tc_fence_after();
      OVB_TRACE(j >= 64 && j < 72 && quad == OVB_TRACE_QUAD && lane == 0, (j - 64) * 32 + 8 + wg * 8 + 0);  // S seen
      float s[128];
      {
        uint32_t r[32];
        tmem_ld32(tS, r);
        tmem_wait_ld();
#pragma unroll
        for (int t = 0; t < 32; ++t) s[t] = __uint_as_float(r[t]);
      }
#pragma unroll
      for (int c = 1; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tS + c * 32, r);
#pragma unroll
        for (int t = 0; t < 32; ++t) s[c * 32 + t] = __uint_as_float(r[t]);
      }
      const int valid = k_len - (kv0 + j) * 128;  // kMask: 1..127
      if constexpr (kMask) {
#pragma unroll
        for (int t = 0; t < 32; ++t) s[t] = t < valid ? s[t] : -INFINITY;
      }
      // columns 32..127 are in flight until this wait; the "+f" operands pin their first use below it
      auto wait_rest = [&]() {
        tmem_wait_ld();
#pragma unroll
        for (int c = 1; c < 4; ++c) reg_fence32(&s[c * 32]);
        if constexpr (kMask) {
#pragma unroll
          for (int t = 32; t < 128; ++t) s[t] = t < valid ? s[t] : -INFINITY;
        }
      };
      OVB_TRACE(j >= 64 && j < 72 && quad == OVB_TRACE_QUAD && lane == 0, (j - 64) * 32 + 8 + wg * 8 + 1);  // first chunk loaded

      const uint64_t scale2 = pack_f32x2(scale, scale);
      uint64_t acc_a = 0ull, acc_b = 0ull;  // (0.f, 0.f)
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
      // 32 keys: exponentials in the log2 domain, scale/shift FMA and row-sum adds packed as fp32x2
      // (FFMA2 / FADD2); bf16 P overwrites columns [16c, 16c+16) of this row's S.  `track`: fold
      // the chunk's own scores — and, for chunks 1 and 2, half of chunk 3's — into the maxima.
      auto exp_chunk = [&](const int c, const uint64_t negm2, const bool track) {
        uint32_t pk[16];
#pragma unroll
        for (int t = 0; t < 32; t += 4) {
          const float s0 = s[c * 32 + t], s1 = s[c * 32 + t + 1], s2 = s[c * 32 + t + 2],
                      s3 = s[c * 32 + t + 3];
          if (track) {
            m0 = fmaxf(m0, s0); m1 = fmaxf(m1, s1); m2 = fmaxf(m2, s2); m3 = fmaxf(m3, s3);
            if (c == 1 || c == 2) {
              m0 = fmaxf(m0, s[96 + (c - 1) * 16 + t / 2]);
              m1 = fmaxf(m1, s[96 + (c - 1) * 16 + t / 2 + 1]);
            }
          }
          float x0, x1, x2, x3;
          unpack_f32x2(ffma2(pack_f32x2(s0, s1), scale2, negm2), x0, x1);
          unpack_f32x2(ffma2(pack_f32x2(s2, s3), scale2, negm2), x2, x3);
          const float e0 = fast_exp2(x0), e1 = fast_exp2(x1), e2 = fast_exp2(x2), e3 = fast_exp2(x3);
          acc_a = fadd2(acc_a, pack_f32x2(e0, e1));
          acc_b = fadd2(acc_b, pack_f32x2(e2, e3));
          pk[t / 2] = pack_bf16x2(e0, e1);
          pk[t / 2 + 1] = pack_bf16x2(e2, e3);
        }
        tmem_st16(tS + c * 16, pk);
      };
      auto publish = [&](uint64_t* bar) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
      };

      bool redo;
      if (j > 0) {
        const float neg_m = -m_used * scale;
        const uint64_t negm2 = pack_f32x2(neg_m, neg_m);
        exp_chunk(0, negm2, true);
        wait_rest();
        exp_chunk(1, negm2, true);
        exp_chunk(2, negm2, true);
        const float m_tile = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        // lazy rescale: only when some row of this warp grew its max by more than 2^8
        redo = __any_sync(0xffffffffu, (m_tile - m_used) * scale > 8.0f);
        OVB_TRACE(j >= 64 && j < 72 && quad == OVB_TRACE_QUAD && lane == 0, (j - 64) * 32 + 8 + wg * 8 + 2);  // 3/4 done
        if (!redo) {
          // Publish keys [0,96).  Chunk 3's shift is made to depend on the arrival's token (masked
          // with a run-time zero): without that, ptxas hoists all of chunk 3's FFMA2/MUFU above
          // the arrive and the early publication is not early at all (seen in the SASS and in
          // the pipeline trace: "P part seen" coincided with "exps issued").
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          const uint32_t tok = mbar_arrive_token(&p_part[wg], lane == 0);
          const float neg_m3 = __uint_as_float(__float_as_uint(neg_m) ^ (tok & rt_zero));
          exp_chunk(3, pack_f32x2(neg_m3, neg_m3), false);
        }
      } else {
        wait_rest();
#pragma unroll
        for (int t = 0; t < 128; t += 4) {
          m0 = fmaxf(m0, s[t]);
          m1 = fmaxf(m1, s[t + 1]);
          m2 = fmaxf(m2, s[t + 2]);
          m3 = fmaxf(m3, s[t + 3]);
        }
        redo = true;
      }
      if (redo) {
        const float m_new = fmaxf(m_used, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        const float alpha = fast_exp2((m_used - m_new) * scale);  // 0 on the first tile
        l_sum *= alpha;
        m_used = m_new;
        if (j > 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld32(tO + c * 32, r);
            tmem_wait_ld();
            const uint64_t alpha2 = pack_f32x2(alpha, alpha);
#pragma unroll
            for (int t = 0; t < 32; t += 2) {
              float y0, y1;
              unpack_f32x2(fmul2(pack_f32x2(__uint_as_float(r[t]), __uint_as_float(r[t + 1])), alpha2), y0, y1);
              r[t] = __float_as_uint(y0);
              r[t + 1] = __float_as_uint(y1);
            }
            tmem_st32(tO + c * 32, r);
          }
        }
        acc_a = 0ull;
        acc_b = 0ull;
        const float neg_m = -m_used * scale;
        const uint64_t negm2 = pack_f32x2(neg_m, neg_m);
#pragma unroll
        for (int c = 0; c < 4; ++c) exp_chunk(c, negm2, false);
        publish(&p_part[wg]);
      }
      float a0, a1, a2, a3;
      unpack_f32x2(acc_a, a0, a1);
      unpack_f32x2(acc_b, a2, a3);
      l_sum += (a0 + a1) + (a2 + a3);
      OVB_TRACE(j >= 64 && j < 72 && quad == OVB_TRACE_QUAD && lane == 0, (j - 64) * 32 + 8 + wg * 8 + 3);  // exps issued
      publish(&p_full[wg]);
      OVB_TRACE(j >= 64 && j < 72 && quad == OVB_TRACE_QUAD && lane == 0, (j - 64) * 32 + 8 + wg * 8 + 4);  // P published
    };
    const int n_full = max(0, min(n_kv, (k_len >> 7) - kv0));  // this CTA's tiles with all 128 keys valid
#pragma unroll 1
    for (int j = 0; j < n_full; ++j) process_tile(j, std::false_type{});
    if (n_full < n_kv) process_tile(n_full, std::true_type{});

    // ---- epilogue
    const bool row_ok = q_row < p.Lq;
    if (slot >= 0) {
      // key-split CTA: unnormalised O, running max (log2 domain) and row sum -> workspace
      const long long prow = static_cast<long long>(slot) * 256 + wg * 128 + quad * 32 + lane;
      float* wo = p.ws_o + prow * 128;
      if (n_kv > 0) {
        mbar_wait(&o_full[wg], 0);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(tO + c * 32, r);
          tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<uint4*>(wo + c * 32 + g * 4) =
                  make_uint4(r[g * 4], r[g * 4 + 1], r[g * 4 + 2], r[g * 4 + 3]);
          }
        }
      }
      // a slice without keys contributes nothing: row sum 0 tells the combine pass to skip it
      if (row_ok) p.ws_ml[prow] = n_kv > 0 ? make_float2(m_used * scale, l_sum) : make_float2(-INFINITY, 0.f);
    } else {
      // O / l -> bf16 -> global (one 256-byte row segment per thread)
      __nv_bfloat16* orow = row_ok ? out_row_ptr(p, batch, q_row, head) : nullptr;
      if (n_kv > 0) {
        mbar_wait(&o_full[wg], 0);
        tc_fence_after();
        const float inv_l = 1.0f / l_sum;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(tO + c * 32, r);
          tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]) * inv_l, __uint_as_float(r[g * 8 + 1]) * inv_l);
              o.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]) * inv_l, __uint_as_float(r[g * 8 + 3]) * inv_l);
              o.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]) * inv_l, __uint_as_float(r[g * 8 + 5]) * inv_l);
              o.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]) * inv_l, __uint_as_float(r[g * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = o;
            }
          }
        }
      } else if (row_ok) {
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int g = 0; g < 16; ++g) *reinterpret_cast<uint4*>(orow + g * 8) = z;
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

// Merges the key-split partials of the last wave (see AttnParams): for every row of every tail item
//   out = sum_p O_p 2^(m_p - m) / sum_p l_p 2^(m_p - m),   m = max_p m_p   (m in the log2 domain).
// One thread per 8 output columns; 16 threads per row.
__global__ void __launch_bounds__(256)
attn_combine_kernel(const AttnParams p, int n_tail_items) {
  const int row_in_grid = blockIdx.x * 16 + (threadIdx.x >> 4);
  const int col = (threadIdx.x & 15) * 8;
  const int t = row_in_grid >> 8;  // tail item
  const int r = row_in_grid & 255;
  if (t >= n_tail_items) return;
  const int item = p.n_whole + t;
  const int q_blk = item % p.nqb;
  const int head = (item / p.nqb) % p.H;
  const int batch = item / (p.nqb * p.H);
  const int q_row = q_blk * 256 + r;
  if (q_row >= p.Lq) return;
  float m = -INFINITY;
  for (int k = 0; k < p.parts; ++k) {
    const float2 ml = p.ws_ml[static_cast<long long>(t * p.parts + k) * 256 + r];
    if (ml.y > 0.f) m = fmaxf(m, ml.x);
  }
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float l = 0.f;
  for (int k = 0; k < p.parts; ++k) {
    const long long prow = static_cast<long long>(t * p.parts + k) * 256 + r;
    const float2 ml = p.ws_ml[prow];
    if (!(ml.y > 0.f)) continue;
    const float w = fast_exp2(ml.x - m);
    l += ml.y * w;
    const float4 a = *reinterpret_cast<const float4*>(p.ws_o + prow * 128 + col);
    const float4 b = *reinterpret_cast<const float4*>(p.ws_o + prow * 128 + col + 4);
    acc[0] += a.x * w; acc[1] += a.y * w; acc[2] += a.z * w; acc[3] += a.w * w;
    acc[4] += b.x * w; acc[5] += b.y * w; acc[6] += b.z * w; acc[7] += b.w * w;
  }
  const float inv_l = l > 0.f ? 1.0f / l : 0.f;
  uint4 o;
  o.x = pack_bf16x2(acc[0] * inv_l, acc[1] * inv_l);
  o.y = pack_bf16x2(acc[2] * inv_l, acc[3] * inv_l);
  o.z = pack_bf16x2(acc[4] * inv_l, acc[5] * inv_l);
  o.w = pack_bf16x2(acc[6] * inv_l, acc[7] * inv_l);
  *reinterpret_cast<uint4*>(out_row_ptr(p, batch, q_row, head) + col) = o;
}

}  // namespace ovb

#ifdef OVB_ATTN_TRACE
extern "C" int ovb_attn_set_trace(long long* dev_buf) {
  return cudaMemcpyToSymbol(ovb::g_attn_trace, &dev_buf, sizeof(dev_buf)) == cudaSuccess ? 0 : -2;
}
#endif

namespace ovb {

struct AttnSplit {
  int n_whole;  // items computed by one CTA each
  int n_tail;   // items of the last, partial wave, cut `parts` ways along the keys
  int parts;
};

// SM count of the CURRENT device (a process may drive several GPUs)
static int sm_count() { return device_sm_count(); }

// One CTA per SM (227 KB of shared memory, all 512 TMEM columns), equal-cost items: T items take
// ceil(T / SMs) rounds.  When the last round is mostly empty, its r items are cut g ways along the
// keys: ceil(r g / SMs) / g rounds instead of 1.  Only done when a workspace was passed, every
// slice keeps >= 8 KV tiles, and at least 4 % of the whole launch is saved.
static AttnSplit plan_split(const ovb_attn_args* a, int min_k_len) {
  const int nqb = (a->Lq + 255) / 256;
  const long long T = static_cast<long long>(nqb) * a->H * a->B;
  const int W = sm_count();
  AttnSplit sp{static_cast<int>(T), 0, 1};
  const int r = static_cast<int>(T % W);
  if (a->workspace == nullptr || r == 0 || a->no_split) return sp;
  const int n_kv = (min_k_len + 127) / 128;
  const double whole_rounds = static_cast<double>(T / W);
  double best = 1.0;
  int best_g = 1;
  for (int g = 2; g <= 8 && n_kv / g >= 8; ++g) {
    const double tail = static_cast<double>((static_cast<long long>(r) * g + W - 1) / W) / g + 0.02 * g;
    if (tail < best) { best = tail; best_g = g; }
  }
  if (best_g == 1 || (1.0 - best) < 0.04 * (whole_rounds + 1.0)) return sp;
  const long long need = static_cast<long long>(r) * best_g * 256 * (128 * 4 + 8);
  if (a->workspace_bytes < need) return sp;
  sp.n_whole = static_cast<int>(T - r);
  sp.n_tail = r;
  sp.parts = best_g;
  return sp;
}

}  // namespace ovb

extern "C" int ovb_attn_plan(const ovb_attn_args* a, int32_t* n_whole, int32_t* n_tail, int32_t* parts) {
  using namespace ovb;
  OVB_CHECK_ARG(a != nullptr && n_whole && n_tail && parts, "ovb_attn_plan: null pointer");
  OVB_CHECK_ARG(a->B > 0 && a->B <= 8 && a->H > 0 && a->Lq > 0 && a->Lk > 0, "ovb_attn_plan: bad shape");
  int min_kl = a->Lk;
  for (int b = 0; b < a->B; ++b) {
    const int kl = a->k_lens != nullptr ? a->k_lens[b] : a->Lk;
    if (kl < min_kl) min_kl = kl;
  }
  const AttnSplit sp = plan_split(a, min_kl);
  *n_whole = sp.n_whole;
  *n_tail = sp.n_tail;
  *parts = sp.parts;
  return OVB_OK;
}

extern "C" int64_t ovb_attn_workspace_size(const ovb_attn_args* a) {
  if (a == nullptr || a->Lq <= 0 || a->H <= 0 || a->B <= 0) return 0;
  const long long T = static_cast<long long>((a->Lq + 255) / 256) * a->H * a->B;
  const long long r = T % ovb::sm_count();
  return r * 8 * 256 * (128 * 4 + 8);  // worst case: 8-way split of the last wave
}

extern "C" int ovb_attn_fwd(const ovb_attn_args* a, void* stream) {
  using namespace ovb;
  OVB_CHECK_ARG(a != nullptr, "ovb_attn_fwd: null args");
  OVB_CHECK_ARG(a->q && a->k && a->v && (a->out || a->out_ptrs), "ovb_attn_fwd: null q/k/v/out");
  OVB_CHECK_ARG(a->B > 0 && a->H > 0 && a->Lq > 0 && a->Lk > 0,
                "ovb_attn_fwd: bad shape B=%d H=%d Lq=%d Lk=%d", a->B, a->H, a->Lq, a->Lk);
  OVB_CHECK_ARG(a->B <= 8, "ovb_attn_fwd: B <= 8 supported (got %d)", a->B);
  OVB_CHECK_ARG(a->q_row_stride % 8 == 0 && a->k_row_stride % 8 == 0 && a->v_row_stride % 8 == 0 &&
                    a->o_row_stride % 8 == 0 && a->q_batch_stride % 8 == 0 &&
                    a->k_batch_stride % 8 == 0 && a->v_batch_stride % 8 == 0 &&
                    a->o_batch_stride % 8 == 0,
                "ovb_attn_fwd: strides must be multiples of 8 elements");
  OVB_CHECK_ARG((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) |
                 reinterpret_cast<uintptr_t>(a->v) | reinterpret_cast<uintptr_t>(a->out)) % 16 == 0,
                "ovb_attn_fwd: pointers must be 16-byte aligned");

  CUtensorMap tmQ, tmK, tmV;
  const uint32_t box[3] = {64, 128, 1};
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(a->H) * 128, static_cast<uint64_t>(a->Lq),
                              static_cast<uint64_t>(a->B)};
    const uint64_t strides[2] = {static_cast<uint64_t>(a->q_row_stride) * 2,
                                 static_cast<uint64_t>(a->B > 1 ? a->q_batch_stride
                                                                : a->q_row_stride * a->Lq) * 2};
    int rc = encode_tmap_bf16(&tmQ, a->q, 3, dims, strides, box);
    if (rc != OVB_OK) return rc;
  }
  {
    // The K/V maps end at the longest valid key length: rows past it are never fetched (TMA zero
    // fill), as in the reference, which trims k/v to k_lens before calling flash-attn
    // (attention.py:75-77).  (Inside a ragged batch, masked rows of shorter samples are still read
    // and multiplied by P = 0, so they must hold finite values.)
    int max_kl = 1;
    for (int b = 0; b < a->B; ++b) {
      const int kl = a->k_lens != nullptr ? a->k_lens[b] : a->Lk;
      if (kl > max_kl) max_kl = kl;
    }
    if (max_kl > a->Lk) max_kl = a->Lk;
    const uint64_t dims[3] = {static_cast<uint64_t>(a->H) * 128, static_cast<uint64_t>(max_kl),
                              static_cast<uint64_t>(a->B)};
    const uint64_t ks[2] = {static_cast<uint64_t>(a->k_row_stride) * 2,
                            static_cast<uint64_t>(a->B > 1 ? a->k_batch_stride
                                                           : a->k_row_stride * a->Lk) * 2};
    int rc = encode_tmap_bf16(&tmK, a->k, 3, dims, ks, box);
    if (rc != OVB_OK) return rc;
    const uint64_t vs[2] = {static_cast<uint64_t>(a->v_row_stride) * 2,
                            static_cast<uint64_t>(a->B > 1 ? a->v_batch_stride
                                                           : a->v_row_stride * a->Lk) * 2};
    rc = encode_tmap_bf16(&tmV, a->v, 3, dims, vs, box);
    if (rc != OVB_OK) return rc;
  }

  AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.o_row_stride = a->o_row_stride;
  p.o_batch_stride = a->o_batch_stride;
  p.Lq = a->Lq;
  p.Lk = a->Lk;
  p.H = a->H;
  const float scale = a->softmax_scale > 0.f ? a->softmax_scale : 0.08838834764831845f;
  p.scale_log2 = scale * 1.4426950408889634f;
  int kls[8];
  int min_kl = a->Lk;
  for (int b = 0; b < 8; ++b) {
    int kl = a->Lk;
    if (a->k_lens != nullptr && b < a->B) kl = a->k_lens[b];
    OVB_CHECK_ARG(kl >= 0 && kl <= a->Lk, "ovb_attn_fwd: k_lens[%d]=%d outside [0, %d]", b, kl,
                  a->Lk);
    kls[b] = kl;
    if (b < a->B && kl < min_kl) min_kl = kl;
  }
  p.kl0 = kls[0]; p.kl1 = kls[1]; p.kl2 = kls[2]; p.kl3 = kls[3];
  p.kl4 = kls[4]; p.kl5 = kls[5]; p.kl6 = kls[6]; p.kl7 = kls[7];

  p.rows_per_out = 0;
  for (int g = 0; g < 8; ++g) p.outp[g] = nullptr;
  if (a->out_ptrs != nullptr) {
    OVB_CHECK_ARG(a->B == 1 && a->rows_per_out > 0 && a->n_out_ptrs >= 1 && a->n_out_ptrs <= 8 &&
                      static_cast<long long>(a->rows_per_out) * a->n_out_ptrs >= a->Lq,
                  "ovb_attn_fwd: scatter form needs B == 1 and n_out_ptrs * rows_per_out >= Lq");
    for (int g = 0; g < a->n_out_ptrs; ++g) {
      OVB_CHECK_ARG(a->out_ptrs[g] != nullptr && reinterpret_cast<uintptr_t>(a->out_ptrs[g]) % 16 == 0,
                    "ovb_attn_fwd: out_ptrs[%d] null or misaligned", g);
      p.outp[g] = reinterpret_cast<__nv_bfloat16*>(a->out_ptrs[g]);
    }
    p.rows_per_out = a->rows_per_out;
  }
  const AttnSplit sp = plan_split(a, min_kl);
  p.nqb = (a->Lq + 255) / 256;
  p.n_whole = sp.n_whole;
  p.parts = sp.parts;
  p.ws_o = nullptr;
  p.ws_ml = nullptr;
  if (sp.n_tail > 0) {
    // workspace layout: [n_tail * parts][256][128] fp32 partial O, then [n_tail * parts][256] float2
    const long long ctas = static_cast<long long>(sp.n_tail) * sp.parts;
    p.ws_o = reinterpret_cast<float*>(a->workspace);
    p.ws_ml = reinterpret_cast<float2*>(p.ws_o + ctas * 256 * 128);
  }

  static bool attr_set[64] = {};
  if (once_per_device(attr_set))
    OVB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kAttnSmem));
  const int grid = sp.n_whole + sp.n_tail * sp.parts;
  attn_fwd_kernel<<<grid, kAttnThreads, kAttnSmem, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  if (sp.n_tail > 0) {
    attn_combine_kernel<<<sp.n_tail * 16, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, sp.n_tail);
    OVB_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  return OVB_OK;
}
