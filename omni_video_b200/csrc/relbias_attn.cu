// Attention with an additive relative-position bias, all heads and samples in ONE launch: the umT5 encoder's
// self-attention (reference omnivideo/modules/t5.py:96-117: scores = q k^T (no 1/sqrt(d)) + position bias
// + key mask, softmax in fp32, P V).  The bias of (query i, key j) depends only on j - i (t5.py:241-263), so the
// caller passes one fp32 vector of 2L-1 values per head instead of a [heads, L, L] tensor.
//
// This is the text encoder: 0.1 TFLOP of attention per 512-token prompt, once per sample, against ~28 PFLOP of
// denoising.  It is written for launch count and exactness of the rounding points (fp32 scores, fp32 softmax,
// bf16 P, fp32 accumulation), not for the tensor-core roofline: warp-level `mma.sync` m16n8k16 on bf16 with a
// flash-style running softmax, 64 query rows per CTA (16 per warp), K and V^T tiles of 64 keys staged in shared
// memory.  The DiT's attention (attn.cu) is the tcgen05 kernel; its 128-wide head tiles do not fit umT5's 64-wide
// heads and it has no bias operand.
#include <cuda_bf16.h>

#include "host_util.h"

namespace ovb {
namespace {

constexpr int RB_ROWS = 64;   // query rows per CTA
constexpr int RB_KEYS = 64;   // keys per staged tile
constexpr int RB_PAD = 8;     // bf16 elements of row padding: B-fragment loads hit 32 distinct banks

struct RelBiasAttnParams {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* out;
  const float* rel_bias;      // [heads][2L-1]
  const int32_t* k_lens;      // [B] or nullptr
  long long ldq, ldk, ldv, ldo;                                // row pitches (elements)
  long long q_batch, k_batch, v_batch, o_batch, bias_head;     // sample / head strides (elements)
  int L, heads;
  float scale_log2;           // softmax_scale * log2(e)
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);   // .x (low half) = lo
  return *reinterpret_cast<uint32_t*>(&t);
}

// Fragment coordinates of mma.m16n8k16 (g = lane / 4, t = lane % 4):
//   A (16x16, row): a0 = (g, 2t..2t+1)  a1 = (g+8, 2t..)  a2 = (g, 2t+8..)  a3 = (g+8, 2t+8..)
//   B (16x8,  col): b0 = (k = 2t..2t+1, n = g)            b1 = (k = 2t+8.., n = g)
//   C (16x8):       c0,c1 = (g, 2t..2t+1)                 c2,c3 = (g+8, 2t..2t+1)
// A C tile pair (n-tiles 2j, 2j+1) is therefore exactly the A fragment of key step j: P never leaves registers.
template <int HD>
__global__ void __launch_bounds__(128) relbias_attn_kernel(const RelBiasAttnParams p) {
  constexpr int KS = HD / 16;          // k-steps of Q K^T
  constexpr int NT = HD / 8;           // n-tiles of P V
  constexpr int VEC = HD / 8;          // 16-byte vectors per row of a head
  __shared__ __align__(16) __nv_bfloat16 Ks[RB_KEYS][HD + RB_PAD];     // [key][d]
  __shared__ __align__(16) __nv_bfloat16 Vt[HD][RB_KEYS + RB_PAD];     // [d][key]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y, b = blockIdx.z;
  const int L = p.L;
  int n_keys = p.k_lens ? p.k_lens[b] : L;
  n_keys = n_keys < 1 ? 1 : (n_keys > L ? L : n_keys);
  const int row0 = blockIdx.x * RB_ROWS + warp * 16 + g, row1 = row0 + 8;

  const __nv_bfloat16* q = p.q + b * p.q_batch + head * HD;
  const __nv_bfloat16* kk = p.k + b * p.k_batch + head * HD;
  const __nv_bfloat16* vv = p.v + b * p.v_batch + head * HD;
  const float* bias = p.rel_bias + head * p.bias_head + (L - 1);      // bias[j - i]

  uint32_t qa[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int c = ks * 16 + 2 * t;
    qa[ks][0] = row0 < L ? *reinterpret_cast<const uint32_t*>(q + row0 * p.ldq + c) : 0u;
    qa[ks][1] = row1 < L ? *reinterpret_cast<const uint32_t*>(q + row1 * p.ldq + c) : 0u;
    qa[ks][2] = row0 < L ? *reinterpret_cast<const uint32_t*>(q + row0 * p.ldq + c + 8) : 0u;
    qa[ks][3] = row1 < L ? *reinterpret_cast<const uint32_t*>(q + row1 * p.ldq + c + 8) : 0u;
  }
  // rows past the end compute on zeros and are never stored; their bias index is clamped into the table
  const int brow0 = row0 < L ? row0 : L - 1, brow1 = row1 < L ? row1 : L - 1;

  float o[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  const int n_tiles = (n_keys + RB_KEYS - 1) / RB_KEYS;
  for (int kt = 0; kt < n_tiles; ++kt) {
    const int key0 = kt * RB_KEYS;
    __syncthreads();                                   // the previous tile has been consumed
    for (int i = threadIdx.x; i < RB_KEYS * VEC; i += 128) {
      const int r = i / VEC, cv = i % VEC, key = key0 + r;
      uint4 kx = make_uint4(0u, 0u, 0u, 0u), vx = kx;
      if (key < L) {
        kx = *reinterpret_cast<const uint4*>(kk + key * p.ldk + cv * 8);
        vx = *reinterpret_cast<const uint4*>(vv + key * p.ldv + cv * 8);
      }
      *reinterpret_cast<uint4*>(&Ks[r][cv * 8]) = kx;
      const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&vx);
#pragma unroll
      for (int e = 0; e < 8; ++e) Vt[cv * 8 + e][r] = ve[e];
    }
    __syncthreads();

    // ---- S = Q K^T for 16 rows x 64 keys
    float s[RB_KEYS / 8][4];
#pragma unroll
    for (int n = 0; n < RB_KEYS / 8; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ks[n * 8 + g][ks * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Ks[n * 8 + g][ks * 16 + 2 * t + 8]);
        mma_bf16_16816(s[n], qa[ks], b0, b1);
      }
    }
    // ---- log2-domain scores: (q.k * scale + bias) * log2(e); keys past the sample's length -> -inf
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < RB_KEYS / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = key0 + n * 8 + 2 * t + e;
        if (j < n_keys) {
          s[n][e] = fmaf(s[n][e], p.scale_log2, __ldg(bias + (j - brow0)) * 1.4426950408889634f);
          s[n][2 + e] = fmaf(s[n][2 + e], p.scale_log2, __ldg(bias + (j - brow1)) * 1.4426950408889634f);
        } else {
          s[n][e] = -INFINITY;
          s[n][2 + e] = -INFINITY;
        }
        mx0 = fmaxf(mx0, s[n][e]);
        mx1 = fmaxf(mx1, s[n][2 + e]);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    // every tile holds at least one valid key (n_tiles is cut at n_keys), so the new maxima are finite
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);          // exp2(-inf) = 0 on the first tile
    m0 = mn0;
    m1 = mn1;
    l0 *= c0;
    l1 *= c1;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1;
    }
    uint32_t pa[RB_KEYS / 16][4];
#pragma unroll
    for (int n = 0; n < RB_KEYS / 8; ++n) {
      const float p00 = exp2f(s[n][0] - m0), p01 = exp2f(s[n][1] - m0);
      const float p10 = exp2f(s[n][2] - m1), p11 = exp2f(s[n][3] - m1);
      l0 += p00 + p01;
      l1 += p10 + p11;
      pa[n >> 1][(n & 1) * 2 + 0] = pack_bf16(p00, p01);
      pa[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p10, p11);
    }
    // ---- O += P V
#pragma unroll
    for (int j = 0; j < RB_KEYS / 16; ++j) {
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Vt[n * 8 + g][j * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Vt[n * 8 + g][j * 16 + 2 * t + 8]);
        mma_bf16_16816(o[n], pa[j], b0, b1);
      }
    }
  }

  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* out = p.out + b * p.o_batch + head * HD;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const int c = n * 8 + 2 * t;
    if (row0 < L) *reinterpret_cast<uint32_t*>(out + row0 * p.ldo + c) = pack_bf16(o[n][0] * i0, o[n][1] * i0);
    if (row1 < L) *reinterpret_cast<uint32_t*>(out + row1 * p.ldo + c) = pack_bf16(o[n][2] * i1, o[n][3] * i1);
  }
}

}  // namespace
}  // namespace ovb

extern "C" int ovb_attn_relbias_fwd(const ovb_attn_relbias_args* a, void* stream) {
  using namespace ovb;
  OVB_CHECK_ARG(a != nullptr && a->q && a->k && a->v && a->out && a->rel_bias, "ovb_attn_relbias_fwd: null pointer");
  OVB_CHECK_ARG(a->B > 0 && a->H > 0 && a->L > 0 && a->B <= 65535 && a->H <= 65535,
                "ovb_attn_relbias_fwd: bad extents B=%d H=%d L=%d", a->B, a->H, a->L);
  OVB_CHECK_ARG(a->head_dim == 64 || a->head_dim == 128, "ovb_attn_relbias_fwd: head_dim=%d (64 or 128)",
                a->head_dim);
  const long long width = static_cast<long long>(a->H) * a->head_dim;
  OVB_CHECK_ARG(a->q_row_stride >= width && a->k_row_stride >= width && a->v_row_stride >= width &&
                    a->o_row_stride >= width,
                "ovb_attn_relbias_fwd: a row pitch is smaller than H*head_dim=%lld", width);
  OVB_CHECK_ARG(a->q_row_stride % 8 == 0 && a->k_row_stride % 8 == 0 && a->v_row_stride % 8 == 0 &&
                    a->o_row_stride % 8 == 0 && a->q_batch_stride % 8 == 0 && a->k_batch_stride % 8 == 0 &&
                    a->v_batch_stride % 8 == 0 && a->o_batch_stride % 8 == 0,
                "ovb_attn_relbias_fwd: pitches must be multiples of 8 elements");
  OVB_CHECK_ARG((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) |
                 reinterpret_cast<uintptr_t>(a->v) | reinterpret_cast<uintptr_t>(a->out)) % 16 == 0,
                "ovb_attn_relbias_fwd: q/k/v/out must be 16-byte aligned");
  OVB_CHECK_ARG(a->bias_head_stride >= 2LL * a->L - 1, "ovb_attn_relbias_fwd: bias_head_stride=%lld < 2L-1",
                static_cast<long long>(a->bias_head_stride));
  RelBiasAttnParams p;
  p.q = static_cast<const __nv_bfloat16*>(a->q);
  p.k = static_cast<const __nv_bfloat16*>(a->k);
  p.v = static_cast<const __nv_bfloat16*>(a->v);
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.rel_bias = a->rel_bias;
  p.k_lens = a->k_lens;
  p.ldq = a->q_row_stride; p.ldk = a->k_row_stride; p.ldv = a->v_row_stride; p.ldo = a->o_row_stride;
  p.q_batch = a->q_batch_stride; p.k_batch = a->k_batch_stride; p.v_batch = a->v_batch_stride;
  p.o_batch = a->o_batch_stride;
  p.bias_head = a->bias_head_stride;
  p.L = a->L;
  p.heads = a->H;
  p.scale_log2 = a->softmax_scale * 1.4426950408889634f;
  const dim3 grid((a->L + RB_ROWS - 1) / RB_ROWS, a->H, a->B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->head_dim == 64) relbias_attn_kernel<64><<<grid, 128, 0, st>>>(p);
  else relbias_attn_kernel<128><<<grid, 128, 0, st>>>(p);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}
