// HBM-bound kernels of the DiT hot path: LayerNorm + AdaLN modulation, full-width RMSNorm + 3-D RoPE,
// patchify / unpatchify, sinusoidal timestep embedding, casts.  All use 16-byte vector accesses,
// one CTA per row (d/8 threads, one 8-element vector per thread) so a row is read exactly once;
// algorithmic bytes per element are listed at each kernel (DESIGN.md §4 / SURVEY.md §8(d)).
#include "host_util.h"
#include "ptx.cuh"

namespace ovb {

__device__ __forceinline__ float block_sum(float v, float* red) {
  // red: >= 32 floats of shared memory
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = (blockDim.x + 31) >> 5;
  __syncthreads();  // protect red from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = lane < nwarps ? red[lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 w = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}
__device__ __forceinline__ void load8_f32(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]);
  o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]);
  o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ void store8_f32(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm + modulation.  Bytes/element: 4 (fp32 x) + 2 (bf16 y) = 6 (+ L2-resident vectors).
struct LnParams {
  const void* x; long long ldx; int x_is_bf16;
  void* out; long long ldo; int out_is_f32;
  int rows, d;
  float eps;
  int mode;
  const void* a; const void* b;
  long long mod_row_stride, mod_sample_stride;
  int rows_per_sample;
  int round_ln_bf16;
};

__global__ void layernorm_kernel(const LnParams p) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const int c0 = threadIdx.x * 8;
  const bool active = c0 < p.d;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (active) {
    if (p.x_is_bf16)
      load8_bf16(reinterpret_cast<const __nv_bfloat16*>(p.x) + row * p.ldx + c0, v);
    else
      load8_f32(reinterpret_cast<const float*>(p.x) + row * p.ldx + c0, v);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  const float mean = block_sum(s, red) / p.d;
  float q = 0.f;
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = v[j] - mean;
      q += t * t;
    }
  }
  const float var = block_sum(q, red) / p.d;
  const float rstd = rsqrtf(var + p.eps);
  if (!active) return;

  float y[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    y[j] = (v[j] - mean) * rstd;
    if (p.round_ln_bf16) y[j] = bf16_round(y[j]);
  }
  if (p.mode == OVB_LN_MODULATE) {
    const int sample = row / p.rows_per_sample;
    const int r_in = row - sample * p.rows_per_sample;
    const long long off = sample * p.mod_sample_stride + r_in * p.mod_row_stride + c0;
    float sc[8], sh[8];
    load8_f32(reinterpret_cast<const float*>(p.a) + off, sc);
    load8_f32(reinterpret_cast<const float*>(p.b) + off, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = y[j] * (1.0f + sc[j]) + sh[j];
  } else {
    float w[8], bb[8];
    load8_bf16(reinterpret_cast<const __nv_bfloat16*>(p.a) + c0, w);
    load8_bf16(reinterpret_cast<const __nv_bfloat16*>(p.b) + c0, bb);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = y[j] * w[j] + bb[j];
  }
  if (p.out_is_f32)
    store8_f32(reinterpret_cast<float*>(p.out) + row * p.ldo + c0, y);
  else
    store8_bf16(reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + c0, y);
}


__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One WARP per row, the row held in registers (VPL float4 vectors per lane, lane l owns vectors
// l, l+32, ...: every warp-wide load is one contiguous 512-byte run).  No shared memory and no
// __syncthreads: the two reductions are five shuffles each.  The launch is persistent: one grid-stride
// pass of sm_count x resident CTAs (two per SM up to d = 2048).  Occupancy (two, three or four CTAs per SM) and the
// vector layout were measured within 3 % of each other at d = 1536 (scripts/norm_time.py): what the kernel was
// short of was issue slots, see FAST below.
// (The one-CTA-per-row form above spends most of its time in the two block barriers: 0.59 / 0.36 of copy
// bandwidth for LayerNorm / RMSNorm at d = 1536.)
__device__ __forceinline__ float4 load4_bf16(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}

// FAST: the shapes the DiT step runs (fp32 residual stream in, bf16 out, d == VPL * 128 so no lane is ever out of
// range) with mode and rounding as template constants.  The general instantiation carries four run-time flags and a
// bounds check per vector — ~1200 SASS instructions per row and lane at d = 1536, 25 per element, which is what made
// a streaming kernel issue-bound (ncu: issue slots 54 % busy); the arithmetic and its order are the same.
template <int VPL, bool FAST = false, int MODE_T = 0, bool ROUND_T = false>
__global__ void __launch_bounds__(256, VPL <= 16 ? 2 : 1) layernorm_warp_kernel(const LnParams p) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * 8;
  const bool x_bf16 = FAST ? false : p.x_is_bf16 != 0;
  const bool out_f32 = FAST ? false : p.out_is_f32 != 0;
  const bool modulate = FAST ? MODE_T == OVB_LN_MODULATE : p.mode == OVB_LN_MODULATE;
  const bool round_ln = FAST ? ROUND_T : p.round_ln_bf16 != 0;
  const int d = FAST ? VPL * 128 : p.d;
  const float inv_d = 1.0f / static_cast<float>(d);
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < p.rows; row += warps) {
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c0 = (i * 32 + lane) * 4;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (FAST || c0 < d) {
        if (x_bf16)
          v[i] = load4_bf16(reinterpret_cast<const __nv_bfloat16*>(p.x) + row * p.ldx + c0);
        else
          v[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.x) + row * p.ldx + c0);
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (FAST || (i * 32 + lane) * 4 < d) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
        q += (a * a + b * b) + (c * c + e * e);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + p.eps);
    const int sample = row / p.rows_per_sample;
    const long long mod_off = sample * p.mod_sample_stride + (row - sample * p.rows_per_sample) * p.mod_row_stride;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c0 = (i * 32 + lane) * 4;
      if (!FAST && c0 >= d) continue;
      float y[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd};
      if (round_ln) {
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = bf16_round(y[j]);
      }
      float4 a, b;
      if (modulate) {
        a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.a) + mod_off + c0);
        b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.b) + mod_off + c0);
        a.x += 1.0f; a.y += 1.0f; a.z += 1.0f; a.w += 1.0f;
      } else {
        a = load4_bf16(reinterpret_cast<const __nv_bfloat16*>(p.a) + c0);
        b = load4_bf16(reinterpret_cast<const __nv_bfloat16*>(p.b) + c0);
      }
      y[0] = y[0] * a.x + b.x; y[1] = y[1] * a.y + b.y; y[2] = y[2] * a.z + b.z; y[3] = y[3] * a.w + b.w;
      if (out_f32)
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * p.ldo + c0) = make_float4(y[0], y[1], y[2], y[3]);
      else
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + c0) =
            make_uint2(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// RMSNorm over full d + RoPE, in place on bf16.  Bytes/element: 2 + 2 = 4.
struct RmsRopeParams {
  __nv_bfloat16* x; long long ldx;
  int rows, d;
  const __nv_bfloat16* weight;
  float eps;
  int rope;
  const float2* cs;  // [table_len][64] (cos, sin)
  int table_len;
  int F, H, W;
  int rows_per_sample, token_offset;
  int cols_per_split; long long col_split_stride;
  __nv_bfloat16* dst[8]; long long dst_ld; int scatter; int copy_only;  // scatter form (<= 8 column groups)
};

__global__ void rmsnorm_rope_kernel(const RmsRopeParams p) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const int c0 = threadIdx.x * 8;
  const bool active = c0 < p.d;
  __nv_bfloat16* ptr = nullptr;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (active) {
    const int seg = c0 / p.cols_per_split;
    ptr = p.x + seg * p.col_split_stride + row * p.ldx + (c0 - seg * p.cols_per_split);
    load8_bf16(ptr, v);
  }
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) q += v[j] * v[j];
  const float ms = block_sum(q, red) / p.d;
  const float rinv = rsqrtf(ms + p.eps);
  if (!active) return;
  float w[8];
  load8_bf16(p.weight + c0, w);
  float y[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) y[j] = bf16_round(bf16_round(v[j] * rinv) * w[j]);

  if (p.rope) {
    const int tok = p.token_offset + row % p.rows_per_sample;
    if (tok < p.F * p.H * p.W) {
      const int hw = p.H * p.W;
      const int f = tok / hw;
      const int rem = tok - f * hw;
      const int h = rem / p.W;
      const int w_ = rem - h * p.W;
      const int pair0 = (c0 & 127) >> 1;  // first of 4 consecutive pairs inside the head
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pr = pair0 + j;
        const int pos = pr < 22 ? f : (pr < 43 ? h : w_);
        const float2 cs = p.cs[pos * 64 + pr];
        const float xe = y[2 * j], xo = y[2 * j + 1];
        y[2 * j] = xe * cs.x - xo * cs.y;
        y[2 * j + 1] = xe * cs.y + xo * cs.x;
      }
    }
  }
  store8_bf16(ptr, y);
}

// Warp-per-row form of the above (see layernorm_warp_kernel).  This kernel is ISSUE-bound, not
// HBM-bound, if written naively (~150 instructions per 16-byte vector), so:
//   * the row stays packed as bf16 in registers between the two passes (4 registers per vector);
//   * bf16(bf16(x * rinv) * w): the second product is one HMUL2.BF16 per pair — the product of two
//     bf16 values is exact in fp32, so its bf16 rounding is the reference's (`.type_as(x) * weight`);
//   * a lane's vectors all sit at the same offset inside a head (vector stride 32 lanes x 8 = 256
//     columns = 2 heads), so its four (cos, sin) pairs are fetched ONCE per row, not per vector.
//   * registers are bounded for three resident CTAs per SM up to d = 1536 (32-bit element offsets; the
//     per-vector destination pointers exist only in the SCATTER instantiation).
template <int VPL, bool SCATTER>
__global__ void __launch_bounds__(256, VPL <= 6 ? 3 : 1) rmsnorm_rope_warp_kernel(const RmsRopeParams p) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * 8;
  const int hw = p.H * p.W;
  const int n_tok = p.F * hw;
  const int pair0 = (lane & 15) * 4;  // ((i*32 + lane) * 8 & 127) / 2 for every i
  // element offset of each of this lane's vectors inside a row (column-split layout): row-invariant
  int off[VPL];               // fits 32 bits: the host checks (col_split - 1) * col_split_stride + d < 2^31
  __nv_bfloat16* dbase[SCATTER ? VPL : 1];  // scatter form: destination of this vector at row 0 (maybe peer memory)
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c0 = (i * 32 + lane) * 8;
    const int seg = c0 / p.cols_per_split;
    off[i] = static_cast<int>(seg * p.col_split_stride) + (c0 - seg * p.cols_per_split);
    if (SCATTER) dbase[i] = c0 < p.d ? p.dst[seg] + (c0 - seg * p.cols_per_split) : nullptr;
  }
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < p.rows; row += warps) {
    uint4 raw[VPL];
    float q = 0.f;
    __nv_bfloat16* xrow = p.x + row * p.ldx;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c0 = (i * 32 + lane) * 8;
      raw[i] = make_uint4(0, 0, 0, 0);
      if (c0 < p.d) raw[i] = *reinterpret_cast<const uint4*>(xrow + off[i]);
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = __uint_as_float(u[j] << 16), hi = __uint_as_float(u[j] & 0xffff0000u);
        q = fmaf(lo, lo, q);
        q = fmaf(hi, hi, q);
      }
    }
    if (SCATTER && p.copy_only) {  // plain scatter of the row (V): no statistics, no rounding
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if ((i * 32 + lane) * 8 < p.d) *reinterpret_cast<uint4*>(dbase[SCATTER ? i : 0] + row * p.dst_ld) = raw[i];
      continue;
    }
    const float rinv = rsqrtf(warp_sum(q) / p.d + p.eps);
    const int tok = p.token_offset + row % p.rows_per_sample;
    const bool rotate = p.rope && tok < n_tok;
    float2 cs[4];
    if (rotate) {
      const int f = tok / hw;
      const int rem = tok - f * hw;
      const int h_ = rem / p.W;
      const int w_ = rem - h_ * p.W;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pr = pair0 + j;
        cs[j] = p.cs[(pr < 22 ? f : (pr < 43 ? h_ : w_)) * 64 + pr];
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c0 = (i * 32 + lane) * 8;
      if (c0 >= p.d) continue;
      const uint4 wv = *reinterpret_cast<const uint4*>(p.weight + c0);
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&raw[i]);
      const uint32_t* wu = reinterpret_cast<const uint32_t*>(&wv);
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = __uint_as_float(u[j] << 16) * rinv, hi = __uint_as_float(u[j] & 0xffff0000u) * rinv;
        __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
        t = __hmul2(t, *reinterpret_cast<const __nv_bfloat162*>(&wu[j]));
        uint32_t y = *reinterpret_cast<uint32_t*>(&t);
        if (rotate) {
          const float xe = __uint_as_float(y << 16), xo = __uint_as_float(y & 0xffff0000u);
          y = pack_bf16x2(xe * cs[j].x - xo * cs[j].y, xe * cs[j].y + xo * cs[j].x);
        }
        o[j] = y;
      }
      __nv_bfloat16* dptr = SCATTER ? dbase[SCATTER ? i : 0] + row * p.dst_ld : xrow + off[i];
      *reinterpret_cast<uint4*>(dptr) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// patchify: x[C,F,H,W] -> tokens[F*(H/2)*(W/2), C*4] bf16, column (c*2+ph)*2+pw.
template <typename T>
__global__ void patchify_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ tok,
                                long long ldt, int C, int F, int H, int W) {
  const int Hp = H / 2, Wp = W / 2;
  const long long total = static_cast<long long>(F) * Hp * Wp * C * 2;  // one (c, ph) pair of pw=0,1
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // i enumerates (c, f, hp, ph, wp) with wp fastest so global reads are coalesced
    long long t = i;
    const int wp = t % Wp; t /= Wp;
    const int ph = t % 2; t /= 2;
    const int hp = t % Hp; t /= Hp;
    const int f = t % F; t /= F;
    const int c = static_cast<int>(t);
    const long long src = ((static_cast<long long>(c) * F + f) * H + (hp * 2 + ph)) * W + wp * 2;
    const float v0 = static_cast<float>(x[src]);
    const float v1 = static_cast<float>(x[src + 1]);
    const long long token = (static_cast<long long>(f) * Hp + hp) * Wp + wp;
    const uint32_t packed = pack_bf16x2(v0, v1);
    *reinterpret_cast<uint32_t*>(tok + token * ldt + (c * 2 + ph) * 2) = packed;
  }
}

// General patch gather: x[C,F,H,W] -> tokens[(F/pt)*(H/ph)*(W/pw), C*pt*ph*pw] bf16, column
// ((c*pt + it)*ph + ih)*pw + iw = the order of Conv3d.weight.flatten(1) (VisualContextAdapter's
// patch_embedding, reference modules/visual_context_adapter.py:56-58, :98).  One thread per output
// element; consecutive threads write consecutive columns of a token (reads are pw-element runs).
template <typename T>
__global__ void patchify3d_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ tok,
                                  long long ldt, int C, int F, int H, int W, int pt, int ph, int pw) {
  const int Fp = F / pt, Hp = H / ph, Wp = W / pw;
  const int cols = C * pt * ph * pw;
  const long long total = static_cast<long long>(Fp) * Hp * Wp * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long token = i / cols;
    int col = static_cast<int>(i - token * cols);
    const int iw = col % pw; col /= pw;
    const int ih = col % ph; col /= ph;
    const int it = col % pt;
    const int c = col / pt;
    long long t = token;
    const int wp = t % Wp; t /= Wp;
    const int hp = t % Hp; t /= Hp;
    const int fp = static_cast<int>(t);
    const long long src = ((static_cast<long long>(c) * F + (fp * pt + it)) * H + (hp * ph + ih)) * W +
                          (wp * pw + iw);
    tok[token * ldt + (i - token * cols)] = __float2bfloat16_rn(static_cast<float>(x[src]));
  }
}

// unpatchify: y[L, 4*C] fp32 (column (ph*2+pw)*C + c) -> out[C, F, 2*Hp, 2*Wp]
__global__ void unpatchify_kernel(const float* __restrict__ y, long long ldy,
                                  float* __restrict__ out, int C, int F, int Hp, int Wp) {
  const int H = Hp * 2, W = Wp * 2;
  const long long total = static_cast<long long>(C) * F * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long t = i;
    const int w = t % W; t /= W;
    const int h = t % H; t /= H;
    const int f = t % F; t /= F;
    const int c = static_cast<int>(t);
    const long long token = (static_cast<long long>(f) * Hp + (h >> 1)) * Wp + (w >> 1);
    out[i] = y[token * ldy + ((h & 1) * 2 + (w & 1)) * C + c];
  }
}

// sinusoidal embedding in fp64 (reference model.py:13-23): out[i, :half] = cos, [half:] = sin
__global__ void sinusoid_kernel(const float* __restrict__ t, float* __restrict__ out, int n,
                                int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x;
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    const double freq = pow(10000.0, -static_cast<double>(j) / half);
    const double ang = static_cast<double>(t[i]) * freq;
    out[static_cast<long long>(i) * dim + j] = static_cast<float>(cos(ang));
    out[static_cast<long long>(i) * dim + half + j] = static_cast<float>(sin(ang));
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ o,
                                     long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v[8];
    load8_f32(x + i * 8, v);
    store8_bf16(o + i * 8, v);
  }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ o,
                                     long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v[8];
    load8_bf16(x + i * 8, v);
    store8_f32(o + i * 8, v);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 linear on bf16-stored weights.  64x64 output tile per CTA, 256 threads, 4x4 per thread.
constexpr int kLT = 64, kLK = 32;
__device__ __forceinline__ float act_silu(float x) { return x / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ Wt,
                  long long ldw, const __nv_bfloat16* __restrict__ bias, float* __restrict__ out,
                  long long ldo, int M, int N, int K, int act_in, int act_out) {
  __shared__ float sx[kLK][kLT + 1];
  __shared__ float sw[kLK][kLT + 1];
  const int m0 = blockIdx.y * kLT, n0 = blockIdx.x * kLT;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += kLK) {
    // 64 rows x 32 k values each for x and W : 2048 elements / 256 threads = 8 per thread
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = threadIdx.x + e * 256;
      const int r = idx / kLK, kk = idx % kLK;
      const int k = k0 + kk;
      float xv = 0.f, wv = 0.f;
      if (k < K) {
        if (m0 + r < M) {
          xv = x[(m0 + r) * ldx + k];
          if (act_in == 1) xv = act_silu(xv);
        }
        if (n0 + r < N) wv = __bfloat162float(Wt[(n0 + r) * ldw + k]);
      }
      sx[kk][r] = xv;
      sw[kk][r] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kLK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sx[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sw[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += __bfloat162float(bias[n]);
      if (act_out == 1) v = act_silu(v);
      out[m * ldo + n] = v;
    }
  }
}

// Tall-skinny form of the above for the DiT head ([65520, 1536] fp32 x [64, 1536]^T): 128 rows x 64 columns per CTA,
// 8 x 4 outputs per thread (12 shared-memory words per 32 FMAs instead of 8 per 16), 16-byte global loads with the
// next K chunk's loads in flight while the current one is multiplied.  Every output is still ONE fp32 FMA chain over
// ascending k, so the result is bit-identical to linear_f32_kernel's.  Needs K % 32 == 0, ldx % 4 == 0, ldw % 8 == 0.
constexpr int kTSM = 128, kTSN = 64, kTSK = 32;
__global__ void __launch_bounds__(256, 2)
linear_f32_tall_kernel(const float* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ Wt, long long ldw,
                       const __nv_bfloat16* __restrict__ bias, float* __restrict__ out, long long ldo, int M, int N,
                       int K, int act_in, int act_out) {
  __shared__ __align__(16) float sx[kTSK][kTSM + 4];
  __shared__ __align__(16) float sw[kTSK][kTSN + 4];
  const int m0 = blockIdx.y * kTSM, n0 = blockIdx.x * kTSN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  // staging roles: x tile = 128 rows x 8 float4 (4 per thread), W tile = 64 rows x 4 uint4 of 8 bf16 (1 per thread)
  const int xr = threadIdx.x / 8, xk = (threadIdx.x % 8) * 4;          // rows xr + 32 e, e = 0..3
  const int wr = threadIdx.x / 4, wk = (threadIdx.x % 4) * 8;
  float4 xv[4];
  uint4 wv;
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = m0 + xr + 32 * e;
      xv[e] = r < M ? *reinterpret_cast<const float4*>(x + r * ldx + k0 + xk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    wv = n0 + wr < N ? *reinterpret_cast<const uint4*>(Wt + (n0 + wr) * ldw + k0 + wk) : make_uint4(0u, 0u, 0u, 0u);
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  fetch(0);
  for (int k0 = 0; k0 < K; k0 += kTSK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float4 v = xv[e];
      if (act_in == 1) v = make_float4(act_silu(v.x), act_silu(v.y), act_silu(v.z), act_silu(v.w));
      const int r = xr + 32 * e;
      sx[xk][r] = v.x; sx[xk + 1][r] = v.y; sx[xk + 2][r] = v.z; sx[xk + 3][r] = v.w;
    }
    {
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&wv);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        sw[wk + 2 * q][wr] = __uint_as_float(u[q] << 16);
        sw[wk + 2 * q + 1][wr] = __uint_as_float(u[q] & 0xffff0000u);
      }
    }
    __syncthreads();
    if (k0 + kTSK < K) fetch(k0 + kTSK);      // in flight during the multiply
#pragma unroll
    for (int kk = 0; kk < kTSK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sx[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sx[kk][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sw[kk][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += __bfloat162float(bias[n]);
      if (act_out == 1) v = act_silu(v);
      out[m * ldo + n] = v;
    }
  }
}

// Few-row form (the time-embedding MLPs: one row per sample, model.py:390-392): a 64-row tile would idle 62 of its
// rows and leave the launch on N/64 CTAs.  One warp per output column, lanes split K (16-byte loads of the bf16
// weight row, the fp32 input rows come from L1), five shuffles per row: the weights are read once at HBM speed.
constexpr int kFewRows = 8;
__global__ void __launch_bounds__(256)
linear_f32_fewrows_kernel(const float* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ Wt,
                          long long ldw, const __nv_bfloat16* __restrict__ bias, float* __restrict__ out, long long ldo,
                          int M, int N, int K, int act_in, int act_out) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  float acc[kFewRows];
#pragma unroll
  for (int m = 0; m < kFewRows; ++m) acc[m] = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    float w[8];
    load8_bf16(Wt + n * ldw + k, w);
#pragma unroll
    for (int m = 0; m < kFewRows; ++m) {
      if (m < M) {
        float v[8];
        load8_f32(x + m * ldx + k, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[m] = fmaf(act_in == 1 ? act_silu(v[j]) : v[j], w[j], acc[m]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < kFewRows; ++m) {
    const float t = warp_sum(acc[m]);
    if (lane == 0 && m < M) {
      float v = t;
      if (bias != nullptr) v += __bfloat162float(bias[n]);
      if (act_out == 1) v = act_silu(v);
      out[m * ldo + n] = v;
    }
  }
}

// modulation table: out[l, b, k, c] = float(mod[l, k, c]) + e[b, e_has_k ? k*d + c : c]
// (reference model.py:238 `modulation + e0` per block, :286 for the head), built once per forward.
__global__ void modulation_table_kernel(const __nv_bfloat16* __restrict__ mod,
                                        const float* __restrict__ e, float* __restrict__ out,
                                        int n_layers, int B, int Kk, int d, int e_has_k) {
  const long long total = static_cast<long long>(n_layers) * B * Kk * d;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long t = i;
    const int c = t % d; t /= d;
    const int k = t % Kk; t /= Kk;
    const int b = t % B; t /= B;
    const int l = static_cast<int>(t);
    const float m = __bfloat162float(mod[(static_cast<long long>(l) * Kk + k) * d + c]);
    const float ev = e[static_cast<long long>(b) * (e_has_k ? Kk * d : d) + (e_has_k ? k * d + c : c)];
    out[i] = m + ev;
  }
}

}  // namespace ovb

using namespace ovb;

// grid for the warp-per-row kernels: 8 rows per CTA pass; persistent at sm_count x the CTAs of this kernel that
// fit one SM (occ: per-device cache of the occupancy query, owned by the call site of each instantiation)
template <typename K>
static inline int warp_row_grid(int rows, K kernel, int (&occ)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev = dev < 0 || dev >= 64 ? 0 : dev;
  if (occ[dev] == 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, 256, 0) != cudaSuccess || nb < 1) nb = 1;
    occ[dev] = nb;
  }
  const long long ctas = (static_cast<long long>(rows) + 7) / 8;
  const long long cap = static_cast<long long>(device_sm_count()) * occ[dev];
  return static_cast<int>(ctas < cap ? ctas : cap);
}

static inline int ew_grid(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(device_sm_count()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

extern "C" int ovb_layernorm(const ovb_ln_args* a, void* stream) {
  OVB_CHECK_ARG(a && a->x && a->out && a->a && a->b, "ovb_layernorm: null pointer");
  OVB_CHECK_ARG(a->rows > 0 && a->d > 0 && a->d % 8 == 0 && a->d <= 8192,
                "ovb_layernorm: d=%d must be a multiple of 8 and <= 8192", a->d);
  OVB_CHECK_ARG(a->ldx % 8 == 0 && a->ldo % 8 == 0, "ovb_layernorm: ld must be multiple of 8");
  OVB_CHECK_ARG(a->mode == OVB_LN_MODULATE || a->mode == OVB_LN_AFFINE, "ovb_layernorm: bad mode");
  OVB_CHECK_ARG(a->mode != OVB_LN_MODULATE ||
                    (a->rows_per_sample > 0 && a->mod_row_stride % 4 == 0 &&
                     a->mod_sample_stride % 4 == 0),
                "ovb_layernorm: modulate needs rows_per_sample > 0 and strides %% 4 == 0");
  LnParams p;
  p.x = a->x; p.ldx = a->ldx; p.x_is_bf16 = a->x_is_bf16;
  p.out = a->out; p.ldo = a->ldo; p.out_is_f32 = a->out_is_f32;
  p.rows = a->rows; p.d = a->d; p.eps = a->eps; p.mode = a->mode;
  p.a = a->a; p.b = a->b;
  p.mod_row_stride = a->mod_row_stride; p.mod_sample_stride = a->mod_sample_stride;
  p.rows_per_sample = a->rows_per_sample > 0 ? a->rows_per_sample : a->rows;
  p.round_ln_bf16 = a->round_ln_bf16;
  const int vpl = (a->d / 4 + 31) / 32;  // float4 vectors per lane with one warp per row
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define OVB_LN_LAUNCH(...)                                                                                         \
  do {                                                                                                             \
    static int occ[64] = {};                                                                                       \
    layernorm_warp_kernel<__VA_ARGS__><<<warp_row_grid(a->rows, layernorm_warp_kernel<__VA_ARGS__>, occ), 256, 0, st>>>(p); \
  } while (0)
#define OVB_LN_FAST(V)                                                            \
  do {                                                                            \
    if (p.mode == OVB_LN_MODULATE) {                                              \
      if (p.round_ln_bf16) OVB_LN_LAUNCH(V, true, OVB_LN_MODULATE, true);         \
      else OVB_LN_LAUNCH(V, true, OVB_LN_MODULATE, false);                        \
    } else {                                                                      \
      if (p.round_ln_bf16) OVB_LN_LAUNCH(V, true, OVB_LN_AFFINE, true);           \
      else OVB_LN_LAUNCH(V, true, OVB_LN_AFFINE, false);                          \
    }                                                                             \
  } while (0)
  // the step's shapes: fp32 in, bf16 out, d = 1536 (OmniVideo2-1.3B) or 5120 (A14B)
  const bool fast = !p.x_is_bf16 && !p.out_is_f32 && getenv("OVB_LN_GENERAL") == nullptr;
  if (fast && a->d == 12 * 128) OVB_LN_FAST(12);
  else if (fast && a->d == 40 * 128) OVB_LN_FAST(40);
  else if (vpl <= 2) OVB_LN_LAUNCH(2);
  else if (vpl <= 4) OVB_LN_LAUNCH(4);
  else if (vpl <= 8) OVB_LN_LAUNCH(8);
  else if (vpl <= 12) OVB_LN_LAUNCH(12);
  else if (vpl <= 16) OVB_LN_LAUNCH(16);
  else if (vpl <= 24) OVB_LN_LAUNCH(24);
  else if (vpl <= 32) OVB_LN_LAUNCH(32);
  else if (vpl <= 40) OVB_LN_LAUNCH(40);
#undef OVB_LN_FAST
#undef OVB_LN_LAUNCH
  else {  // d > 5120: the row no longer fits one warp's registers in fp32
    const int threads = ((a->d / 8 + 31) / 32) * 32;
    layernorm_kernel<<<a->rows, threads, 0, st>>>(p);
  }
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_rmsnorm_rope(const ovb_rmsnorm_rope_args* a, void* stream) {
  OVB_CHECK_ARG(a && a->x && a->weight, "ovb_rmsnorm_rope: null pointer");
  OVB_CHECK_ARG(a->rows > 0 && a->d > 0 && a->d % 8 == 0 && a->d <= 8192,
                "ovb_rmsnorm_rope: d=%d must be a multiple of 8 and <= 8192", a->d);
  OVB_CHECK_ARG(a->ldx % 8 == 0, "ovb_rmsnorm_rope: ldx must be a multiple of 8");
  const int col_split = a->col_split > 1 ? a->col_split : 1;
  OVB_CHECK_ARG(a->d % col_split == 0 && (a->d / col_split) % 8 == 0,
                "ovb_rmsnorm_rope: d/col_split must be a multiple of 8");
  if (a->rope) {
    OVB_CHECK_ARG(a->d % 128 == 0, "ovb_rmsnorm_rope: RoPE needs head_dim 128 (d %% 128 == 0)");
    OVB_CHECK_ARG(a->cs_table && a->F > 0 && a->H > 0 && a->W > 0 && a->rows_per_sample > 0,
                  "ovb_rmsnorm_rope: RoPE needs cs_table, grid sizes and rows_per_sample");
    OVB_CHECK_ARG(a->F <= a->table_len && a->H <= a->table_len && a->W <= a->table_len,
                  "ovb_rmsnorm_rope: grid (%d,%d,%d) exceeds table_len %d", a->F, a->H, a->W,
                  a->table_len);
  }
  RmsRopeParams p;
  p.x = reinterpret_cast<__nv_bfloat16*>(a->x); p.ldx = a->ldx;
  p.rows = a->rows; p.d = a->d;
  p.weight = reinterpret_cast<const __nv_bfloat16*>(a->weight);
  p.eps = a->eps; p.rope = a->rope;
  p.cs = reinterpret_cast<const float2*>(a->cs_table); p.table_len = a->table_len;
  p.F = a->F; p.H = a->H; p.W = a->W;
  p.rows_per_sample = a->rows_per_sample > 0 ? a->rows_per_sample : a->rows;
  p.token_offset = a->token_offset;
  p.cols_per_split = a->d / col_split;
  p.col_split_stride = col_split > 1 ? a->col_split_stride : 0;
  p.scatter = a->dst_ptrs != nullptr;
  p.copy_only = a->copy_only;
  p.dst_ld = a->dst_ld;
  for (int g = 0; g < 8; ++g) p.dst[g] = nullptr;
  if (p.scatter) {
    OVB_CHECK_ARG(col_split <= 8 && a->dst_ld % 8 == 0 && a->dst_ld >= a->d / col_split,
                  "ovb_rmsnorm_rope: scatter form needs col_split <= 8 and dst_ld %% 8 == 0");
    OVB_CHECK_ARG(a->d / 8 <= 20 * 32, "ovb_rmsnorm_rope: scatter form supports d <= 5120");
    for (int g = 0; g < col_split; ++g) {
      OVB_CHECK_ARG(a->dst_ptrs[g] != nullptr && reinterpret_cast<uintptr_t>(a->dst_ptrs[g]) % 16 == 0,
                    "ovb_rmsnorm_rope: dst_ptrs[%d] null or misaligned", g);
      p.dst[g] = reinterpret_cast<__nv_bfloat16*>(a->dst_ptrs[g]);
    }
  } else {
    OVB_CHECK_ARG(!a->copy_only, "ovb_rmsnorm_rope: copy_only needs dst_ptrs");
  }
  const int vpl = (a->d / 8 + 31) / 32;
  OVB_CHECK_ARG(static_cast<long long>(col_split - 1) * a->col_split_stride + a->d < (1LL << 31),
                "ovb_rmsnorm_rope: column-split layout spans more than 2^31 elements");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define OVB_RMS_LAUNCH(V)                                                                                        \
  do {                                                                                                           \
    static int occ0[64] = {}, occ1[64] = {};                                                                     \
    if (p.scatter)                                                                                               \
      rmsnorm_rope_warp_kernel<V, true><<<warp_row_grid(a->rows, rmsnorm_rope_warp_kernel<V, true>, occ1), 256, 0, st>>>(p);  \
    else                                                                                                         \
      rmsnorm_rope_warp_kernel<V, false><<<warp_row_grid(a->rows, rmsnorm_rope_warp_kernel<V, false>, occ0), 256, 0, st>>>(p); \
  } while (0)
  if (vpl <= 1) OVB_RMS_LAUNCH(1);
  else if (vpl <= 2) OVB_RMS_LAUNCH(2);
  else if (vpl <= 4) OVB_RMS_LAUNCH(4);
  else if (vpl <= 6) OVB_RMS_LAUNCH(6);
  else if (vpl <= 8) OVB_RMS_LAUNCH(8);
  else if (vpl <= 16) OVB_RMS_LAUNCH(16);
  else if (vpl <= 20) OVB_RMS_LAUNCH(20);
#undef OVB_RMS_LAUNCH
  else {  // d > 5120
    const int threads = ((a->d / 8 + 31) / 32) * 32;
    rmsnorm_rope_kernel<<<a->rows, threads, 0, st>>>(p);
  }
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_patchify(const void* x, int32_t x_is_bf16, void* tokens, int64_t ldt, int32_t C,
                            int32_t F, int32_t H, int32_t W, void* stream) {
  OVB_CHECK_ARG(x && tokens, "ovb_patchify: null pointer");
  OVB_CHECK_ARG(C > 0 && F > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && ldt >= C * 4 &&
                    ldt % 2 == 0,
                "ovb_patchify: bad shape C=%d F=%d H=%d W=%d ldt=%lld", C, F, H, W,
                static_cast<long long>(ldt));
  const long long total = static_cast<long long>(F) * (H / 2) * (W / 2) * C * 2;
  const int grid = ew_grid(total, 256);
  if (x_is_bf16)
    patchify_kernel<__nv_bfloat16><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(tokens), ldt, C,
        F, H, W);
  else
    patchify_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float*>(x), reinterpret_cast<__nv_bfloat16*>(tokens), ldt, C, F, H,
        W);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_patchify3d(const void* x, int32_t x_is_bf16, void* tokens, int64_t ldt, int32_t C,
                              int32_t F, int32_t H, int32_t W, int32_t pt, int32_t ph, int32_t pw,
                              void* stream) {
  OVB_CHECK_ARG(x && tokens, "ovb_patchify3d: null pointer");
  OVB_CHECK_ARG(C > 0 && F > 0 && H > 0 && W > 0 && pt > 0 && ph > 0 && pw > 0,
                "ovb_patchify3d: bad shape C=%d F=%d H=%d W=%d patch=(%d,%d,%d)", C, F, H, W, pt, ph, pw);
  // a strided Conv3d floors: trailing voxels that do not fill a patch are dropped
  OVB_CHECK_ARG(F >= pt && H >= ph && W >= pw,
                "ovb_patchify3d: F,H,W = %d,%d,%d smaller than patch (%d,%d,%d)", F, H, W, pt, ph, pw);
  OVB_CHECK_ARG(ldt >= static_cast<int64_t>(C) * pt * ph * pw, "ovb_patchify3d: ldt=%lld too small",
                static_cast<long long>(ldt));
  const long long total = static_cast<long long>(F / pt) * (H / ph) * (W / pw) * C * pt * ph * pw;
  const int grid = ew_grid(total, 256);
  if (x_is_bf16)
    patchify3d_kernel<__nv_bfloat16><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(tokens), ldt, C, F, H,
        W, pt, ph, pw);
  else
    patchify3d_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float*>(x), reinterpret_cast<__nv_bfloat16*>(tokens), ldt, C, F, H, W, pt,
        ph, pw);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_unpatchify(const float* y, int64_t ldy, float* out, int32_t C, int32_t F,
                              int32_t Hp, int32_t Wp, void* stream) {
  OVB_CHECK_ARG(y && out, "ovb_unpatchify: null pointer");
  OVB_CHECK_ARG(C > 0 && F > 0 && Hp > 0 && Wp > 0 && ldy >= 4 * C, "ovb_unpatchify: bad shape");
  const long long total = static_cast<long long>(C) * F * Hp * Wp * 4;
  unpatchify_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, ldy, out, C, F, Hp, Wp);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_linear_f32(const float* x, int64_t ldx, const void* W, int64_t ldw,
                              const void* bias, float* out, int64_t ldo, int32_t M, int32_t N,
                              int32_t K, int32_t act_in, int32_t act_out, void* stream) {
  OVB_CHECK_ARG(x && W && out, "ovb_linear_f32: null pointer");
  OVB_CHECK_ARG(M > 0 && N > 0 && K > 0, "ovb_linear_f32: bad shape");
  const bool tall = M >= 4096 && K % kTSK == 0 && ldx % 4 == 0 && ldw % 8 == 0 &&
                    reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(W) % 16 == 0 &&
                    getenv("OVB_LINEAR_F32_GENERAL") == nullptr;
  const bool few = M <= kFewRows && K % 8 == 0 && ldx % 4 == 0 && ldw % 8 == 0 &&
                   reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(W) % 16 == 0 &&
                   getenv("OVB_LINEAR_F32_GENERAL") == nullptr;
  if (few) {
    linear_f32_fewrows_kernel<<<(N + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ldx, reinterpret_cast<const __nv_bfloat16*>(W), ldw,
        reinterpret_cast<const __nv_bfloat16*>(bias), out, ldo, M, N, K, act_in, act_out);
  } else if (tall) {
    dim3 grid((N + kTSN - 1) / kTSN, (M + kTSM - 1) / kTSM);
    linear_f32_tall_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ldx, reinterpret_cast<const __nv_bfloat16*>(W), ldw,
        reinterpret_cast<const __nv_bfloat16*>(bias), out, ldo, M, N, K, act_in, act_out);
  } else {
    dim3 grid((N + kLT - 1) / kLT, (M + kLT - 1) / kLT);
    linear_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ldx, reinterpret_cast<const __nv_bfloat16*>(W), ldw,
        reinterpret_cast<const __nv_bfloat16*>(bias), out, ldo, M, N, K, act_in, act_out);
  }
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_sinusoidal_embedding(const float* t, float* out, int32_t n, int32_t dim,
                                        void* stream) {
  OVB_CHECK_ARG(t && out && n > 0 && dim > 0 && dim % 2 == 0, "ovb_sinusoidal_embedding: bad args");
  sinusoid_kernel<<<n, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, out, n, dim);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_cast_f32_to_bf16(const float* x, void* out, int64_t n, void* stream) {
  OVB_CHECK_ARG(x && out && n > 0 && n % 8 == 0, "ovb_cast_f32_to_bf16: n must be a multiple of 8");
  cast_f32_bf16_kernel<<<ew_grid(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(out), n / 8);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}
extern "C" int ovb_cast_bf16_to_f32(const void* x, float* out, int64_t n, void* stream) {
  OVB_CHECK_ARG(x && out && n > 0 && n % 8 == 0, "ovb_cast_bf16_to_f32: n must be a multiple of 8");
  cast_bf16_f32_kernel<<<ew_grid(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), out, n / 8);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_modulation_table(const void* mod, const float* e, float* out, int32_t n_layers,
                                    int32_t B, int32_t K, int32_t d, int32_t e_has_k, void* stream) {
  OVB_CHECK_ARG(mod && e && out && n_layers > 0 && B > 0 && K > 0 && d > 0,
                "ovb_modulation_table: bad args");
  const long long total = static_cast<long long>(n_layers) * B * K * d;
  modulation_table_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(mod), e, out, n_layers, B, K, d, e_has_k);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

// ------------------------------------------------------------------------------------------------
// out[i] = sum_k c[k] * x_k[i]   (up to 5 terms, fp32 math, fp32 or bf16 tensors).
// The sampler side of a denoising step is nothing but such combinations of the latent-sized tensors
// (classifier-free guidance, x0 prediction, UniPC corrector and predictor: reference
// x2x_gen_unified.py:484, utils/fm_solvers_unipc.py:319-323, :467-475, :610-623); one launch per
// combination instead of ~6 ATen kernels each.  round_mask bit k rounds term k to bf16 before the
// sum (what torch's type promotion does for `0-dim fp32 tensor * bf16 tensor`).
namespace ovb {
struct LinCombParams {
  const void* x[5];
  float c[5];
  int is_bf16[5];
  int n_terms;
  int round_mask;
  void* out;
  int out_bf16;
  long long n;
};

__global__ void lincomb_kernel(const LinCombParams p) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= p.n) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (k >= p.n_terms) break;
    float v[4];
    if (p.is_bf16[k]) {
      const uint2 w = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.x[k]) + i);
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
      const float4 w = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.x[k]) + i);
      v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = p.c[k] * v[j];
      if ((p.round_mask >> k) & 1) t = bf16_round(t);
      acc[j] = k == 0 ? t : acc[j] + t;
    }
  }
  if (p.out_bf16) {
    uint2 o;
    o.x = pack_bf16x2(acc[0], acc[1]);
    o.y = pack_bf16x2(acc[2], acc[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + i) = o;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + i) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}
}  // namespace ovb

extern "C" int ovb_lincomb(const ovb_lincomb_args* a, void* stream) {
  using namespace ovb;
  OVB_CHECK_ARG(a != nullptr && a->out != nullptr, "ovb_lincomb: null args");
  OVB_CHECK_ARG(a->n_terms >= 1 && a->n_terms <= 5, "ovb_lincomb: n_terms=%d outside [1,5]", a->n_terms);
  OVB_CHECK_ARG(a->n > 0 && a->n % 4 == 0, "ovb_lincomb: n must be a positive multiple of 4");
  LinCombParams p;
  uintptr_t align = reinterpret_cast<uintptr_t>(a->out);
  for (int k = 0; k < 5; ++k) {
    p.x[k] = k < a->n_terms ? a->x[k] : nullptr;
    p.c[k] = a->c[k];
    p.is_bf16[k] = a->is_bf16[k];
    if (k < a->n_terms) {
      OVB_CHECK_ARG(a->x[k] != nullptr, "ovb_lincomb: term %d is null", k);
      align |= reinterpret_cast<uintptr_t>(a->x[k]);
    }
  }
  OVB_CHECK_ARG(align % 16 == 0, "ovb_lincomb: pointers must be 16-byte aligned");
  p.n_terms = a->n_terms;
  p.round_mask = a->round_mask;
  p.out = a->out;
  p.out_bf16 = a->out_is_bf16;
  p.n = a->n;
  lincomb_kernel<<<ew_grid(a->n / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}


// ------------------------------------------------------------------------------------------------
// Cross-GPU signal / wait over peer-mapped flag pads (include/ovb200.h "peer transport").
namespace ovb {
struct PeerPads { uint32_t* pad[8]; };

__global__ void peer_signal_wait_kernel(const PeerPads pads, int rank, int world, uint32_t epoch_arg,
                                        uint32_t* epoch_ctr, uint32_t* status) {
  const int r = threadIdx.x;
  // The epoch normally lives in device memory and advances by one per launch, so that the launch
  // can sit inside a replayed CUDA graph; epoch_arg is used when no counter is given.
  const uint32_t epoch = epoch_ctr != nullptr ? *epoch_ctr + 1u : epoch_arg;
  __syncwarp();
  if (r < world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(pads.pad[r] + rank), "r"(epoch) : "memory");
    const uint32_t* mine = pads.pad[rank] + r;
    const long long t0 = clock64();
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(mine) : "memory");
      if (v >= epoch) break;
      // ~30 s at 2 GHz: a peer died or the epochs diverged.  Continuing would let the next kernel
      // read half-written receive buffers, so this is FATAL: record the reason and trap (the host
      // sees a launch failure at its next synchronisation; UlyssesPeerTransport.check() names it).
      if (clock64() - t0 > 60000000000LL) {
        if (status != nullptr) {
          *status = 1u;
          __threadfence_system();
        }
        __trap();
      }
    } while (true);
  }
  __syncwarp();
  if (r == 0 && epoch_ctr != nullptr) *epoch_ctr = epoch;
}
}  // namespace ovb

extern "C" int ovb_peer_signal_wait(uint32_t* const* peer_pads, int32_t rank, int32_t world,
                                    uint32_t epoch, uint32_t* epoch_counter, uint32_t* status,
                                    void* stream) {
  using namespace ovb;
  OVB_CHECK_ARG(peer_pads != nullptr && world >= 1 && world <= 8 && rank >= 0 && rank < world,
                "ovb_peer_signal_wait: bad rank/world %d/%d", rank, world);
  PeerPads pads;
  for (int r = 0; r < 8; ++r) pads.pad[r] = nullptr;
  for (int r = 0; r < world; ++r) {
    OVB_CHECK_ARG(peer_pads[r] != nullptr, "ovb_peer_signal_wait: pad %d is null", r);
    pads.pad[r] = peer_pads[r];
  }
  peer_signal_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(pads, rank, world, epoch,
                                                                           epoch_counter, status);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_enable_peer_access(int32_t peer_device) {
  int cur = -1;
  OVB_CHECK_CUDA(cudaGetDevice(&cur));
  if (cur == peer_device) return OVB_OK;
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    (void)cudaGetLastError();
    return OVB_OK;
  }
  OVB_CHECK_CUDA(e);
  return OVB_OK;
}
