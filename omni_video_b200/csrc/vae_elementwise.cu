// HBM-bound kernels of the Wan2.1 VAE path: channel RMS_norm (+SiLU) on channels-last voxels,
// planar <-> channels-last converters with the latent (de)normalisation folded in, and the row
// softmax of the single-head d=384 attention.  16-byte vector accesses throughout.
#include "host_util.h"
#include "ptx.cuh"

namespace ovb {

__device__ __forceinline__ void unpack8(const uint4& w, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}

// ------------------------------------------------------------------------------------------------
// RMS_norm over C (+SiLU).  Bytes/element: 2 (bf16 in) + 2 (bf16 out) = 4.
// A group of LPV lanes (16 or 32) owns one voxel: lane l handles the 16-byte chunks l, l+LPV, ...
// of the voxel's C*2 contiguous bytes (C/8 chunks: 12 / 24 / 48 for C = 96 / 192 / 384), so a warp
// reads 2 (or 1) whole voxels = one contiguous span, and the sum of squares is a shuffle reduction
// inside the group.  MAXC chunks per lane are kept in registers so x is read exactly once.
template <int LPV, int MAXC>
__global__ void __launch_bounds__(256) rmsnorm_silu_cl_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const float* __restrict__ gamma,
                                                               __nv_bfloat16* __restrict__ out,
                                                               long long n_vox, int C, int silu) {
  constexpr int kU = MAXC >= 2 ? 2 : 4;  // voxels per group per iteration: 16-byte loads in flight per lane
  const int lane = threadIdx.x & (LPV - 1);
  constexpr int kGroupsPerWarp = 32 / LPV;
  const int sub = (threadIdx.x & 31) / LPV;
  const long long warp_id = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int chunks_all = C >> 3;
  const float sqrt_c = sqrtf(static_cast<float>(C));
  constexpr int kVoxPerWarp = kGroupsPerWarp * kU;
  // gamma chunks of this lane stay in registers for the whole kernel
  float g[MAXC][8];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + i * LPV;
#pragma unroll
    for (int j = 0; j < 8; ++j) g[i][j] = c < chunks_all ? gamma[c * 8 + j] : 0.f;
  }
  // warp-uniform trip count: the shuffles below need all 32 lanes
  for (long long v0 = warp_id * kVoxPerWarp; v0 < n_vox; v0 += n_warps * kVoxPerWarp) {
    uint4 raw[kU][MAXC];
    int chunks[kU];
    // all loads of the iteration first: kU * MAXC independent 16-byte requests per lane
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long v = v0 + u * kGroupsPerWarp + sub;
      chunks[u] = v < n_vox ? chunks_all : 0;
      const uint4* src = reinterpret_cast<const uint4*>(x + v * C);
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int c = lane + i * LPV;
        raw[u][i] = make_uint4(0u, 0u, 0u, 0u);
        if (c < chunks[u]) raw[u][i] = src[c];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long v = v0 + u * kGroupsPerWarp + sub;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        float f[8];
        unpack8(raw[u][i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
      }
#pragma unroll
      for (int o = LPV / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float inv = sqrt_c / fmaxf(sqrtf(ss), 1e-12f);
      uint4* dst = reinterpret_cast<uint4*>(out + v * C);
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int c = lane + i * LPV;
        if (c < chunks[u]) {
          float f[8];
          unpack8(raw[u][i], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float y = f[j] * inv * g[i][j];
            // SiLU(y) = y * sigmoid(y) = h + h * tanh(h), h = y/2: one MUFU op and no division
            // (tanh.approx relative error ~2^-11, below the bf16 rounding of the result)
            if (silu) {
              const float h = 0.5f * y;
              float th;
              asm("tanh.approx.f32 %0, %1;\n" : "=f"(th) : "f"(h));
              y = fmaf(h, th, h);
            }
            f[j] = y;
          }
          uint4 o;
          o.x = pack_bf16x2(f[0], f[1]);
          o.y = pack_bf16x2(f[2], f[3]);
          o.z = pack_bf16x2(f[4], f[5]);
          o.w = pack_bf16x2(f[6], f[7]);
          dst[c] = o;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// planar fp32 [C][n] -> channels-last bf16 [n][Cpad] (affine per channel, zero padded channels).
// One thread per voxel; reads are coalesced along n per channel plane, writes are Cpad*2 bytes
// contiguous per thread.  Only used on the small boundary tensors (latent z, input video).
__global__ void planar_to_cl_kernel(const float* __restrict__ src, const float* __restrict__ scale,
                                    const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
                                    int C, int Cpad, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int c0 = 0; c0 < Cpad; c0 += 8) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      f[j] = c < C ? src[c * n + i] * scale[c] + shift[c] : 0.f;
    }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + i * Cpad + c0) = o;
  }
}

__global__ void cl_to_planar_kernel(const __nv_bfloat16* __restrict__ src, int Cs,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    float* __restrict__ out, int C, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int c = 0; c < C; ++c)
    out[c * n + i] = __bfloat162float(src[i * Cs + c]) * scale[c] + shift[c];
}

// ------------------------------------------------------------------------------------------------
// Row softmax: fp32 scores -> bf16 probabilities.  One CTA per row, three passes over a row that
// stays in L1/L2 (rows are <= 14400 floats = 56 KB).
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, long long lds,
                                                            __nv_bfloat16* __restrict__ P, long long ldp,
                                                            int cols, float scale_log2) {
  __shared__ float red[32];
  const float* s = S + blockIdx.x * lds;
  __nv_bfloat16* p = P + blockIdx.x * ldp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, s[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) sum += exp2f((s[c] - m) * scale_log2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) sum += red[i];
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x; c < cols; c += blockDim.x)
    p[c] = __float2bfloat16_rn(exp2f((s[c] - m) * scale_log2) * inv);
}

// Decoded video -> displayable frames (reference utils/utils.py:177-185, `cache_video`):
//   out[t, y0 + y, x0 + x, c] = uint8( (clamp(x[c, t, y, x], lo, hi) - lo) / max(hi - lo, 1e-5) * 255 )   (truncation)
// planar fp32 [C, T, H, W] in, interleaved uint8 out with arbitrary frame/row pitch and origin, so that several
// videos can be laid out as torchvision's make_grid does.  Same fp32 operation order as torch (clamp, sub, div,
// mul, cast), IEEE division: results are bit-identical.  One thread per pixel; 12 B read, 3 B written per pixel.
__global__ void video_to_uint8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int C, int T, int H,
                                      int W, long long out_frame_stride, long long out_row_stride, float lo, float hi,
                                      float denom) {
  const long long n = static_cast<long long>(T) * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xw = static_cast<int>(i % W);
    const long long r = i / W;
    const int y = static_cast<int>(r % H);
    const int t = static_cast<int>(r / H);
    uint8_t* o = out + t * out_frame_stride + y * out_row_stride + static_cast<long long>(xw) * C;
    for (int c = 0; c < C; ++c) {
      float v = x[c * n + i];
      v = fminf(fmaxf(v, lo), hi);
      v = __fmul_rn(__fdiv_rn(__fsub_rn(v, lo), denom), 255.0f);
      o[c] = static_cast<uint8_t>(static_cast<int>(v));
    }
  }
}

// Sum of per-tap partial products (the second half of a convolution with very few output channels, see
// ovb_tap_sum in ovb200.h):  out[c][t][h][w] = bias[c] + sum_k y[k][t + dt_k][h + dh_k][w + dw_k][c],
// y bf16 [n_taps][Ty][Hy][Wy][4], terms outside y's extents are zero (causal / spatial zero padding).
// One thread per output voxel, consecutive lanes = consecutive w: every tap is one 8-byte load per lane and
// 256 contiguous bytes per warp; every y element is read exactly once (8 B x n_taps per voxel in, 4 B x C out).
struct TapSumParams {
  const uint2* y;
  const float* bias;
  float* out;
  int n_taps, Ty, Hy, Wy, C, To, Ho, Wo, clamp;
  signed char dt[27], dh[27], dw[27];
};

__global__ void __launch_bounds__(256) tap_sum_kernel(const __grid_constant__ TapSumParams p) {
  const long long n = static_cast<long long>(p.To) * p.Ho * p.Wo;
  const long long plane = static_cast<long long>(p.Ty) * p.Hy * p.Wy;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % p.Wo);
    const long long r = i / p.Wo;
    const int h = static_cast<int>(r % p.Ho);
    const int t = static_cast<int>(r / p.Ho);
    uint2 v[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      v[k] = make_uint2(0u, 0u);
      if (k < p.n_taps) {
        const int tt = t + p.dt[k], hh = h + p.dh[k], ww = w + p.dw[k];
        if (tt >= 0 && tt < p.Ty && hh >= 0 && hh < p.Hy && ww >= 0 && ww < p.Wy)
          v[k] = __ldcs(p.y + k * plane + (static_cast<long long>(tt) * p.Hy + hh) * p.Wy + ww);
      }
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 27; ++k) {  // fixed order: deterministic
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[k].x));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v[k].y));
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < p.C) {
        float o = acc[c] + (p.bias != nullptr ? p.bias[c] : 0.f);
        if (p.clamp) o = fminf(fmaxf(o, -1.0f), 1.0f);
        p.out[c * n + i] = o;
      }
    }
  }
}

}  // namespace ovb

extern "C" int ovb_rmsnorm_silu_cl(const void* x, const float* gamma, void* out, int64_t n_vox,
                                   int32_t C, int32_t silu, void* stream) {
  using namespace ovb;
  OVB_CHECK_ARG(x && gamma && out, "ovb_rmsnorm_silu_cl: null pointer");
  OVB_CHECK_ARG(n_vox > 0 && C > 0 && C % 8 == 0 && C <= 1024,
                "ovb_rmsnorm_silu_cl: bad shape n_vox=%lld C=%d (C %% 8 == 0, C <= 1024)",
                static_cast<long long>(n_vox), C);
  OVB_CHECK_ARG((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                 reinterpret_cast<uintptr_t>(gamma)) % 16 == 0,
                "ovb_rmsnorm_silu_cl: pointers must be 16-byte aligned");
  const int chunks = C / 8;
  const int threads = 256;
  const int sms = device_sm_count();
  auto s = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* xo = reinterpret_cast<__nv_bfloat16*>(out);
  if (chunks <= 16) {
    const long long groups_needed = n_vox;
    long long blocks = (groups_needed * 16 + threads - 1) / threads;
    if (blocks > sms * 16) blocks = sms * 16;
    rmsnorm_silu_cl_kernel<16, 1><<<static_cast<int>(blocks), threads, 0, s>>>(xi, gamma, xo, n_vox, C, silu);
  } else {
    long long blocks = (n_vox * 32 + threads - 1) / threads;
    if (blocks > sms * 16) blocks = sms * 16;
    if (chunks <= 32)
      rmsnorm_silu_cl_kernel<32, 1><<<static_cast<int>(blocks), threads, 0, s>>>(xi, gamma, xo, n_vox, C, silu);
    else if (chunks <= 64)
      rmsnorm_silu_cl_kernel<32, 2><<<static_cast<int>(blocks), threads, 0, s>>>(xi, gamma, xo, n_vox, C, silu);
    else
      rmsnorm_silu_cl_kernel<32, 4><<<static_cast<int>(blocks), threads, 0, s>>>(xi, gamma, xo, n_vox, C, silu);
  }
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_planar_to_cl(const float* src, const float* scale, const float* shift, void* out,
                                int32_t C, int32_t Cpad, int64_t n, void* stream) {
  using namespace ovb;
  OVB_CHECK_ARG(src && scale && shift && out, "ovb_planar_to_cl: null pointer");
  OVB_CHECK_ARG(C > 0 && Cpad >= C && Cpad % 8 == 0 && n > 0, "ovb_planar_to_cl: bad shape C=%d Cpad=%d", C, Cpad);
  const int threads = 256;
  const long long blocks = (n + threads - 1) / threads;
  planar_to_cl_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      src, scale, shift, reinterpret_cast<__nv_bfloat16*>(out), C, Cpad, n);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_cl_to_planar(const void* src, int32_t Cs, const float* scale, const float* shift,
                                float* out, int32_t C, int64_t n, void* stream) {
  using namespace ovb;
  OVB_CHECK_ARG(src && scale && shift && out, "ovb_cl_to_planar: null pointer");
  OVB_CHECK_ARG(C > 0 && Cs >= C && n > 0, "ovb_cl_to_planar: bad shape C=%d Cs=%d", C, Cs);
  const int threads = 256;
  const long long blocks = (n + threads - 1) / threads;
  cl_to_planar_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), Cs, scale, shift, out, C, n);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_softmax_rows(const float* S, int64_t lds, void* P, int64_t ldp, int32_t rows,
                                int32_t cols, float scale, void* stream) {
  using namespace ovb;
  OVB_CHECK_ARG(S && P && rows > 0 && cols > 0 && lds >= cols && ldp >= cols, "ovb_softmax_rows: bad arguments");
  softmax_rows_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      S, lds, reinterpret_cast<__nv_bfloat16*>(P), ldp, cols, scale * 1.4426950408889634f);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_tap_sum(const ovb_tap_sum_args* a, void* stream) {
  using namespace ovb;
  OVB_CHECK_ARG(a != nullptr && a->y && a->out && a->tap_dt && a->tap_dh && a->tap_dw, "ovb_tap_sum: null pointer");
  OVB_CHECK_ARG(a->n_taps >= 1 && a->n_taps <= 27, "ovb_tap_sum: n_taps=%d outside [1,27]", a->n_taps);
  OVB_CHECK_ARG(a->C >= 1 && a->C <= 4, "ovb_tap_sum: C=%d outside [1,4]", a->C);
  OVB_CHECK_ARG(a->Ty > 0 && a->Hy > 0 && a->Wy > 0 && a->To > 0 && a->Ho > 0 && a->Wo > 0,
                "ovb_tap_sum: bad extents y [%d,%d,%d] out [%d,%d,%d]", a->Ty, a->Hy, a->Wy, a->To, a->Ho, a->Wo);
  OVB_CHECK_ARG(reinterpret_cast<uintptr_t>(a->y) % 8 == 0, "ovb_tap_sum: y must be 8-byte aligned");
  TapSumParams p;
  p.y = reinterpret_cast<const uint2*>(a->y);
  p.bias = a->bias;
  p.out = a->out;
  p.n_taps = a->n_taps;
  p.Ty = a->Ty; p.Hy = a->Hy; p.Wy = a->Wy;
  p.C = a->C;
  p.To = a->To; p.Ho = a->Ho; p.Wo = a->Wo;
  p.clamp = a->clamp;
  for (int k = 0; k < 27; ++k) {
    p.dt[k] = k < a->n_taps ? a->tap_dt[k] : 0;
    p.dh[k] = k < a->n_taps ? a->tap_dh[k] : 0;
    p.dw[k] = k < a->n_taps ? a->tap_dw[k] : 0;
  }
  const long long n = static_cast<long long>(a->To) * a->Ho * a->Wo;
  const long long blocks = (n + 255) / 256;
  const long long cap = static_cast<long long>(device_sm_count()) * 64;
  tap_sum_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}

extern "C" int ovb_video_to_uint8(const float* x, uint8_t* out, int32_t C, int32_t T, int32_t H, int32_t W,
                                  int64_t out_frame_stride, int64_t out_row_stride, float lo, float hi,
                                  void* stream) {
  using namespace ovb;
  OVB_CHECK_ARG(x && out, "ovb_video_to_uint8: null pointer");
  OVB_CHECK_ARG(C >= 1 && C <= 4 && T > 0 && H > 0 && W > 0, "ovb_video_to_uint8: bad shape C=%d T=%d H=%d W=%d", C,
                T, H, W);
  OVB_CHECK_ARG(out_row_stride >= static_cast<int64_t>(W) * C && out_frame_stride >= out_row_stride * H,
                "ovb_video_to_uint8: output pitches too small");
  OVB_CHECK_ARG(hi >= lo, "ovb_video_to_uint8: value range (%g, %g)", lo, hi);
  const float denom = fmaxf(hi - lo, 1e-5f);
  const long long n = static_cast<long long>(T) * H * W;
  const long long blocks = (n + 255) / 256;
  const unsigned grid = static_cast<unsigned>(blocks < 148LL * 32 ? blocks : 148LL * 32);
  video_to_uint8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, C, T, H, W, out_frame_stride,
                                                                             out_row_stride, lo, hi, denom);
  OVB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return OVB_OK;
}
