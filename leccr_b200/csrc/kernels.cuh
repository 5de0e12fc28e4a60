// CUDA-core kernels around the tensor-core similarity pass: operand preparation
// (cast / L2-normalise / split), finalisation of the streamed reductions, and the
// HBM-bound ranking of an already materialised score matrix.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include "ptx.cuh"

namespace leccr {

// -------------------------------------------------------------------------------- 16-bit helpers
template <int FMT>
__device__ __forceinline__ uint16_t f32_to_16(float x) {
  if (FMT == 0) {
    __half h = __float2half_rn(x);
    return *reinterpret_cast<uint16_t*>(&h);
  } else {
    __nv_bfloat16 h = __float2bfloat16_rn(x);
    return *reinterpret_cast<uint16_t*>(&h);
  }
}
template <int FMT>
__device__ __forceinline__ float f16_to_32(uint16_t u) {
  if (FMT == 0) {
    return __half2float(*reinterpret_cast<__half*>(&u));
  } else {
    return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
  }
}

// Order-preserving float <-> uint mapping so atomicMax/atomicMin work for any sign.
__device__ __forceinline__ unsigned f32_ordered(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_f32(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Per-tensor statistics written by the prep kernels (all atomically max-combined; zero-init).
//   [0] max_i ||hi_i||   [1] max_i ||x_i - hi_i||   [2] max |x|   [3] non-finite / fp16-overflow flag
constexpr int kStatWords = 4;
// max-combine a non-negative float into a per-tensor statistic.  Thousands of rows update the same
// word, so look first: after the first few rows the atomic is almost never needed.
__device__ __forceinline__ void stat_max(float* addr, float v) {
  if (v > *reinterpret_cast<volatile float*>(addr)) atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

// Combine the per-row statistics of a block's warps in shared memory and publish once per block.
// Every thread of the block must call this (rows past the end contribute zeros).
__device__ __forceinline__ void publish_stats(float* stats, float a, float b, float amax, bool bad) {
  __shared__ float s_st[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (stats == nullptr) return;  // uniform
  if (lane == 0) {
    s_st[0][warp] = a;
    s_st[1][warp] = b;
    s_st[2][warp] = amax;
    s_st[3][warp] = bad ? 1.f : 0.f;
  }
  __syncthreads();
  if (warp == 0 && lane < 4) {
    float m = 0.f;
    for (int w = 0; w < nw; ++w) m = fmaxf(m, s_st[lane][w]);
    if (m > 0.f) stat_max(stats + lane, m);
  }
}

// --------------------------------------------------------------------------------
// prep_rows: fp32 rows -> 16-bit tensor-core operand.  One warp per row.
//   normalize : F.normalize(x, dim=-1) (eps 1e-12) fused in front of the cast
//               (models/xvlm.py:245-256 is the step right before the path).
//   layout 0  : [hi]                 K = D
//   layout 1  : [hi | lo | hi]       K = 3D, "rows" role of the split product
//   layout 2  : [hi | hi | lo]       K = 3D, "cols" role:  rows.cols^T = hi.hi + lo.hi + hi.lo
// lo = cast(x - hi): the product then carries ~fp32 accuracy through 16-bit tensor cores.
// --------------------------------------------------------------------------------
template <int FMT>
__global__ void prep_rows_kernel(const float* __restrict__ src, long long ld_src, int n, int D,
                                 int normalize, int layout, uint16_t* __restrict__ dst,
                                 long long ld_dst, float* __restrict__ rn_hi,
                                 float* __restrict__ rn_lo, float* __restrict__ stats) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= n) {
    publish_stats(stats, 0.f, 0.f, 0.f, false);
    return;
  }
  const float* x = src + static_cast<long long>(row) * ld_src;
  float inv = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float t = x[d];
      ss = fmaf(t, t, ss);
    }
    ss = warp_sum(ss);
    inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  }
  uint16_t* o = dst + static_cast<long long>(row) * ld_dst;
  float nh = 0.f, nl = 0.f, amax = 0.f;
  bool bad = false;
  for (int d = lane; d < D; d += 32) {
    const float t = normalize ? x[d] * inv : x[d];
    const uint16_t h = f32_to_16<FMT>(t);
    const float hf = f16_to_32<FMT>(h);
    const float r = t - hf;
    bad |= !isfinite(hf);
    amax = fmaxf(amax, fabsf(t));
    nh = fmaf(hf, hf, nh);
    if (layout == 0) {
      o[d] = h;
      nl = fmaf(r, r, nl);
    } else {
      const uint16_t l = f32_to_16<FMT>(r);
      const float rr = r - f16_to_32<FMT>(l);
      nl = fmaf(rr, rr, nl);
      o[d] = h;
      if (layout == 1) {
        o[D + d] = l;
        o[2 * D + d] = h;
      } else {
        o[D + d] = h;
        o[2 * D + d] = l;
      }
    }
  }
  nh = warp_sum(nh);
  nl = warp_sum(nl);
  amax = warp_max(amax);
  const unsigned anybad = __ballot_sync(0xffffffffu, bad);
  const float a = sqrtf(nh), b = sqrtf(nl);
  if (lane == 0) {
    if (rn_hi) rn_hi[row] = a;
    if (rn_lo) rn_lo[row] = b;
  }
  publish_stats(stats, a, b, amax, anybad != 0);
}

// Fast path of prep_rows for D % 128 == 0, D <= 1024, 16-byte aligned rows: one warp per row, each
// lane holds D/128 float4 in registers (one pass over HBM even when normalising), 8-byte stores.
template <int FMT>
__device__ __forceinline__ void prep_rows_vec_body(const float* __restrict__ src, long long ld_src, int n, int D,
                                                   int normalize, int layout, uint16_t* __restrict__ dst,
                                                   long long ld_dst, float* __restrict__ rn_hi,
                                                   float* __restrict__ rn_lo, float* __restrict__ stats, int row) {
  const int lane = threadIdx.x & 31;
  if (row >= n) {
    publish_stats(stats, 0.f, 0.f, 0.f, false);
    return;
  }
  const int nv = D >> 7;  // float4 per lane
  const float4* x4 = reinterpret_cast<const float4*>(src + static_cast<long long>(row) * ld_src);
  float4 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < nv) x[i] = __ldg(x4 + lane + 32 * i);
  float inv = 1.f;
  if (normalize) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) ss += x[i].x * x[i].x + x[i].y * x[i].y + x[i].z * x[i].z + x[i].w * x[i].w;
    ss = warp_sum(ss);
    inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  }
  uint16_t* o = dst + static_cast<long long>(row) * ld_dst;
  float nh = 0.f, nl = 0.f, amax = 0.f;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < nv) {
      const float t[4] = {x[i].x * inv, x[i].y * inv, x[i].z * inv, x[i].w * inv};
      uint16_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = f32_to_16<FMT>(t[j]);
        const float hf = f16_to_32<FMT>(h[j]);
        float r = t[j] - hf;
        bad |= !isfinite(hf);
        amax = fmaxf(amax, fabsf(t[j]));
        nh = fmaf(hf, hf, nh);
        l[j] = 0;
        if (layout != 0) {
          l[j] = f32_to_16<FMT>(r);
          r -= f16_to_32<FMT>(l[j]);
        }
        nl = fmaf(r, r, nl);
      }
      const int d = 4 * (lane + 32 * i);
      const uint2 hv = make_uint2(h[0] | (static_cast<uint32_t>(h[1]) << 16), h[2] | (static_cast<uint32_t>(h[3]) << 16));
      const uint2 lv = make_uint2(l[0] | (static_cast<uint32_t>(l[1]) << 16), l[2] | (static_cast<uint32_t>(l[3]) << 16));
      *reinterpret_cast<uint2*>(o + d) = hv;
      if (layout == 1) {
        *reinterpret_cast<uint2*>(o + D + d) = lv;
        *reinterpret_cast<uint2*>(o + 2 * D + d) = hv;
      } else if (layout == 2) {
        *reinterpret_cast<uint2*>(o + D + d) = hv;
        *reinterpret_cast<uint2*>(o + 2 * D + d) = lv;
      }
    }
  }
  nh = warp_sum(nh);
  nl = warp_sum(nl);
  amax = warp_max(amax);
  const unsigned anybad = __ballot_sync(0xffffffffu, bad);
  const float a = sqrtf(nh), b = sqrtf(nl);
  if (lane == 0) {
    if (rn_hi) rn_hi[row] = a;
    if (rn_lo) rn_lo[row] = b;
  }
  publish_stats(stats, a, b, amax, anybad != 0);
}

template <int FMT>
__global__ void prep_rows_vec_kernel(const float* __restrict__ src, long long ld_src, int n, int D,
                                     int normalize, int layout, uint16_t* __restrict__ dst,
                                     long long ld_dst, float* __restrict__ rn_hi,
                                     float* __restrict__ rn_lo, float* __restrict__ stats) {
  prep_rows_vec_body<FMT>(src, ld_src, n, D, normalize, layout, dst, ld_dst, rn_hi, rn_lo, stats,
                          blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
}

// Two tensors (the two embedding sets of an evaluation) in one launch: blocks [0, blocks0) cast the first.
struct PrepTensor {
  const float* src;
  long long ld_src;
  int n;
  uint16_t* dst;
  long long ld_dst;
  float* rn_hi;
  float* rn_lo;
  float* stats;
};
template <int FMT>
__global__ void prep_rows_vec_pair_kernel(const PrepTensor t0, const PrepTensor t1, int blocks0, int D, int normalize,
                                          int layout) {
  const bool second = static_cast<int>(blockIdx.x) >= blocks0;  // block-uniform
  const PrepTensor& t = second ? t1 : t0;
  const int row = (blockIdx.x - (second ? blocks0 : 0)) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  prep_rows_vec_body<FMT>(t.src, t.ld_src, t.n, D, normalize, layout, t.dst, t.ld_dst, t.rn_hi, t.rn_lo, t.stats, row);
}

// --------------------------------------------------------------------------------
// normalize_rows: y = x / max(||x||_2, 1e-12) per row == F.normalize(x, dim=-1), the last step of
// XVLMBase.get_features (models/xvlm.py:245-256, models/xvlm_video.py:264-277), as a training op: fp32 result
// (what the reference hands to its losses), 1 / max(||x||, eps) kept for the backward, and -- optionally, in the
// same pass -- the 16-bit tensor-core operand of y, so the similarity stage needs no second cast.
// One warp per row.  Backward: dx = inv * (g - y (y . g)) where ||x|| > eps, g * inv below (clamp_min passes no
// gradient to the norm).
// --------------------------------------------------------------------------------
template <int FMT>
__global__ void normalize_rows_kernel(const float* __restrict__ x, long long ld_x, int n, int D, float* __restrict__ y,
                                      long long ld_y, float* __restrict__ inv_out, uint16_t* __restrict__ y16,
                                      long long ld_y16) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* xr = x + static_cast<long long>(row) * ld_x;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) ss = fmaf(xr[d], xr[d], ss);
  const float nrm = sqrtf(warp_sum(ss));
  const float inv = 1.f / fmaxf(nrm, 1e-12f);
  if (lane == 0 && inv_out != nullptr) inv_out[row] = nrm > 1e-12f ? inv : -inv;  // sign bit marks a clamped row
  float* yr = y + static_cast<long long>(row) * ld_y;
  for (int d = lane; d < D; d += 32) {
    const float v = xr[d] * inv;
    yr[d] = v;
    if (y16 != nullptr) y16[static_cast<long long>(row) * ld_y16 + d] = f32_to_16<FMT>(v);
  }
}

__global__ void normalize_rows_bwd_kernel(const float* __restrict__ y, long long ld_y, const float* __restrict__ inv_in,
                                          const float* __restrict__ g, long long ld_g, int n, int D,
                                          float* __restrict__ dx, long long ld_dx) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* yr = y + static_cast<long long>(row) * ld_y;
  const float* gr = g + static_cast<long long>(row) * ld_g;
  const float iv = inv_in[row];
  const bool clamped = iv < 0.f;
  const float inv = fabsf(iv);
  float dot = 0.f;
  if (!clamped)
    for (int d = lane; d < D; d += 32) dot = fmaf(yr[d], gr[d], dot);
  dot = warp_sum(dot);
  float* o = dx + static_cast<long long>(row) * ld_dx;
  for (int d = lane; d < D; d += 32) o[d] = inv * (gr[d] - yr[d] * dot);
}

// --------------------------------------------------------------------------------
// prep_push: the cast prologue fused with its all-gather.  Every rank casts ITS rows of the fp32
// features to the 16-bit operand format and stores them straight into EVERY rank's gathered operand
// buffer (peer pointers over NVLink, own pointer included) at its row offset -- the exchange of
// models/xvlm.py:53-59 happens inside the kernel that produces the data, no separate collective.
// dsts: device array of `world` base pointers (symmetric buffers, one per rank).  One warp per row.
// --------------------------------------------------------------------------------
template <int FMT>
__global__ void prep_push_kernel(const float* __restrict__ src, long long ld_src, int n, int D, int normalize,
                                 uint16_t* const* __restrict__ dsts, int world, long long dst_row0,
                                 long long dst_col0, long long ld_dst) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* x = src + static_cast<long long>(row) * ld_src;
  float inv = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) ss = fmaf(x[d], x[d], ss);
    inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  }
  const long long off = (dst_row0 + row) * ld_dst + dst_col0;
  const bool vec = (D & 3) == 0 && (ld_src & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                   (ld_dst & 3) == 0 && (dst_col0 & 3) == 0;
  if (vec) {
    for (int d = 4 * lane; d < D; d += 128) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(x + d));
      const uint16_t h0 = f32_to_16<FMT>(t.x * inv), h1 = f32_to_16<FMT>(t.y * inv);
      const uint16_t h2 = f32_to_16<FMT>(t.z * inv), h3 = f32_to_16<FMT>(t.w * inv);
      const uint2 hv = make_uint2(h0 | (static_cast<uint32_t>(h1) << 16), h2 | (static_cast<uint32_t>(h3) << 16));
      for (int p = 0; p < world; ++p) *reinterpret_cast<uint2*>(dsts[p] + off + d) = hv;  // peer stores
    }
  } else {
    for (int d = lane; d < D; d += 32) {
      const uint16_t h = f32_to_16<FMT>(x[d] * inv);
      for (int p = 0; p < world; ++p) dsts[p][off + d] = h;
    }
  }
}

// The whole contrastive exchange of one rank in ONE launch: blockIdx.y = 0 casts + pushes the image rows to
// column 0 of every rank's [n][2D] buffer, 1 the text rows to column D, 2 pushes the idx words (optional);
// block (0, 0) also clears `zero_words` doubles at `zero` (scratch of the forward that follows).
template <int FMT>
__global__ void itc_push_kernel(const float* __restrict__ img, long long ld_img, const float* __restrict__ txt,
                                long long ld_txt, const unsigned long long* __restrict__ idx, int B, int D,
                                uint16_t* const* __restrict__ dsts, unsigned long long* const* __restrict__ idx_dsts,
                                uint16_t* __restrict__ own_rows, unsigned long long* __restrict__ own_idx,
                                int world, long long row0, double* __restrict__ zero, int zero_words) {
  // world == 1 (dsts == nullptr): the only destination is this rank's private buffer
  if (blockIdx.x == 0 && blockIdx.y == 0 && zero != nullptr && threadIdx.x < zero_words) zero[threadIdx.x] = 0.0;
  if (blockIdx.y == 2) {
    if (idx == nullptr) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
      const unsigned long long w = idx[i];
      if (idx_dsts == nullptr) own_idx[row0 + i] = w;
      else
        for (int p = 0; p < world; ++p) idx_dsts[p][row0 + i] = w;
    }
    return;
  }
  const float* src = blockIdx.y == 0 ? img : txt;
  const long long ld_src = blockIdx.y == 0 ? ld_img : ld_txt;
  const long long col0 = blockIdx.y == 0 ? 0 : D;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* x = src + static_cast<long long>(row) * ld_src;
  const long long off = (row0 + row) * (2LL * D) + col0;
  const bool vec = (D & 3) == 0 && (ld_src & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  if (vec) {
    for (int d = 4 * lane; d < D; d += 128) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(x + d));
      const uint16_t h0 = f32_to_16<FMT>(t.x), h1 = f32_to_16<FMT>(t.y);
      const uint16_t h2 = f32_to_16<FMT>(t.z), h3 = f32_to_16<FMT>(t.w);
      const uint2 hv = make_uint2(h0 | (static_cast<uint32_t>(h1) << 16), h2 | (static_cast<uint32_t>(h3) << 16));
      if (dsts == nullptr) *reinterpret_cast<uint2*>(own_rows + off + d) = hv;
      else
        for (int p = 0; p < world; ++p) *reinterpret_cast<uint2*>(dsts[p] + off + d) = hv;  // peer stores
    }
  } else {
    for (int d = lane; d < D; d += 32) {
      const uint16_t h = f32_to_16<FMT>(x[d]);
      if (dsts == nullptr) own_rows[off + d] = h;
      else
        for (int p = 0; p < world; ++p) dsts[p][off + d] = h;
    }
  }
}

// Push a contiguous block of 8-byte words (the idx column, models/xvlm.py:285) to every rank's buffer.
__global__ void push_words_kernel(const unsigned long long* __restrict__ src, long long n_words,
                                  unsigned long long* const* __restrict__ dsts, int world, long long dst_word0) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_words;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned long long w = src[i];
    for (int p = 0; p < world; ++p) dsts[p][dst_word0 + i] = w;
  }
}

// Row norms / maxima of an operand that is already 16-bit (e.g. a bf16-stored gallery).
template <int FMT>
__global__ void stats_rows16_kernel(const uint16_t* __restrict__ src, long long ld_src, int n, int D,
                                    float* __restrict__ rn_hi, float* __restrict__ rn_lo,
                                    float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) {
    publish_stats(stats, 0.f, 0.f, 0.f, false);
    return;
  }
  const uint16_t* x = src + static_cast<long long>(row) * ld_src;
  float nh = 0.f, amax = 0.f;
  bool bad = false;
  for (int d = lane; d < D; d += 32) {
    const float t = f16_to_32<FMT>(x[d]);
    bad |= !isfinite(t);
    amax = fmaxf(amax, fabsf(t));
    nh = fmaf(t, t, nh);
  }
  nh = warp_sum(nh);
  amax = warp_max(amax);
  const unsigned anybad = __ballot_sync(0xffffffffu, bad);
  const float a = sqrtf(nh);
  if (lane == 0) {
    if (rn_hi) rn_hi[row] = a;
    if (rn_lo) rn_lo[row] = 0.f;
  }
  publish_stats(stats, a, 0.f, amax, anybad != 0);
}

// 16-bit transpose [n][D] -> [D][ldT] (operand of the gradient products: K runs over rows).
__global__ void transpose16_kernel(const uint16_t* __restrict__ src, long long ld_src, int n, int D,
                                   uint16_t* __restrict__ dst, long long ld_dst) {
  __shared__ uint16_t t[32][34];
  const int r0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, d = d0 + threadIdx.x;
    t[i][threadIdx.x] = (r < n && d < D) ? src[static_cast<long long>(r) * ld_src + d] : 0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int d = d0 + i, r = r0 + threadIdx.x;
    if (d < D && r < ld_dst) dst[static_cast<long long>(d) * ld_dst + r] = (r < n) ? t[threadIdx.x][i] : 0;
  }
}

// Deterministic split-K reduction: out[i] = sum_s parts[s][i] in a fixed order.
__global__ void splitk_reduce_kernel(const float* __restrict__ parts, int n_splits, long long plane,
                                     float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < plane;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int sp = 0; sp < n_splits; ++sp) acc += parts[static_cast<long long>(sp) * plane + i];
    out[i] = acc;
  }
}

// Two equally shaped problems in one launch (blockIdx.y selects), plus an optional scalar product
// o = a * b (dL/dtemp of the contrastive backward) so the backward needs no launch of its own for it.
__global__ void splitk_reduce2_kernel(const float* __restrict__ parts0, const float* __restrict__ parts1, int n_splits,
                                      long long plane, float* __restrict__ out0, float* __restrict__ out1,
                                      const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o) {
  const float* parts = blockIdx.y == 0 ? parts0 : parts1;
  float* out = blockIdx.y == 0 ? out0 : out1;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < plane;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int sp = 0; sp < n_splits; ++sp) acc += parts[static_cast<long long>(sp) * plane + i];
    out[i] = acc;
  }
  if (o != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *o = *a * *b;
}

// Both transposes of the contrastive backward in one launch (blockIdx.z selects the operand).
__global__ void transpose16_pair_kernel(const uint16_t* __restrict__ src0, const uint16_t* __restrict__ src1,
                                        long long ld_src, int n, int D, uint16_t* __restrict__ dst0,
                                        uint16_t* __restrict__ dst1, long long ld_dst) {
  __shared__ uint16_t t[32][34];
  const uint16_t* src = blockIdx.z == 0 ? src0 : src1;
  uint16_t* dst = blockIdx.z == 0 ? dst0 : dst1;
  const int r0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, d = d0 + threadIdx.x;
    t[i][threadIdx.x] = (r < n && d < D) ? src[static_cast<long long>(r) * ld_src + d] : 0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int d = d0 + i, r = r0 + threadIdx.x;
    if (d < D && r < ld_dst) dst[static_cast<long long>(d) * ld_dst + r] = (r < n) ? t[threadIdx.x][i] : 0;
  }
}

// --------------------------------------------------------------------------------
// infonce_finalize: merge the per-chunk partials of EpiLse, emit per-row lse (log2 units) and
// 1/cnt for the backward, and the scalars
//   loss  = (loss_i2t + loss_t2i) / 2                         models/xvlm.py:292
//   dtemp = d loss / d temp = -(1/temp) * 1/2 * sum_p mean_r (E_r[z] - pz_r / cnt_r)
// One block.  Orientation p has n[p] rows and nch[p] chunks.
// --------------------------------------------------------------------------------
struct FinalizeLseParams {
  const float* part[2];
  int n[2];
  int nch[2];
  float* lse2[2];
  float* rcnt[2];
  const float* temp;
  float* out;        // [0] loss, [1] dtemp, [2] loss_p0, [3] loss_p1
  double* scratch;   // [4] partial sums + ticket counter (as a 5th double slot), zeroed by the host side
};

// One thread per (orientation, row); block partial sums are combined with double atomics and the
// last block to finish (ticket) writes the scalars.
__global__ void infonce_finalize_kernel(const FinalizeLseParams P) {
  __shared__ double red[4][8];
  __shared__ bool is_last;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};  // loss_p0, loss_p1, dt_p0, dt_p1
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = gid >= P.n[0] ? 1 : 0;
  const int r = p ? gid - P.n[0] : gid;
  if (r < P.n[p]) {
    const float* q = P.part[p] + static_cast<long long>(r) * P.nch[p] * 5;
    float m = -CUDART_INF_F;
    for (int c = 0; c < P.nch[p]; ++c) m = fmaxf(m, q[5 * c]);
    float l = 0.f, w = 0.f, pz = 0.f, cnt = 0.f;
    for (int c = 0; c < P.nch[p]; ++c) {
      const float s = exp2f(q[5 * c] - m);
      l = fmaf(q[5 * c + 1], s, l);
      w = fmaf(q[5 * c + 2], s, w);
      pz += q[5 * c + 3];
      cnt += q[5 * c + 4];
    }
    const float lse2 = m + log2f(l);
    const float rc = cnt > 0.f ? 1.f / cnt : 0.f;
    P.lse2[p][r] = lse2;
    P.rcnt[p][r] = rc;
    // natural-log units: multiply log2-domain quantities by ln 2
    acc[p] = 0.6931471805599453 * (static_cast<double>(lse2) - static_cast<double>(pz) * rc);
    acc[2 + p] = 0.6931471805599453 * (static_cast<double>(w) / l - static_cast<double>(pz) * rc);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < 4; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    for (int k = 0; k < 4; ++k) {
      double t = 0.0;
      for (int w = 0; w < nw; ++w) t += red[k][w];
      atomicAdd(P.scratch + k, t);
    }
    __threadfence();
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(P.scratch + 4), 1u);
    is_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    volatile double* sc = P.scratch;
    const double l0 = P.n[0] > 0 ? sc[0] / P.n[0] : 0.0;
    const double l1 = P.n[1] > 0 ? sc[1] / P.n[1] : 0.0;
    const double d0 = P.n[0] > 0 ? sc[2] / P.n[0] : 0.0;
    const double d1 = P.n[1] > 0 ? sc[3] / P.n[1] : 0.0;
    const double temp = static_cast<double>(*P.temp);
    P.out[0] = static_cast<float>(0.5 * (l0 + l1));
    P.out[1] = static_cast<float>(-0.5 * (d0 + d1) / temp);
    P.out[2] = static_cast<float>(l0);
    P.out[3] = static_cast<float>(l1);
    P.out[4] = static_cast<float>(-d0 / temp);  // d loss_i2t / d temp, d loss_t2i / d temp (one-directional losses)
    P.out[5] = static_cast<float>(-d1 / temp);
  }
}

// --------------------------------------------------------------------------------
// Strip forward (world > 1): a rank runs the tensor-core pass for ITS rows only -- rows [row_begin, +row_count)
// of both orientations, all n columns -- because a row's log-sum-exp, positives and E_softmax[z] depend on that
// row alone (models/xvlm.py:279-290 are row-wise).  What the backward and the scalar loss need from the other
// ranks is only the per-row statistics: infonce_finalize_local merges this rank's chunk partials and stores
// lse2 / rcnt of its rows into EVERY rank's statistics slot through peer pointers, plus this rank's four partial
// sums; after one barrier infonce_reduce adds the partial sums of all ranks in rank order (same bits on every
// rank) and copies the statistics into the private buffers the backward reads.
// Statistics slot layout: float lse2[2][n] | float rcnt[2][n] | double partial[world][4].
// --------------------------------------------------------------------------------
struct FinalizeLocalParams {
  const float* part[2];
  int nch[2];
  int n, row_begin, row_count;
  float* const* stat_ptrs;  // [world] device pointers to every rank's slot (own rank included)
  int world, rank;
  double* scratch;          // local: 4 partial sums + ticket, zeroed before the launch
};

__global__ void infonce_finalize_local_kernel(const FinalizeLocalParams P) {
  __shared__ double red[4][8];
  __shared__ bool is_last;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = gid >= P.row_count ? 1 : 0;
  const int lr = p ? gid - P.row_count : gid;
  if (lr < P.row_count) {
    const int r = P.row_begin + lr;
    const float* q = P.part[p] + static_cast<long long>(r) * P.nch[p] * 5;
    float m = -CUDART_INF_F;
    for (int c = 0; c < P.nch[p]; ++c) m = fmaxf(m, q[5 * c]);
    float l = 0.f, w = 0.f, pz = 0.f, cnt = 0.f;
    for (int c = 0; c < P.nch[p]; ++c) {
      const float s = exp2f(q[5 * c] - m);
      l = fmaf(q[5 * c + 1], s, l);
      w = fmaf(q[5 * c + 2], s, w);
      pz += q[5 * c + 3];
      cnt += q[5 * c + 4];
    }
    const float lse2 = m + log2f(l);
    const float rc = cnt > 0.f ? 1.f / cnt : 0.f;
    const long long o = static_cast<long long>(p) * P.n + r;
    for (int w2 = 0; w2 < P.world; ++w2) {
      float* st = P.stat_ptrs[w2];
      st[o] = lse2;
      st[2LL * P.n + o] = rc;
    }
    acc[p] = 0.6931471805599453 * (static_cast<double>(lse2) - static_cast<double>(pz) * rc);
    acc[2 + p] = 0.6931471805599453 * (static_cast<double>(w) / l - static_cast<double>(pz) * rc);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < 4; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    for (int k = 0; k < 4; ++k) {
      double t = 0.0;
      for (int w = 0; w < nw; ++w) t += red[k][w];
      atomicAdd(P.scratch + k, t);
    }
    __threadfence();
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(P.scratch + 4), 1u);
    is_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x < 4) {
    __threadfence();
    const double v = *(reinterpret_cast<volatile double*>(P.scratch) + threadIdx.x);
    for (int w2 = 0; w2 < P.world; ++w2) {
      double* part = reinterpret_cast<double*>(P.stat_ptrs[w2] + 4LL * P.n);
      part[4 * P.rank + threadIdx.x] = v;
    }
  }
}

// After the barrier: scalars from the partial sums of all ranks (fixed order), statistics -> private buffers.
__global__ void infonce_reduce_kernel(const float* __restrict__ stat, int n, int world, const float* __restrict__ temp,
                                      float* __restrict__ out, float* __restrict__ lse2, float* __restrict__ rcnt) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nt = gridDim.x * blockDim.x;
  for (int i = tid; i < 2 * n; i += nt) {
    lse2[i] = stat[i];
    rcnt[i] = stat[2LL * n + i];
  }
  if (tid == 0) {
    const double* part = reinterpret_cast<const double*>(stat + 4LL * n);
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int w = 0; w < world; ++w)
      for (int k = 0; k < 4; ++k) s[k] += part[4 * w + k];
    const double l0 = s[0] / n, l1 = s[1] / n, d0 = s[2] / n, d1 = s[3] / n;
    const double t = static_cast<double>(*temp);
    out[0] = static_cast<float>(0.5 * (l0 + l1));
    out[1] = static_cast<float>(-0.5 * (d0 + d1) / t);
    out[2] = static_cast<float>(l0);
    out[3] = static_cast<float>(l1);
    out[4] = static_cast<float>(-d0 / t);
    out[5] = static_cast<float>(-d1 / t);
  }
}

// --------------------------------------------------------------------------------
// topk_finalize: one warp per query row.
//   1. select the row's best KP among the raw candidate lists EpiTopK left per column chunk
//      (threshold by warp-wide bisection on the score, then a 32-wide rank sort)
//   2. emit the top-k (score, column)
//   3. optional exact Recall support: for every ground-truth column g of the row compute the exact
//      score t_g (fp32 dot of the original inputs) and decide rank_g = #{j : t_j > t_g} from the
//      candidate list, re-scoring only candidates whose approximate score is within the rigorous
//      error bound eps of t_g.  rank = min_g rank_g, as image_Retrieval_caption.py:274-278.
//      If the list cannot decide (GT score inside the band of the list's last entry while fewer
//      than kRankCap definitely-greater items are known) the row is flagged for exact_rank_rows.
// --------------------------------------------------------------------------------
constexpr int kRankCap = 10;   // Recall@1/5/10 only ever asks whether rank < 10
constexpr int kMaxSlots = 16;   // candidate slots per lane in topk_finalize: one per list (lists leave the epilogue with <= 32 entries)
constexpr int kFinalizeWarps = 8;

struct TopkFinalizeParams {
  const float* cand_val;  // [n_rows][n_chunks][list_cap]
  const int* cand_idx;
  const int* cand_cnt;    // [n_rows][n_chunks]
  int n_rows, n_cols, n_chunks, KP, k;
  int list_cap;           // stride between lists in entries (EpiTopK::C); the epilogue leaves at most 32 entries in each
  float* topk_val;  // [n_rows][k]
  int* topk_idx;
  // exact recall (all optional; gt_off == nullptr disables)
  const int* gt_off;  // CSR [n_rows + 1]
  const int* gt_ids;
  const void* rows_x;  // original row operand  [n_rows][D]  (fp32 or 16-bit)
  const void* cols_x;  // original col operand  [n_cols][D]
  long long ld_rows, ld_cols;
  int D;
  int x_dtype;  // 0 fp32, 1 fp16, 2 bf16
  const float* rn_hi;       // per-row norms of the 16-bit row operand and of its rounding residual
  const float* rn_lo;
  const float* col_stats;   // kStatWords of the column operand
  float acc_slack;          // fp32 accumulation slack coefficient (times ||hi_r|| ||hi_c||max)
  int* rank;                // [n_rows]  exact when < kRankCap (else a lower bound >= kRankCap)
  int* flag_list;           // rows the list could not decide (need exact_rank_rows)
  int* flag_count;          // [1] zeroed before the launch
  float* gt_score;          // [nnz] exact score of every ground-truth pair (optional)
};

__device__ __forceinline__ float load_x(const void* base, long long off, int dtype) {
  if (dtype == 0) return __ldg(reinterpret_cast<const float*>(base) + off);
  const uint16_t u = __ldg(reinterpret_cast<const uint16_t*>(base) + off);
  return dtype == 1 ? f16_to_32<0>(u) : f16_to_32<1>(u);
}

// A row of the original operands held across a warp: lane l keeps elements 4*(l + 32 i) .. +3.
// Fast path for fp32 with D % 128 == 0, D <= 512 and 16-byte aligned rows.
struct RowRegs {
  float4 q[4];
};
__device__ __forceinline__ bool vec_ok(const void* base, long long ld, int D, int dtype) {
  return dtype == 0 && (D & 127) == 0 && D <= 512 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0;
}
__device__ __forceinline__ void load_row_regs(RowRegs& r, const void* base, long long ld, int row, int D, int lane) {
  const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + static_cast<long long>(row) * ld);
#pragma unroll
  for (int i = 0; i < 4; ++i) r.q[i] = (i * 128 < D) ? __ldg(p + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ float dot_row_regs(const RowRegs& a, const RowRegs& b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s = fmaf(a.q[i].x, b.q[i].x, s);
    s = fmaf(a.q[i].y, b.q[i].y, s);
    s = fmaf(a.q[i].z, b.q[i].z, s);
    s = fmaf(a.q[i].w, b.q[i].w, s);
  }
  return warp_sum(s);
}

// exact fp32 dot of row r of rows_x with row c of cols_x, cooperatively by one warp
__device__ __forceinline__ float warp_dot(const void* rows_x, long long ld_rows, int r, const void* cols_x,
                                          long long ld_cols, int c, int D, int dtype, int lane) {
  float s = 0.f;
  for (int d = lane; d < D; d += 32)
    s = fmaf(load_x(rows_x, static_cast<long long>(r) * ld_rows + d, dtype),
             load_x(cols_x, static_cast<long long>(c) * ld_cols + d, dtype), s);
  return warp_sum(s);
}

// SLOTS = candidate slots per lane this instantiation handles (>= n_lists * list_cap / 32): the
// selection loops are fully unrolled over it, so short merges do not pay for long ones.
template <int SLOTS>
__device__ __forceinline__ void topk_finalize_body(const TopkFinalizeParams& P, int block, float (*s_val)[32],
                                                   int (*s_idx)[32]) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int row = block * kFinalizeWarps + wib;
  if (row >= P.n_rows) return;
  const int KP = P.KP;
  const int sh = 0;                           // slots per list and lane, as a shift: lists arrive with <= 32 entries
  const int per_lane = P.n_chunks << sh;
  const long long cbase = static_cast<long long>(row) * P.n_chunks;
  // all loads are issued before any is consumed: list lengths (one lane each), then every slot
  const int my_cnt = lane < P.n_chunks ? __ldg(P.cand_cnt + cbase + lane) : 0;
  float v[SLOTS];
  int id[SLOTS];
#pragma unroll
  for (int t = 0; t < SLOTS; ++t) {
    v[t] = -CUDART_INF_F;
    id[t] = -1;
    if (t < per_lane) {
      const long long o = (cbase + (t >> sh)) * P.list_cap + lane + 32 * (t & sh);
      v[t] = __ldg(P.cand_val + o);
      id[t] = __ldg(P.cand_idx + o);
    }
  }
  float gmax = -CUDART_INF_F, gmin = CUDART_INF_F;
  int n_loc = 0;
#pragma unroll
  for (int t = 0; t < SLOTS; ++t) {
    const int cnt_t = __shfl_sync(0xffffffffu, my_cnt, t >> sh);
    if (t < per_lane && lane + 32 * (t & sh) < cnt_t) {
      gmax = fmaxf(gmax, v[t]);
      gmin = fminf(gmin, v[t]);
      ++n_loc;
    } else {
      v[t] = -CUDART_INF_F;
      id[t] = -1;
    }
  }
  gmax = warp_max(gmax);
  gmin = warp_min(gmin);
  const int n_tot = __reduce_add_sync(0xffffffffu, n_loc);
  // threshold: KP <= #{v > lo} <= 32 (or everything when there are at most 32 candidates)
  uint32_t lo_k = f32_key(-CUDART_INF_F), hi_k = f32_key(gmax);
  int c_lo = n_tot;
  bool first = true;
  for (int it = 0; it < 40; ++it) {
    if (c_lo <= 32 || hi_k - lo_k <= 1u) break;
    uint32_t mid_k = lo_k + ((hi_k - lo_k) >> 1);
    if (first) {
      const uint32_t vk = f32_key(gmin);
      if (vk > lo_k && vk < hi_k) mid_k = vk;
      first = false;
    }
    const float mid = key_f32(mid_k);
    int cm = 0;
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) cm += (id[t] >= 0 && v[t] > mid) ? 1 : 0;
    cm = __reduce_add_sync(0xffffffffu, cm);
    if (cm >= KP) {
      lo_k = mid_k;
      c_lo = cm;
    } else {
      hi_k = mid_k;
    }
  }
  const bool tie = c_lo > 32;
  const float lo = key_f32(lo_k);
  const float keep_thr = tie ? key_f32(hi_k) : lo;
  s_val[wib][lane] = -CUDART_INF_F;
  s_idx[wib][lane] = -1;
  __syncwarp();
  int base = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int t = 0; t < SLOTS; ++t) {
    const bool sel = id[t] >= 0 && v[t] > keep_thr;
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    const int pos = base + __popc(m & lt_mask);
    if (sel && pos < 32) {
      s_val[wib][pos] = v[t];
      s_idx[wib][pos] = id[t];
    }
    base += __popc(m);
  }
  if (tie) {  // more than 32 equal scores at the threshold: any of them completes the list
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) {
      const bool sel = id[t] >= 0 && !(v[t] > keep_thr) && v[t] > lo;
      const unsigned m = __ballot_sync(0xffffffffu, sel);
      const int pos = base + __popc(m & lt_mask);
      if (sel && pos < 32) {
        s_val[wib][pos] = v[t];
        s_idx[wib][pos] = id[t];
      }
      base += __popc(m);
    }
  }
  __syncwarp();
  const float cv = s_val[wib][lane];
  const int ci = s_idx[wib][lane];
  int rk = 0;
#pragma unroll
  for (int l = 0; l < 32; ++l) {
    const float ov = __shfl_sync(0xffffffffu, cv, l);
    const int oi = __shfl_sync(0xffffffffu, ci, l);
    // descending score; ties: lower column first; empty slots (idx -1) last
    const bool before = (oi >= 0) && (ci < 0 || ov > cv || (ov == cv && oi < ci));
    rk += (before && l != lane) ? 1 : 0;
  }
  if (ci < 0) rk = 32 + lane;  // keep empties out of the way
  __syncwarp();
  if (rk < 32) {
    s_val[wib][rk] = cv;
    s_idx[wib][rk] = ci;
  }
  __syncwarp();
  const int n_sorted = __popc(__ballot_sync(0xffffffffu, ci >= 0));
  float mv = (lane < n_sorted && lane < KP) ? s_val[wib][lane] : -CUDART_INF_F;
  int mi = (lane < n_sorted && lane < KP) ? s_idx[wib][lane] : -1;
  if (lane < P.k) {
    P.topk_val[static_cast<long long>(row) * P.k + lane] = mv;
    P.topk_idx[static_cast<long long>(row) * P.k + lane] = mi;
  }
  if (P.gt_off == nullptr) return;

  // ---- exact rank of the ground truth
  const int m_cnt = __popc(__ballot_sync(0xffffffffu, mi >= 0));  // list length
  const bool exhaustive = m_cnt < KP || P.n_cols <= KP;           // every column is in the list
  const float v_last = __shfl_sync(0xffffffffu, mv, KP - 1);
  const float ch = P.col_stats[0], cl = P.col_stats[1];
  const float rh = P.rn_hi[row], rl = P.rn_lo[row];
  const float eps = rl * ch + rh * cl + rl * cl + P.acc_slack * rh * ch;
  int best = 0x7fffffff;
  int undecided = 0;
  const int g0 = P.gt_off[row], g1 = P.gt_off[row + 1];
  const bool fast = vec_ok(P.rows_x, P.ld_rows, P.D, P.x_dtype) && vec_ok(P.cols_x, P.ld_cols, P.D, P.x_dtype);
  RowRegs qrow;
  if (fast) load_row_regs(qrow, P.rows_x, P.ld_rows, row, P.D, lane);
  for (int gi = g0; gi < g1; ++gi) {
    const int g = P.gt_ids[gi];
    float tg;
    if (fast) {
      RowRegs crow;
      load_row_regs(crow, P.cols_x, P.ld_cols, g, P.D, lane);
      tg = dot_row_regs(qrow, crow);
    } else {
      tg = warp_dot(P.rows_x, P.ld_rows, row, P.cols_x, P.ld_cols, g, P.D, P.x_dtype, lane);
    }
    if (P.gt_score != nullptr && lane == 0) P.gt_score[gi] = tg;
    const bool in_list = (lane < KP) && (mi >= 0) && (mi != g);
    const bool def_gt = in_list && (mv > tg + eps);
    const bool amb = in_list && !def_gt && (mv >= tg - eps);
    int cnt = __popc(__ballot_sync(0xffffffffu, def_gt));
    unsigned amb_mask = __ballot_sync(0xffffffffu, amb);
    const bool gt_inside = exhaustive || (tg - eps > v_last);  // nothing outside the list can beat it
    if (!gt_inside && cnt < kRankCap) {
      undecided = 1;  // cannot bound the competitors outside the list
      continue;
    }
    if (gt_inside) {
      while (amb_mask) {  // re-score the ambiguous candidates exactly
        const int src = __ffs(amb_mask) - 1;
        amb_mask &= amb_mask - 1;
        const int j = __shfl_sync(0xffffffffu, mi, src);
        float tj;
        if (fast) {
          RowRegs crow;
          load_row_regs(crow, P.cols_x, P.ld_cols, j, P.D, lane);
          tj = dot_row_regs(qrow, crow);
        } else {
          tj = warp_dot(P.rows_x, P.ld_rows, row, P.cols_x, P.ld_cols, j, P.D, P.x_dtype, lane);
        }
        cnt += (tj > tg) ? 1 : 0;
      }
    }
    best = min(best, cnt);
  }
  if (lane == 0) {
    // a decided GT makes undecided siblings irrelevant only if it already has rank 0
    const bool flagged = undecided && best > 0;
    const int rk_out = (best == 0x7fffffff) ? kRankCap : best;
    P.rank[row] = rk_out;
    if (flagged) P.flag_list[atomicAdd(P.flag_count, 1)] = row;
  }
}

// --------------------------------------------------------------------------------
// Recall-only evaluation (EpiRank): before the tensor-core pass, one warp per row computes the EXACT score of every
// ground-truth column of the row (fp32 dots of the original inputs), keeps the best one t and the row's rigorous
// bound eps on |16-bit-operand score - exact score| (the one topk_finalize uses), and publishes the band
// (t - eps, t + eps].  Rows without ground truth get rank kRankCap and an empty band.  After the pass,
// rank_resolve re-scores the (row, column) pairs that fell inside a band exactly: one warp per pair.
// --------------------------------------------------------------------------------
struct GtBestParams {
  const int* gt_off;
  const int* gt_ids;
  const void* rows_x;
  const void* cols_x;
  long long ld_rows, ld_cols;
  int D, x_dtype, n_rows;
  const float* rn_hi;
  const float* rn_lo;
  const float* col_stats;
  float acc_slack;
  float* best;      // [n_rows] exact best ground-truth score
  float* lo;        // [n_rows]
  float* hi;        // [n_rows]
  float* gt_score;  // [nnz] or null
  int* rank;        // [n_rows] initialised here: 0, or kRankCap for rows without ground truth
  int* row_flag;    // [n_rows] zeroed here
  int* flag_words;  // 4 words (flag count, grid-barrier word of rank_post, ...) zeroed here
  int* amb_count;   // zeroed here
};

// blockIdx.y = problem (one launch serves both directions of an evaluation)
__global__ void gt_best_kernel(const GtBestParams P0, const GtBestParams P1) {
  const GtBestParams& P = blockIdx.y == 0 ? P0 : P1;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x < 4) {
    P.flag_words[threadIdx.x] = 0;
    if (threadIdx.x == 0) *P.amb_count = 0;
  }
  if (row >= P.n_rows) return;
  if (lane == 0) P.row_flag[row] = 0;
  const int g0 = P.gt_off[row], g1 = P.gt_off[row + 1];
  const bool fast = vec_ok(P.rows_x, P.ld_rows, P.D, P.x_dtype) && vec_ok(P.cols_x, P.ld_cols, P.D, P.x_dtype);
  RowRegs qrow;
  if (fast) load_row_regs(qrow, P.rows_x, P.ld_rows, row, P.D, lane);
  float best = -CUDART_INF_F;
  for (int gi = g0; gi < g1; ++gi) {
    const int g = P.gt_ids[gi];
    float tg;
    if (fast) {
      RowRegs crow;
      load_row_regs(crow, P.cols_x, P.ld_cols, g, P.D, lane);
      tg = dot_row_regs(qrow, crow);
    } else {
      tg = warp_dot(P.rows_x, P.ld_rows, row, P.cols_x, P.ld_cols, g, P.D, P.x_dtype, lane);
    }
    if (P.gt_score != nullptr && lane == 0) P.gt_score[gi] = tg;
    best = fmaxf(best, tg);
  }
  if (lane == 0) {
    const float ch = P.col_stats[0], cl = P.col_stats[1];
    const float rh = P.rn_hi[row], rl = P.rn_lo[row];
    const float eps = rl * ch + rh * cl + rl * cl + P.acc_slack * rh * ch;
    const bool has = g1 > g0;
    P.best[row] = has ? best : CUDART_INF_F;
    P.lo[row] = has ? best - eps : CUDART_INF_F;
    P.hi[row] = has ? best + eps : CUDART_INF_F;
    P.rank[row] = has ? 0 : kRankCap;
  }
}

struct ResolveParams {
  const int2* pairs;
  const int* n_pairs;
  int cap;
  const void* rows_x;
  const void* cols_x;
  long long ld_rows, ld_cols;
  int D, x_dtype;
  const float* best;
  int* rank;
};

// blockIdx.y = problem
__global__ void rank_resolve_kernel(const ResolveParams P0, const ResolveParams P1) {
  const ResolveParams& P = blockIdx.y == 0 ? P0 : P1;
  const int lane = threadIdx.x & 31;
  const int n = min(*P.n_pairs, P.cap);
  // the SAME dot routine (same summation order) as gt_best_kernel: a ground-truth column must tie with itself
  const bool fast = vec_ok(P.rows_x, P.ld_rows, P.D, P.x_dtype) && vec_ok(P.cols_x, P.ld_cols, P.D, P.x_dtype);
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += gridDim.x * (blockDim.x >> 5)) {
    const int2 pr = P.pairs[i];
    if (P.rank[pr.x] >= kRankCap) continue;  // decided by the definite counts alone (monotone: pairs only add)
    float tj;
    if (fast) {
      RowRegs qrow, crow;
      load_row_regs(qrow, P.rows_x, P.ld_rows, pr.x, P.D, lane);
      load_row_regs(crow, P.cols_x, P.ld_cols, pr.y, P.D, lane);
      tj = dot_row_regs(qrow, crow);
    } else {
      tj = warp_dot(P.rows_x, P.ld_rows, pr.x, P.cols_x, P.ld_cols, pr.y, P.D, P.x_dtype, lane);
    }
    if (lane == 0 && tj > P.best[pr.x]) atomicAdd(P.rank + pr.x, 1);
  }
}

template <int SLOTS>
__global__ void __launch_bounds__(kFinalizeWarps * 32) topk_finalize_kernel(const TopkFinalizeParams P) {
  __shared__ float s_val[kFinalizeWarps][32];
  __shared__ int s_idx[kFinalizeWarps][32];
  topk_finalize_body<SLOTS>(P, blockIdx.x, s_val, s_idx);
}

// Both problems of a launch in one grid: blocks [0, blocks0) finalize problem 0 with S0 slots, the rest problem 1
// with S1 (one launch instead of two, and the short one no longer runs a wave of its own).
template <int S0, int S1>
__global__ void __launch_bounds__(kFinalizeWarps * 32)
topk_finalize_pair_kernel(const TopkFinalizeParams P0, const TopkFinalizeParams P1, int blocks0) {
  __shared__ float s_val[kFinalizeWarps][32];
  __shared__ int s_idx[kFinalizeWarps][32];
  if (static_cast<int>(blockIdx.x) < blocks0) topk_finalize_body<S0>(P0, blockIdx.x, s_val, s_idx);
  else topk_finalize_body<S1>(P1, blockIdx.x - blocks0, s_val, s_idx);
}

// Exact fallback for flagged rows: rank = min_g #{j : t_j > t_g} with every score an fp32 dot.
// One block per flagged row (grid-stride over the list; normally the list is empty).
__device__ __forceinline__ void exact_rank_rows(const TopkFinalizeParams& P, float* s_tg, int* s_cnt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int n_flagged = *P.flag_count;
  for (int fi = blockIdx.x; fi < n_flagged; fi += gridDim.x) {
    const int row = P.flag_list[fi];
    const int g0 = P.gt_off[row];
    const int ng_all = P.gt_off[row + 1] - g0;
    int best = 0x7fffffff;  // thread 0's running minimum over the ground-truth chunks
    // any number of ground-truth entries per row (image_Retrieval_caption.py:274-278): chunks of kGtChunk
    for (int gb = 0; gb < ng_all; gb += 16) {
      const int ng = min(ng_all - gb, 16);
      __syncthreads();
      if (warp == 0) {
        for (int gi = 0; gi < ng; ++gi) {
          const float tg = warp_dot(P.rows_x, P.ld_rows, row, P.cols_x, P.ld_cols, P.gt_ids[g0 + gb + gi],
                                    P.D, P.x_dtype, lane);
          if (lane == 0) {
            s_tg[gi] = tg;
            s_cnt[gi] = 0;
          }
        }
      }
      __syncthreads();
      int cnt[16];
#pragma unroll
      for (int gi = 0; gi < 16; ++gi) cnt[gi] = 0;
      for (int j = warp; j < P.n_cols; j += nw) {
        const float tj = warp_dot(P.rows_x, P.ld_rows, row, P.cols_x, P.ld_cols, j, P.D, P.x_dtype, lane);
#pragma unroll
        for (int gi = 0; gi < 16; ++gi)
          if (gi < ng && tj > s_tg[gi]) ++cnt[gi];
      }
      if (lane == 0) {
#pragma unroll
        for (int gi = 0; gi < 16; ++gi)
          if (gi < ng && cnt[gi]) atomicAdd(&s_cnt[gi], cnt[gi]);
      }
      __syncthreads();
      if (threadIdx.x == 0)
        for (int gi = 0; gi < ng; ++gi) best = min(best, s_cnt[gi]);
    }
    if (threadIdx.x == 0) P.rank[row] = best == 0x7fffffff ? kRankCap : best;
  }
}

__global__ void exact_rank_rows_kernel(const TopkFinalizeParams P) {
  __shared__ float s_tg[16];
  __shared__ int s_cnt[16];
  exact_rank_rows(P, s_tg, s_cnt);
}

// What follows topk_finalize, for up to two problems in ONE launch (blockIdx.y = problem): the exact fallback for
// the flagged rows (normally none), then #{rank < 1, 5, 10}.  Counting has to see the fallback's ranks: only
// when rows WERE flagged the blocks of a problem meet at a grid barrier (flag_count[1], zeroed with the flag
// count; every launch keeps its grid co-resident: at most 2 blocks of 256 threads per SM).
__global__ void rank_post_kernel(const TopkFinalizeParams P0, const TopkFinalizeParams P1, int* __restrict__ counts0,
                                 int* __restrict__ counts1) {
  __shared__ float s_tg[16];
  __shared__ int s_cnt[16];
  const TopkFinalizeParams& P = blockIdx.y == 0 ? P0 : P1;
  int* counts = blockIdx.y == 0 ? counts0 : counts1;
  if (P.gt_off == nullptr) return;
  const int n_flagged = *P.flag_count;
  if (n_flagged > 0) {
    exact_rank_rows(P, s_tg, s_cnt);
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      unsigned* bar = reinterpret_cast<unsigned*>(P.flag_count) + 1;
      atomicAdd(bar, 1u);
      unsigned seen;
      long long spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
        if (++spins > (1LL << 28)) __trap();
      } while (seen < gridDim.x);
    }
    __syncthreads();
  }
  if (counts == nullptr) return;
  int c1 = 0, c5 = 0, c10 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n_rows; i += gridDim.x * blockDim.x) {
    const int r = P.rank[i];
    c1 += r < 1;
    c5 += r < 5;
    c10 += r < 10;
  }
  for (int o = 16; o > 0; o >>= 1) {
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    c5 += __shfl_xor_sync(0xffffffffu, c5, o);
    c10 += __shfl_xor_sync(0xffffffffu, c10, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c1) atomicAdd(counts + 0, c1);
    if (c5) atomicAdd(counts + 1, c5);
    if (c10) atomicAdd(counts + 2, c10);
  }
}

// --------------------------------------------------------------------------------
// rank_rows / rank_cols: ranking of an already materialised fp32 score matrix (the drop-in
// itm_eval, image_Retrieval_caption.py:261-295).  HBM-bound: one pass over the matrix.
//   rank_rows: for row r, rank = min_{g in gt(r)} #{c : S[r][c] > S[r][g]}         (i2t)
//   rank_cols: for column c, rank = min_{g in gt(c)} #{r : S[r][c] > S[g][c]}      (t2i read
//              from the same row-major matrix; the reference's t2i matrix is its transpose view)
// --------------------------------------------------------------------------------
constexpr int kGtChunk = 16;  // ground-truth entries handled per sweep; longer lists take more sweeps

__global__ void rank_rows_kernel(const float* __restrict__ S, long long ld, int R, int C,
                                 const int* __restrict__ gt_off, const int* __restrict__ gt_ids,
                                 int* __restrict__ rank) {
  __shared__ float s_t[kGtChunk];
  __shared__ int s_c[kGtChunk];
  const int row = blockIdx.x;
  if (row >= R) return;
  const float* s = S + static_cast<long long>(row) * ld;
  const int g0 = gt_off[row];
  const int ng_all = gt_off[row + 1] - g0;
  int best = 0x7fffffff;  // thread 0's minimum over all ground-truth entries (:274-278 takes the min over every one)
  for (int gb = 0; gb < ng_all; gb += kGtChunk) {
    const int ng = min(ng_all - gb, kGtChunk);
    __syncthreads();
    if (threadIdx.x < ng) {
      s_t[threadIdx.x] = s[gt_ids[g0 + gb + threadIdx.x]];
      s_c[threadIdx.x] = 0;
    }
    __syncthreads();
    int cnt[kGtChunk];
#pragma unroll
    for (int g = 0; g < kGtChunk; ++g) cnt[g] = 0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float v = __ldg(s + c);
#pragma unroll
      for (int g = 0; g < kGtChunk; ++g)
        if (g < ng && v > s_t[g]) ++cnt[g];
    }
#pragma unroll
    for (int g = 0; g < kGtChunk; ++g) {
      if (g < ng) {
        int v = cnt[g];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_c[g], v);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int g = 0; g < ng; ++g) best = min(best, s_c[g]);
  }
  if (threadIdx.x == 0) rank[row] = best == 0x7fffffff ? C : best;
}

// One thread per column, rows split over blockIdx.y; partial counts combined with atomicAdd into
// cnt[nnz] (zero-initialised), then rank_cols_min takes the min over each column's GT.
__global__ void rank_cols_count_kernel(const float* __restrict__ S, long long ld, int R, int C,
                                       const int* __restrict__ gt_off, const int* __restrict__ gt_ids,
                                       int rows_per_block, int* __restrict__ cnt_nnz) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int g0 = gt_off[c];
  const int ng_all = gt_off[c + 1] - g0;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  for (int gb = 0; gb < ng_all; gb += kGtChunk) {  // one sweep per kGtChunk ground-truth entries (normally one)
    const int ng = min(ng_all - gb, kGtChunk);
    float t[kGtChunk];
    int cnt[kGtChunk];
#pragma unroll
    for (int g = 0; g < kGtChunk; ++g) {
      cnt[g] = 0;
      t[g] = (g < ng) ? __ldg(S + static_cast<long long>(gt_ids[g0 + gb + g]) * ld + c) : CUDART_INF_F;
    }
    for (int r = r0; r < r1; ++r) {
      const float v = __ldg(S + static_cast<long long>(r) * ld + c);
#pragma unroll
      for (int g = 0; g < kGtChunk; ++g) cnt[g] += (v > t[g]) ? 1 : 0;
    }
#pragma unroll
    for (int g = 0; g < kGtChunk; ++g)
      if (g < ng && cnt[g]) atomicAdd(cnt_nnz + g0 + gb + g, cnt[g]);
  }
}

__global__ void rank_cols_min_kernel(int C, int R, const int* __restrict__ gt_off,
                                     const int* __restrict__ cnt_nnz, int* __restrict__ rank) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  int best = 0x7fffffff;
  const int g1 = gt_off[c + 1];
  for (int g = gt_off[c]; g < g1; ++g) best = min(best, cnt_nnz[g]);
  rank[c] = best == 0x7fffffff ? R : best;
}

// Recall@1/5/10 counts from ranks: out[0..2] += #{rank < 1, 5, 10}.
__global__ void recall_count_kernel(const int* __restrict__ rank, int n, int* __restrict__ out) {
  int c1 = 0, c5 = 0, c10 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = rank[i];
    c1 += r < 1;
    c5 += r < 5;
    c10 += r < 10;
  }
  for (int o = 16; o > 0; o >>= 1) {
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    c5 += __shfl_xor_sync(0xffffffffu, c5, o);
    c10 += __shfl_xor_sync(0xffffffffu, c10, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c1) atomicAdd(out + 0, c1);
    if (c5) atomicAdd(out + 1, c5);
    if (c10) atomicAdd(out + 2, c10);
  }
}

// --------------------------------------------------------------------------------
// double_sim (video_Retrieval_caption_double_sim.py:87-91,170-179)
//   C = max_n Cn ; global min / max of S and of C ; fused = w1 * norm(S) + w2 * norm(C)
//   norm(x) = -(( -x - min(-x)) / max(-x - min(-x))) == (x - max x) / (max x - min x)
// mm[0..3] = ordered-uint {max S, min S, max C, min C}; init {0, ~0, 0, ~0}.
// --------------------------------------------------------------------------------
__global__ void capmax_minmax_kernel(const float* __restrict__ S, const float* __restrict__ Cn,
                                     float* __restrict__ Cmax, int n_cap, long long numel,
                                     unsigned* __restrict__ mm) {
  float smax = -CUDART_INF_F, smin = CUDART_INF_F, cmax = -CUDART_INF_F, cmin = CUDART_INF_F;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float s = S[i];
    smax = fmaxf(smax, s);
    smin = fminf(smin, s);
    if (Cn != nullptr) {
      float c = Cn[i];
      for (int k = 1; k < n_cap; ++k) c = fmaxf(c, Cn[static_cast<long long>(k) * numel + i]);
      Cmax[i] = c;
      cmax = fmaxf(cmax, c);
      cmin = fminf(cmin, c);
    }
  }
  smax = warp_max(smax);
  smin = warp_min(smin);
  cmax = warp_max(cmax);
  cmin = warp_min(cmin);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(mm + 0, f32_ordered(smax));
    atomicMin(mm + 1, f32_ordered(smin));
    if (Cn != nullptr) {
      atomicMax(mm + 2, f32_ordered(cmax));
      atomicMin(mm + 3, f32_ordered(cmin));
    }
  }
}

__global__ void mm_init_kernel(unsigned* mm) {
  if (threadIdx.x < 4) mm[threadIdx.x] = (threadIdx.x & 1) ? 0xffffffffu : 0u;
}

// mode 1: norm fusion, mode 2: raw fusion (image_Retrieval_caption.py:244-246).
// Same fp32 operation order as the reference (no FMA contraction).
__global__ void fuse_scores_kernel(float* __restrict__ S, const float* __restrict__ Cmax, long long numel,
                                   const unsigned* __restrict__ mm, float w1, float w2, int mode) {
  float smax = 0.f, sden = 1.f, cmax = 0.f, cden = 1.f;
  if (mode == 1) {
    smax = ordered_f32(mm[0]);
    sden = __fsub_rn(smax, ordered_f32(mm[1]));
    cmax = ordered_f32(mm[2]);
    cden = __fsub_rn(cmax, ordered_f32(mm[3]));
  }
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float a = S[i], b = Cmax[i];
    if (mode == 1) {
      a = -__fdiv_rn(__fsub_rn(smax, a), sden);
      b = -__fdiv_rn(__fsub_rn(cmax, b), cden);
    }
    S[i] = __fadd_rn(__fmul_rn(w1, a), __fmul_rn(w2, b));
  }
}

// Ground-truth scores of the fused double_sim evaluation (between its two tensor-core passes): thread t < n_txt
// fuses (S, max C) of text t at its ground-truth video -> the score text t has to beat in the t2i direction;
// thread n_txt + i takes the best of video i's ground-truth texts (rank = min over the ground truth =
// number of scores above the BEST ground-truth score, image_Retrieval_caption.py:274-278).
__global__ void ds_gt_scores_kernel(const float* __restrict__ gt_s, const float* __restrict__ gt_c, int n_txt,
                                    const int* __restrict__ vid_off, const int* __restrict__ vid_ids, int n_vid,
                                    const unsigned* __restrict__ mm, float w1, float w2, int mode,
                                    float* __restrict__ txt_gt, float* __restrict__ vid_best) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_txt + n_vid) return;
  const DsFuse fuse = ds_fuse_load(mm, w1, w2, mode);
  if (t < n_txt) {
    txt_gt[t] = fuse(gt_s[t], gt_c[t]);
  } else {
    const int i = t - n_txt;
    float best = -CUDART_INF_F;
    for (int e = vid_off[i]; e < vid_off[i + 1]; ++e) {
      const int g = vid_ids[e];
      best = fmaxf(best, fuse(gt_s[g], gt_c[g]));
    }
    vid_best[i] = vid_off[i + 1] > vid_off[i] ? best : CUDART_INF_F;
  }
}

// --------------------------------------------------------------------------------
// Cross-rank exchange helpers over peer memory (one node, NVLink / NVSwitch).
// --------------------------------------------------------------------------------
// Barrier over peer-visible flag words.  flags[p] -> rank p's block of `world` uint32 (slot r is written by
// rank r only).  Thread p publishes `epoch` into peer p's slot [rank] (release, system scope: everything
// this stream did before -- including peer stores of earlier kernels -- is visible to whoever acquires it)
// and waits until peer p's epoch has arrived in the own block.  Epochs only grow.  The wait is bounded by
// WALL TIME (timeout_ns, %globaltimer; 0 = wait for ever, as NCCL would): ranks of a training job may be
// minutes apart (a rank-local checkpoint write, image_Retrieval_caption.py:478-499, has no dist.barrier after
// it), so the default is minutes, not seconds; only a peer that is really gone traps.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// epoch == 0: the epoch comes from a device-resident counter (word kBarrierCounterWord of the rank's own flag
// block, advanced by the kernel), so the launch has no per-call argument and a training step that contains
// barriers can be captured in a CUDA graph and replayed.
constexpr int kBarrierCounterWord = 48;
__global__ void peer_barrier_kernel(unsigned* const* __restrict__ flags, int world, int rank, unsigned epoch,
                                    unsigned long long timeout_ns) {
  __shared__ unsigned s_epoch;
  const int p = threadIdx.x;
  if (epoch == 0u) {
    if (p == 0) {
      unsigned* counter = flags[rank] + kBarrierCounterWord;
      s_epoch = *counter + 1u;
      *counter = s_epoch;
    }
    __syncthreads();
    epoch = s_epoch;
  }
  if (p >= world) return;
  __threadfence_system();
  unsigned* theirs = flags[p] + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
  const unsigned* mine = flags[rank] + p;
  unsigned seen = 0;
  unsigned spins = 0;
  const unsigned long long t0 = global_ns();
  for (;;) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
    if (static_cast<int>(seen - epoch) >= 0) break;
    if ((++spins & 1023u) == 0 && timeout_ns != 0 && global_ns() - t0 > timeout_ns) {
      printf("leccr: peer barrier timed out (rank %d waiting for rank %d, epoch %u, seen %u)\n", rank, p, epoch, seen);
      __trap();
    }
    __nanosleep(spins < 4096u ? 64 : 1000);
  }
}

// Gather + merge in one kernel for the row-partitioned gallery: every rank keeps its per-query local top-k
// lists (descending, ties by lower column) in a peer-visible buffer; after a barrier each rank PULLS the
// lists of its query slice from all ranks (coalesced peer loads staged through shared memory, one warp per
// 32 consecutive queries) and each thread W-way merges one query.  Global column = local + off[p].
// Ties across ranks: lower global column first (the order a single pass over the whole gallery gives).
constexpr int kMergeMaxWorld = 8;
struct MergeOffsets {
  int off[kMergeMaxWorld];
};
__global__ void topk_merge_peers_kernel(const float* const* __restrict__ vals, const int* const* __restrict__ idxs,
                                        int world, int k_in, long long q_begin, long long q_count, MergeOffsets offs,
                                        int k_out, float* __restrict__ out_val, int* __restrict__ out_idx) {
  extern __shared__ float merge_smem[];
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q0 = (static_cast<long long>(blockIdx.x) * warps + warp) * 32;  // first query of this warp (slice-relative)
  if (q0 >= q_count) return;
  const int nq = static_cast<int>(min(32LL, q_count - q0));
  const int seg = 32 * k_in;  // words per (rank, warp) segment
  float* sv = merge_smem + static_cast<size_t>(warp) * world * seg * 2;
  int* si = reinterpret_cast<int*>(sv + static_cast<size_t>(world) * seg);
  for (int p = 0; p < world; ++p) {
    const float* gv = vals[p] + (q_begin + q0) * k_in;
    const int* gi = idxs[p] + (q_begin + q0) * k_in;
    for (int t = lane; t < nq * k_in; t += 32) {
      sv[p * seg + t] = gv[t];
      si[p * seg + t] = gi[t] + offs.off[p];
    }
  }
  __syncwarp();
  if (lane >= nq) return;
  int pos[kMergeMaxWorld];
  float hv[kMergeMaxWorld];
  int hi[kMergeMaxWorld];
#pragma unroll
  for (int p = 0; p < kMergeMaxWorld; ++p) {
    pos[p] = 0;
    const bool ok = p < world;
    hv[p] = ok ? sv[p * seg + lane * k_in] : -CUDART_INF_F;
    hi[p] = ok ? si[p * seg + lane * k_in] : 0x7fffffff;
  }
  const long long o = (q0 + lane) * k_out;
  for (int s = 0; s < k_out; ++s) {
    int best = 0;
    float bv = hv[0];
    int bi = hi[0];
#pragma unroll
    for (int p = 1; p < kMergeMaxWorld; ++p) {
      const bool better = hv[p] > bv || (hv[p] == bv && hi[p] < bi);
      if (better) {
        best = p;
        bv = hv[p];
        bi = hi[p];
      }
    }
    out_val[o + s] = bv;
    out_idx[o + s] = bi;
#pragma unroll
    for (int p = 0; p < kMergeMaxWorld; ++p) {
      if (p == best) {
        const int np = ++pos[p];
        const bool ok = np < k_in && p < world;
        hv[p] = ok ? sv[p * seg + lane * k_in + np] : -CUDART_INF_F;
        hi[p] = ok ? si[p * seg + lane * k_in + np] : 0x7fffffff;
      }
    }
  }
}

__global__ void scalar_product_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o) {
  if (threadIdx.x == 0) *o = *a * *b;
}

// --------------------------------------------------------------------------------
// Caption contrastive loss: models/model_retrieval_caption.py:145-152
//   sim    = caption.reshape(n * B, d) @ text.T          (tensor cores, split precision, materialised: n*B*B small)
//   logits = max_n sim / temp ; labels = arange(B) ; loss = (CE(logits) + CE(logits.T)) / 2
// The kernels below take the materialised sim [n][B][B] and produce the max / argmax, both log-sum-exp
// families (rows and columns, log2 units) with E_softmax[z] for d loss / d temp, the loss, and in the backward
// the gradient strip G' routed to the arg-max plane (what autograd's max does).
// --------------------------------------------------------------------------------
__device__ __forceinline__ void lse_combine(float& m, float& l, float& w, float m2, float l2, float w2) {
  const float mn = fmaxf(m, m2);
  const float a = (m == -CUDART_INF_F) ? 0.f : exp2f(m - mn), b = (m2 == -CUDART_INF_F) ? 0.f : exp2f(m2 - mn);
  l = l * a + l2 * b;
  w = w * a + w2 * b;
  m = mn;
}

// One block per row i: L[i][j] = max_a S[a][i][j], amax = first arg max (torch.max semantics on ties),
// row log-sum-exp of z2 = L * log2e / temp and sum_j softmax_ij * z2_ij.
__global__ void capmax_rows_kernel(const float* __restrict__ S, int n_cap, int B, const float* __restrict__ temp,
                                   float* __restrict__ L, unsigned char* __restrict__ amax,
                                   float* __restrict__ lse_row, float* __restrict__ e_row) {
  __shared__ float sm[8], sl[8], sw[8];
  const int i = blockIdx.x;
  const float sc = 1.4426950408889634f / __ldg(temp);
  float m = -CUDART_INF_F, l = 0.f, w = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    float best = S[(static_cast<long long>(0) * B + i) * B + j];
    int arg = 0;
    for (int a = 1; a < n_cap; ++a) {
      const float v = S[(static_cast<long long>(a) * B + i) * B + j];
      if (v > best) {
        best = v;
        arg = a;
      }
    }
    L[static_cast<long long>(i) * B + j] = best;
    amax[static_cast<long long>(i) * B + j] = static_cast<unsigned char>(arg);
    const float z = best * sc;
    lse_combine(m, l, w, z, 1.f, z);
  }
  for (int o = 16; o > 0; o >>= 1)
    lse_combine(m, l, w, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, l, o),
                __shfl_xor_sync(0xffffffffu, w, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sm[warp] = m;
    sl[warp] = l;
    sw[warp] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (blockDim.x >> 5); ++k) lse_combine(m, l, w, sm[k], sl[k], sw[k]);
    lse_row[i] = m + log2f(l);
    e_row[i] = w / l;
  }
}

// Columns: block = 32 columns x 8 row phases; same statistics down the columns of L.
__global__ void capmax_cols_kernel(const float* __restrict__ L, int B, const float* __restrict__ temp,
                                   float* __restrict__ lse_col, float* __restrict__ e_col) {
  __shared__ float sm[8][33], sl[8][33], sw[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const float sc = 1.4426950408889634f / __ldg(temp);
  float m = -CUDART_INF_F, l = 0.f, w = 0.f;
  if (j < B)
    for (int i = threadIdx.y; i < B; i += blockDim.y) {
      const float z = L[static_cast<long long>(i) * B + j] * sc;
      lse_combine(m, l, w, z, 1.f, z);
    }
  sm[threadIdx.y][threadIdx.x] = m;
  sl[threadIdx.y][threadIdx.x] = l;
  sw[threadIdx.y][threadIdx.x] = w;
  __syncthreads();
  if (threadIdx.y == 0 && j < B) {
    for (int k = 1; k < blockDim.y; ++k) lse_combine(m, l, w, sm[k][threadIdx.x], sl[k][threadIdx.x], sw[k][threadIdx.x]);
    lse_col[j] = m + log2f(l);
    e_col[j] = w / l;
  }
}

// One block: loss and d loss / d temp (per unit upstream gradient), fp64 sums.
//   loss  = 1/(2B) sum_i [(lse_row_i - z_ii) + (lse_col_i - z_ii)]
//   dtemp = -(1/temp) [ 1/(2B) (sum_i E_row_i[z] + sum_j E_col_j[z]) - 1/B sum_i z_ii ]
__global__ void caploss_finalize_kernel(const float* __restrict__ L, int B, const float* __restrict__ temp,
                                        const float* __restrict__ lse_row, const float* __restrict__ lse_col,
                                        const float* __restrict__ e_row, const float* __restrict__ e_col,
                                        float* __restrict__ out) {
  __shared__ double r0[8], r1[8];
  const double sc = 1.4426950408889634 / static_cast<double>(*temp);
  double a0 = 0.0, a1 = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const double zii = static_cast<double>(L[static_cast<long long>(i) * B + i]) * sc;
    a0 += static_cast<double>(lse_row[i]) + static_cast<double>(lse_col[i]) - 2.0 * zii;
    a1 += static_cast<double>(e_row[i]) + static_cast<double>(e_col[i]) - 2.0 * zii;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    r0[threadIdx.x >> 5] = a0;
    r1[threadIdx.x >> 5] = a1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int k = 0; k < (blockDim.x >> 5); ++k) {
      t0 += r0[k];
      t1 += r1[k];
    }
    const double ln2 = 0.6931471805599453;
    out[0] = static_cast<float>(ln2 * t0 / (2.0 * B));
    out[1] = static_cast<float>(-ln2 * t1 / (2.0 * B) / static_cast<double>(*temp));
  }
}

// Backward strip: G'_ij = softmax_row_ij + softmax_col_ij - 2 delta_ij in [-2, 2], stored 16-bit in the plane of
// the arg max ([n][B][ld], other planes and the pad columns zero); scale[0] = grad_out / (2 B temp) is applied
// in fp32 by the gradient products' epilogue.
template <int FMT>
__global__ void capgrad_kernel(const float* __restrict__ L, const unsigned char* __restrict__ amax, int n_cap, int B,
                               int ld, const float* __restrict__ temp, const float* __restrict__ lse_row,
                               const float* __restrict__ lse_col, const float* __restrict__ grad_out,
                               uint16_t* __restrict__ G, float* __restrict__ scale) {
  const int i = blockIdx.x;
  const float t = __ldg(temp);
  const float sc = 1.4426950408889634f / t;
  if (i == 0 && threadIdx.x == 0) scale[0] = __ldg(grad_out) / (2.f * static_cast<float>(B) * t);
  const float lr = lse_row[i];
  for (int j = threadIdx.x; j < ld; j += blockDim.x) {
    float g = 0.f;
    int arg = -1;
    if (j < B) {
      const float z = L[static_cast<long long>(i) * B + j] * sc;
      g = exp2f(z - lr) + exp2f(z - lse_col[j]) - (i == j ? 2.f : 0.f);
      arg = amax[static_cast<long long>(i) * B + j];
    }
    for (int a = 0; a < n_cap; ++a)
      G[(static_cast<long long>(a) * B + i) * ld + j] = f32_to_16<FMT>(a == arg ? g : 0.f);
  }
}

// --------------------------------------------------------------------------------
// Top-k of the rows (or, through the strides, the columns) of a MATERIALISED score matrix: the double_sim
// fusion produces one (video_Retrieval_caption_double_sim.py:178; cfg4 = 1000 x 1000), and fused_eval returns
// top-k lists for it like for the plain similarity.  One warp per row: every lane keeps a sorted top-KMAX of
// its strided share, then k rounds of a warp arg-max pop the winners.  Order: score descending, ties by
// lower column (the rule of topk_finalize).
// --------------------------------------------------------------------------------
constexpr int kDenseTopkMax = 16;
__global__ void topk_dense_kernel(const float* __restrict__ S, long long ld_r, long long ld_c, int R, int C, int k,
                                  float* __restrict__ out_val, int* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  float v[kDenseTopkMax];
  int id[kDenseTopkMax];
#pragma unroll
  for (int i = 0; i < kDenseTopkMax; ++i) {
    v[i] = -CUDART_INF_F;
    id[i] = 0x7fffffff;
  }
  const float* base = S + static_cast<long long>(row) * ld_r;
  for (int c = lane; c < C; c += 32) {
    const float x = base[static_cast<long long>(c) * ld_c];
    if (x > v[kDenseTopkMax - 1]) {  // columns arrive in increasing order: a tie never displaces an earlier column
      v[kDenseTopkMax - 1] = x;
      id[kDenseTopkMax - 1] = c;
#pragma unroll
      for (int i = kDenseTopkMax - 1; i > 0; --i) {
        const bool sw = v[i] > v[i - 1];
        const float tv = v[i];
        const int ti = id[i];
        if (sw) {
          v[i] = v[i - 1];
          id[i] = id[i - 1];
          v[i - 1] = tv;
          id[i - 1] = ti;
        }
      }
    }
  }
  for (int s = 0; s < k; ++s) {
    // the warp's best head: (score desc, column asc)
    float bv = v[0];
    int bi = id[0];
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      out_val[static_cast<long long>(row) * k + s] = bv;
      out_idx[static_cast<long long>(row) * k + s] = bi == 0x7fffffff ? -1 : bi;
    }
    if (id[0] == bi && bi != 0x7fffffff) {  // the owner pops its head (columns are unique across lanes)
#pragma unroll
      for (int i = 0; i < kDenseTopkMax - 1; ++i) {
        v[i] = v[i + 1];
        id[i] = id[i + 1];
      }
      v[kDenseTopkMax - 1] = -CUDART_INF_F;
      id[kDenseTopkMax - 1] = 0x7fffffff;
    }
  }
}

// --------------------------------------------------------------------------------
// dstl_loss: models/model_retrieval_caption.py:94-116 on the gathered tensors
//   labels = softmax_rows(alpha * norm(text_s image^T) + (1 - alpha) * norm(max_n caption_n text_s^T))
//   loss   = KL(labels || softmax_rows(text_t image^T)), reduction batchmean
// F (the fused label logits, from leccr_double_sim_fuse: its norm differs from the reference's positive variant
// by the constant +1 per term, which a row softmax cancels) and TV = text_t image^T are materialised [N][N].
// --------------------------------------------------------------------------------
// One block per row: both log-sum-exps (natural units), then the row's KL term.
__global__ void dstl_rows_kernel(const float* __restrict__ Fm, const float* __restrict__ TV, int N,
                                 float* __restrict__ lse_f, float* __restrict__ lse_t, float* __restrict__ row_loss) {
  __shared__ float sm[2][8], sl[2][8];
  __shared__ float s_lse[2];
  __shared__ float s_red[8];
  const int r = blockIdx.x;
  const float* f = Fm + static_cast<long long>(r) * N;
  const float* t = TV + static_cast<long long>(r) * N;
  const float L2E = 1.4426950408889634f;
  float mf = -CUDART_INF_F, lf = 0.f, mt = -CUDART_INF_F, lt = 0.f, dummy = 0.f, d2 = 0.f;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    lse_combine(mf, lf, dummy, f[c] * L2E, 1.f, 0.f);
    lse_combine(mt, lt, d2, t[c] * L2E, 1.f, 0.f);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lse_combine(mf, lf, dummy, __shfl_xor_sync(0xffffffffu, mf, o), __shfl_xor_sync(0xffffffffu, lf, o), 0.f);
    lse_combine(mt, lt, d2, __shfl_xor_sync(0xffffffffu, mt, o), __shfl_xor_sync(0xffffffffu, lt, o), 0.f);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sm[0][warp] = mf;
    sl[0][warp] = lf;
    sm[1][warp] = mt;
    sl[1][warp] = lt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (blockDim.x >> 5); ++k) {
      lse_combine(mf, lf, dummy, sm[0][k], sl[0][k], 0.f);
      lse_combine(mt, lt, d2, sm[1][k], sl[1][k], 0.f);
    }
    s_lse[0] = (mf + log2f(lf)) * 0.6931471805599453f;
    s_lse[1] = (mt + log2f(lt)) * 0.6931471805599453f;
    lse_f[r] = s_lse[0];
    lse_t[r] = s_lse[1];
  }
  __syncthreads();
  const float LF = s_lse[0], LT = s_lse[1];
  float acc = 0.f;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    const float lf_c = f[c] - LF;  // log labels
    acc += __expf(lf_c) * (lf_c - (t[c] - LT));
  }
  acc = warp_sum(acc);
  if (lane == 0) s_red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int k = 0; k < (blockDim.x >> 5); ++k) tot += s_red[k];
    row_loss[r] = tot;
  }
}

__global__ void dstl_finalize_kernel(const float* __restrict__ row_loss, int N, float* __restrict__ out) {
  __shared__ double red[8];
  double a = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) a += static_cast<double>(row_loss[i]);
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (blockDim.x >> 5); ++k) t += red[k];
    out[0] = static_cast<float>(t / N);
  }
}

// Backward strips for the local rows [row0, row0 + nloc): G''_rc = (softmax(TV)_rc - labels_rc) * N  (O(1)),
// rows -> Gr [nloc][ld] (for d text_t = Gr image), columns -> GcT [nloc][ld] (GcT[c - row0][r], for
// d image = GcT text_t); scale[0] = grad_out / N^2 is applied in fp32 by the products' epilogue.
template <int FMT>
__global__ void dstl_grad_kernel(const float* __restrict__ Fm, const float* __restrict__ TV, int N, int ld,
                                 const float* __restrict__ lse_f, const float* __restrict__ lse_t, int row0, int nloc,
                                 const float* __restrict__ grad_out, uint16_t* __restrict__ Gr,
                                 uint16_t* __restrict__ GcT, float* __restrict__ scale) {
  const int r = blockIdx.x;
  if (r == 0 && threadIdx.x == 0) scale[0] = __ldg(grad_out) / (static_cast<float>(N) * static_cast<float>(N));
  const float LF = lse_f[r], LT = lse_t[r];
  const float* f = Fm + static_cast<long long>(r) * N;
  const float* t = TV + static_cast<long long>(r) * N;
  const bool local_row = r >= row0 && r < row0 + nloc;
  const float fn = static_cast<float>(N);
  if (local_row) {
    uint16_t* o = Gr + static_cast<long long>(r - row0) * ld;
    for (int c = threadIdx.x; c < ld; c += blockDim.x)
      o[c] = f32_to_16<FMT>(c < N ? (__expf(t[c] - LT) - __expf(f[c] - LF)) * fn : 0.f);
  }
  for (int c = row0 + threadIdx.x; c < row0 + nloc; c += blockDim.x)
    GcT[static_cast<long long>(c - row0) * ld + r] = f32_to_16<FMT>((__expf(t[c] - LT) - __expf(f[c] - LF)) * fn);
  if (r == 0)  // pad columns of the transposed strip
    for (int i = threadIdx.x; i < nloc * (ld - N); i += blockDim.x)
      GcT[static_cast<long long>(i / (ld - N)) * ld + N + i % (ld - N)] = 0;
}

}  // namespace leccr
