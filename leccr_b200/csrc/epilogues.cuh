// Epilogue policies for sim_gemm_kernel (gemm_sm100.cuh).  Each epilogue thread owns one row of the
// 128 x 256 accumulator tile, so the row-wise reductions of the reference
//   - np.argsort(score)[::-1] per row            image_Retrieval_caption.py:268,289
//   - F.log_softmax(logits, dim=1) / cross_entropy  models/xvlm.py:279-290
// become thread-local streaming reductions over the column tiles of a work item.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <math_constants.h>
#include "gemm_sm100.cuh"

namespace leccr {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ------------------------------------------------------------------------------------------
// EpiStore: materialise S (fp32), optionally scaled and/or accumulated (split-K partial sums).
// Used for the reference's small score matrices (image_Retrieval_caption.py:151) and for the
// gradient products dA = G B, dB = G^T A of the contrastive backward.
// ------------------------------------------------------------------------------------------
struct EpiStore {
  struct Params {
    float* out[2];
    long long ld[2];
    float scale[2];
    const float* scale_ptr[2];  // optional device scalar multiplied into scale (upstream grad)
    const float* div_ptr[2];    // optional device scalar divided out of scale (temperature)
    int accumulate;             // 1: red.add (split-K), 0: plain store
  };
  static constexpr int kSmemBytes = 0;
  struct State {};
  __device__ static void begin(State&, const Params&, const ItemCtx&) {}
  __device__ static void end(State&, const Params&, const ItemCtx&) {}
  __device__ static void tile(State&, const Params& P, const ItemCtx& c, uint32_t taddr, int col0) {
    float scale = P.scale[c.p];
    if (P.scale_ptr[c.p] != nullptr) scale *= __ldg(P.scale_ptr[c.p]);
    if (P.div_ptr[c.p] != nullptr) scale /= __ldg(P.div_ptr[c.p]);
    float* orow = P.out[c.p] + static_cast<long long>(c.row) * P.ld[c.p];
    const bool vec_ok = ((P.ld[c.p] & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.out[c.p]) & 15) == 0);
#pragma unroll 1
    for (int cb = 0; cb < BN; cb += 32) {
      const int col = col0 + cb;
      if (col >= c.n_cols) break;  // warp-uniform
      float v[32];
      tmem_ld_32x32(taddr + cb, v);
      tmem_wait_ld();
      if (c.row < c.n_rows) {
        if (P.accumulate) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (col + e < c.n_cols) atomicAdd(orow + col + e, v[e] * scale);
        } else if (vec_ok && col + 32 <= c.n_cols) {
          float4* o4 = reinterpret_cast<float4*>(orow + col);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            o4[e] = make_float4(v[4 * e] * scale, v[4 * e + 1] * scale, v[4 * e + 2] * scale,
                                v[4 * e + 3] * scale);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (col + e < c.n_cols) orow[col + e] = v[e] * scale;
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// EpiTopK<KP>: streaming per-row top-KP of the approximate (16-bit operand) scores.
// Common case per 8 columns: one max-reduce and one compare against the row's threshold.
// Rare case: append to the row's list in shared memory; when a list is nearly full the warp
// sorts that one row cooperatively (rank selection) and raises the row's threshold.
// Output per (row, column chunk): KP (score, column) pairs sorted by score descending
// (ties: lower column first), padded with (-inf, -1).
// ------------------------------------------------------------------------------------------
template <int KP>
struct EpiTopK {
  static constexpr int C = 32;  // list capacity per row == warp width (one entry per lane when sorting)
  static_assert(KP <= 24, "KP must leave room for one 8-column group");
  static constexpr int LDS = C + 1;  // +1 word: conflict-free both for per-thread append and per-row sort
  struct Params {
    float* out_val[2];  // [n_rows][n_chunks][KP]
    int* out_idx[2];
    int n_chunks[2];
  };
  static constexpr int kSmemBytes = 2 * kEpiThreads * LDS * 4;
  struct State {
    float thr;
    int cnt;
  };

  __device__ static float* vals(const ItemCtx& c) { return reinterpret_cast<float*>(c.smem); }
  __device__ static int* idxs(const ItemCtx& c) {
    return reinterpret_cast<int*>(c.smem) + kEpiThreads * LDS;
  }

  __device__ static void begin(State& st, const Params&, const ItemCtx&) {
    st.thr = -CUDART_INF_F;
    st.cnt = 0;
  }

  // Warp-cooperative: sort row `src_lane`'s list, keep the best KP.  Returns (via out params on
  // every lane) the lane's element and its rank; the caller decides where the sorted list goes.
  __device__ static void rank_row(const ItemCtx& c, int src_lane, int cnt_src, float& v, int& id,
                                  int& rank) {
    const int et_src = c.warp_q * 32 + src_lane;
    const float* vr = vals(c) + et_src * LDS;
    const int* ir = idxs(c) + et_src * LDS;
    const bool have = c.lane < cnt_src;
    v = have ? vr[c.lane] : -CUDART_INF_F;
    id = have ? ir[c.lane] : -1;
    rank = 0;
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const float ov = __shfl_sync(0xffffffffu, v, l);
      rank += (ov > v || (ov == v && l < c.lane)) ? 1 : 0;  // slots are in arrival (column) order
    }
  }

  __device__ static void compact_row(State& st, const ItemCtx& c, int src_lane) {
    __syncwarp();  // the owner's appends must be visible to the whole warp
    const int cnt_src = __shfl_sync(0xffffffffu, st.cnt, src_lane);
    float v;
    int id, rank;
    rank_row(c, src_lane, cnt_src, v, id, rank);
    __syncwarp();
    const int et_src = c.warp_q * 32 + src_lane;
    if (rank < KP) {
      vals(c)[et_src * LDS + rank] = v;
      idxs(c)[et_src * LDS + rank] = id;
    }
    const unsigned m = __ballot_sync(0xffffffffu, rank == KP - 1);
    const float new_thr = __shfl_sync(0xffffffffu, v, __ffs(m) - 1);
    __syncwarp();
    if (c.lane == src_lane) {
      st.thr = new_thr;
      st.cnt = KP;
    }
  }

  __device__ static void tile(State& st, const Params&, const ItemCtx& c, uint32_t taddr, int col0) {
    float* myv = vals(c) + c.et * LDS;
    int* myi = idxs(c) + c.et * LDS;
#pragma unroll 1
    for (int cb = 0; cb < BN; cb += 32) {
      const int col = col0 + cb;
      if (col >= c.n_cols) break;  // warp-uniform
      float v[32];
      tmem_ld_32x32(taddr + cb, v);
      tmem_wait_ld();
      if (col + 32 > c.n_cols) {  // ragged last columns (TMA zero-filled): exclude them
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (col + e >= c.n_cols) v[e] = -CUDART_INF_F;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float m = fmaxf(fmaxf(fmaxf(v[8 * g], v[8 * g + 1]), fmaxf(v[8 * g + 2], v[8 * g + 3])),
                        fmaxf(fmaxf(v[8 * g + 4], v[8 * g + 5]), fmaxf(v[8 * g + 6], v[8 * g + 7])));
        if (m > st.thr) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (v[8 * g + e] > st.thr) {
              myv[st.cnt] = v[8 * g + e];
              myi[st.cnt] = col + 8 * g + e;
              ++st.cnt;
            }
          }
        }
        unsigned need = __ballot_sync(0xffffffffu, st.cnt > C - 8);
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          compact_row(st, c, src);
        }
      }
    }
  }

  __device__ static void end(State& st, const Params& P, const ItemCtx& c) {
    __syncwarp();
    const int nch = P.n_chunks[c.p];
#pragma unroll 1
    for (int src = 0; src < 32; ++src) {
      const int row = c.rb * BM + c.warp_q * 32 + src;
      if (row >= c.n_rows) break;  // warp-uniform
      const int cnt_src = __shfl_sync(0xffffffffu, st.cnt, src);
      float v;
      int id, rank;
      rank_row(c, src, cnt_src, v, id, rank);
      if (rank < KP) {
        const long long o = (static_cast<long long>(row) * nch + c.cc) * KP + rank;
        P.out_val[c.p][o] = v;
        P.out_idx[c.p][o] = id;
      }
    }
    __syncwarp();
  }
};

// ------------------------------------------------------------------------------------------
// EpiLse: forward of the symmetric InfoNCE loss (models/xvlm.py:273-290) for one orientation:
// per row i, online over the columns j with z_ij = s_ij / temp:
//   m = max_j z, l = sum_j exp(z - m), w = sum_j exp(z - m) z       (log-sum-exp and E_softmax[z])
//   pz = sum_j pos_ij z_ij, cnt = sum_j pos_ij,  pos_ij = (idx_i == idx_j)  (or i == j when idx is null)
// written per (row, column chunk); infonce_finalize merges chunks and forms loss and dtemp.
// Everything is kept in log2 units (zt = z * log2 e) so the exponentials are single EX2s.
// ------------------------------------------------------------------------------------------
struct EpiLse {
  struct Params {
    const float* temp;             // device scalar (nn.Parameter self.temp, models/xvlm.py:177)
    const long long* idx_rows[2];  // may be null: identity labels
    const long long* idx_cols[2];
    float* part[2];                // [n_rows][n_chunks][5] = m2, l, w2, pz2, cnt
    int n_chunks[2];
  };
  static constexpr int kSmemBytes = 0;
  struct State {
    float m, l, w, pz, cnt, sc;
    long long my_idx;
  };
  __device__ static void begin(State& st, const Params& P, const ItemCtx& c) {
    st.m = -CUDART_INF_F;
    st.l = 0.f;
    st.w = 0.f;
    st.pz = 0.f;
    st.cnt = 0.f;
    st.sc = kLog2e / __ldg(P.temp);
    const long long* ir = P.idx_rows[c.p];
    st.my_idx = (ir != nullptr && c.row < c.n_rows) ? __ldg(ir + c.row) : static_cast<long long>(c.row);
  }
  __device__ static void tile(State& st, const Params& P, const ItemCtx& c, uint32_t taddr, int col0) {
    const long long* ic = P.idx_cols[c.p];
#pragma unroll 1
    for (int cb = 0; cb < BN; cb += 32) {
      const int col = col0 + cb;
      if (col >= c.n_cols) break;  // warp-uniform
      float v[32];
      tmem_ld_32x32(taddr + cb, v);
      tmem_wait_ld();
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        v[e] = (col + e < c.n_cols) ? v[e] * st.sc : -CUDART_INF_F;
        mx = fmaxf(mx, v[e]);
      }
      const float m_new = fmaxf(st.m, mx);
      const float corr = exp2f(st.m - m_new);  // exp2(-inf) = 0 on the first chunk
      float l = 0.f, w = 0.f;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float p = exp2f(v[e] - m_new);
        l += p;
        w = fmaf(p, (col + e < c.n_cols) ? v[e] : 0.f, w);
      }
      st.l = fmaf(st.l, corr, l);
      st.w = fmaf(st.w, corr, w);
      st.m = m_new;
      if (ic != nullptr) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (col + e < c.n_cols && __ldg(ic + col + e) == st.my_idx) {
            st.pz += v[e];
            st.cnt += 1.f;
          }
        }
      } else {
        const int d = c.row - col;  // identity labels: only the diagonal element
        if (d >= 0 && d < 32 && c.row < c.n_cols) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (e == d) {
              st.pz += v[e];
              st.cnt += 1.f;
            }
        }
      }
    }
  }
  __device__ static void end(State& st, const Params& P, const ItemCtx& c) {
    if (c.row < c.n_rows) {
      float* o = P.part[c.p] + (static_cast<long long>(c.row) * P.n_chunks[c.p] + c.cc) * 5;
      o[0] = st.m;
      o[1] = st.l;
      o[2] = st.w;
      o[3] = st.pz;
      o[4] = st.cnt;
    }
  }
};

// ------------------------------------------------------------------------------------------
// EpiGrad: backward of the loss w.r.t. the logits, for the LOCAL rows of one orientation only
// (AllGather.backward keeps rows [B*rank, B*(rank+1)), models/xvlm.py:62-67):
//   G'_ij = softmax_row(z)_ij + softmax_col(z)_ij - pos_ij (1/cnt_i + 1/cnt_j)      in [-2, 2]
// written as a 16-bit strip [local rows][n_cols] that feeds the gradient product on the tensor
// cores; the 1/(2 N temp) factor is applied in that product's epilogue in fp32.
// lse / rcnt are in log2 units / reciprocals, produced by infonce_finalize.
// ------------------------------------------------------------------------------------------
struct EpiGrad {
  struct Params {
    const float* temp;
    const long long* idx_rows[2];
    const long long* idx_cols[2];
    const float* lse_rows[2];   // log2-domain lse of this orientation's rows
    const float* lse_cols[2];   // log2-domain lse of the other orientation (indexed by column)
    const float* rcnt_rows[2];  // 1 / cnt
    const float* rcnt_cols[2];
    void* strip[2];             // 16-bit [nrow][ld]
    long long ld[2];
    int row0[2];                // first absolute row of the strip (need not be a multiple of 128)
    int nrow[2];                // rows in the strip
    int fmt;                    // 0 fp16, 1 bf16
  };
  static constexpr int kSmemBytes = 0;
  struct State {
    float sc, lse, rc;
    long long my_idx;
    bool ok;
  };
  __device__ static void begin(State& st, const Params& P, const ItemCtx& c) {
    st.sc = kLog2e / __ldg(P.temp);
    const bool ok = c.row < c.n_rows && c.row >= P.row0[c.p] && c.row < P.row0[c.p] + P.nrow[c.p];
    st.ok = ok;
    st.lse = ok ? __ldg(P.lse_rows[c.p] + c.row) : 0.f;
    st.rc = ok ? __ldg(P.rcnt_rows[c.p] + c.row) : 0.f;
    const long long* ir = P.idx_rows[c.p];
    st.my_idx = (ir != nullptr && ok) ? __ldg(ir + c.row) : static_cast<long long>(c.row);
  }
  __device__ static void end(State&, const Params&, const ItemCtx&) {}
  __device__ static void tile(State& st, const Params& P, const ItemCtx& c, uint32_t taddr, int col0) {
    const long long* ic = P.idx_cols[c.p];
    const float* lc = P.lse_cols[c.p];
    const float* rc = P.rcnt_cols[c.p];
    const long long ld = P.ld[c.p];
    uint16_t* srow = reinterpret_cast<uint16_t*>(P.strip[c.p]) +
                     static_cast<long long>(c.row - P.row0[c.p]) * ld;
#pragma unroll 1
    for (int cb = 0; cb < BN; cb += 32) {
      const int col = col0 + cb;
      if (col >= ld) break;  // warp-uniform (ld = n_cols rounded up to 8; pad columns written as 0)
      float v[32];
      tmem_ld_32x32(taddr + cb, v);
      tmem_wait_ld();
      uint32_t packed[16];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int j = col + e;
        float g = 0.f;
        if (st.ok && j < c.n_cols) {
          const float z = v[e] * st.sc;
          g = exp2f(z - st.lse) + exp2f(z - __ldg(lc + j));
          const long long cj = (ic != nullptr) ? __ldg(ic + j) : static_cast<long long>(j);
          if (cj == st.my_idx) g -= st.rc + __ldg(rc + j);
        }
        v[e] = g;
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if (P.fmt == 0) {
          __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
          packed[e] = *reinterpret_cast<uint32_t*>(&h);
        } else {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          packed[e] = *reinterpret_cast<uint32_t*>(&h);
        }
      }
      if (st.ok) {
        if (col + 32 <= ld) {
          uint4* o = reinterpret_cast<uint4*>(srow + col);  // ld % 8 == 0 and col % 32 == 0: 16-byte aligned
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = make_uint4(packed[4 * e], packed[4 * e + 1], packed[4 * e + 2], packed[4 * e + 3]);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (col + e < ld)
              srow[col + e] = static_cast<uint16_t>((packed[e >> 1] >> ((e & 1) * 16)) & 0xffffu);
        }
      }
    }
  }
};

}  // namespace leccr
