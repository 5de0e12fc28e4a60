// Epilogue policies for sim_gemm_kernel (gemm_sm100.cuh).  Each epilogue thread owns one row of the
// 128 x 256 accumulator tile, so the row-wise reductions of the reference
//   - np.argsort(score)[::-1] per row               image_Retrieval_caption.py:268,289
//   - F.log_softmax(logits, dim=1) / cross_entropy  models/xvlm.py:279-290
// become thread-local streaming reductions over the column tiles of a work item.
// Everything a thread does per element is branch-free or warp-uniform; the only per-thread
// divergence is the rare list maintenance of the top-k epilogue, and that contains no collectives.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include "gemm_sm100.cuh"

namespace leccr {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// Walk the 256 accumulator columns of this thread's row in chunks of 32, with the TMEM load of
// the next chunk in flight while the current one is processed.  f(v, col, n_valid) must be
// warp-uniform in its use of collectives.  n_limit: first column index that does not exist.
template <class F>
__device__ __forceinline__ void for_each_chunk(uint32_t taddr, int col0, int n_limit, F&& f, int max_chunks = BN / 32) {
  const int nch = min(max_chunks, (n_limit - col0 + 31) / 32);  // warp-uniform
  if (nch <= 0) return;
  float va[32], vb[32];
  tmem_ld_32x32(taddr, va);
  tmem_wait_ld_regs(va);
#pragma unroll 1
  for (int ch = 0; ch < nch; ch += 2) {
    const bool has_b = ch + 1 < nch;
    const bool has_a2 = ch + 2 < nch;
    if (has_b) tmem_ld_32x32(taddr + (ch + 1) * 32, vb);
    f(va, col0 + ch * 32);
    if (has_b) {
      tmem_wait_ld_regs(vb);
      if (has_a2) tmem_ld_32x32(taddr + (ch + 2) * 32, va);
      f(vb, col0 + (ch + 1) * 32);
      if (has_a2) tmem_wait_ld_regs(va);
    }
  }
}

// ------------------------------------------------------------------------------------------
// EpiStore: materialise S (fp32), optionally scaled; in split-K mode each split writes its own
// partial plane (out + split * split_stride) and splitk_reduce_kernel sums them in a fixed order.
// Used for the reference's small score matrices (image_Retrieval_caption.py:151) and for the
// gradient products dA = G B, dB = G^T A of the contrastive backward.
// Each warp transposes its 32 x 32 sub-tile through shared memory so global stores are full
// 128-byte lines.
// ------------------------------------------------------------------------------------------
struct EpiStore {
  struct Params {
    float* out[2];
    long long ld[2];
    float scale[2];
    const float* scale_ptr[2];  // optional device scalar multiplied into scale (upstream grad)
    const float* div_ptr[2];    // optional device scalar divided out of scale (temperature)
    long long split_stride[2];  // elements between split-K partial planes (0 when not split)
  };
  static constexpr int kWGs = 2;
  static constexpr bool kSplitCols = false;
  static constexpr int kTileLd = 33;
  static constexpr int kSmemBytes = 4 * 32 * kTileLd * 4;
  struct State {};
  __device__ static void begin(State&, const Params&, const ItemCtx&) {}
  __device__ static void end(State&, const Params&, const ItemCtx&) {}
  __device__ static void prefetch(State&, const Params&, const ItemCtx&) {}
  __device__ static void tile(State&, const Params& P, const ItemCtx& c, uint32_t taddr, int col0) {
    float scale = P.scale[c.p];
    if (P.scale_ptr[c.p] != nullptr) scale *= __ldg(P.scale_ptr[c.p]);
    if (P.div_ptr[c.p] != nullptr) scale /= __ldg(P.div_ptr[c.p]);
    float* out = P.out[c.p] + static_cast<long long>(c.cc) * P.split_stride[c.p];  // cc == K split here
    const long long ld = P.ld[c.p];
    const uint32_t tbase = smem_u32(c.smem) + static_cast<uint32_t>(c.warp_q) * 32 * kTileLd * 4;
    const int row0 = c.rb * BM + c.warp_q * 32;
    const int lane = c.lane;
    const int n_rows = c.n_rows, n_cols = c.n_cols;
    for_each_chunk(taddr, col0, n_cols, [&](float(&v)[32], int col) {
#pragma unroll
      for (int e = 0; e < 32; ++e) sts_f32(tbase + (lane * kTileLd + e) * 4, v[e] * scale);
      __syncwarp();
      float* o = out + static_cast<long long>(row0) * ld + col + lane;
      if (row0 + 32 <= n_rows && col + 32 <= n_cols) {  // interior: no bounds checks (warp-uniform)
#pragma unroll
        for (int r = 0; r < 32; ++r) o[static_cast<long long>(r) * ld] = lds_f32(tbase + (r * kTileLd + lane) * 4);
      } else {
        const bool col_ok = col + lane < n_cols;
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const float x = lds_f32(tbase + (r * kTileLd + lane) * 4);
          if (col_ok && row0 + r < n_rows) o[static_cast<long long>(r) * ld] = x;
        }
      }
      __syncwarp();
    });
  }
};

// ------------------------------------------------------------------------------------------
// EpiTopK<KP>: streaming per-row candidates for the top-KP of the approximate (16-bit operand)
// scores.  Per thread (= row): a threshold `thr` with the invariant "at least KP seen scores are
// > thr (or thr = -inf)", and a list (shared memory, capacity C) of every seen score > thr.
//   common case, per 8 columns : one max-reduce, one compare with thr
//   rare                        : append (score, column) to the list
//   when a list is nearly full  : raise thr by bisection on the score value until KP..KEEP of the
//                                 listed scores exceed it, drop the rest.  Purely per-thread
//                                 (no shuffles), so all 32 rows of a warp shrink side by side.
// Nothing <= thr can belong to the row's top-KP, so the union of a row's lists over its column
// chunks contains the exact top-KP of the approximate scores; topk_finalize selects them.
// ------------------------------------------------------------------------------------------
// Compare-exchange and bitonic networks on register arrays (all indices are compile-time constants
// after unrolling): branch-free, high ILP -- unlike counting loops they are not latency-bound.
__device__ __forceinline__ void cex_desc(float& a, float& b) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  a = hi;
  b = lo;
}
template <int N>
__device__ __forceinline__ void bitonic_merge_desc(float* a) {  // a bitonic -> sorted descending
#pragma unroll
  for (int j = N >> 1; j > 0; j >>= 1)
#pragma unroll
    for (int i = 0; i < N; ++i)
      if ((i ^ j) > i) cex_desc(a[i], a[i ^ j]);
}
template <int N>
__device__ __forceinline__ void bitonic_sort_desc(float* a) {
#pragma unroll
  for (int k = 2; k <= N; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int l = i ^ j;
        if (l > i) {
          if ((i & k) == 0) cex_desc(a[i], a[l]);
          else cex_desc(a[l], a[i]);
        }
      }
}
// top-16 (as a set, in a[0..16)) of two descending-sorted 16-lists a, b; sorted again if `resort`
__device__ __forceinline__ void top16_of_two(float* a, const float* b, bool resort) {
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = fmaxf(a[i], b[15 - i]);
  if (resort) bitonic_merge_desc<16>(a);
}

// Shapes: <16, 64, 1> one warpgroup with 64-entry lists (66 KB), <16, 64, 2> two warpgroups (alternate tiles),
// each with its own 64-entry lists (133 KB, beside a pipeline of half-depth stages).  The two threads that own
// the same row (and the CTAs that own other column chunks of it) cooperate through row_thr.
// MODE 0: the filter / dense choice is a run-time parameter (Params::dense); MODE 1: always dense, the
// filter code is not even compiled in (half the instruction footprint: the single epilogue warp per
// scheduler cannot hide instruction-cache misses or branch resolution).
template <int KP, int C_, int WGS, int MODE = 0>
struct EpiTopK {
  static constexpr int kWGs = WGS;
  static constexpr bool kSplitCols = WGS > 1;  // two warpgroups: each takes half the columns of every tile
  static constexpr int C = C_;           // list capacity
  static constexpr int ES = 8;           // bytes per list entry: (score, column) interleaved so an append is ONE 64-bit store
  // row pitch in words: even (8-byte aligned entries) and twice an odd number (64-bit per-thread accesses of a
  // half-warp then fall on 16 distinct bank pairs)
  static constexpr int LDSW = (C % 2 == 1) ? 2 * C : 2 * C + 2;
  static constexpr int TRIG = C - 8;     // a group of 8 columns must always fit
  static constexpr int G = (C + 15) / 16; // groups of 16 the shrink sorts
  static constexpr int KEEP = KP + 2;    // the bisection fallback stops once this few remain
  static constexpr int JOIN = KP + 4;    // every row fuller than this shrinks whenever any row of the warp must:
                                          // one round refreshes (nearly) all 32 thresholds, so rounds stay rare
  static constexpr int CP = (C + 15) / 16 * 16;  // list length padded to whole groups of 16 for the sorting networks
  static_assert(KP == 16 && (CP == 64 || CP == 32), "the shrink sorts 2 or 4 groups of 16");
  static_assert(KP <= KEEP && KEEP < JOIN && JOIN <= TRIG, "inconsistent list policy");
  struct Params {
    float* out_val[2];  // [n_rows][n_sub][C]
    int* out_idx[2];
    int* out_cnt[2];    // [n_rows][n_sub]
    int n_sub[2];       // partial lists per row: n_chunks * kWGs (over all calls of a streamed evaluation)
    int sub_base[2];    // first list slot this launch writes (streamed evaluation: one column window per call)
    int col_base[2];    // global index of the launch's first column (added to the stored column indices)
    unsigned* row_thr[2];  // [n_rows] shared per-row threshold keys (zeroed per launch), or null
    unsigned* row_h8[2];   // [n_rows][2] (two warpgroups only) key of the 8th best score seen by ANY list of
                           // warpgroup 0 / 1 for the row.  Lists of different warpgroups hold disjoint columns,
                           // so min(h8[0], h8[1]) is a valid threshold for the row (8 + 8 scores reach it): each
                           // warpgroup filters with (nearly) the 16th best of the UNION instead of the 16th
                           // best of its own half (measured: appends per row 441 -> see profiles/README.md)
    int debug_mode;     // measurement aid: 1 = threshold +inf (filter only), 2 = skip the tile entirely
    unsigned long long* debug_counters;  // measurement aid: [chunks, hit chunks, hit groups, shrink rounds, appends]
    int trig;           // a shrink round starts when some row of the warp holds more than this (<= TRIG)
    int dense;          // While nearly every 8-column group holds a candidate for one of the warp's 32 rows
                        // (threshold warm-up, short rows) the filter and its branches only cost: such tiles
                        // append (predicated) all 32 columns of every chunk.  0 never, 1 always, 2 adaptive:
                        // the first tiles of a work item, then again whenever >= 6 of a tile's 8 chunks hit.
  };
  static constexpr int kSmemBytes = kEpiThreads * LDSW * 4;
  static constexpr int kIdxOff = 4;  // byte offset from a value slot to its index slot
  struct State {
    float thr;
    int cnt;
    uint32_t vb, ib;  // shared-window addresses of this thread's value / index list
    unsigned dc[5];   // debug counters (per thread; lane 0's are warp-level events)
    int dense_tiles;  // warp-uniform: upcoming tiles to run without the filter (see Params::dense)
    unsigned pre_key; // shared threshold key fetched while waiting for the accumulator (see prefetch)
    unsigned pre_h8;  // the other warpgroup's 8th-best key, fetched alongside
    unsigned own_h8;  // this list's 8th-best key (after its last shrink)
  };
  static constexpr int DTRIG = C - 16;   // dense mode appends 16 columns between checks

  __device__ static void begin(State& st, const Params& P, const ItemCtx& c) {
    st.thr = P.debug_mode == 1 ? CUDART_INF_F : -CUDART_INF_F;
    st.cnt = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) st.dc[i] = 0;
    st.dense_tiles = (MODE == 0 && P.dense == 2) ? 4 : 0;
    st.pre_key = 0u;
    st.pre_h8 = 0u;
    st.own_h8 = 0u;
    st.vb = smem_u32(c.smem) + static_cast<uint32_t>(c.et) * LDSW * 4;
    st.ib = st.vb + kIdxOff;
  }

  __device__ static int count_above(uint32_t vb, int n, float t) {
    int cgt = 0;
#pragma unroll 8
    for (int s = 0; s < C; ++s) {
      const float x = lds_f32(vb + s * ES);
      cgt += (s < n && x > t) ? 1 : 0;
    }
    return cgt;
  }
  // The shrinks run rarely and are large: keep them out of line (one copy, small hot loop) and pass
  // the state by value so it stays in registers.  Result: (thr bits << 32) | cnt.
  __device__ static unsigned long long pack_state(float thr, int cnt) {
    return (static_cast<unsigned long long>(__float_as_uint(thr)) << 32) | static_cast<unsigned>(cnt);
  }

  // Branch-free shrink: tau = the exact 16th largest listed score, found with sorting networks in
  // registers (sort the groups of 16, merge keeping the top 16); keep the entries >= tau, i.e. 16
  // unless scores tie.  (A cheaper bound -- the minimum of 16 strided quad maxima -- keeps 36 of 64
  // on average: the threshold then sits at the 36th best and twice as many later scores pass it.)
  // Ties that leave the list too full fall through to the bisection shrink.
  // h8 (optional, global): receives atomicMax of the key of this list's 8th best score (see Params::row_h8).
  __device__ __noinline__ static unsigned long long quad_shrink(float thr_in, int n, uint32_t vb, uint32_t ib,
                                                               unsigned* h8 = nullptr) {
    float x[CP];
#pragma unroll
    for (int s = 0; s < CP; ++s) {
      x[s] = s < C ? lds_f32(vb + s * ES) : -CUDART_INF_F;
      if (s >= n) x[s] = -CUDART_INF_F;
    }
    // exact 16th largest of the list: sort the C/16 groups of 16, merge keeping the top 16
#pragma unroll
    for (int q = 0; q < G; ++q) bitonic_sort_desc<16>(x + 16 * q);
    if (G == 4) {
      top16_of_two(x, x + 16, true);
      top16_of_two(x + 32, x + 48, true);
      top16_of_two(x, x + 32, false);
    } else {
      top16_of_two(x, x + 16, false);
    }
    float tau;
    if (WGS > 1) {
      bitonic_merge_desc<16>(x);  // top16_of_two leaves a bitonic sequence: one merge sorts it
      tau = x[15];
      if (h8 != nullptr && x[7] > -CUDART_INF_F) atomicMax(h8, f32_key(x[7]));
    } else {
      tau = x[0];
#pragma unroll
      for (int k = 1; k < 16; ++k) tau = fminf(tau, x[k]);
    }
    // an adopted (shared) threshold may already exceed tau: then everything below it is dead too
    const float te = fmaxf(tau, thr_in);
    int j = 0;
#pragma unroll 8
    for (int s = 0; s < C; ++s) {  // the network permuted x: re-read the list
      const float xs = lds_f32(vb + s * ES);
      const int id = lds_s32(ib + s * ES);
      if (s < n && xs >= te) {
        sts_f32(vb + j * ES, xs);
        sts_s32(ib + j * ES, id);
        ++j;
      }
    }
    const int cnt = (te == -CUDART_INF_F) ? n : j;
    if (cnt > JOIN) return shrink(te, cnt, vb, ib);
    return pack_state(te, cnt);
  }

  __device__ __noinline__ static unsigned long long shrink(float thr_in, int n, uint32_t vb, uint32_t ib) {
    float vmax = -CUDART_INF_F, vmin = CUDART_INF_F;
#pragma unroll 8
    for (int s = 0; s < C; ++s) {
      const float x = lds_f32(vb + s * ES);
      if (s < n) {
        vmax = fmaxf(vmax, x);
        vmin = fminf(vmin, x);
      }
    }
    // keys: lo_k has c_lo >= KP listed scores above it, hi_k has c_hi < KP
    uint32_t lo_k = f32_key(thr_in), hi_k = f32_key(vmax);
    int c_lo = n, c_hi = 0;
    bool first = thr_in == -CUDART_INF_F;  // -inf .. vmin is empty: probe vmin first
#pragma unroll 1
    for (int it = 0; it < 40; ++it) {
      if (c_lo <= KEEP || hi_k - lo_k <= 1u) break;
      uint32_t mid_k = lo_k + ((hi_k - lo_k) >> 1);
      if (first) {
        const uint32_t vk = f32_key(vmin);
        if (vk > lo_k && vk < hi_k) mid_k = vk;
        first = false;
      }
      const int cm = count_above(vb, n, key_f32(mid_k));
      if (cm >= KP) {
        lo_k = mid_k;
        c_lo = cm;
      } else {
        hi_k = mid_k;
        c_hi = cm;
      }
    }
    // Normal: keep scores > lo.  Tie block (adjacent keys, more than KEEP equal scores at hi):
    // keep scores > hi plus as many == hi as needed to hold KP; then thr = hi.
    const bool tie = c_lo > KEEP;
    const float lo = key_f32(lo_k);
    const float keep_thr = tie ? key_f32(hi_k) : lo;
    int quota = tie ? KP - c_hi : 0;
    int j = 0;
#pragma unroll 4
    for (int s = 0; s < C; ++s) {
      const float x = lds_f32(vb + s * ES);
      const int id = lds_s32(ib + s * ES);
      bool keep = s < n && x > keep_thr;
      if (!keep && s < n && quota > 0 && x > lo) {
        keep = true;
        --quota;
      }
      if (keep) {
        sts_f32(vb + j * ES, x);
        sts_s32(ib + j * ES, id);
        ++j;
      }
    }
    return pack_state(keep_thr, j);
  }

  // The unfiltered treatment of one chunk of 32 values (dense mode; also the consumer of the fused double_sim
  // scores, EpiDsTopK): every value is offered to the list, four at a time, with a (warp-uniform) check for a
  // shrink round every 16.  Invariant: cnt <= DTRIG before every 16 values.
  __device__ static void dense_chunk(State& st, const float (&v)[32], int col, unsigned* my_h8, unsigned* shared_thr) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float thr = st.thr;
      uint32_t cur = st.vb + st.cnt * ES;  // list cursor as an address: one predicated add per element
#pragma unroll
      for (int e = 0; e < 16; e += 4)
        cur = append4_if_gt(v[16 * h + e], v[16 * h + e + 1], v[16 * h + e + 2], v[16 * h + e + 3], thr, cur,
                            col + 16 * h + e);
      const int cnt = static_cast<int>(cur - st.vb) / ES;
      st.dc[4] += cnt - st.cnt;
      st.cnt = cnt;
      if (__any_sync(0xffffffffu, cnt > DTRIG)) {
        ++st.dc[3];
        if (cnt > JOIN) {
          const unsigned long long r = quad_shrink(st.thr, cnt, st.vb, st.ib, my_h8);
          st.thr = __uint_as_float(static_cast<unsigned>(r >> 32));
          st.cnt = static_cast<int>(r & 0xffffffffu);
          if (shared_thr != nullptr) atomicMax(shared_thr, f32_key(st.thr));
        }
      }
    }
  }

  // Called before the wait for the accumulator: the (L2) read of the row's shared threshold overlaps it.
  // A threshold proven valid by ANY column chunk of this row is valid for every chunk; stale reads are fine.
  __device__ static void prefetch(State& st, const Params& P, const ItemCtx& c) {
    if (P.row_thr[c.p] != nullptr && c.row < c.n_rows)
      st.pre_key = *reinterpret_cast<volatile unsigned*>(P.row_thr[c.p] + c.row);
    if (WGS > 1 && P.row_h8[c.p] != nullptr && c.row < c.n_rows) {
      const volatile unsigned* h = P.row_h8[c.p] + 2 * static_cast<long long>(c.row);
      st.pre_h8 = h[c.wg ^ 1];
      st.own_h8 = max(st.own_h8, h[c.wg]);  // other column chunks of the row publish here too
    }
  }

  __device__ static void tile(State& st, const Params& P, const ItemCtx& c, uint32_t taddr, int col0) {
    const int n_cols = c.n_cols;
    if (P.debug_mode == 2) return;
    // A threshold proven valid by ANY column chunk of this row (>= 16 scores of the row are at least
    // that large) is valid for every chunk: adopt the best one published so far.  Stale reads are fine.
    unsigned* shared_thr = (P.row_thr[c.p] != nullptr && c.row < c.n_rows) ? P.row_thr[c.p] + c.row : nullptr;
    if (st.pre_key > f32_key(st.thr)) st.thr = key_f32(st.pre_key);  // fetched by prefetch() during the wait
    if (WGS > 1) {
      const unsigned u = min(st.own_h8, st.pre_h8);  // 8 scores here + 8 scores there reach this key
      if (u > f32_key(st.thr)) st.thr = key_f32(u);
    }
    unsigned* my_h8 = (WGS > 1 && P.row_h8[c.p] != nullptr && c.row < c.n_rows)
                          ? P.row_h8[c.p] + 2 * static_cast<long long>(c.row) + c.wg : nullptr;
    // warp-uniform; lists shorter than 64 entries cannot take 16 unfiltered columns between checks
    const bool dense = C >= 64 && (MODE == 1 || P.dense == 1 || st.dense_tiles > 0);
    if (st.dense_tiles > 0) --st.dense_tiles;
    if (dense && __any_sync(0xffffffffu, st.cnt > DTRIG)) {  // the dense path appends up to 16 between checks
      if (st.cnt > JOIN) {
        const unsigned long long r = quad_shrink(st.thr, st.cnt, st.vb, st.ib, my_h8);
        st.thr = __uint_as_float(static_cast<unsigned>(r >> 32));
        st.cnt = static_cast<int>(r & 0xffffffffu);
        if (shared_thr != nullptr) atomicMax(shared_thr, f32_key(st.thr));
      }
    }
    int hit_chunks = 0;  // warp-uniform
    const int cb = P.col_base[c.p];
    for_each_chunk(taddr, col0, n_cols, [&](float(&v)[32], int lcol) {
      if (lcol + 32 > n_cols) {  // ragged last columns (TMA zero-filled): exclude them
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (lcol + e >= n_cols) v[e] = -CUDART_INF_F;
      }
      const int col = lcol + cb;  // stored indices are global
      if (dense) {  // warp-uniform; invariant: cnt <= DTRIG = C - 16 before every 16 columns
        dense_chunk(st, v, col, my_h8, shared_thr);
        return;
      }
      if (MODE == 1) return;  // compile-time: nothing below exists in the dense-only kernel
      // level 1: one maximum over the whole chunk (3-input FMNMX3 tree, ~16 instructions) and one vote;
      // on long rows five chunks out of six end here
      float cm[8];
#pragma unroll
      for (int g = 0; g < 8; ++g)
        cm[g] = fmaxf(fmaxf(v[4 * g], v[4 * g + 1]), fmaxf(v[4 * g + 2], v[4 * g + 3]));
      const float cmax = fmaxf(fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])),
                               fmaxf(fmaxf(cm[4], cm[5]), fmaxf(cm[6], cm[7])));
      ++st.dc[0];
      if (!__any_sync(0xffffffffu, cmax > st.thr)) return;  // warp-uniform
      // level 2: which of the four 8-column groups hold a candidate for some row
      float gm[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) gm[g] = fmaxf(cm[2 * g], cm[2 * g + 1]);
      // ONE warp collective decides the whole chunk: which of the four 8-column groups contain a
      // candidate for some row (votes and branches are the latency that bounds this epilogue)
      unsigned hm = (gm[0] > st.thr ? 1u : 0u) | (gm[1] > st.thr ? 2u : 0u) | (gm[2] > st.thr ? 4u : 0u) |
                    (gm[3] > st.thr ? 8u : 0u);
      hm = __reduce_or_sync(0xffffffffu, hm);
      ++st.dc[1];
      ++hit_chunks;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (hm & (1u << g)) {  // warp-uniform
          ++st.dc[2];
          const float thr = st.thr;
          // predicated appends, four at a time: straight-line code beats any branch here
          uint32_t cur = st.vb + st.cnt * ES;
          cur = append4_if_gt(v[8 * g], v[8 * g + 1], v[8 * g + 2], v[8 * g + 3], thr, cur, col + 8 * g);
          cur = append4_if_gt(v[8 * g + 4], v[8 * g + 5], v[8 * g + 6], v[8 * g + 7], thr, cur, col + 8 * g + 4);
          const int cnt = static_cast<int>(cur - st.vb) / ES;
          st.dc[4] += cnt - st.cnt;
          st.cnt = cnt;
          if (__any_sync(0xffffffffu, cnt > P.trig)) {
            ++st.dc[3];
            if (cnt > JOIN) {
              const unsigned long long r = quad_shrink(st.thr, cnt, st.vb, st.ib, my_h8);
              st.thr = __uint_as_float(static_cast<unsigned>(r >> 32));
              st.cnt = static_cast<int>(r & 0xffffffffu);
              if (shared_thr != nullptr) atomicMax(shared_thr, f32_key(st.thr));
            }
          }
        }
      }
    }, BN / 32 / kWGs);
    if (MODE == 0 && P.dense == 2 && hit_chunks >= 6) st.dense_tiles = 4;
  }

  // Dump the lists; rows of a warp are written one after another so stores coalesce.  A last shrink round
  // (only when some row of the warp holds more than 32 entries) leaves every list with at most JOIN <= 32
  // entries: half the write-out and half the candidate slots topk_finalize has to merge.
  __device__ static void end(State& st, const Params& P, const ItemCtx& c) {
    __syncwarp();
    if (__any_sync(0xffffffffu, st.cnt > 32)) {
      if (st.cnt > JOIN) {
        unsigned* shared_thr = (P.row_thr[c.p] != nullptr && c.row < c.n_rows) ? P.row_thr[c.p] + c.row : nullptr;
        const unsigned long long r = quad_shrink(st.thr, st.cnt, st.vb, st.ib);
        st.thr = __uint_as_float(static_cast<unsigned>(r >> 32));
        st.cnt = static_cast<int>(r & 0xffffffffu);
        if (shared_thr != nullptr) atomicMax(shared_thr, f32_key(st.thr));
      }
    }
    __syncwarp();
    if (P.debug_counters != nullptr) {
      const unsigned app = __reduce_add_sync(0xffffffffu, st.dc[4]);
      if (c.lane == 0) {
        for (int i = 0; i < 4; ++i) atomicAdd(P.debug_counters + i, static_cast<unsigned long long>(st.dc[i]));
        atomicAdd(P.debug_counters + 4, static_cast<unsigned long long>(app));
      }
    }
    const int nch = P.n_sub[c.p];
    const uint32_t wbase = smem_u32(c.smem) + static_cast<uint32_t>(c.warp_q * 32) * LDSW * 4;
#pragma unroll 4
    for (int src = 0; src < 32; ++src) {
      const int row = c.rb * BM + c.warp_q * 32 + src;
      if (row >= c.n_rows) break;  // warp-uniform
      const int cnt_src = __shfl_sync(0xffffffffu, st.cnt, src);
      const long long o = (static_cast<long long>(row) * nch + P.sub_base[c.p] + c.sub) * C;
      const uint32_t vb = wbase + static_cast<uint32_t>(src) * LDSW * 4;
      const uint32_t ib = vb + kIdxOff;
      {
        const int s = c.lane;  // cnt_src <= 32 after the final round
        if (s < cnt_src) {
          P.out_val[c.p][o + s] = lds_f32(vb + s * ES);
          P.out_idx[c.p][o + s] = lds_s32(ib + s * ES);
        }
      }
      if (c.lane == 0) P.out_cnt[c.p][static_cast<long long>(row) * nch + P.sub_base[c.p] + c.sub] = cnt_src;
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
// The labels of a tile's 256 columns are staged in shared memory once per tile.
// ------------------------------------------------------------------------------------------
struct EpiLse {
  struct Params {
    const float* temp;             // device scalar (nn.Parameter self.temp, models/xvlm.py:177)
    const long long* idx_rows[2];  // may be null: identity labels
    const long long* idx_cols[2];
    float* part[2];                // [n_rows][n_sub][5] = m2, l, w2, pz2, cnt
    int n_sub[2];                  // partials per row: n_chunks * kWGs
  };
  static constexpr int kWGs = 2;
  static constexpr bool kSplitCols = false;
  static constexpr int kSmemBytes = 2 * BN * 8;
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
  __device__ static void prefetch(State&, const Params&, const ItemCtx&) {}
  __device__ static void tile(State& st, const Params& P, const ItemCtx& c, uint32_t taddr, int col0) {
    const long long* ic = P.idx_cols[c.p];
    const int n_cols = c.n_cols;
    const uint32_t sidx = smem_u32(c.smem) + ((c.tile_n >> (kWGs - 1)) & 1u) * BN * 8;
    if (ic != nullptr) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = col0 + c.et + h * kEpiThreads;
        sts_s64(sidx + (c.et + h * kEpiThreads) * 8, j < n_cols ? __ldg(ic + j) : 0);
      }
      epi_bar_sync(c.wg);  // one barrier per tile; the staging buffers alternate tile by tile
    }
    const bool ragged = col0 + BN > n_cols;
    const float sc = st.sc;
    const long long my = st.my_idx;
    const int row = c.row;
    for_each_chunk(taddr, col0, n_cols, [&](float(&v)[32], int col) {
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        v[e] *= sc;
        if (ragged && col + e >= n_cols) v[e] = -CUDART_INF_F;
        mx = fmaxf(mx, v[e]);
      }
      const float m_new = fmaxf(st.m, mx);
      const float corr = ex2_approx(st.m - m_new);  // ex2(-inf) = 0 on the first chunk
      float l[4] = {0.f, 0.f, 0.f, 0.f}, w[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float p = ex2_approx(v[e] - m_new);
        l[e & 3] += p;
        w[e & 3] = fmaf(p, (ragged && col + e >= n_cols) ? 0.f : v[e], w[e & 3]);
      }
      st.l = fmaf(st.l, corr, (l[0] + l[1]) + (l[2] + l[3]));
      st.w = fmaf(st.w, corr, (w[0] + w[1]) + (w[2] + w[3]));
      st.m = m_new;
      if (ic != nullptr) {
        const uint32_t sa = sidx + (col - col0) * 8;
        float pz[2] = {0.f, 0.f}, pc[2] = {0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const bool pos = lds_s64(sa + e * 8) == my && !(ragged && col + e >= n_cols);
          pz[e & 1] += pos ? v[e] : 0.f;
          pc[e & 1] += pos ? 1.f : 0.f;
        }
        st.pz += pz[0] + pz[1];
        st.cnt += pc[0] + pc[1];
      } else {
        const int d = row - col;  // identity labels: only the diagonal element
        if (d >= 0 && d < 32 && row < n_cols) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (e == d) {
              st.pz += v[e];
              st.cnt += 1.f;
            }
        }
      }
    });
  }
  __device__ static void end(State& st, const Params& P, const ItemCtx& c) {
    if (c.row < c.n_rows) {
      float* o = P.part[c.p] + (static_cast<long long>(c.row) * P.n_sub[c.p] + c.sub) * 5;
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
// lse / rcnt are in log2 units / reciprocals, produced by infonce_finalize.  The column-side
// label, lse and 1/cnt of a tile's 256 columns are staged in shared memory once per tile.
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
    float row_w[2], col_w[2];   // weights of the row / column softmax terms: (1, 1) for the symmetric loss; the
                                // one-directional loss -sum(log_softmax(sim, 1) * labels) keeps only the softmax
                                // over the rows of orientation 0: (1, 0) there and (0, 1) in orientation 1
  };
  static constexpr int kWGs = 2;
  static constexpr bool kSplitCols = false;
  static constexpr int kBufBytes = BN * 16;  // idx (8) + lse (4) + rcnt (4) per column
  static constexpr int kSmemBytes = 2 * kBufBytes;
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
  __device__ static void prefetch(State&, const Params&, const ItemCtx&) {}
  __device__ static void tile(State& st, const Params& P, const ItemCtx& c, uint32_t taddr, int col0) {
    const long long* ic = P.idx_cols[c.p];
    const float* lc = P.lse_cols[c.p];
    const float* rcp = P.rcnt_cols[c.p];
    const int ld = static_cast<int>(P.ld[c.p]);
    const int n_cols = c.n_cols;
    const uint32_t sbuf = smem_u32(c.smem) + ((c.tile_n >> (kWGs - 1)) & 1u) * kBufBytes;
    const uint32_t s_idx = sbuf, s_lse = sbuf + BN * 8, s_rc = sbuf + BN * 12;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = c.et + h * kEpiThreads;
      const int j = col0 + t;
      const bool in = j < n_cols;
      sts_s64(s_idx + t * 8, in ? (ic != nullptr ? __ldg(ic + j) : static_cast<long long>(j)) : 0);
      sts_f32(s_lse + t * 4, in ? __ldg(lc + j) : CUDART_INF_F);  // exp2(z - inf) = 0 for pad columns
      sts_f32(s_rc + t * 4, in ? __ldg(rcp + j) : 0.f);
    }
    epi_bar_sync(c.wg);
    uint16_t* srow = reinterpret_cast<uint16_t*>(P.strip[c.p]) +
                     static_cast<long long>(c.row - P.row0[c.p]) * ld;
    const bool ragged = col0 + BN > n_cols;
    const float sc = st.sc, lse_r = st.lse;
    const float rw = P.row_w[c.p], cw = P.col_w[c.p];
    const float rc_r = st.rc * rw;
    const long long my = st.my_idx;
    const bool ok = st.ok;
    const int fmt = P.fmt;
    // ld = n_cols rounded up to 8; pad columns are written as 0
    for_each_chunk(taddr, col0, ld, [&](float(&v)[32], int col) {
      const int cl = col - col0;
      uint32_t packed[16];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float z = v[e] * sc;
        float g = rw * ex2_approx(z - lse_r) + cw * ex2_approx(z - lds_f32(s_lse + (cl + e) * 4));
        const bool pos = lds_s64(s_idx + (cl + e) * 8) == my;
        g -= pos ? (rc_r + cw * lds_f32(s_rc + (cl + e) * 4)) : 0.f;
        if (ragged && col + e >= n_cols) g = 0.f;
        v[e] = g;
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if (fmt == 0) {
          __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
          packed[e] = *reinterpret_cast<uint32_t*>(&h);
        } else {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          packed[e] = *reinterpret_cast<uint32_t*>(&h);
        }
      }
      if (ok) {
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
    });
  }
};

// ------------------------------------------------------------------------------------------
// double_sim in the epilogue (video_Retrieval_caption_double_sim.py:170-179; raw variant
// image_Retrieval_caption.py:239-246):   F = w1 * f(S) + w2 * f(max_n C_n),  S = V T^T,  C_n = cap_n T^T.
// The video rows and their n caption queries are INTERLEAVED into one operand "VC" in groups of G rows
// (G = 2, 4, 8 >= n + 1; row G i = video i, rows G i + 1 .. G i + n = its captions, the rest repeat the last
// caption), so ONE accumulator tile holds S and every C_n of a block of videos and the plain mainloop serves:
//   orientation B (rows = texts, columns = VC): a thread finds S and the C_n of a (text, video) pair in G
//                 adjacent registers;
//   orientation A (rows = VC, columns = texts): in G adjacent lanes -- log2 G shuffles.
// Pass 1 (EpiDsStats, orientation B, nothing stored but 2 floats per text): global min / max of S and of
// max_n C_n (norm_score, :87-91) and, for every text, S and max C at its ground-truth video.  Pass 2
// (EpiDsTopK, both orientations in one launch): recompute, fuse with the reference's fp32 operation order,
// count the scores above the row's best ground-truth score (rank = number of strictly greater, as everywhere
// in this library) and feed the per-row top-k lists.  The N x M matrices never reach HBM.
// ------------------------------------------------------------------------------------------
struct DsFuse {
  float smax, sden, cmax, cden, w1, w2;
  int mode;  // 1 norm, 2 raw
  __device__ __forceinline__ float operator()(float s, float c) const {
    if (mode == 1) {
      s = -__fdiv_rn(__fsub_rn(smax, s), sden);
      c = -__fdiv_rn(__fsub_rn(cmax, c), cden);
    }
    return __fadd_rn(__fmul_rn(w1, s), __fmul_rn(w2, c));
  }
};
__device__ __forceinline__ float ds_ordered_f32(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ DsFuse ds_fuse_load(const unsigned* mm, float w1, float w2, int mode) {
  DsFuse f;
  f.w1 = w1;
  f.w2 = w2;
  f.mode = mode;
  f.smax = f.cmax = 0.f;
  f.sden = f.cden = 1.f;
  if (mode == 1) {
    f.smax = ds_ordered_f32(mm[0]);
    f.sden = __fsub_rn(f.smax, ds_ordered_f32(mm[1]));
    f.cmax = ds_ordered_f32(mm[2]);
    f.cden = __fsub_rn(f.cmax, ds_ordered_f32(mm[3]));
  }
  return f;
}

template <int G>
struct EpiDsStats {
  struct Params {
    int n_cap;              // caption queries per video (1 <= n_cap < G)
    int n_groups;           // videos
    const int* gt_group;    // [n_rows] ground-truth video of every text, or null
    float* gt_s;            // [n_rows] S at the ground truth
    float* gt_c;            // [n_rows] max_n C_n at the ground truth
    unsigned* mm;           // ordered-uint {max S, min S, max C, min C} (mm_init_kernel)
  };
  static constexpr int kWGs = 2;
  static constexpr bool kSplitCols = false;
  static constexpr int kSmemBytes = 16;
  struct State {
    float smax, smin, cmax, cmin;
    int gt;
  };
  __device__ static void begin(State& st, const Params& P, const ItemCtx& c) {
    st.smax = st.cmax = -CUDART_INF_F;
    st.smin = st.cmin = CUDART_INF_F;
    st.gt = (P.gt_group != nullptr && c.row < c.n_rows) ? __ldg(P.gt_group + c.row) : -1;
  }
  __device__ static void prefetch(State&, const Params&, const ItemCtx&) {}
  __device__ static void tile(State& st, const Params& P, const ItemCtx& c, uint32_t taddr, int col0) {
    const bool row_ok = c.row < c.n_rows;
    const int n_cap = P.n_cap;
    for_each_chunk(taddr, col0, c.n_cols, [&](float(&v)[32], int col) {
#pragma unroll
      for (int q = 0; q < 32 / G; ++q) {
        const int grp = col / G + q;
        const float s = v[q * G];
        float cm = v[q * G + 1];
#pragma unroll
        for (int k = 2; k < G; ++k)
          if (k <= n_cap) cm = fmaxf(cm, v[q * G + k]);
        if (row_ok && grp < P.n_groups) {
          st.smax = fmaxf(st.smax, s);
          st.smin = fminf(st.smin, s);
          st.cmax = fmaxf(st.cmax, cm);
          st.cmin = fminf(st.cmin, cm);
          if (grp == st.gt) {
            P.gt_s[c.row] = s;
            P.gt_c[c.row] = cm;
          }
        }
      }
    });
  }
  __device__ static void end(State& st, const Params& P, const ItemCtx& c) {
    const float a = warp_max(st.smax), b = warp_min(st.smin), d = warp_max(st.cmax), e = warp_min(st.cmin);
    if (c.lane == 0 && a >= b) {  // a < b: this warp saw no valid element
      atomicMax(P.mm + 0, f32_key(a));
      atomicMin(P.mm + 1, f32_key(b));
      atomicMax(P.mm + 2, f32_key(d));
      atomicMin(P.mm + 3, f32_key(e));
    }
  }
};

// Pass 2.  Problem 1 = orientation B (rows texts, columns VC; stored column = video index): BOTH rank
// directions are counted here, from the same accumulator bits pass 1 took the ground-truth scores from (the
// other orientation sums its split-precision blocks in another order and may differ in the last bit, which
// would let a ground-truth element beat itself): a text's own row count is thread-local, a video's count is an
// atomicAdd per score that beats the video's best ground truth (a handful per video).  Problem 0 = orientation
// A (rows VC, columns texts) exists only for the per-video top-k lists (list rows are VC rows, only rows G i
// carry data) and is left out when no lists are wanted.  Lists, thresholds, shrink rounds and the write-out
// are EpiTopK's (two warpgroups, dense).
// (the template parameter must not be called G: the non-dependent base has a member G that would hide it)
template <int GRP>
struct EpiDsTopK : EpiTopK<16, 64, 2, 1> {
  using Base = EpiTopK<16, 64, 2, 1>;
  struct Params : Base::Params {
    int n_cap, n_groups;
    const unsigned* mm;
    float w1, w2;
    int mode;
    int b_problem;           // index of the orientation-B problem in the launch (0 when A is left out)
    const float* txt_best;   // [n_texts] fused score at the text's ground-truth video, or null (no ranks)
    const float* vid_best;   // [n_groups] best fused ground-truth score of the video
    int* rank_txt;           // [n_texts]  zeroed, atomically accumulated
    int* rank_vid;           // [n_groups]
  };
  struct State : Base::State {
    DsFuse fuse;
    float gt;
    int above;
  };
  __device__ static void begin(State& st, const Params& P, const ItemCtx& c) {
    Base::begin(st, P, c);
    st.fuse = ds_fuse_load(P.mm, P.w1, P.w2, P.mode);
    st.above = 0;
    st.gt = (c.p == P.b_problem && P.txt_best != nullptr && c.row < c.n_rows) ? __ldg(P.txt_best + c.row) : CUDART_INF_F;
  }
  __device__ static void prefetch(State& st, const Params& P, const ItemCtx& c) { Base::prefetch(st, P, c); }
  __device__ static void tile(State& st, const Params& P, const ItemCtx& c, uint32_t taddr, int col0) {
    const int n_cols = c.n_cols;
    unsigned* shared_thr = (P.row_thr[c.p] != nullptr && c.row < c.n_rows) ? P.row_thr[c.p] + c.row : nullptr;
    if (st.pre_key > f32_key(st.thr)) st.thr = key_f32(st.pre_key);
    {
      const unsigned u = min(st.own_h8, st.pre_h8);
      if (u > f32_key(st.thr)) st.thr = key_f32(u);
    }
    unsigned* my_h8 = (P.row_h8[c.p] != nullptr && c.row < c.n_rows)
                          ? P.row_h8[c.p] + 2 * static_cast<long long>(c.row) + c.wg : nullptr;
    if (__any_sync(0xffffffffu, st.cnt > Base::DTRIG)) {
      if (st.cnt > Base::JOIN) {
        const unsigned long long r = Base::quad_shrink(st.thr, st.cnt, st.vb, st.ib, my_h8);
        st.thr = __uint_as_float(static_cast<unsigned>(r >> 32));
        st.cnt = static_cast<int>(r & 0xffffffffu);
        if (shared_thr != nullptr) atomicMax(shared_thr, f32_key(st.thr));
      }
    }
    const int n_cap = P.n_cap;
    const bool row_ok = c.row < c.n_rows;
    if (c.p != P.b_problem) {
      // orientation A: lane = VC row; the GRP lanes of a video combine with shuffles
      const int m = c.row % GRP;                       // 0 video, 1..n_cap captions, above: padding
      const bool head = m == 0 && row_ok;
      const int base_lane = c.lane & ~(GRP - 1);
      for_each_chunk(taddr, col0, n_cols, [&](float(&v)[32], int lcol) {
        float w[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float x = (m >= 1 && m <= n_cap) ? v[e] : -CUDART_INF_F;
#pragma unroll
          for (int o = 1; o < GRP; o <<= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
          const float s = __shfl_sync(0xffffffffu, v[e], base_lane);
          w[e] = (head && lcol + e < n_cols) ? st.fuse(s, x) : -CUDART_INF_F;
        }
        Base::dense_chunk(st, w, lcol, my_h8, shared_thr);
      }, BN / 32 / kWGs);
    } else {
      // orientation B: GRP adjacent columns = one video
      const int n_groups = P.n_groups;
      const bool ranks = P.txt_best != nullptr;
      for_each_chunk(taddr, col0, n_cols, [&](float(&v)[32], int lcol) {
        float w[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) w[e] = -CUDART_INF_F;
#pragma unroll
        for (int q = 0; q < 32 / GRP; ++q) {
          const float s = v[q * GRP];
          float cm = v[q * GRP + 1];
#pragma unroll
          for (int k = 2; k < GRP; ++k)
            if (k <= n_cap) cm = fmaxf(cm, v[q * GRP + k]);
          const float f = st.fuse(s, cm);
          const int grp = lcol / GRP + q;
          const bool ok = row_ok && grp < n_groups;
          w[q] = ok ? f : -CUDART_INF_F;
          if (ranks && ok) {
            st.above += f > st.gt ? 1 : 0;
            if (f > __ldg(P.vid_best + grp)) atomicAdd(P.rank_vid + grp, 1);
          }
        }
        Base::dense_chunk(st, w, lcol / GRP, my_h8, shared_thr);   // stored column = video index (+ e)
      }, BN / 32 / kWGs);
    }
  }
  __device__ static void end(State& st, const Params& P, const ItemCtx& c) {
    if (c.p == P.b_problem && P.rank_txt != nullptr && st.above > 0 && c.row < c.n_rows)
      atomicAdd(P.rank_txt + c.row, st.above);
    Base::end(st, P, c);
  }
};

// ------------------------------------------------------------------------------------------
// EpiRank: Recall@K without candidate lists -- what the reference's itm_eval actually returns
// (image_Retrieval_caption.py:261-317: per row the position of its best ground-truth column, then
// Recall@1/5/10; no top-k list leaves that function).  The row's best ground-truth score t is known EXACTLY
// before the pass (fp32 dots of the original inputs, gt_best_kernel), with the rigorous bound eps on the error of
// a 16-bit-operand score.  Per element the epilogue only counts: s > t + eps is definitely greater, s <= t - eps
// definitely not; the rare scores inside the band are written as (row, column) pairs and re-scored exactly
// afterwards (rank_resolve_kernel).  ~4 ALU instructions per element, no shared memory, no per-element store --
// the short-row evaluations (cfg1 / cfg2), which the list epilogue runs at a third of the tensor peak, run at the
// mainloop's pace.  Both warpgroups drain every tile (half its columns each).
// ------------------------------------------------------------------------------------------
struct EpiRank {
  struct Params {
    const float* lo[2];     // [n_rows] t - eps   (+inf: the row has no ground truth)
    const float* hi[2];     // [n_rows] t + eps
    int* rank[2];           // [n_rows] += number of definitely greater scores
    int* amb_count[2];      // pairs written so far
    int2* amb_list[2];      // (row, column) pairs inside the band
    int amb_cap[2];
    int* row_flag[2];       // [n_rows] 0/1: the pair list overflowed for this row -> exact fallback
    int* flag_count[2];
    int* flag_list[2];
    int debug_mode;         // measurement aid: 1 = empty band (count only), 2 = skip the tile entirely
  };
  static constexpr int kWGs = 2;
  static constexpr bool kSplitCols = true;
  // Band pairs are collected per WARP in shared memory and flushed with ONE global atomic per ~batch: a global
  // atomic per pair (its return value is the slot) stalled the epilogue warp for a round trip on a fifth of the
  // chunks of cfg2, whose ground-truth scores sit in the bulk's tail (265 us against 123 for the mainloop).
  static constexpr int kWarpCap = 96;                                  // pairs buffered per warp
  static constexpr int kRankCapEpi = 10;                               // == kRankCap (kernels.cuh), LECCR_RANK_CAP
  static constexpr int kWarpBytes = 16 + kWarpCap * 8;                 // counter (padded) + pairs
  static constexpr int kSmemBytes = 4 * kWarpBytes;                    // per warpgroup
  struct State {
    float lo, hi;
    int above;
    uint32_t wbuf;  // shared-window address of this warp's buffer
  };
  __device__ static void begin(State& st, const Params& P, const ItemCtx& c) {
    const bool ok = c.row < c.n_rows;
    st.lo = ok ? __ldg(P.lo[c.p] + c.row) : CUDART_INF_F;
    st.hi = ok ? __ldg(P.hi[c.p] + c.row) : CUDART_INF_F;
    if (P.debug_mode == 1) st.lo = st.hi;
    st.above = 0;
    st.wbuf = smem_u32(c.smem) + static_cast<uint32_t>(c.warp_q) * kWarpBytes;
    if (c.lane == 0) sts_s32(st.wbuf, 0);
    __syncwarp();
  }
  __device__ static void prefetch(State&, const Params&, const ItemCtx&) {}
  __device__ static void flag_row(const Params& P, int p, int row) {
    if (atomicExch(P.row_flag[p] + row, 1) == 0) P.flag_list[p][atomicAdd(P.flag_count[p], 1)] = row;
  }
  // Warp-collective: move the buffered pairs to the global list (one atomic), reset the buffer.
  __device__ __noinline__ static void flush(const State& st, const Params& P, int p, int lane) {
    __syncwarp();
    const int n = min(lds_s32(st.wbuf), kWarpCap);
    if (n > 0) {
      int base = 0;
      if (lane == 0) base = atomicAdd(P.amb_count[p], n);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int i = lane; i < n; i += 32) {
        const int row = lds_s32(st.wbuf + 16 + i * 8);
        const int col = lds_s32(st.wbuf + 20 + i * 8);
        if (base + i < P.amb_cap[p]) P.amb_list[p][base + i] = make_int2(row, col);
        else flag_row(P, p, row);
      }
    }
    __syncwarp();
    if (lane == 0) sts_s32(st.wbuf, 0);
    __syncwarp();
  }
  __device__ __forceinline__ static void push_band(const State& st, const Params& P, int p, int row, int col) {
    int slot;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(st.wbuf) : "memory");
    if (slot < kWarpCap) {
      sts_s32(st.wbuf + 16 + slot * 8, row);
      sts_s32(st.wbuf + 20 + slot * 8, col);
    } else {
      flag_row(P, p, row);  // more than kWarpCap pairs of one warp between two flushes: heavy ties
    }
  }
  template <int E>
  __device__ __forceinline__ static void mask_all(const float (&v)[32], float hi, float lo, unsigned (&mh)[4], unsigned (&ml)[4]) {
    if constexpr (E < 32) {
      mask_gt2<(1u << E)>(v[E], hi, lo, mh[E & 3], ml[E & 3]);
      mask_all<E + 1>(v, hi, lo, mh, ml);
    }
  }
  __device__ static void tile(State& st, const Params& P, const ItemCtx& c, uint32_t taddr, int col0) {
    const int n_cols = c.n_cols;
    const float lo = st.lo, hi = st.hi;
    if (P.debug_mode == 2) return;
    for_each_chunk(taddr, col0, n_cols, [&](float(&v)[32], int lcol) {
      if (lcol + 32 > n_cols) {  // ragged last columns (TMA zero-filled): exclude them
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (lcol + e >= n_cols) v[e] = -CUDART_INF_F;
      }
      unsigned mh[4] = {0u, 0u, 0u, 0u}, ml[4] = {0u, 0u, 0u, 0u};  // four partial masks: short dependency chains
      mask_all<0>(v, hi, lo, mh, ml);
      const unsigned m_hi = (mh[0] | mh[1]) | (mh[2] | mh[3]);
      unsigned band = ((ml[0] | ml[1]) | (ml[2] | ml[3])) & ~m_hi;
      st.above += __popc(m_hi);
      // a row that already has kRankCap definitely greater scores is decided for Recall@1/5/10 (the rank contract is
      // "exact below LECCR_RANK_CAP"): its band elements -- such rows have the densest bands -- need no re-scoring
      if (st.above >= kRankCapEpi) band = 0u;
      if (__any_sync(0xffffffffu, band != 0u)) {  // some score of the warp's 32 x 32 block sits inside a band
        while (band != 0u) {                      // per lane: its own band elements only (column = bit index)
          const int e = __ffs(band) - 1;
          band &= band - 1u;
          push_band(st, P, c.p, c.row, lcol + e);
        }
        __syncwarp();
        if (lds_s32(st.wbuf) > kWarpCap - 32) flush(st, P, c.p, c.lane);  // warp-uniform (same word for every lane)
      }
    }, BN / 32 / kWGs);
  }
  __device__ static void end(State& st, const Params& P, const ItemCtx& c) {
    if (st.above > 0 && c.row < c.n_rows) atomicAdd(P.rank[c.p] + c.row, st.above);
    flush(st, P, c.p, c.lane);
  }
};

}  // namespace leccr
