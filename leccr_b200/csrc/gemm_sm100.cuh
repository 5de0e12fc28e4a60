// Similarity mainloop for sm_100a: S = R * C^T over 16-bit (fp16 / bf16) K-major operands.
//
//   R ("rows")  : the entities being ranked / whose softmax rows are formed  [n_rows, K]
//   C ("cols")  : the gallery / the other modality                           [n_cols, K]
//
// One persistent CTA per SM.  Warp 0 feeds shared memory with TMA (128-byte swizzle), warp 1
// issues tcgen05.mma (128 x 256 x 16, fp32 accumulators in TMEM, two 256-column buffers so the
// epilogue of tile t overlaps the MMAs of tile t+1), the remaining warps are the epilogue: one or
// two warpgroups (Epi::kWGs); with two, either warpgroup g drains the tiles whose TMEM buffer is g
// (kSplitCols false) or both drain every tile, half its columns each (kSplitCols true: the next tile's
// MMAs still overlap), so a single warp per scheduler never has to keep pace with the tensor core alone.  Epilogue thread
// <-> one row of the tile (TMEM lane), so every per-row reduction the reference performs
// (top-k, log-sum-exp, min/max, label sums) is thread-local and the N x M matrix never has to
// exist in HBM.  The epilogue is a policy class (see epilogues.cuh).
//
// A work item is (problem, row block of 128, column chunk); a launch can carry two problems
// (the two orientations i2t / t2i of the same pair of embedding sets).
#pragma once
#include "ptx.cuh"

namespace leccr {

constexpr int BM = 128;       // rows per tile   (TMEM lanes)
constexpr int BN = 256;       // columns per tile (TMEM columns per accumulator buffer)
constexpr int BK = 64;        // K elements per pipeline stage = one 128-byte swizzle atom
constexpr int UMMA_K = 16;    // K per tcgen05.mma for 16-bit operands
constexpr int kEpiThreads = 128;  // per epilogue warpgroup
constexpr int gemm_threads(int wgs) { return 64 + kEpiThreads * wgs; }
constexpr int kAStageBytes = BM * BK * 2;
constexpr int kBStageBytes = BN * BK * 2;
constexpr int kStageBytes = kAStageBytes + kBStageBytes;
constexpr int kTmemCols = 512;
// The K depth of a pipeline stage is a kernel parameter: 64 (one 128-byte swizzle atom per row, the default)
// or 32 (64-byte swizzle): half-size stages leave room for a second epilogue warpgroup's shared memory.
template <int kBK>
struct StageGeom {
  static_assert(kBK == 64 || kBK == 32, "stage depth is one 128-byte or one 64-byte swizzle row");
  static constexpr int kA = BM * kBK * 2;
  static constexpr int kB = BN * kBK * 2;
  static constexpr int kBytes = kA + kB;
  static constexpr int kAlign = kBK == 64 ? 1024 : 512;  // swizzle atom: 8 rows x row bytes
};

struct SimProblem {
  CUtensorMap tm_rows;  // box {BK, BM}
  CUtensorMap tm_cols;  // box {BK, BN}
  int n_rows, n_cols;
  int row_block_begin;  // first row block this launch covers (local strips in the backward)
  int row_blocks;       // number of row blocks this launch covers
  int col_tiles;        // ceil(n_cols / BN)
  int n_chunks;         // column chunks per row block
  int tiles_per_chunk;
  int item_base;        // first work-item id of this problem
};

struct SimLaunch {
  SimProblem prob[2];
  int n_prob;
  int n_items;
  int k_chunks;      // ceil(K / BK)
  int fmt;           // 0 fp16, 1 bf16
  int k_splits;      // > 1: a work item's "chunk" is a K range (split-K), all column tiles
  int kc_per_split;  // K chunks per split
};

// What an epilogue thread knows about the work item it is in.
struct ItemCtx {
  int p;        // problem index
  int rb;       // row block (absolute)
  int cc;       // column chunk
  int sub;      // partial-result slot of this (chunk, warpgroup): cc * kWGs + wg
  int wg;       // epilogue warpgroup
  int row;      // absolute row owned by this thread
  int n_rows, n_cols;
  int et;       // epilogue thread id 0..127 (== row within tile)
  int warp_q;   // TMEM lane quadrant of this warp
  int lane;
  uint32_t tile_n;  // running tile counter of this CTA (parity selects double buffers)
  uint8_t* smem;    // epilogue-private shared memory (Epi::kSmemBytes)
};

__device__ __forceinline__ void decode_item(const SimLaunch& L, int item, int& p, int& rb, int& ct0,
                                            int& ct1, int& cc, int& kc0, int& kc1) {
  p = (L.n_prob > 1 && item >= L.prob[1].item_base) ? 1 : 0;
  const SimProblem& P = L.prob[p];
  // chunk-major order: the column chunks of a row block are spread over successive waves, so later
  // chunks start from the per-row thresholds the earlier ones have published (EpiTopK::row_thr)
  int local = item - P.item_base;
  cc = local / P.row_blocks;
  int r = local - cc * P.row_blocks;
  rb = P.row_block_begin + r;
  if (L.k_splits > 1) {
    ct0 = 0;
    ct1 = P.col_tiles;
    kc0 = cc * L.kc_per_split;
    kc1 = min(L.k_chunks, kc0 + L.kc_per_split);
  } else {
    ct0 = cc * P.tiles_per_chunk;
    ct1 = min(P.col_tiles, ct0 + P.tiles_per_chunk);
    kc0 = 0;
    kc1 = L.k_chunks;
  }
}

// kARes: the row-block operand of a work item (all of K, at most kAResChunks stage-deep chunks) stays
// resident in shared memory for the item's whole column chunk and only the column operand streams through
// the stages: a third less L2 -> SM traffic per tile (the mainloop's limit when the epilogue is cheap).
// Chunk kc of the next item's row block is loaded as soon as the last tile of the current item has consumed
// chunk kc (a_empty / a_full barriers per chunk), so items follow each other without draining the pipeline.
constexpr int kAResChunks = 4;

// kBMN: the column operand is MN-major -- its tensor map describes the row-major [K][n_cols] source (the contraction
// index is the slow axis) and a stage holds BN / 64 boxes of {64 columns, kBK K-rows} (ptx.cuh make_sw128_mnmajor_desc).
template <class Epi, int kStages, int kBK = BK, bool kARes = false, bool kBMN = false>
__global__ void __launch_bounds__(gemm_threads(Epi::kWGs), 1)
sim_gemm_kernel(const __grid_constant__ SimLaunch L, const __grid_constant__ typename Epi::Params EP) {
  extern __shared__ uint8_t smem_raw[];
  using SG = StageGeom<kBK>;
  static_assert(!kBMN || (kBK == 64 && !kARes), "MN-major column operand: 128-byte swizzle, streamed stages");
  constexpr uint32_t kMnBoxBytes = 64 * kBK * 2;  // one {64 columns, kBK rows} box
  constexpr int kAStageBytes = SG::kA;
  constexpr int kStageBytes = kARes ? SG::kB : SG::kBytes;   // resident A: the stages hold B only
  constexpr int kAResBytes = kARes ? kAResChunks * SG::kA : 0;
  // swizzle atoms need 1024-byte (128 B rows) / 512-byte (64 B rows) alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + (SG::kAlign - 1)) &
                                             ~static_cast<uintptr_t>(SG::kAlign - 1));
  uint8_t* a_res = smem;                       // resident row-block chunks (kARes)
  uint8_t* stage_base = smem + kAResBytes;
  constexpr int kWGs = Epi::kWGs;
  uint8_t* epi_smem = stage_base + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + kWGs * Epi::kSmemBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tfull_bar = bars + 2 * kStages;
  uint64_t* tempty_bar = bars + 2 * kStages + 2;
  uint64_t* afull_bar = bars + 2 * kStages + 4;                 // kARes only
  uint64_t* aempty_bar = bars + 2 * kStages + 4 + kAResChunks;  // kARes only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4 + (kARes ? 2 * kAResChunks : 0));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int p = 0; p < L.n_prob; ++p) {
      tma_prefetch_desc(&L.prob[p].tm_rows);
      tma_prefetch_desc(&L.prob[p].tm_cols);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      // split-column policies: both warpgroups drain every tile (half the columns each)
      mbar_init(&tempty_bar[b], kEpiThreads * (Epi::kSplitCols ? Epi::kWGs : 1));
    }
    if (kARes) {
      for (int kc = 0; kc < kAResChunks; ++kc) {
        mbar_init(&afull_bar[kc], 1);
        mbar_init(&aempty_bar[kc], 1);
      }
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t item_n = 0;
      for (int item = blockIdx.x; item < L.n_items; item += gridDim.x, ++item_n) {
        int p, rb, ct0, ct1, cc, kc0, kc1;
        decode_item(L, item, p, rb, ct0, ct1, cc, kc0, kc1);
        const CUtensorMap* tmr = &L.prob[p].tm_rows;
        const CUtensorMap* tmc = &L.prob[p].tm_cols;
        for (int ct = ct0; ct < ct1; ++ct) {
          for (int kc = kc0; kc < kc1; ++kc) {
            if (kARes && ct == ct0) {  // this item's row-block chunk kc, once the previous item is done with the slot
              if (item_n > 0) mbar_wait(&aempty_bar[kc], (item_n - 1) & 1u, 500 + kc, 64);
              mbar_arrive_expect_tx(&afull_bar[kc], kAStageBytes);
              tma_load_2d(a_res + kc * kAStageBytes, tmr, &afull_bar[kc], kc * kBK, rb * BM);
            }
            mbar_wait(&empty_bar[stage], phase ^ 1u, 100 + stage, 64);
            uint8_t* sa = stage_base + stage * kStageBytes;
            uint8_t* sb = kARes ? sa : sa + kAStageBytes;
            mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
            if (!kARes) tma_load_2d(sa, tmr, &full_bar[stage], kc * kBK, rb * BM);
            if (kBMN) {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(sb + j * kMnBoxBytes, tmc, &full_bar[stage], ct * BN + j * 64, kc * kBK);
            } else {
              tma_load_2d(sb, tmc, &full_bar[stage], kc * kBK, ct * BN);
            }
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc_f16(L.fmt, BM, BN, kBMN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_n = 0;
      uint32_t item_n = 0;
      for (int item = blockIdx.x; item < L.n_items; item += gridDim.x, ++item_n) {
        int p, rb, ct0, ct1, cc, kc0, kc1;
        decode_item(L, item, p, rb, ct0, ct1, cc, kc0, kc1);
        for (int ct = ct0; ct < ct1; ++ct, ++tile_n) {
          const uint32_t buf = tile_n & 1u;
          const uint32_t use = tile_n >> 1;
          mbar_wait(&tempty_bar[buf], (use & 1u) ^ 1u, 200 + buf, 64);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BN;
          for (int kc = kc0; kc < kc1; ++kc) {
            if (kARes && ct == ct0) mbar_wait(&afull_bar[kc], item_n & 1u, 600 + kc);
            mbar_wait(&full_bar[stage], phase, 300 + stage);
            tc_fence_after();
            const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
            const uint64_t adesc = make_kmajor_desc<kBK>(kARes ? smem_u32(a_res + kc * kAStageBytes) : sa);
            const uint64_t bdesc = kBMN ? make_sw128_mnmajor_desc(sa + kAStageBytes, kMnBoxBytes)
                                        : make_kmajor_desc<kBK>(kARes ? sa : sa + kAStageBytes);
            // K-major: +32 bytes per K step inside the swizzle atom == +2 in the (addr >> 4) field;
            // MN-major: 16 K-rows of 128 bytes == +128
            constexpr uint32_t kBStep = kBMN ? 128u : 2u;
#pragma unroll
            for (int k = 0; k < kBK / UMMA_K; ++k) {
              umma_f16(d_tmem, adesc + 2u * k, bdesc + kBStep * k, idesc, (kc > kc0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // stage reusable once these MMAs have read it
            if (kARes && ct == ct1 - 1) umma_commit(&aempty_bar[kc]);  // the item's last use of row chunk kc
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit(&tfull_bar[buf]);  // accumulator complete
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warpgroups
    ItemCtx c;
    c.warp_q = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    c.lane = lane;
    c.et = c.warp_q * 32 + lane;
    c.wg = (warp - 2) >> 2;
    c.smem = epi_smem + c.wg * Epi::kSmemBytes;
    uint32_t tile_n = 0;
    typename Epi::State st;
    for (int item = blockIdx.x; item < L.n_items; item += gridDim.x) {
      int ct0, ct1, kc0, kc1;
      decode_item(L, item, c.p, c.rb, ct0, ct1, c.cc, kc0, kc1);
      c.n_rows = L.prob[c.p].n_rows;
      c.n_cols = L.prob[c.p].n_cols;
      c.row = c.rb * BM + c.et;
      c.sub = c.cc * kWGs + c.wg;
      Epi::begin(st, EP, c);
      for (int ct = ct0; ct < ct1; ++ct, ++tile_n) {
        if (!Epi::kSplitCols && kWGs == 2 && (tile_n & 1u) != static_cast<uint32_t>(c.wg)) continue;
        const uint32_t buf = tile_n & 1u;
        const uint32_t use = tile_n >> 1;
        Epi::prefetch(st, EP, c);  // loads whose latency should hide behind the wait
        mbar_wait(&tfull_bar[buf], use & 1u, 400 + buf, 32);
        tc_fence_after();
        const int col_off = Epi::kSplitCols ? c.wg * (BN / kWGs) : 0;  // this warpgroup's columns of the tile
        const uint32_t taddr = tmem_base + buf * BN + col_off + (static_cast<uint32_t>(c.warp_q * 32) << 16);
        c.tile_n = tile_n;
        Epi::tile(st, EP, c, taddr, ct * BN + col_off);
        tc_fence_before();
        mbar_arrive(&tempty_bar[buf]);
      }
      Epi::end(st, EP, c);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <class Epi, int kStages, int kBK = BK, bool kARes = false>
constexpr size_t sim_gemm_smem_bytes() {
  return StageGeom<kBK>::kAlign +
         (kARes ? static_cast<size_t>(kAResChunks) * StageGeom<kBK>::kA + static_cast<size_t>(kStages) * StageGeom<kBK>::kB
                : static_cast<size_t>(kStages) * StageGeom<kBK>::kBytes) +
         Epi::kWGs * Epi::kSmemBytes + (2 * kStages + 4 + (kARes ? 2 * kAResChunks : 0)) * 8 + 16;
}

}  // namespace leccr
