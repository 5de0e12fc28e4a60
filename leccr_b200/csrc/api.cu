// Host side of the C ABI declared in include/leccr_b200.h: argument checking, TMA descriptor
// construction, work decomposition and kernel launches.  No device allocation, no synchronisation.
#include "../../include/leccr_b200.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <atomic>
#include <mutex>

#include "epilogues.cuh"
#include "gemm_sm100.cuh"
#include "kernels.cuh"

using namespace leccr;

namespace {

thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", where, cudaGetErrorString(e));
  return LECCR_ERR_CUDA;
}
#define CUDA_TRY(expr)                                   \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return cuda_fail(_e, #expr);  \
  } while (0)
#define LAUNCH_CHECK(name)                                   \
  do {                                                       \
    cudaError_t _e = cudaGetLastError();                     \
    if (_e != cudaSuccess) return cuda_fail(_e, name);       \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      sms = v;
    else
      return 148;
  }
  return sms;
}

// cuTensorMapEncodeTiled is a driver entry point: it needs the device's primary context current on the calling
// thread.  A thread whose first CUDA call is ours (torch's autograd thread running a backward that starts with a
// tensor-core launch) has none yet -- runtime calls that only query the device do not bind it.
void bind_primary_context() {
  thread_local bool bound = false;
  if (!bound) {
    cudaFree(nullptr);
    bound = true;
  }
}

// 2-D K-major 16-bit operand [n][K] (ld elements) with a {BK, box_rows} box and 128-byte swizzle.
int make_tmap(CUtensorMap* tm, const void* base, int64_t n, int64_t K, int64_t ld, int fmt, int box_rows, int bk = BK) {
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld & 7) != 0) return LECCR_ERR_ALIGN;
  if (n <= 0 || K <= 0 || ld < K) return LECCR_ERR_ARG;
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return LECCR_ERR_DRIVER;
  bind_primary_context();
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(n)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(bk), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, fmt == LECCR_FMT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                   2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return LECCR_ERR_DRIVER;
  }
  return LECCR_OK;
}

// The MN-major column operand of a product whose contraction index is the slow axis of the source: row-major
// [K rows][n_mn] (ld elements), {64 elements, BK rows} boxes, 128-byte swizzle (gemm_sm100.cuh kBMN).
int make_tmap_mn(CUtensorMap* tm, const void* base, int64_t k_rows, int64_t n_mn, int64_t ld, int fmt) {
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld & 7) != 0) return LECCR_ERR_ALIGN;
  if (k_rows <= 0 || n_mn <= 0 || ld < n_mn) return LECCR_ERR_ARG;
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return LECCR_ERR_DRIVER;
  bind_primary_context();
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(n_mn), static_cast<cuuint64_t>(k_rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(BK)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, fmt == LECCR_FMT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                   2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled (MN-major) failed with CUresult %d", (int)r);
    return LECCR_ERR_DRIVER;
  }
  return LECCR_OK;
}

struct Plan {
  int row_block_begin, row_blocks, col_tiles, n_chunks, tiles_per_chunk;
};

Plan plan_problem(int64_t n_cols, int64_t row_begin, int64_t row_count, int tiles_per_chunk) {
  Plan p;
  p.row_block_begin = static_cast<int>(row_begin / BM);
  p.row_blocks = static_cast<int>((row_begin + row_count + BM - 1) / BM) - p.row_block_begin;
  p.col_tiles = static_cast<int>((n_cols + BN - 1) / BN);
  int tpc = std::max(1, std::min(tiles_per_chunk, p.col_tiles));
  p.n_chunks = (p.col_tiles + tpc - 1) / tpc;
  p.tiles_per_chunk = tpc;
  return p;
}

// Column-chunk length shared by the problems of one launch: aim at ~4 work items per SM so the
// static round-robin schedule balances, but keep chunks long enough to amortise per-item state.
int auto_tiles_per_chunk(int64_t total_tiles, int min_tpc) {
  int64_t t = total_tiles / (4LL * num_sms());
  return static_cast<int>(std::max<int64_t>(min_tpc, std::min<int64_t>(t, 64)));
}

int fill_problem(SimProblem& P, const void* rows16, int64_t ld_rows, const void* cols16, int64_t ld_cols,
                 int64_t n_rows, int64_t n_cols, int K, int fmt, const Plan& pl, int item_base, int bk = BK) {
  int rc = make_tmap(&P.tm_rows, rows16, n_rows, K, ld_rows, fmt, BM, bk);
  if (rc != LECCR_OK) return rc;
  rc = make_tmap(&P.tm_cols, cols16, n_cols, K, ld_cols, fmt, BN, bk);
  if (rc != LECCR_OK) return rc;
  P.n_rows = static_cast<int>(n_rows);
  P.n_cols = static_cast<int>(n_cols);
  P.row_block_begin = pl.row_block_begin;
  P.row_blocks = pl.row_blocks;
  P.col_tiles = pl.col_tiles;
  P.n_chunks = pl.n_chunks;
  P.tiles_per_chunk = pl.tiles_per_chunk;
  P.item_base = item_base;
  return LECCR_OK;
}

// Development aid (leccr_profile_*): CUDA events around every tensor-core launch, recorded on the launch's own
// stream WITHOUT synchronising, so a timed loop can keep them on; leccr_profile_read resolves them.
bool g_profile = false;
struct EventPair {
  cudaEvent_t begin, end;
};
constexpr int kProfilePairs = 256;
EventPair g_pairs[kProfilePairs];
int g_pairs_made = 0, g_pairs_used = 0;
double g_profile_ms = 0.0;
int g_profile_launches = 0;

cudaError_t profile_resolve() {
  for (int i = 0; i < g_pairs_used; ++i) {
    cudaError_t e = cudaEventSynchronize(g_pairs[i].end);
    if (e != cudaSuccess) return e;
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, g_pairs[i].begin, g_pairs[i].end);
    if (e != cudaSuccess) return e;
    g_profile_ms += ms;
    ++g_profile_launches;
  }
  g_pairs_used = 0;
  return cudaSuccess;
}

// Development aid (LECCR_STAGE_PROFILE=1): events between the launches of one C-ABI call, printed by
// leccr_profile_read.  Not used by the product path.
struct StageMark {
  const char* name;
  cudaEvent_t ev;
};
StageMark g_marks[64];
int g_n_marks = 0;
bool stage_profile_on() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LECCR_STAGE_PROFILE");
    v = (e != nullptr && atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}
void prof_mark(const char* name, cudaStream_t stream) {
  if (!stage_profile_on() || g_n_marks >= 64) return;
  StageMark& m = g_marks[g_n_marks];
  if (m.ev == nullptr) cudaEventCreate(&m.ev);
  m.name = name;
  cudaEventRecord(m.ev, stream);
  ++g_n_marks;
}
void prof_dump() {
  if (!stage_profile_on() || g_n_marks == 0) return;
  cudaEventSynchronize(g_marks[g_n_marks - 1].ev);
  for (int i = 1; i < g_n_marks; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_marks[i - 1].ev, g_marks[i].ev);
    fprintf(stderr, "  stage %-28s %8.1f us\n", g_marks[i].name, ms * 1e3f);
  }
  g_n_marks = 0;
}

// As many pipeline stages as fit beside the epilogue's own shared memory (227 KB per CTA).
template <class Epi, int kBK = BK, bool kARes = false>
constexpr int stages_for() {
  return sim_gemm_smem_bytes<Epi, 4, kBK, kARes>() <= 232448 ? 4 : 3;
}

template <class Epi, int kBK = BK, bool kARes = false, bool kBMN = false>
int launch_gemm(const SimLaunch& L, const typename Epi::Params& EP, cudaStream_t stream) {
  constexpr int kStages = stages_for<Epi, kBK, kARes>();
  auto kern = sim_gemm_kernel<Epi, kStages, kBK, kARes, kBMN>;
  constexpr size_t smem = sim_gemm_smem_bytes<Epi, kStages, kBK, kARes>();
  static_assert(smem <= 232448, "exceeds the 227 KB shared memory limit of sm_100");
  // function attributes are per device: one flag per device ordinal (a process may drive several GPUs)
  static std::atomic<int> attr_set[64];
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || attr_set[dev].load(std::memory_order_acquire) == 0) {
    const cudaError_t attr_err =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(sim_gemm_kernel)");
    if (dev >= 0 && dev < 64) attr_set[dev].store(1, std::memory_order_release);
  }
  if (L.n_items <= 0) return LECCR_OK;
  const int grid = std::min(L.n_items, num_sms());
  EventPair* pair = nullptr;
  if (g_profile) {
    if (g_pairs_used == kProfilePairs) CUDA_TRY(profile_resolve());  // pool exhausted: drain (synchronises)
    if (g_pairs_used == g_pairs_made) {
      CUDA_TRY(cudaEventCreate(&g_pairs[g_pairs_made].begin));
      CUDA_TRY(cudaEventCreate(&g_pairs[g_pairs_made].end));
      ++g_pairs_made;
    }
    pair = &g_pairs[g_pairs_used++];
    CUDA_TRY(cudaEventRecord(pair->begin, stream));
  }
  kern<<<grid, gemm_threads(Epi::kWGs), smem, stream>>>(L, EP);
  LAUNCH_CHECK("sim_gemm_kernel");
  if (pair != nullptr) CUDA_TRY(cudaEventRecord(pair->end, stream));
  return LECCR_OK;
}

bool bad_fmt(int fmt) { return fmt != LECCR_FMT_F16 && fmt != LECCR_FMT_BF16; }

}  // namespace

extern "C" {

const char* leccr_strerror(int code) {
  switch (code) {
    case LECCR_OK: return "ok";
    case LECCR_ERR_ARG: return "invalid argument (shape, null pointer or unsupported option)";
    case LECCR_ERR_ALIGN: return "operand violates TMA alignment (16-byte base, ld % 8 == 0)";
    case LECCR_ERR_ARCH: return "device is not sm_100 (B200); leccr_b200 has no other backend";
    case LECCR_ERR_CUDA: return "CUDA runtime error (see leccr_last_cuda_error)";
    case LECCR_ERR_WORKSPACE: return "workspace too small";
    case LECCR_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    case LECCR_ERR_NCCL: return "libnccl.so.2 unavailable or an NCCL call failed (see leccr_last_cuda_error)";
    default: return "unknown leccr error code";
  }
}
const char* leccr_last_cuda_error(void) { return g_cuda_err; }
int leccr_abi_version(void) { return 2; }

void leccr_profile_enable(int on) {
  g_profile = on != 0;
  g_profile_ms = 0.0;
  g_profile_launches = 0;
  g_pairs_used = 0;
}
int leccr_profile_read(double* total_ms, int* launches) {
  if (total_ms == nullptr || launches == nullptr) return LECCR_ERR_ARG;
  CUDA_TRY(profile_resolve());
  *total_ms = g_profile_ms;
  *launches = g_profile_launches;
  prof_dump();
  return LECCR_OK;
}

int leccr_check_device(void) {
  int dev = 0, major = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? LECCR_OK : LECCR_ERR_ARCH;
}

// ------------------------------------------------------------------------------------ prep
int leccr_prep(const float* src, int64_t n, int D, int64_t ld_src, int normalize, int fmt, int layout,
               void* dst16, int64_t ld_dst, float* rn_hi, float* rn_lo, float* stats,
               leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (src == nullptr || dst16 == nullptr || n <= 0 || D <= 0 || bad_fmt(fmt) || layout < 0 || layout > 2)
    return LECCR_ERR_ARG;
  const int64_t K = layout == 0 ? D : 3LL * D;
  if (ld_src < D || ld_dst < K) return LECCR_ERR_ARG;
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((n + wpb - 1) / wpb);
  uint16_t* dst = static_cast<uint16_t*>(dst16);
  prof_mark("prep:begin", stream);
  const bool vec = (D % 128 == 0) && D <= 1024 && (ld_src % 4 == 0) && (ld_dst % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(src) % 16 == 0) && (reinterpret_cast<uintptr_t>(dst16) % 8 == 0);
  if (vec) {
    if (fmt == LECCR_FMT_F16)
      prep_rows_vec_kernel<0><<<grid, wpb * 32, 0, stream>>>(src, ld_src, (int)n, D, normalize, layout, dst,
                                                            ld_dst, rn_hi, rn_lo, stats);
    else
      prep_rows_vec_kernel<1><<<grid, wpb * 32, 0, stream>>>(src, ld_src, (int)n, D, normalize, layout, dst,
                                                            ld_dst, rn_hi, rn_lo, stats);
  } else if (fmt == LECCR_FMT_F16) {
    prep_rows_kernel<0><<<grid, wpb * 32, 0, stream>>>(src, ld_src, (int)n, D, normalize, layout, dst, ld_dst,
                                                      rn_hi, rn_lo, stats);
  } else {
    prep_rows_kernel<1><<<grid, wpb * 32, 0, stream>>>(src, ld_src, (int)n, D, normalize, layout, dst, ld_dst,
                                                      rn_hi, rn_lo, stats);
  }
  LAUNCH_CHECK("prep_rows_kernel");
  prof_mark("prep", stream);
  return LECCR_OK;
}

int leccr_prep_pair(const float* src0, int64_t n0, int64_t ld_src0, void* dst16_0, int64_t ld_dst0, float* rn_hi0,
                    float* rn_lo0, float* stats0, const float* src1, int64_t n1, int64_t ld_src1, void* dst16_1,
                    int64_t ld_dst1, float* rn_hi1, float* rn_lo1, float* stats1, int D, int normalize, int fmt,
                    int layout, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (src0 == nullptr || src1 == nullptr || dst16_0 == nullptr || dst16_1 == nullptr || n0 <= 0 || n1 <= 0 || D <= 0 ||
      bad_fmt(fmt) || layout < 0 || layout > 2)
    return LECCR_ERR_ARG;
  const int64_t K = layout == 0 ? D : 3LL * D;
  auto vec_ok = [&](const float* src, int64_t ld_src, void* dst, int64_t ld_dst) {
    return ld_src >= D && ld_dst >= K && (D % 128 == 0) && D <= 1024 && (ld_src % 4 == 0) && (ld_dst % 4 == 0) &&
           (reinterpret_cast<uintptr_t>(src) % 16 == 0) && (reinterpret_cast<uintptr_t>(dst) % 8 == 0);
  };
  if (!vec_ok(src0, ld_src0, dst16_0, ld_dst0) || !vec_ok(src1, ld_src1, dst16_1, ld_dst1)) {  // general path: two launches
    int rc = leccr_prep(src0, n0, D, ld_src0, normalize, fmt, layout, dst16_0, ld_dst0, rn_hi0, rn_lo0, stats0, stream_);
    if (rc != LECCR_OK) return rc;
    return leccr_prep(src1, n1, D, ld_src1, normalize, fmt, layout, dst16_1, ld_dst1, rn_hi1, rn_lo1, stats1, stream_);
  }
  const int wpb = 8;
  const int b0 = static_cast<int>((n0 + wpb - 1) / wpb), b1 = static_cast<int>((n1 + wpb - 1) / wpb);
  PrepTensor t0 = {src0, ld_src0, static_cast<int>(n0), static_cast<uint16_t*>(dst16_0), ld_dst0, rn_hi0, rn_lo0, stats0};
  PrepTensor t1 = {src1, ld_src1, static_cast<int>(n1), static_cast<uint16_t*>(dst16_1), ld_dst1, rn_hi1, rn_lo1, stats1};
  if (fmt == LECCR_FMT_F16)
    prep_rows_vec_pair_kernel<0><<<b0 + b1, wpb * 32, 0, stream>>>(t0, t1, b0, D, normalize, layout);
  else
    prep_rows_vec_pair_kernel<1><<<b0 + b1, wpb * 32, 0, stream>>>(t0, t1, b0, D, normalize, layout);
  LAUNCH_CHECK("prep_rows_vec_pair_kernel");
  return LECCR_OK;
}

int leccr_prep_push(const float* src, int64_t n, int D, int64_t ld_src, int normalize, int fmt,
                    void* const* dst_ptrs_dev, int world, int64_t dst_row0, int64_t dst_col0, int64_t ld_dst,
                    leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (src == nullptr || dst_ptrs_dev == nullptr || n <= 0 || D <= 0 || bad_fmt(fmt) || world < 1 || ld_src < D ||
      dst_row0 < 0 || dst_col0 < 0 || ld_dst < dst_col0 + D)
    return LECCR_ERR_ARG;
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((n + wpb - 1) / wpb);
  uint16_t* const* dsts = reinterpret_cast<uint16_t* const*>(dst_ptrs_dev);
  if (fmt == LECCR_FMT_F16)
    prep_push_kernel<0><<<grid, wpb * 32, 0, stream>>>(src, ld_src, (int)n, D, normalize, dsts, world, dst_row0,
                                                      dst_col0, ld_dst);
  else
    prep_push_kernel<1><<<grid, wpb * 32, 0, stream>>>(src, ld_src, (int)n, D, normalize, dsts, world, dst_row0,
                                                      dst_col0, ld_dst);
  LAUNCH_CHECK("prep_push_kernel");
  return LECCR_OK;
}

int leccr_push_words(const void* src, int64_t n_words, void* const* dst_ptrs_dev, int world, int64_t dst_word0,
                     leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (src == nullptr || dst_ptrs_dev == nullptr || n_words <= 0 || world < 1 || dst_word0 < 0) return LECCR_ERR_ARG;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((n_words + 255) / 256, 2LL * num_sms()));
  push_words_kernel<<<grid, 256, 0, stream>>>(static_cast<const unsigned long long*>(src), n_words,
                                             reinterpret_cast<unsigned long long* const*>(dst_ptrs_dev), world,
                                             dst_word0);
  LAUNCH_CHECK("push_words_kernel");
  return LECCR_OK;
}

int leccr_stats16(const void* src16, int fmt, int64_t n, int D, int64_t ld_src, float* rn_hi, float* rn_lo,
                  float* stats, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (src16 == nullptr || n <= 0 || D <= 0 || bad_fmt(fmt) || ld_src < D) return LECCR_ERR_ARG;
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((n + wpb - 1) / wpb);
  const uint16_t* s = static_cast<const uint16_t*>(src16);
  if (fmt == LECCR_FMT_F16)
    stats_rows16_kernel<0><<<grid, wpb * 32, 0, stream>>>(s, ld_src, (int)n, D, rn_hi, rn_lo, stats);
  else
    stats_rows16_kernel<1><<<grid, wpb * 32, 0, stream>>>(s, ld_src, (int)n, D, rn_hi, rn_lo, stats);
  LAUNCH_CHECK("stats_rows16_kernel");
  return LECCR_OK;
}

int leccr_transpose16(const void* src16, int64_t n, int D, int64_t ld_src, void* dst16, int64_t ld_dst,
                      leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (src16 == nullptr || dst16 == nullptr || n <= 0 || D <= 0 || ld_src < D || ld_dst < n) return LECCR_ERR_ARG;
  dim3 grid(static_cast<unsigned>((ld_dst + 31) / 32), static_cast<unsigned>((D + 31) / 32));
  dim3 block(32, 8);
  transpose16_kernel<<<grid, block, 0, stream>>>(static_cast<const uint16_t*>(src16), ld_src, (int)n, D,
                                                 static_cast<uint16_t*>(dst16), ld_dst);
  LAUNCH_CHECK("transpose16_kernel");
  return LECCR_OK;
}

// ------------------------------------------------------------------------------------ sim_f32
// Shared by leccr_sim_f32 and the gradient products: up to two problems, optional split-K into
// partial planes `parts` ([k_splits][n_rows * ld_out] per problem) followed by a fixed-order reduce.
struct StoreProblem {
  const void* rows16;
  const void* cols16;
  int64_t ld_rows, ld_cols, n_rows, n_cols;
  float* out;
  int64_t ld_out;
  float* parts;  // split-K partial planes (nullptr when k_splits <= 1)
  // cols16 is the row-major [K][n_cols] source itself (MN-major operand) instead of a K-major [n_cols][K] matrix:
  // no transposed copy of an operand whose contraction index is its row index (all problems of a launch alike)
  bool cols_mn;
};

static int launch_store(const StoreProblem* sp, int n_prob, int K, int fmt, float scale, const float* scale_dev,
                        const float* div_dev, int k_splits, cudaStream_t stream, const float* prod_a = nullptr,
                        const float* prod_b = nullptr, float* prod_out = nullptr) {
  SimLaunch L;
  memset(&L, 0, sizeof(L));
  L.n_prob = n_prob;
  L.fmt = fmt;
  L.k_chunks = (K + BK - 1) / BK;
  EpiStore::Params EP;
  memset(&EP, 0, sizeof(EP));
  const bool split = k_splits > 1;
  if (split) {
    L.kc_per_split = (L.k_chunks + k_splits - 1) / k_splits;
    L.k_splits = (L.k_chunks + L.kc_per_split - 1) / L.kc_per_split;
    if (L.k_splits < 2) {  // keep the split decode (all column tiles per item) with one plane
      L.k_splits = 2;
      L.kc_per_split = L.k_chunks;
    }
  }
  int item_base = 0;
  int n_planes = 1;
  for (int p = 0; p < n_prob; ++p) {
    Plan pl;
    if (split) {
      pl = plan_problem(sp[p].n_cols, 0, sp[p].n_rows, 1 << 30);
      pl.n_chunks = (L.k_chunks + L.kc_per_split - 1) / L.kc_per_split;
      n_planes = pl.n_chunks;
    } else {
      int64_t tiles = 0;
      for (int q = 0; q < n_prob; ++q)
        tiles += ((sp[q].n_rows + BM - 1) / BM) * ((sp[q].n_cols + BN - 1) / BN);
      pl = plan_problem(sp[p].n_cols, 0, sp[p].n_rows, auto_tiles_per_chunk(tiles, 1));
    }
    if (sp[p].cols_mn != sp[0].cols_mn) return LECCR_ERR_ARG;
    int rc = fill_problem(L.prob[p], sp[p].rows16, sp[p].ld_rows, sp[p].cols_mn ? sp[p].rows16 : sp[p].cols16,
                          sp[p].cols_mn ? sp[p].ld_rows : sp[p].ld_cols, sp[p].n_rows, sp[p].cols_mn ? sp[p].n_rows : sp[p].n_cols,
                          K, fmt, pl, item_base);
    if (rc == LECCR_OK && sp[p].cols_mn) {
      rc = make_tmap_mn(&L.prob[p].tm_cols, sp[p].cols16, K, sp[p].n_cols, sp[p].ld_cols, fmt);
      L.prob[p].n_cols = static_cast<int>(sp[p].n_cols);
    }
    if (rc != LECCR_OK) return rc;
    item_base += pl.row_blocks * pl.n_chunks;
    EP.out[p] = split ? sp[p].parts : sp[p].out;
    EP.ld[p] = sp[p].ld_out;
    EP.scale[p] = scale;
    EP.scale_ptr[p] = scale_dev;
    EP.div_ptr[p] = div_dev;
    EP.split_stride[p] = split ? sp[p].n_rows * sp[p].ld_out : 0;
  }
  L.n_items = item_base;
  int rc = sp[0].cols_mn ? launch_gemm<EpiStore, BK, false, true>(L, EP, stream) : launch_gemm<EpiStore>(L, EP, stream);
  if (rc != LECCR_OK) return rc;
  bool prod_done = false;
  if (split) {
    const long long plane0 = sp[0].n_rows * sp[0].ld_out;
    if (n_prob == 2 && sp[1].n_rows * sp[1].ld_out == plane0) {  // one launch for both (and the scalar product)
      dim3 g(static_cast<unsigned>(std::min<long long>((plane0 + 255) / 256, 2LL * num_sms())), 2);
      splitk_reduce2_kernel<<<g, 256, 0, stream>>>(sp[0].parts, sp[1].parts, n_planes, plane0, sp[0].out, sp[1].out,
                                                  prod_a, prod_b, prod_out);
      LAUNCH_CHECK("splitk_reduce2_kernel");
      prod_done = true;
    } else {
      for (int p = 0; p < n_prob; ++p) {
        const long long plane = sp[p].n_rows * sp[p].ld_out;
        const unsigned g = static_cast<unsigned>(std::min<long long>((plane + 255) / 256, 4LL * num_sms()));
        splitk_reduce_kernel<<<g, 256, 0, stream>>>(sp[p].parts, n_planes, plane, sp[p].out);
        LAUNCH_CHECK("splitk_reduce_kernel");
      }
    }
  }
  if (prod_out != nullptr && !prod_done) {
    scalar_product_kernel<<<1, 32, 0, stream>>>(prod_a, prod_b, prod_out);
    LAUNCH_CHECK("scalar_product_kernel");
  }
  return LECCR_OK;
}

int leccr_sim_f32(const void* rows16, int64_t ld_rows, const void* cols16, int64_t ld_cols, int64_t n_rows,
                  int64_t n_cols, int K, int fmt, float* S, int64_t ld_S, float scale, const float* scale_dev,
                  leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows16 == nullptr || cols16 == nullptr || S == nullptr || n_rows <= 0 || n_cols <= 0 || K <= 0 ||
      bad_fmt(fmt) || ld_S < n_cols)
    return LECCR_ERR_ARG;
  int rc = leccr_check_device();
  if (rc != LECCR_OK) return rc;
  StoreProblem sp = {rows16, cols16, ld_rows, ld_cols, n_rows, n_cols, S, ld_S, nullptr};
  return launch_store(&sp, 1, K, fmt, scale, scale_dev, nullptr, 1, stream);
}

// ------------------------------------------------------------------------------------ sim_topk
// Work decomposition of a top-k launch.  topk_finalize holds 2 * n_chunks candidates per lane,
// so a row's columns are split into at most kMaxTopkChunks chunks.  Each problem gets a
// share of ~4 work items per SM proportional to its tile count.
// Two epilogue shapes (see EpiTopK): one warpgroup with 64-entry lists, or two with 32-entry lists.
using TopK1 = EpiTopK<LECCR_TOPK_KP, 64, 1>;
using TopK1D = EpiTopK<LECCR_TOPK_KP, 64, 1, 1>;  // dense-only build of TopK1 (short column chunks)
// Short column chunks are bound by the issue latency of the single epilogue warp per scheduler (ncu: issue
// slots 36 % busy).  TopK2D runs TWO epilogue warpgroups (alternate tiles, each with its own 64-entry lists);
// their 130 KB of lists fit beside the pipeline because its stages are half as deep (K = 32, 64-byte swizzle).
using TopK2D = EpiTopK<LECCR_TOPK_KP, 64, 2, 1>;
constexpr int kBK2 = 32;
// Long column chunks (cfg5): the filter epilogue on ONE warp per scheduler costs 6-8 % against the mainloop and
// its hit handling another 10 %; TopK2F runs it on two warpgroups (half the columns of every tile each) with
// 31-entry lists, which fit beside the resident row block and three K = 64 stages (hits are rare on long rows,
// so short lists cost few extra shrink rounds).
using TopK2F = EpiTopK<LECCR_TOPK_KP, 31, 2, 0>;
constexpr int kMaxTopkChunks = 8;  // topk_finalize holds n_chunks * kWGs * (C / 32) <= 16 slots per lane

static bool topk_two_wgs_allowed() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LECCR_TOPK_WGS");  // measurement aid: 1 keeps the single-warpgroup shapes
    v = (e != nullptr && atoi(e) == 1) ? 0 : 1;
  }
  return v == 1;
}
static bool topk_a_resident() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LECCR_TOPK_ARES");  // measurement aid: 0 streams the row block with the gallery
    v = (e != nullptr && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}
// Dense mode (and with it the two-warpgroup shape) is chosen when every tensor-core problem of the launch has
// short column chunks; decided from the plans alone so that workspace sizing and launch agree.
static bool topk_dense(const Plan* plans, int n_prob, const int* gemm_mask) {
  static int never = -1;
  if (never < 0) {
    const char* e = getenv("LECCR_TOPK_WGS");  // measurement aid: 3 = filter shapes even for short chunks
    never = (e != nullptr && atoi(e) == 3) ? 1 : 0;
  }
  if (never) return false;
  for (int p = 0; p < n_prob; ++p)
    if ((gemm_mask == nullptr || gemm_mask[p]) && plans[p].tiles_per_chunk > 32) return false;
  return true;
}

static int topk_plan(const leccr_topk_problem* probs, int n_prob, int tiles_per_chunk, Plan* plans) {
  int64_t total_tiles = 0;
  for (int p = 0; p < n_prob; ++p)
    total_tiles += ((probs[p].n_rows + BM - 1) / BM) * ((probs[p].n_cols + BN - 1) / BN);
  for (int p = 0; p < n_prob; ++p) {
    const int64_t row_blocks = (probs[p].n_rows + BM - 1) / BM;
    const int64_t col_tiles = (probs[p].n_cols + BN - 1) / BN;
    int tpc = tiles_per_chunk;
    if (tpc <= 0) {
      const double share = 4.0 * num_sms() * static_cast<double>(row_blocks * col_tiles) / static_cast<double>(total_tiles);
      int64_t chunks = static_cast<int64_t>(share / static_cast<double>(row_blocks) + 0.5);
      // (More chunks than load balance needs do not pay: later chunks inherit a threshold from the
      // earlier ones, but a stale one, and the total number of list inserts per row grows.)
      chunks = std::max<int64_t>(1, std::min<int64_t>(chunks, std::min<int64_t>(kMaxTopkChunks, col_tiles)));
      if (n_prob == 1 && col_tiles >= 256) {
        // Long rows, one problem: the items are equally long and the grid is persistent, so what matters is that
        // the last wave is full (782 row blocks x 1 chunk = 5.3 waves of 148 would idle 12 % of the machine).
        // Take the smallest chunk count >= the balance rule's whose waves are >= 97 % full, else the fullest.
        const int64_t sms = num_sms();
        int64_t best_c = chunks;
        double best_eff = 0.0;
        for (int64_t c = chunks; c <= std::min<int64_t>(kMaxTopkChunks, col_tiles); ++c) {
          const int64_t items = row_blocks * c;
          const double eff = static_cast<double>(items) / static_cast<double>((items + sms - 1) / sms * sms);
          if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best_c = c;
          }
          if (eff >= 0.97) break;
        }
        chunks = best_c;
      }
      tpc = static_cast<int>((col_tiles + chunks - 1) / chunks);
    }
    plans[p] = plan_problem(probs[p].n_cols, 0, probs[p].n_rows, tpc);
    if (plans[p].n_chunks > kMaxTopkChunks) return LECCR_ERR_ARG;
  }
  return LECCR_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// Per-problem workspace: candidate lists [n_rows][subs][64] (score, column), counts [n_rows][subs],
// undecided-row list, shared per-row thresholds.
static size_t topk_ws_bytes(int64_t n_rows, int subs, bool two = true) {
  // lists of 32 entries: one warpgroup keeps 64 entries per (row, column chunk), two keep 128
  const size_t lists = static_cast<size_t>(n_rows) * subs * (two ? 4 : 2);
  return align256(lists * 32 * 4) * 2 + align256(lists * 4) + align256(static_cast<size_t>(n_rows) * 4 + 16) +
         align256(static_cast<size_t>(n_rows) * 12);  // row_thr [n_rows] + row_h8 [n_rows][2]
}

size_t leccr_sim_topk_workspace(const leccr_topk_problem* probs, int n_prob, int tiles_per_chunk) {
  if (probs == nullptr || n_prob < 1 || n_prob > 2) return 0;
  Plan plans[2];
  if (topk_plan(probs, n_prob, tiles_per_chunk, plans) != LECCR_OK) return 0;
  size_t bytes = 0;
  const bool two = topk_two_wgs_allowed() && topk_dense(plans, n_prob, nullptr);
  for (int p = 0; p < n_prob; ++p) bytes += topk_ws_bytes(probs[p].n_rows, plans[p].n_chunks, two);
  return bytes;
}

size_t leccr_sim_topk_stream_workspace(int64_t n_rows, int sub_total) {
  if (n_rows <= 0 || sub_total < 1 || sub_total > kMaxTopkChunks) return 0;
  return topk_ws_bytes(n_rows, sub_total, true);
}

// The engine behind leccr_sim_topk and leccr_sim_topk_stream: per problem an optional tensor-core phase over
// the columns given (writing candidate-list slots [sub_begin, sub_begin + chunks) of the problem's
// persistent workspace) and an optional finalize phase over all sub_total slots.
static int topk_core(const leccr_topk_problem* probs, const leccr_topk_stream* so, const Plan* plans, int n_prob, int D,
                     int fmt, int k, int two_hint, cudaStream_t stream) {
  SimLaunch L;
  memset(&L, 0, sizeof(L));
  L.fmt = fmt;
  int gemm_mask[2] = {0, 0};
  for (int p = 0; p < n_prob; ++p) gemm_mask[p] = (so[p].phases & LECCR_TOPK_GEMM) ? 1 : 0;
  const bool dense_launch = topk_dense(plans, n_prob, gemm_mask);
  // a problem's lists must have ONE shape over all its calls: streamed problems (two_hint) say so themselves
  const bool two = two_hint >= 0 ? (two_hint == 1) : (topk_two_wgs_allowed() && dense_launch);
  const bool two_f = !two && topk_two_wgs_allowed();  // filter epilogue on two warpgroups
  const int wgs = (two || two_f) ? 2 : 1;
  const int cap = two_f ? TopK2F::C : TopK1::C;
  const int bk = two ? kBK2 : BK;
  L.k_chunks = (D + bk - 1) / bk;
  TopK1::Params EP;  // all shapes share the parameter layout
  static_assert(sizeof(TopK1::Params) == sizeof(TopK2D::Params) && sizeof(TopK1::Params) == sizeof(TopK2F::Params),
                "parameter layouts must agree");
  memset(&EP, 0, sizeof(EP));
  if (const char* dbg = getenv("LECCR_TOPK_DEBUG")) EP.debug_mode = atoi(dbg);  // measurement aid only
  // shrink rounds start when a list holds more than `trig`; long chunks prefer fresher thresholds
  // (fewer candidates pass), short ones fewer rounds (measured optimum is flat between 36 and 56)
  EP.trig = two_f ? TopK2F::TRIG : (plans[0].tiles_per_chunk >= 64 ? 40 : TopK1::TRIG);
  // short column chunks never leave the warm-up regime: run them without the filter (measured: cfg2
  // 342 -> 304 us; long chunks are better off filtering: 12,500 x 1M 1202 vs 832 TFLOP/s)
  EP.dense = dense_launch ? 1 : 0;
  if (const char* dn = getenv("LECCR_TOPK_DENSE")) EP.dense = atoi(dn);  // measurement aid
  if (two) EP.dense = 1;  // the two-warpgroup shape for short chunks exists as a dense-only build
  if (two_f) EP.dense = 0;
  if (const char* tg = getenv("LECCR_TOPK_TRIG")) EP.trig = std::min(EP.trig, std::max(LECCR_TOPK_KP + 4, atoi(tg)));
  if (const char* dc = getenv("LECCR_TOPK_COUNTERS")) {  // measurement aid only: device address (hex) of 5 x u64
    EP.debug_counters = reinterpret_cast<unsigned long long*>(strtoull(dc, nullptr, 16));
  }
  float* cand_val[2];
  int* cand_idx[2];
  int* cand_cnt[2];
  int* flag[2];
  int item_base = 0;
  int n_launch = 0;  // problems taking part in the tensor-core launch
  for (int p = 0; p < n_prob; ++p) {
    const leccr_topk_problem& q = probs[p];
    const leccr_topk_stream& o = so[p];
    if (o.workspace == nullptr || o.workspace_bytes < topk_ws_bytes(q.n_rows, o.sub_total, two)) return LECCR_ERR_WORKSPACE;
    uint8_t* ws = static_cast<uint8_t*>(o.workspace);
    const size_t lists = static_cast<size_t>(q.n_rows) * o.sub_total * (two ? 4 : 2);
    cand_val[p] = reinterpret_cast<float*>(ws);
    ws += align256(lists * 32 * 4);
    cand_idx[p] = reinterpret_cast<int*>(ws);
    ws += align256(lists * 32 * 4);
    cand_cnt[p] = reinterpret_cast<int*>(ws);
    ws += align256(lists * 4);
    flag[p] = reinterpret_cast<int*>(ws);  // [0] count, [4..] list of undecided rows
    ws += align256(static_cast<size_t>(q.n_rows) * 4 + 16);
    unsigned* row_thr = reinterpret_cast<unsigned*>(ws);
    // does this one call fill every slot?  (fewer column chunks than slots leave empty lists behind)
    const bool whole = (o.phases & LECCR_TOPK_GEMM) && o.sub_begin == 0 && plans[p].n_chunks == o.sub_total;
    if (o.phases & LECCR_TOPK_INIT) {
      CUDA_TRY(cudaMemsetAsync(flag[p], 0, 16, stream));
      CUDA_TRY(cudaMemsetAsync(row_thr, 0, static_cast<size_t>(q.n_rows) * 12, stream));
      if (!whole) CUDA_TRY(cudaMemsetAsync(cand_cnt[p], 0, lists * 4, stream));
    }
    if (o.phases & LECCR_TOPK_GEMM) {
      const int g = n_launch++;
      int rc = fill_problem(L.prob[g], q.rows16, q.ld_rows16, q.cols16, q.ld_cols16, q.n_rows, q.n_cols, D, fmt,
                            plans[p], item_base, bk);
      if (rc != LECCR_OK) return rc;
      item_base += plans[p].row_blocks * plans[p].n_chunks;
      // every owner of a row (column chunks of this call, earlier calls, the other warpgroup) cooperates
      // through a shared per-row threshold
      if (o.sub_total > 1 || wgs == 2) EP.row_thr[g] = row_thr;
      if (wgs == 2) EP.row_h8[g] = row_thr + q.n_rows;
      EP.out_val[g] = cand_val[p];
      EP.out_idx[g] = cand_idx[p];
      EP.out_cnt[g] = cand_cnt[p];
      EP.n_sub[g] = o.sub_total * wgs;
      EP.sub_base[g] = o.sub_begin * wgs;
      EP.col_base[g] = static_cast<int>(o.col_begin);
    }
  }
  L.n_prob = n_launch;
  L.n_items = item_base;
  int rc = LECCR_OK;
  if (n_launch > 0) {
    prof_mark("topk:begin", stream);
    if (two) {
      TopK2D::Params EP2;
      memcpy(&EP2, &EP, sizeof(EP2));
      rc = launch_gemm<TopK2D, kBK2>(L, EP2, stream);
    } else if (EP.dense == 1 && EP.debug_mode == 0 && EP.debug_counters == nullptr) {
      TopK1D::Params EPD;
      static_assert(sizeof(TopK1D::Params) == sizeof(TopK1::Params), "parameter layouts must agree");
      memcpy(&EPD, &EP, sizeof(EPD));
      rc = launch_gemm<TopK1D>(L, EPD, stream);
    } else if (two_f) {
      TopK2F::Params EPF;
      memcpy(&EPF, &EP, sizeof(EPF));
      if (L.k_chunks <= kAResChunks && topk_a_resident()) rc = launch_gemm<TopK2F, BK, true>(L, EPF, stream);
      else rc = launch_gemm<TopK2F>(L, EPF, stream);
    } else if (L.k_chunks <= kAResChunks && topk_a_resident()) {
      rc = launch_gemm<TopK1, BK, true>(L, EP, stream);  // row block resident, only the gallery streams
    } else {
      rc = launch_gemm<TopK1>(L, EP, stream);
    }
    if (rc != LECCR_OK) return rc;
    prof_mark("topk:gemm", stream);
  }

  TopkFinalizeParams post[2], fin[2];
  int32_t* post_counts[2] = {nullptr, nullptr};
  int fin_slots[2] = {0, 0};
  int n_post = 0, n_fin = 0;
  for (int p = 0; p < n_prob; ++p) {
    if (!(so[p].phases & LECCR_TOPK_FINALIZE)) continue;
    const leccr_topk_problem& q = probs[p];
    TopkFinalizeParams F;
    memset(&F, 0, sizeof(F));
    F.cand_val = cand_val[p];
    F.cand_idx = cand_idx[p];
    F.cand_cnt = cand_cnt[p];
    F.n_rows = static_cast<int>(q.n_rows);
    F.n_cols = static_cast<int>(so[p].n_cols_total > 0 ? so[p].n_cols_total : q.n_cols);
    F.n_chunks = so[p].sub_total * wgs;
    F.list_cap = cap;
    F.KP = LECCR_TOPK_KP;
    F.k = k;
    F.topk_val = q.topk_val;
    F.topk_idx = q.topk_idx;
    F.gt_off = q.gt_off;
    F.gt_ids = q.gt_ids;
    F.rows_x = q.rows_x;
    F.cols_x = q.cols_x;
    F.ld_rows = q.ld_rows_x;
    F.ld_cols = q.ld_cols_x;
    F.D = D;
    F.x_dtype = q.x_dtype;
    F.rn_hi = q.rn_hi;
    F.rn_lo = q.rn_lo;
    F.col_stats = q.col_stats;
    // fp32 accumulation of K products in the tensor core: 2 ulp per product, conservatively
    F.acc_slack = 2.0f * 1.1920929e-7f * static_cast<float>(D);
    F.rank = q.rank;
    F.flag_count = flag[p];
    F.flag_list = flag[p] + 4;
    F.gt_score = q.gt_score;
    fin[n_fin] = F;
    fin_slots[n_fin] = F.n_chunks;  // one 32-entry slot per list
    ++n_fin;
    if (q.gt_off != nullptr) {
      post[n_post] = F;
      post_counts[n_post] = q.recall_counts;
      ++n_post;
    }
  }
  if (n_fin == 2) {
    // one launch for both problems; the pair kernel is instantiated for S0 >= S1
    if (fin_slots[0] < fin_slots[1]) {
      std::swap(fin[0], fin[1]);
      std::swap(fin_slots[0], fin_slots[1]);
    }
    const int b0 = (fin[0].n_rows + kFinalizeWarps - 1) / kFinalizeWarps;
    const int b1 = (fin[1].n_rows + kFinalizeWarps - 1) / kFinalizeWarps;
    auto cls = [](int sl) { return sl <= 2 ? 0 : sl <= 4 ? 1 : sl <= 8 ? 2 : 3; };
    const int c0 = cls(fin_slots[0]), c1 = cls(fin_slots[1]);
#define LECCR_FIN_PAIR(A, B) topk_finalize_pair_kernel<A, B><<<b0 + b1, kFinalizeWarps * 32, 0, stream>>>(fin[0], fin[1], b0)
    switch (c0 * 4 + c1) {
      case 0: LECCR_FIN_PAIR(2, 2); break;
      case 4: LECCR_FIN_PAIR(4, 2); break;
      case 5: LECCR_FIN_PAIR(4, 4); break;
      case 8: LECCR_FIN_PAIR(8, 2); break;
      case 9: LECCR_FIN_PAIR(8, 4); break;
      case 10: LECCR_FIN_PAIR(8, 8); break;
      case 12: LECCR_FIN_PAIR(kMaxSlots, 2); break;
      case 13: LECCR_FIN_PAIR(kMaxSlots, 4); break;
      case 14: LECCR_FIN_PAIR(kMaxSlots, 8); break;
      default: LECCR_FIN_PAIR(kMaxSlots, kMaxSlots); break;
    }
#undef LECCR_FIN_PAIR
    LAUNCH_CHECK("topk_finalize_pair_kernel");
    prof_mark("topk:finalize", stream);
  } else if (n_fin == 1) {
    const TopkFinalizeParams& F = fin[0];
    const unsigned grid = static_cast<unsigned>((F.n_rows + kFinalizeWarps - 1) / kFinalizeWarps);
    const int slots = fin_slots[0];
    if (slots <= 2) topk_finalize_kernel<2><<<grid, kFinalizeWarps * 32, 0, stream>>>(F);
    else if (slots <= 4) topk_finalize_kernel<4><<<grid, kFinalizeWarps * 32, 0, stream>>>(F);
    else if (slots <= 8) topk_finalize_kernel<8><<<grid, kFinalizeWarps * 32, 0, stream>>>(F);
    else topk_finalize_kernel<kMaxSlots><<<grid, kFinalizeWarps * 32, 0, stream>>>(F);
    LAUNCH_CHECK("topk_finalize_kernel");
    prof_mark("topk:finalize", stream);
  }
  if (n_post > 0) {  // exact fallback for flagged rows + Recall counts of every finalised problem: one launch
    if (n_post == 1) memset(&post[1], 0, sizeof(post[1]));  // gt_off == nullptr: the second slice exits at once
    dim3 g(static_cast<unsigned>(num_sms()), static_cast<unsigned>(n_post));
    rank_post_kernel<<<g, 256, 0, stream>>>(post[0], post[1], post_counts[0], post_counts[1]);
    LAUNCH_CHECK("rank_post_kernel");
    prof_mark("topk:rank_post", stream);
  }
  return LECCR_OK;
}

static int topk_check_problem(const leccr_topk_problem& q, bool need_finalize) {
  if (q.rows16 == nullptr || q.cols16 == nullptr || q.n_rows <= 0 || q.n_cols <= 0) return LECCR_ERR_ARG;
  if (need_finalize && (q.topk_val == nullptr || q.topk_idx == nullptr)) return LECCR_ERR_ARG;
  if (need_finalize && q.gt_off != nullptr &&
      (q.gt_ids == nullptr || q.rows_x == nullptr || q.cols_x == nullptr || q.rn_hi == nullptr ||
       q.rn_lo == nullptr || q.col_stats == nullptr || q.rank == nullptr || q.x_dtype < 0 || q.x_dtype > 2))
    return LECCR_ERR_ARG;
  return LECCR_OK;
}

int leccr_sim_topk(const leccr_topk_problem* probs, int n_prob, int D, int fmt, int k, int tiles_per_chunk,
                   void* workspace, size_t workspace_bytes, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (probs == nullptr || n_prob < 1 || n_prob > 2 || D <= 0 || bad_fmt(fmt) || k < 1 || k > LECCR_TOPK_KP)
    return LECCR_ERR_ARG;
  for (int p = 0; p < n_prob; ++p)
    if (topk_check_problem(probs[p], true) != LECCR_OK) return LECCR_ERR_ARG;
  int rc = leccr_check_device();
  if (rc != LECCR_OK) return rc;
  Plan plans[2];
  rc = topk_plan(probs, n_prob, tiles_per_chunk, plans);
  if (rc != LECCR_OK) return rc;  // tiles_per_chunk too small: more than kMaxTopkChunks column chunks per row
  if (workspace == nullptr || workspace_bytes < leccr_sim_topk_workspace(probs, n_prob, tiles_per_chunk))
    return LECCR_ERR_WORKSPACE;
  leccr_topk_stream so[2];
  memset(so, 0, sizeof(so));
  const bool two = topk_two_wgs_allowed() && topk_dense(plans, n_prob, nullptr);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  for (int p = 0; p < n_prob; ++p) {
    so[p].phases = LECCR_TOPK_INIT | LECCR_TOPK_GEMM | LECCR_TOPK_FINALIZE;
    so[p].sub_begin = 0;
    so[p].sub_count = so[p].sub_total = plans[p].n_chunks;
    so[p].workspace = ws;
    so[p].workspace_bytes = topk_ws_bytes(probs[p].n_rows, plans[p].n_chunks, two);
    ws += so[p].workspace_bytes;
  }
  return topk_core(probs, so, plans, n_prob, D, fmt, k, two ? 1 : 0, stream);
}

int leccr_sim_topk_stream(const leccr_topk_problem* probs, const leccr_topk_stream* streams, int n_prob, int D,
                          int fmt, int k, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (probs == nullptr || streams == nullptr || n_prob < 1 || n_prob > 2 || D <= 0 || bad_fmt(fmt) || k < 1 ||
      k > LECCR_TOPK_KP)
    return LECCR_ERR_ARG;
  Plan plans[2];
  for (int p = 0; p < n_prob; ++p) {
    const leccr_topk_stream& o = streams[p];
    if ((o.phases & ~(LECCR_TOPK_INIT | LECCR_TOPK_GEMM | LECCR_TOPK_FINALIZE | LECCR_TOPK_LONG)) != 0 ||
        (o.phases & ~LECCR_TOPK_LONG) == 0 || ((o.phases ^ streams[0].phases) & LECCR_TOPK_LONG) != 0 ||
        o.sub_total < 1 || o.sub_total > kMaxTopkChunks || o.sub_begin < 0 || o.sub_count < 1 ||
        o.sub_begin + o.sub_count > o.sub_total || o.col_begin < 0 || o.col_begin > 0x7fffffffLL)
      return LECCR_ERR_ARG;
    if (topk_check_problem(probs[p], (o.phases & LECCR_TOPK_FINALIZE) != 0) != LECCR_OK) return LECCR_ERR_ARG;
    const int64_t col_tiles = (probs[p].n_cols + BN - 1) / BN;
    const int tpc = static_cast<int>((col_tiles + o.sub_count - 1) / o.sub_count);
    plans[p] = plan_problem(probs[p].n_cols, 0, probs[p].n_rows, tpc);
  }
  int rc = leccr_check_device();
  if (rc != LECCR_OK) return rc;
  // A streamed problem's lists keep one shape from INIT to FINALIZE: the two-warpgroup dense shape, which
  // requires every window's column chunks to be short (<= 32 tiles): choose sub_count accordingly -- or, when
  // every call of the problem carries LECCR_TOPK_LONG (gallery windows of a large-gallery search), the filter
  // shape of the one-shot path's long chunks.
  const bool two = topk_two_wgs_allowed() && !(streams[0].phases & LECCR_TOPK_LONG);
  if (two) {
    for (int p = 0; p < n_prob; ++p)
      if ((streams[p].phases & LECCR_TOPK_GEMM) && plans[p].tiles_per_chunk > 32) return LECCR_ERR_ARG;
  }
  return topk_core(probs, streams, plans, n_prob, D, fmt, k, two ? 1 : 0, stream);
}

// ------------------------------------------------------------------------------------ Recall only (EpiRank)
// Similarity + exact ranks of the ground truth + Recall@1/5/10 with NO candidate lists: what itm_eval returns.
static size_t rank_ws_bytes(int64_t n_rows) {
  const size_t cap = std::max<size_t>(4096, static_cast<size_t>(n_rows) * 64);
  return 3 * align256(static_cast<size_t>(n_rows) * 4) /* best, lo, hi */ + align256(static_cast<size_t>(n_rows) * 4) /* row_flag */ +
         align256(static_cast<size_t>(n_rows) * 4 + 16) /* flag count + list */ + 256 /* pair count */ + align256(cap * 8);
}

size_t leccr_sim_rank_workspace(const leccr_topk_problem* probs, int n_prob) {
  if (probs == nullptr || n_prob < 1 || n_prob > 2) return 0;
  size_t b = 0;
  for (int p = 0; p < n_prob; ++p) b += rank_ws_bytes(probs[p].n_rows);
  return b;
}

int leccr_sim_rank(const leccr_topk_problem* probs, int n_prob, int D, int fmt, void* workspace, size_t workspace_bytes,
                   leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (probs == nullptr || n_prob < 1 || n_prob > 2 || D <= 0 || bad_fmt(fmt)) return LECCR_ERR_ARG;
  for (int p = 0; p < n_prob; ++p) {
    const leccr_topk_problem& q = probs[p];
    if (q.rows16 == nullptr || q.cols16 == nullptr || q.n_rows <= 0 || q.n_cols <= 0 || q.gt_off == nullptr ||
        q.gt_ids == nullptr || q.rows_x == nullptr || q.cols_x == nullptr || q.rn_hi == nullptr || q.rn_lo == nullptr ||
        q.col_stats == nullptr || q.rank == nullptr || q.x_dtype < 0 || q.x_dtype > 2)
      return LECCR_ERR_ARG;
  }
  int rc = leccr_check_device();
  if (rc != LECCR_OK) return rc;
  if (workspace == nullptr || workspace_bytes < leccr_sim_rank_workspace(probs, n_prob)) return LECCR_ERR_WORKSPACE;
  // work decomposition: ~6 equally long items per SM over both problems (counts are accumulated atomically, so a
  // row may be split into any number of column chunks)
  int64_t total_tiles = 0;
  for (int p = 0; p < n_prob; ++p)
    total_tiles += ((probs[p].n_rows + BM - 1) / BM) * ((probs[p].n_cols + BN - 1) / BN);
  int tpc = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(total_tiles / (6LL * num_sms()), 64)));
  if (const char* e = getenv("LECCR_RANK_TPC")) tpc = std::max(1, atoi(e));  // measurement aid
  SimLaunch L;
  memset(&L, 0, sizeof(L));
  L.n_prob = n_prob;
  L.fmt = fmt;
  L.k_chunks = (D + BK - 1) / BK;
  EpiRank::Params EP;
  memset(&EP, 0, sizeof(EP));
  TopkFinalizeParams post[2];
  memset(post, 0, sizeof(post));
  float* best[2];
  GtBestParams gtp[2];
  memset(gtp, 0, sizeof(gtp));
  int64_t max_rows = 0;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int item_base = 0;
  for (int p = 0; p < n_prob; ++p) {
    const leccr_topk_problem& q = probs[p];
    const size_t nb = align256(static_cast<size_t>(q.n_rows) * 4);
    const int cap = static_cast<int>(std::max<size_t>(4096, static_cast<size_t>(q.n_rows) * 64));
    best[p] = reinterpret_cast<float*>(ws);
    float* lo = reinterpret_cast<float*>(ws + nb);
    float* hi = reinterpret_cast<float*>(ws + 2 * nb);
    ws += 3 * nb;
    int* row_flag = reinterpret_cast<int*>(ws);
    ws += nb;
    int* flag = reinterpret_cast<int*>(ws);
    ws += align256(static_cast<size_t>(q.n_rows) * 4 + 16);
    int* amb_count = reinterpret_cast<int*>(ws);
    ws += 256;
    int2* amb_list = reinterpret_cast<int2*>(ws);
    ws += align256(static_cast<size_t>(cap) * 8);
    GtBestParams& G = gtp[p];
    G.gt_off = q.gt_off;
    G.gt_ids = q.gt_ids;
    G.rows_x = q.rows_x;
    G.cols_x = q.cols_x;
    G.ld_rows = q.ld_rows_x;
    G.ld_cols = q.ld_cols_x;
    G.D = D;
    G.x_dtype = q.x_dtype;
    G.n_rows = static_cast<int>(q.n_rows);
    G.rn_hi = q.rn_hi;
    G.rn_lo = q.rn_lo;
    G.col_stats = q.col_stats;
    G.acc_slack = 2.0f * 1.1920929e-7f * static_cast<float>(D);
    G.best = best[p];
    G.lo = lo;
    G.hi = hi;
    G.gt_score = q.gt_score;
    G.rank = q.rank;
    G.row_flag = row_flag;
    G.flag_words = flag;
    G.amb_count = amb_count;
    max_rows = std::max<int64_t>(max_rows, q.n_rows);
    const Plan pl = plan_problem(q.n_cols, 0, q.n_rows, tpc);
    rc = fill_problem(L.prob[p], q.rows16, q.ld_rows16, q.cols16, q.ld_cols16, q.n_rows, q.n_cols, D, fmt, pl, item_base);
    if (rc != LECCR_OK) return rc;
    item_base += pl.row_blocks * pl.n_chunks;
    EP.lo[p] = lo;
    EP.hi[p] = hi;
    EP.rank[p] = q.rank;
    EP.amb_count[p] = amb_count;
    EP.amb_list[p] = amb_list;
    EP.amb_cap[p] = cap;
    EP.row_flag[p] = row_flag;
    EP.flag_count[p] = flag;
    EP.flag_list[p] = flag + 4;
    TopkFinalizeParams& F = post[p];
    F.n_rows = static_cast<int>(q.n_rows);
    F.n_cols = static_cast<int>(q.n_cols);
    F.gt_off = q.gt_off;
    F.gt_ids = q.gt_ids;
    F.rows_x = q.rows_x;
    F.cols_x = q.cols_x;
    F.ld_rows = q.ld_rows_x;
    F.ld_cols = q.ld_cols_x;
    F.D = D;
    F.x_dtype = q.x_dtype;
    F.rank = q.rank;
    F.flag_count = flag;
    F.flag_list = flag + 4;
    F.gt_score = nullptr;
  }
  L.n_items = item_base;
  {  // exact best ground-truth scores, bands, zeroed flags and counters: one launch for both problems
    dim3 gg(static_cast<unsigned>((max_rows + 7) / 8), static_cast<unsigned>(n_prob));
    gt_best_kernel<<<gg, 256, 0, stream>>>(gtp[0], gtp[1]);
    LAUNCH_CHECK("gt_best_kernel");
  }
  if (const char* dbg = getenv("LECCR_RANK_DEBUG")) EP.debug_mode = atoi(dbg);  // measurement aid only
  static int ares = -1;
  if (ares < 0) {
    const char* e = getenv("LECCR_RANK_ARES");  // measurement aid: 0 streams the row block with the gallery
    ares = (e != nullptr && atoi(e) == 0) ? 0 : 1;
  }
  if (L.k_chunks <= kAResChunks && ares) rc = launch_gemm<EpiRank, BK, true>(L, EP, stream);
  else rc = launch_gemm<EpiRank>(L, EP, stream);
  if (rc != LECCR_OK) return rc;
  {
    ResolveParams rp[2];
    memset(rp, 0, sizeof(rp));
    for (int p = 0; p < n_prob; ++p) {
      const leccr_topk_problem& q = probs[p];
      rp[p].pairs = EP.amb_list[p];
      rp[p].n_pairs = EP.amb_count[p];
      rp[p].cap = EP.amb_cap[p];
      rp[p].rows_x = q.rows_x;
      rp[p].cols_x = q.cols_x;
      rp[p].ld_rows = q.ld_rows_x;
      rp[p].ld_cols = q.ld_cols_x;
      rp[p].D = D;
      rp[p].x_dtype = q.x_dtype;
      rp[p].best = best[p];
      rp[p].rank = q.rank;
    }
    dim3 gr(static_cast<unsigned>(4 * num_sms()), static_cast<unsigned>(n_prob));
    rank_resolve_kernel<<<gr, 256, 0, stream>>>(rp[0], rp[1]);
    LAUNCH_CHECK("rank_resolve_kernel");
  }
  if (n_prob == 1) memset(&post[1], 0, sizeof(post[1]));
  dim3 g(static_cast<unsigned>(num_sms()), static_cast<unsigned>(n_prob));
  rank_post_kernel<<<g, 256, 0, stream>>>(post[0], post[1], probs[0].recall_counts,
                                          n_prob > 1 ? probs[1].recall_counts : nullptr);
  LAUNCH_CHECK("rank_post_kernel");
  return LECCR_OK;
}

// ------------------------------------------------------------------------------------ double_sim, fused
// Two tensor-core passes over the interleaved video/caption operand (EpiDsStats, EpiDsTopK); no N x M buffer.
// Problem order of the plans: [0] orientation A (rows VC, columns texts), [1] orientation B.
static int ds_plan(int64_t n_vid, int64_t n_txt, int G, Plan* plans) {
  leccr_topk_problem pr[2];
  memset(pr, 0, sizeof(pr));
  pr[0].n_rows = G * n_vid;
  pr[0].n_cols = n_txt;
  pr[1].n_rows = n_txt;
  pr[1].n_cols = G * n_vid;
  return topk_plan(pr, 2, 0, plans);
}

size_t leccr_double_sim_topk_workspace(int64_t n_vid, int64_t n_txt, int G) {
  if (n_vid <= 0 || n_txt <= 0 || (G != 2 && G != 4 && G != 8)) return 0;
  Plan plans[2];
  if (ds_plan(n_vid, n_txt, G, plans) != LECCR_OK) return 0;
  return topk_ws_bytes(G * n_vid, plans[0].n_chunks, true) + topk_ws_bytes(n_txt, plans[1].n_chunks, true) +
         align256(static_cast<size_t>(n_txt) * 4) * 3 + align256(static_cast<size_t>(n_vid) * 4) + 256;
}

extern "C++" {
template <int G>
static int ds_launch(const void* vc16, const void* t16, int64_t n_vid, int64_t n_txt, int K, int fmt, int n_cap,
                     float w1, float w2, int mode, const int32_t* txt_gt, const int32_t* vid_gt_off,
                     const int32_t* vid_gt_ids, int k, float* topk_val_vc, int32_t* topk_idx_vc, float* topk_val_txt,
                     int32_t* topk_idx_txt, int32_t* rank_vid, int32_t* rank_txt, int32_t* recall_counts,
                     void* workspace, cudaStream_t stream) {
  Plan plans[2];
  int rc = ds_plan(n_vid, n_txt, G, plans);
  if (rc != LECCR_OK) return rc;
  const int64_t n_vc = G * n_vid;
  const bool ranks = txt_gt != nullptr;
  const bool lists_a = topk_val_vc != nullptr && topk_idx_vc != nullptr;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint8_t* ws_list[2] = {ws, ws + topk_ws_bytes(n_vc, plans[0].n_chunks, true)};
  ws = ws_list[1] + topk_ws_bytes(n_txt, plans[1].n_chunks, true);
  float* gt_s = reinterpret_cast<float*>(ws);
  ws += align256(static_cast<size_t>(n_txt) * 4);
  float* gt_c = reinterpret_cast<float*>(ws);
  ws += align256(static_cast<size_t>(n_txt) * 4);
  float* txt_best = reinterpret_cast<float*>(ws);
  ws += align256(static_cast<size_t>(n_txt) * 4);
  float* vid_best = reinterpret_cast<float*>(ws);
  ws += align256(static_cast<size_t>(n_vid) * 4);
  unsigned* mm = reinterpret_cast<unsigned*>(ws);
  mm_init_kernel<<<1, 32, 0, stream>>>(mm);
  LAUNCH_CHECK("mm_init_kernel");
  // ---- pass 1 (orientation B): min / max of S and max_n C_n, (S, max C) at every text's ground-truth video
  if (mode == LECCR_FUSE_NORM || ranks) {
    SimLaunch L;
    memset(&L, 0, sizeof(L));
    L.n_prob = 1;
    L.fmt = fmt;
    L.k_chunks = (K + BK - 1) / BK;
    const int64_t tiles = ((n_txt + BM - 1) / BM) * ((n_vc + BN - 1) / BN);
    const Plan pl = plan_problem(n_vc, 0, n_txt, auto_tiles_per_chunk(tiles, 1));
    rc = fill_problem(L.prob[0], t16, K, vc16, K, n_txt, n_vc, K, fmt, pl, 0);
    if (rc != LECCR_OK) return rc;
    L.n_items = pl.row_blocks * pl.n_chunks;
    typename EpiDsStats<G>::Params EP;
    memset(&EP, 0, sizeof(EP));
    EP.n_cap = n_cap;
    EP.n_groups = static_cast<int>(n_vid);
    EP.gt_group = txt_gt;
    EP.gt_s = gt_s;
    EP.gt_c = gt_c;
    EP.mm = mm;
    rc = launch_gemm<EpiDsStats<G>>(L, EP, stream);
    if (rc != LECCR_OK) return rc;
  }
  if (ranks) {
    const unsigned g = static_cast<unsigned>((n_txt + n_vid + 255) / 256);
    ds_gt_scores_kernel<<<g, 256, 0, stream>>>(gt_s, gt_c, (int)n_txt, vid_gt_off, vid_gt_ids, (int)n_vid, mm, w1, w2, mode,
                                               txt_best, vid_best);
    LAUNCH_CHECK("ds_gt_scores_kernel");
    CUDA_TRY(cudaMemsetAsync(rank_vid, 0, static_cast<size_t>(n_vid) * 4, stream));
    CUDA_TRY(cudaMemsetAsync(rank_txt, 0, static_cast<size_t>(n_txt) * 4, stream));
    CUDA_TRY(cudaMemsetAsync(recall_counts, 0, 6 * 4, stream));
  }
  // ---- pass 2: orientation B (ranks of both directions, per-text lists) and, for per-video lists, orientation A
  SimLaunch L;
  memset(&L, 0, sizeof(L));
  L.fmt = fmt;
  L.k_chunks = (K + kBK2 - 1) / kBK2;
  typename EpiDsTopK<G>::Params EP;
  memset(&EP, 0, sizeof(EP));
  EP.dense = 1;
  EP.trig = TopK2D::TRIG;
  const void* rows16[2] = {vc16, t16};
  const void* cols16[2] = {t16, vc16};
  const int64_t nr[2] = {n_vc, n_txt}, nc[2] = {n_txt, n_vc};
  float* cand_val[2] = {nullptr, nullptr};
  int* cand_idx[2] = {nullptr, nullptr};
  int* cand_cnt[2] = {nullptr, nullptr};
  int item_base = 0, g = 0;
  for (int p = lists_a ? 0 : 1; p < 2; ++p, ++g) {
    uint8_t* w = ws_list[p];
    const size_t lists = static_cast<size_t>(nr[p]) * plans[p].n_chunks * 4;
    cand_val[p] = reinterpret_cast<float*>(w);
    w += align256(lists * 32 * 4);
    cand_idx[p] = reinterpret_cast<int*>(w);
    w += align256(lists * 32 * 4);
    cand_cnt[p] = reinterpret_cast<int*>(w);
    w += align256(lists * 4);
    w += align256(static_cast<size_t>(nr[p]) * 4 + 16);  // (flag list of the exact-rank fallback: unused here)
    unsigned* row_thr = reinterpret_cast<unsigned*>(w);
    CUDA_TRY(cudaMemsetAsync(row_thr, 0, static_cast<size_t>(nr[p]) * 12, stream));
    rc = fill_problem(L.prob[g], rows16[p], K, cols16[p], K, nr[p], nc[p], K, fmt, plans[p], item_base, kBK2);
    if (rc != LECCR_OK) return rc;
    item_base += plans[p].row_blocks * plans[p].n_chunks;
    EP.row_thr[g] = row_thr;
    EP.row_h8[g] = row_thr + nr[p];
    EP.out_val[g] = cand_val[p];
    EP.out_idx[g] = cand_idx[p];
    EP.out_cnt[g] = cand_cnt[p];
    EP.n_sub[g] = plans[p].n_chunks * 2;
  }
  L.n_prob = g;
  L.n_items = item_base;
  EP.b_problem = g - 1;
  EP.n_cap = n_cap;
  EP.n_groups = static_cast<int>(n_vid);
  EP.mm = mm;
  EP.w1 = w1;
  EP.w2 = w2;
  EP.mode = mode;
  EP.txt_best = ranks ? txt_best : nullptr;
  EP.vid_best = ranks ? vid_best : nullptr;
  EP.rank_txt = ranks ? rank_txt : nullptr;
  EP.rank_vid = ranks ? rank_vid : nullptr;
  rc = launch_gemm<EpiDsTopK<G>, kBK2>(L, EP, stream);
  if (rc != LECCR_OK) return rc;
  // ---- top-k lists out of the candidate lists (no exact re-scoring: the scores ARE the fp32-faithful fused ones)
  float* tv[2] = {topk_val_vc, topk_val_txt};
  int32_t* ti[2] = {topk_idx_vc, topk_idx_txt};
  for (int p = 0; p < 2; ++p) {
    if (tv[p] == nullptr || ti[p] == nullptr || cand_val[p] == nullptr) continue;
    TopkFinalizeParams F;
    memset(&F, 0, sizeof(F));
    F.cand_val = cand_val[p];
    F.cand_idx = cand_idx[p];
    F.cand_cnt = cand_cnt[p];
    F.n_rows = static_cast<int>(nr[p]);
    F.n_cols = static_cast<int>(p == 0 ? n_txt : n_vid);
    F.n_chunks = plans[p].n_chunks * 2;
    F.list_cap = TopK2D::C;
    F.KP = LECCR_TOPK_KP;
    F.k = k;
    F.topk_val = tv[p];
    F.topk_idx = ti[p];
    const unsigned grid = static_cast<unsigned>((F.n_rows + kFinalizeWarps - 1) / kFinalizeWarps);
    const int slots = F.n_chunks;
    if (slots <= 2) topk_finalize_kernel<2><<<grid, kFinalizeWarps * 32, 0, stream>>>(F);
    else if (slots <= 4) topk_finalize_kernel<4><<<grid, kFinalizeWarps * 32, 0, stream>>>(F);
    else if (slots <= 8) topk_finalize_kernel<8><<<grid, kFinalizeWarps * 32, 0, stream>>>(F);
    else topk_finalize_kernel<kMaxSlots><<<grid, kFinalizeWarps * 32, 0, stream>>>(F);
    LAUNCH_CHECK("topk_finalize_kernel");
  }
  if (ranks) {
    const unsigned g0 = static_cast<unsigned>(std::min<int64_t>((n_vid + 255) / 256, 2LL * num_sms()));
    recall_count_kernel<<<g0, 256, 0, stream>>>(rank_vid, (int)n_vid, recall_counts);
    LAUNCH_CHECK("recall_count_kernel");
    const unsigned g1 = static_cast<unsigned>(std::min<int64_t>((n_txt + 255) / 256, 2LL * num_sms()));
    recall_count_kernel<<<g1, 256, 0, stream>>>(rank_txt, (int)n_txt, recall_counts + 3);
    LAUNCH_CHECK("recall_count_kernel");
  }
  return LECCR_OK;
}
}  // extern "C++"

int leccr_double_sim_topk(const void* vc16, const void* t16, int64_t n_vid, int64_t n_txt, int K, int fmt, int G,
                          int n_cap, float w1, float w2, int mode, const int32_t* txt_gt, const int32_t* vid_gt_off,
                          const int32_t* vid_gt_ids, int k, float* topk_val_vc, int32_t* topk_idx_vc,
                          float* topk_val_txt, int32_t* topk_idx_txt, int32_t* rank_vid, int32_t* rank_txt,
                          int32_t* recall_counts, void* workspace, size_t workspace_bytes, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (vc16 == nullptr || t16 == nullptr || n_vid <= 0 || n_txt <= 0 || K <= 0 || (K & 7) != 0 || bad_fmt(fmt) ||
      (G != 2 && G != 4 && G != 8) || n_cap < 1 || n_cap >= G || (mode != LECCR_FUSE_NORM && mode != LECCR_FUSE_RAW) ||
      k < 1 || k > LECCR_TOPK_KP)
    return LECCR_ERR_ARG;
  if (txt_gt != nullptr && (vid_gt_off == nullptr || vid_gt_ids == nullptr || rank_vid == nullptr || rank_txt == nullptr ||
                            recall_counts == nullptr))
    return LECCR_ERR_ARG;
  int rc = leccr_check_device();
  if (rc != LECCR_OK) return rc;
  const size_t need = leccr_double_sim_topk_workspace(n_vid, n_txt, G);
  if (need == 0) return LECCR_ERR_ARG;
  if (workspace == nullptr || workspace_bytes < need) return LECCR_ERR_WORKSPACE;
#define LECCR_DS(GG)                                                                                                  \
  ds_launch<GG>(vc16, t16, n_vid, n_txt, K, fmt, n_cap, w1, w2, mode, txt_gt, vid_gt_off, vid_gt_ids, k, topk_val_vc, \
                topk_idx_vc, topk_val_txt, topk_idx_txt, rank_vid, rank_txt, recall_counts, workspace, stream)
  if (G == 2) return LECCR_DS(2);
  if (G == 4) return LECCR_DS(4);
  return LECCR_DS(8);
#undef LECCR_DS
}

// ------------------------------------------------------------------------------------ InfoNCE
static Plan infonce_plan(int64_t n, int tiles_per_chunk) {
  const int64_t tiles = 2 * ((n + BM - 1) / BM) * ((n + BN - 1) / BN);
  const int tpc = tiles_per_chunk > 0 ? tiles_per_chunk : auto_tiles_per_chunk(tiles, 4);
  return plan_problem(n, 0, n, tpc);
}

size_t leccr_infonce_fwd_workspace(int64_t n, int tiles_per_chunk) {
  if (n <= 0) return 0;
  const Plan pl = infonce_plan(n, tiles_per_chunk);
  return 2 * align256(static_cast<size_t>(n) * pl.n_chunks * EpiLse::kWGs * 5 * 4) + 256;
}

// scratch (8 doubles of block sums + ticket) sits behind the two partial planes of the workspace
static double* infonce_fwd_scratch(int64_t n, int tiles_per_chunk, void* workspace) {
  const Plan pl = infonce_plan(n, tiles_per_chunk);
  const size_t part_bytes = align256(static_cast<size_t>(n) * pl.n_chunks * EpiLse::kWGs * 5 * 4);
  return reinterpret_cast<double*>(static_cast<uint8_t*>(workspace) + 2 * part_bytes);
}

static int infonce_fwd_impl(const void* a16, const void* b16, int64_t ld16, const int64_t* idx, int64_t n, int D,
                            int fmt, const float* temp, float* out, float* lse2, float* rcnt, int tiles_per_chunk,
                            void* workspace, size_t workspace_bytes, bool scratch_is_zero, leccr_stream_t stream_);

int leccr_infonce_fwd(const void* a16, const void* b16, int64_t ld16, const int64_t* idx, int64_t n, int D,
                      int fmt, const float* temp, float* out, float* lse2, float* rcnt, int tiles_per_chunk,
                      void* workspace, size_t workspace_bytes, leccr_stream_t stream_) {
  return infonce_fwd_impl(a16, b16, ld16, idx, n, D, fmt, temp, out, lse2, rcnt, tiles_per_chunk, workspace,
                          workspace_bytes, false, stream_);
}

static int infonce_fwd_impl(const void* a16, const void* b16, int64_t ld16, const int64_t* idx, int64_t n, int D,
                            int fmt, const float* temp, float* out, float* lse2, float* rcnt, int tiles_per_chunk,
                            void* workspace, size_t workspace_bytes, bool scratch_is_zero, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (a16 == nullptr || b16 == nullptr || temp == nullptr || out == nullptr || lse2 == nullptr || rcnt == nullptr ||
      n <= 0 || D <= 0 || bad_fmt(fmt))
    return LECCR_ERR_ARG;
  int rc = leccr_check_device();
  if (rc != LECCR_OK) return rc;
  if (workspace == nullptr || workspace_bytes < leccr_infonce_fwd_workspace(n, tiles_per_chunk))
    return LECCR_ERR_WORKSPACE;
  const Plan pl = infonce_plan(n, tiles_per_chunk);
  SimLaunch L;
  memset(&L, 0, sizeof(L));
  L.n_prob = 2;
  L.fmt = fmt;
  L.k_chunks = (D + BK - 1) / BK;
  const int items = pl.row_blocks * pl.n_chunks;
  rc = fill_problem(L.prob[0], a16, ld16, b16, ld16, n, n, D, fmt, pl, 0);
  if (rc != LECCR_OK) return rc;
  rc = fill_problem(L.prob[1], b16, ld16, a16, ld16, n, n, D, fmt, pl, items);
  if (rc != LECCR_OK) return rc;
  L.n_items = 2 * items;
  const int n_sub = pl.n_chunks * EpiLse::kWGs;
  const size_t part_bytes = align256(static_cast<size_t>(n) * n_sub * 5 * 4);
  float* part0 = static_cast<float*>(workspace);
  float* part1 = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + part_bytes);
  double* scratch = reinterpret_cast<double*>(static_cast<uint8_t*>(workspace) + 2 * part_bytes);
  if (!scratch_is_zero) CUDA_TRY(cudaMemsetAsync(scratch, 0, 64, stream));
  EpiLse::Params EP;
  memset(&EP, 0, sizeof(EP));
  EP.temp = temp;
  for (int p = 0; p < 2; ++p) {
    EP.idx_rows[p] = reinterpret_cast<const long long*>(idx);
    EP.idx_cols[p] = reinterpret_cast<const long long*>(idx);
    EP.n_sub[p] = n_sub;
  }
  EP.part[0] = part0;
  EP.part[1] = part1;
  rc = launch_gemm<EpiLse>(L, EP, stream);
  if (rc != LECCR_OK) return rc;
  FinalizeLseParams F;
  memset(&F, 0, sizeof(F));
  F.part[0] = part0;
  F.part[1] = part1;
  F.n[0] = F.n[1] = static_cast<int>(n);
  F.nch[0] = F.nch[1] = n_sub;
  F.scratch = scratch;
  F.lse2[0] = lse2;
  F.lse2[1] = lse2 + n;
  F.rcnt[0] = rcnt;
  F.rcnt[1] = rcnt + n;
  F.temp = temp;
  F.out = out;
  infonce_finalize_kernel<<<static_cast<unsigned>((2 * n + 255) / 256), 256, 0, stream>>>(F);
  LAUNCH_CHECK("infonce_finalize_kernel");
  return LECCR_OK;
}

// Strip forward of a distributed step: tensor-core pass over this rank's rows only, statistics exchanged
// through the peer-mapped slots (see infonce_finalize_local_kernel).  Workspace as leccr_infonce_fwd_workspace.
static Plan infonce_strip_plan(int64_t n, int64_t row_begin, int64_t row_count) {
  const int64_t col_tiles = (n + BN - 1) / BN;
  const int64_t row_blocks = (row_begin + row_count + BM - 1) / BM - row_begin / BM;
  // one launch carries both orientations: spread 2 * row_blocks * chunks items over the SMs
  int64_t chunks = std::max<int64_t>(1, num_sms() / std::max<int64_t>(1, 2 * row_blocks));
  chunks = std::min<int64_t>(chunks, col_tiles);
  const int tpc = static_cast<int>((col_tiles + chunks - 1) / chunks);
  return plan_problem(n, row_begin, row_count, tpc);
}
// row_blocks * chunks <= max(num_sms / 2, row_blocks of the whole matrix) for every strip of an n-row problem
static size_t strip_part_bytes(int64_t n) {
  const int64_t rb_chunks = std::max<int64_t>(num_sms() / 2 + 1, (n + BM - 1) / BM + 1);
  return align256(static_cast<size_t>(rb_chunks) * BM * EpiLse::kWGs * 5 * 4);
}
static double* strip_scratch(int64_t n, void* workspace) {
  return reinterpret_cast<double*>(static_cast<uint8_t*>(workspace) + 2 * strip_part_bytes(n));
}

static int infonce_fwd_strips(const void* a16, const void* b16, int64_t ld16, const int64_t* idx, int64_t n, int D,
                              int fmt, const float* temp, float* out, float* lse2, float* rcnt, int64_t row_begin,
                              int64_t row_count, float* const* stat_ptrs_dev, const float* local_stat, int world,
                              int rank, uint32_t* const* flag_ptrs_dev, uint32_t epoch, void* workspace,
                              size_t workspace_bytes, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (workspace == nullptr || workspace_bytes < 2 * strip_part_bytes(n) + 256) return LECCR_ERR_WORKSPACE;
  const Plan pl = infonce_strip_plan(n, row_begin, row_count);
  SimLaunch L;
  memset(&L, 0, sizeof(L));
  L.n_prob = 2;
  L.fmt = fmt;
  L.k_chunks = (D + BK - 1) / BK;
  const int items = pl.row_blocks * pl.n_chunks;
  int rc = fill_problem(L.prob[0], a16, ld16, b16, ld16, n, n, D, fmt, pl, 0);
  if (rc != LECCR_OK) return rc;
  rc = fill_problem(L.prob[1], b16, ld16, a16, ld16, n, n, D, fmt, pl, items);
  if (rc != LECCR_OK) return rc;
  L.n_items = 2 * items;
  const int n_sub = pl.n_chunks * EpiLse::kWGs;
  // partial planes hold the strip's row blocks only; the kernels address rows absolutely, so the plane
  // pointers are biased by the strip's first row
  const size_t part_bytes = strip_part_bytes(n);
  const long long bias = static_cast<long long>(pl.row_block_begin) * BM * n_sub * 5;
  float* part0 = static_cast<float*>(workspace) - bias;
  float* part1 = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + part_bytes) - bias;
  double* scratch = strip_scratch(n, workspace);
  EpiLse::Params EP;
  memset(&EP, 0, sizeof(EP));
  EP.temp = temp;
  for (int p = 0; p < 2; ++p) {
    EP.idx_rows[p] = reinterpret_cast<const long long*>(idx);
    EP.idx_cols[p] = reinterpret_cast<const long long*>(idx);
    EP.n_sub[p] = n_sub;
  }
  EP.part[0] = part0;
  EP.part[1] = part1;
  rc = launch_gemm<EpiLse>(L, EP, stream);
  if (rc != LECCR_OK) return rc;
  FinalizeLocalParams F;
  memset(&F, 0, sizeof(F));
  F.part[0] = part0;
  F.part[1] = part1;
  F.nch[0] = F.nch[1] = n_sub;
  F.n = static_cast<int>(n);
  F.row_begin = static_cast<int>(row_begin);
  F.row_count = static_cast<int>(row_count);
  F.stat_ptrs = stat_ptrs_dev;
  F.world = world;
  F.rank = rank;
  F.scratch = scratch;
  infonce_finalize_local_kernel<<<static_cast<unsigned>((2 * row_count + 255) / 256), 256, 0, stream>>>(F);
  LAUNCH_CHECK("infonce_finalize_local_kernel");
  rc = leccr_peer_barrier(flag_ptrs_dev, world, rank, epoch, stream_);
  if (rc != LECCR_OK) return rc;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(32, (2 * n + 255) / 256));
  infonce_reduce_kernel<<<blocks, 256, 0, stream>>>(local_stat, static_cast<int>(n), world, temp, out, lse2, rcnt);
  LAUNCH_CHECK("infonce_reduce_kernel");
  return LECCR_OK;
}

static int64_t round_up8(int64_t x) { return (x + 7) & ~static_cast<int64_t>(7); }

// The gradient products read the gathered operands as MN-major column operands (no transposed copies);
// LECCR_BWD_MN=0 keeps the K-major products over transposed copies (development comparison).
static bool bwd_mn_major() {
  static const bool on = [] {
    const char* e = getenv("LECCR_BWD_MN");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}

// split-K factor of the gradient products: fill the machine, at least 2 K-chunks per split
static int bwd_splits(int k_chunks, int row_blocks) {
  const int want = std::max(1, num_sms() / std::max(1, 2 * row_blocks));
  return std::max(1, std::min(want, (k_chunks + 1) / 2));
}

size_t leccr_infonce_bwd_workspace(int64_t n, int64_t row_count, int D) {
  if (n <= 0 || row_count <= 0 || D <= 0) return 0;
  const int k_chunks = static_cast<int>((n + BK - 1) / BK);
  const int row_blocks = static_cast<int>((row_count + BM - 1) / BM);
  const size_t parts = 2 * static_cast<size_t>(bwd_splits(k_chunks, row_blocks)) * row_count * D * 4;
  return 2 * align256(static_cast<size_t>(row_count) * round_up8(n) * 2) + align256(parts);
}

static int infonce_bwd_impl(const void* a16, const void* b16, int64_t ld16, const void* aT16, const void* bT16,
                            int64_t ldT, const int64_t* idx, int64_t n, int D, int fmt, const float* temp,
                            const float* lse2, const float* rcnt, int64_t row_begin, int64_t row_count,
                            const float* grad_out, float* dA, float* dB, void* workspace, size_t workspace_bytes,
                            const float* prod_b, float* prod_out, int one_dir, leccr_stream_t stream_);

int leccr_infonce_bwd(const void* a16, const void* b16, int64_t ld16, const void* aT16, const void* bT16,
                      int64_t ldT, const int64_t* idx, int64_t n, int D, int fmt, const float* temp,
                      const float* lse2, const float* rcnt, int64_t row_begin, int64_t row_count,
                      const float* grad_out, float* dA, float* dB, void* workspace, size_t workspace_bytes,
                      leccr_stream_t stream_) {
  return infonce_bwd_impl(a16, b16, ld16, aT16, bT16, ldT, idx, n, D, fmt, temp, lse2, rcnt, row_begin, row_count,
                          grad_out, dA, dB, workspace, workspace_bytes, nullptr, nullptr, 0, stream_);
}

// prod_out (optional) = grad_out * *prod_b, computed by the last launch.  one_dir: gradient of the i2t half only,
// loss = -mean_i sum_j log_softmax(a b^T / temp, 1)_ij labels_ij  (models/model_retrieval_caption.py:141).
static int infonce_bwd_impl(const void* a16, const void* b16, int64_t ld16, const void* aT16, const void* bT16,
                            int64_t ldT, const int64_t* idx, int64_t n, int D, int fmt, const float* temp,
                            const float* lse2, const float* rcnt, int64_t row_begin, int64_t row_count,
                            const float* grad_out, float* dA, float* dB, void* workspace, size_t workspace_bytes,
                            const float* prod_b, float* prod_out, int one_dir, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool mn = bwd_mn_major();
  if (a16 == nullptr || b16 == nullptr || (!mn && (aT16 == nullptr || bT16 == nullptr)) || temp == nullptr ||
      lse2 == nullptr || rcnt == nullptr || dA == nullptr || dB == nullptr || n <= 0 || D <= 0 || bad_fmt(fmt) ||
      row_begin < 0 || row_count <= 0 || row_begin + row_count > n)
    return LECCR_ERR_ARG;
  int rc = leccr_check_device();
  if (rc != LECCR_OK) return rc;
  if (workspace == nullptr || workspace_bytes < leccr_infonce_bwd_workspace(n, row_count, D))
    return LECCR_ERR_WORKSPACE;
  const int64_t ldS = round_up8(n);
  const size_t strip_bytes = align256(static_cast<size_t>(row_count) * ldS * 2);
  void* strip0 = workspace;
  void* strip1 = static_cast<uint8_t*>(workspace) + strip_bytes;

  // 1. recompute the logits of the local row strips on the tensor cores, emit G' (16-bit)
  {
    const int64_t tiles = 2 * ((row_count + BM - 1) / BM + 1) * ((n + BN - 1) / BN);
    const Plan pl = plan_problem(n, row_begin, row_count, auto_tiles_per_chunk(tiles, 2));
    SimLaunch L;
    memset(&L, 0, sizeof(L));
    L.n_prob = 2;
    L.fmt = fmt;
    L.k_chunks = (D + BK - 1) / BK;
    const int items = pl.row_blocks * pl.n_chunks;
    rc = fill_problem(L.prob[0], a16, ld16, b16, ld16, n, n, D, fmt, pl, 0);
    if (rc != LECCR_OK) return rc;
    rc = fill_problem(L.prob[1], b16, ld16, a16, ld16, n, n, D, fmt, pl, items);
    if (rc != LECCR_OK) return rc;
    L.n_items = 2 * items;
    EpiGrad::Params EP;
    memset(&EP, 0, sizeof(EP));
    EP.temp = temp;
    EP.fmt = fmt;
    for (int p = 0; p < 2; ++p) {
      EP.idx_rows[p] = reinterpret_cast<const long long*>(idx);
      EP.idx_cols[p] = reinterpret_cast<const long long*>(idx);
      EP.lse_rows[p] = lse2 + p * n;
      EP.lse_cols[p] = lse2 + (1 - p) * n;
      EP.rcnt_rows[p] = rcnt + p * n;
      EP.rcnt_cols[p] = rcnt + (1 - p) * n;
      EP.ld[p] = ldS;
      EP.row0[p] = static_cast<int>(row_begin);
      EP.nrow[p] = static_cast<int>(row_count);
      EP.row_w[p] = (one_dir && p == 1) ? 0.f : 1.f;
      EP.col_w[p] = (one_dir && p == 0) ? 0.f : 1.f;
    }
    EP.strip[0] = strip0;
    EP.strip[1] = strip1;
    rc = launch_gemm<EpiGrad>(L, EP, stream);
    if (rc != LECCR_OK) return rc;
  }
  // 2. dA_loc = G'[loc,:] B / (2 n temp),  dB_loc = G'[:,loc]^T A / (2 n temp): split-K products
  {
    const int k_chunks = static_cast<int>((n + BK - 1) / BK);
    const int row_blocks = static_cast<int>((row_count + BM - 1) / BM);
    const int splits = bwd_splits(k_chunks, row_blocks);
    float* parts0 = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + 2 * strip_bytes);
    float* parts1 = parts0 + static_cast<size_t>(splits) * row_count * D;
    StoreProblem sp[2] = {
        {strip0, mn ? b16 : bT16, ldS, mn ? ld16 : ldT, row_count, D, dA, D, parts0, mn},
        {strip1, mn ? a16 : aT16, ldS, mn ? ld16 : ldT, row_count, D, dB, D, parts1, mn},
    };
    rc = launch_store(sp, 2, static_cast<int>(n), fmt, 1.0f / ((one_dir ? 1.0f : 2.0f) * static_cast<float>(n)), grad_out, temp,
                      splits, stream, grad_out, prod_b, prod_out);
    if (rc != LECCR_OK) return rc;
  }
  return LECCR_OK;
}

// ------------------------------------------------------------------------------------ ranking
int leccr_rank_rows(const float* S, int64_t ld, int64_t R, int64_t C, const int32_t* gt_off,
                    const int32_t* gt_ids, int32_t* rank, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (S == nullptr || gt_off == nullptr || gt_ids == nullptr || rank == nullptr || R <= 0 || C <= 0 || ld < C)
    return LECCR_ERR_ARG;
  rank_rows_kernel<<<static_cast<unsigned>(R), 256, 0, stream>>>(S, ld, (int)R, (int)C, gt_off, gt_ids, rank);
  LAUNCH_CHECK("rank_rows_kernel");
  return LECCR_OK;
}

int leccr_rank_cols(const float* S, int64_t ld, int64_t R, int64_t C, const int32_t* gt_off,
                    const int32_t* gt_ids, int32_t* scratch_nnz, int32_t* rank, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (S == nullptr || gt_off == nullptr || gt_ids == nullptr || rank == nullptr || scratch_nnz == nullptr ||
      R <= 0 || C <= 0 || ld < C)
    return LECCR_ERR_ARG;
  const unsigned gx = static_cast<unsigned>((C + 127) / 128);
  // enough row slabs to fill the machine
  int slabs = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((R + 63) / 64, (4LL * num_sms() + gx - 1) / gx)));
  const int rows_per_block = static_cast<int>((R + slabs - 1) / slabs);
  slabs = static_cast<int>((R + rows_per_block - 1) / rows_per_block);
  rank_cols_count_kernel<<<dim3(gx, slabs), 128, 0, stream>>>(S, ld, (int)R, (int)C, gt_off, gt_ids,
                                                             rows_per_block, scratch_nnz);
  LAUNCH_CHECK("rank_cols_count_kernel");
  rank_cols_min_kernel<<<static_cast<unsigned>((C + 255) / 256), 256, 0, stream>>>((int)C, (int)R, gt_off,
                                                                                    scratch_nnz, rank);
  LAUNCH_CHECK("rank_cols_min_kernel");
  return LECCR_OK;
}

int leccr_recall_counts(const int32_t* rank, int64_t n, int32_t* counts, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rank == nullptr || counts == nullptr || n <= 0) return LECCR_ERR_ARG;
  const unsigned g = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 2LL * num_sms()));
  recall_count_kernel<<<g, 256, 0, stream>>>(rank, (int)n, counts);
  LAUNCH_CHECK("recall_count_kernel");
  return LECCR_OK;
}

// ------------------------------------------------------------------------------------ double_sim
int leccr_double_sim_fuse(float* S, const float* Cn, int n_cap, int64_t numel, float* Cmax, uint32_t* mm,
                          float w1, float w2, int mode, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (S == nullptr || Cn == nullptr || Cmax == nullptr || mm == nullptr || n_cap < 1 || numel <= 0 ||
      (mode != LECCR_FUSE_NORM && mode != LECCR_FUSE_RAW))
    return LECCR_ERR_ARG;
  const unsigned g = static_cast<unsigned>(std::min<int64_t>((numel + 255) / 256, 8LL * num_sms()));
  mm_init_kernel<<<1, 32, 0, stream>>>(mm);
  LAUNCH_CHECK("mm_init_kernel");
  capmax_minmax_kernel<<<g, 256, 0, stream>>>(S, Cn, Cmax, n_cap, numel, mm);
  LAUNCH_CHECK("capmax_minmax_kernel");
  fuse_scores_kernel<<<g, 256, 0, stream>>>(S, Cmax, numel, mm, w1, w2, mode);
  LAUNCH_CHECK("fuse_scores_kernel");
  return LECCR_OK;
}

int leccr_peer_barrier(uint32_t* const* flag_ptrs_dev, int world, int rank, uint32_t epoch, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (flag_ptrs_dev == nullptr || world < 1 || world > 32 || rank < 0 || rank >= world) return LECCR_ERR_ARG;
  // LECCR_PEER_TIMEOUT_S: seconds a rank waits for its peers before it gives up (default 600; 0 = for ever)
  static long long timeout_s = -1;
  if (timeout_s < 0) {
    const char* e = getenv("LECCR_PEER_TIMEOUT_S");
    timeout_s = (e != nullptr && atoll(e) >= 0) ? atoll(e) : 600;
  }
  peer_barrier_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<unsigned* const*>(flag_ptrs_dev), world, rank, epoch,
                                            static_cast<unsigned long long>(timeout_s) * 1000000000ull);
  LAUNCH_CHECK("peer_barrier_kernel");
  return LECCR_OK;
}

int leccr_normalize_fwd(const float* x, int64_t n, int D, int64_t ld_x, float* y, int64_t ld_y, float* inv,
                        void* y16, int64_t ld_y16, int fmt, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (x == nullptr || y == nullptr || n <= 0 || D <= 0 || ld_x < D || ld_y < D || (y16 != nullptr && (bad_fmt(fmt) || ld_y16 < D)))
    return LECCR_ERR_ARG;
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((n + wpb - 1) / wpb);
  if (fmt == LECCR_FMT_BF16)
    normalize_rows_kernel<1><<<grid, wpb * 32, 0, stream>>>(x, ld_x, (int)n, D, y, ld_y, inv, static_cast<uint16_t*>(y16), ld_y16);
  else
    normalize_rows_kernel<0><<<grid, wpb * 32, 0, stream>>>(x, ld_x, (int)n, D, y, ld_y, inv, static_cast<uint16_t*>(y16), ld_y16);
  LAUNCH_CHECK("normalize_rows_kernel");
  return LECCR_OK;
}

int leccr_normalize_bwd(const float* y, int64_t ld_y, const float* inv, const float* g, int64_t ld_g, int64_t n, int D,
                        float* dx, int64_t ld_dx, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (y == nullptr || inv == nullptr || g == nullptr || dx == nullptr || n <= 0 || D <= 0 || ld_y < D || ld_g < D || ld_dx < D)
    return LECCR_ERR_ARG;
  const int wpb = 8;
  normalize_rows_bwd_kernel<<<static_cast<unsigned>((n + wpb - 1) / wpb), wpb * 32, 0, stream>>>(y, ld_y, inv, g, ld_g, (int)n, D,
                                                                                              dx, ld_dx);
  LAUNCH_CHECK("normalize_rows_bwd_kernel");
  return LECCR_OK;
}

int leccr_memcpy_peer_async(void* dst, const void* src, size_t bytes, leccr_stream_t stream_) {
  if (dst == nullptr || src == nullptr) return LECCR_ERR_ARG;
  if (bytes == 0) return LECCR_OK;
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream_)));
  return LECCR_OK;
}

int leccr_topk_merge_peers(const float* const* val_ptrs_dev, const int32_t* const* idx_ptrs_dev, int world, int k_in,
                           int64_t q_begin, int64_t q_count, const int64_t* col_offset_host, int k_out,
                           float* out_val, int32_t* out_idx, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (val_ptrs_dev == nullptr || idx_ptrs_dev == nullptr || world < 1 || world > kMergeMaxWorld || k_in < 1 ||
      k_in > 64 || k_out < 1 || k_out > world * k_in || q_begin < 0 || q_count < 0 || out_val == nullptr ||
      out_idx == nullptr)
    return LECCR_ERR_ARG;
  if (q_count == 0) return LECCR_OK;
  MergeOffsets offs;
  for (int p = 0; p < kMergeMaxWorld; ++p) {
    const int64_t o = (col_offset_host != nullptr && p < world) ? col_offset_host[p] : 0;
    if (o < 0 || o > 0x7fffffffLL) return LECCR_ERR_ARG;
    offs.off[p] = static_cast<int>(o);
  }
  const int warps = 2;
  const size_t smem = static_cast<size_t>(warps) * world * 32 * k_in * 8;
  if (smem > 200 * 1024) return LECCR_ERR_ARG;
  // function attributes are per device: set it on every call that needs it (cheap) rather than once per process
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(topk_merge_peers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const unsigned grid = static_cast<unsigned>((q_count + warps * 32 - 1) / (warps * 32));
  topk_merge_peers_kernel<<<grid, warps * 32, smem, stream>>>(val_ptrs_dev, reinterpret_cast<const int* const*>(idx_ptrs_dev),
                                                            world, k_in, q_begin, q_count, offs, k_out, out_val, out_idx);
  LAUNCH_CHECK("topk_merge_peers_kernel");
  return LECCR_OK;
}

// ------------------------------------------------------------------------------------ one-call contrastive step
// The Python shim's per-launch overhead (tens of microseconds per torch / ctypes call) dominated the
// training step; these two entries issue the whole forward / backward launch sequence from C++.
size_t leccr_itc_fwd_workspace(int64_t n, int tiles_per_chunk) {
  if (n <= 0) return 0;
  return std::max(leccr_infonce_fwd_workspace(n, tiles_per_chunk), 2 * strip_part_bytes(n) + 256);
}

int leccr_itc_forward(const float* image_feat, int64_t ld_img, const float* text_feat, int64_t ld_txt,
                      const int64_t* idx, int64_t B, int D, int fmt, int rank, int world,
                      void* const* rows_ptrs_dev, void* const* idx_ptrs_dev, uint32_t* const* flag_ptrs_dev,
                      uint32_t epoch, const void* local_slot, size_t local_slot_bytes, void* const* stat_ptrs_dev,
                      const void* local_stat_slot, void* both16, int64_t* idx_all, const float* temp, float* out,
                      float* lse2, float* rcnt, void* workspace, size_t workspace_bytes, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (image_feat == nullptr || text_feat == nullptr || both16 == nullptr || B <= 0 || D <= 0 || (D & 7) != 0 ||
      bad_fmt(fmt) || world < 1 || rank < 0 || rank >= world || (idx != nullptr && idx_all == nullptr) ||
      ld_img < D || ld_txt < D)
    return LECCR_ERR_ARG;
  const int64_t n = B * world;
  if (workspace == nullptr || workspace_bytes < leccr_itc_fwd_workspace(n, 0)) return LECCR_ERR_WORKSPACE;
  const bool strips = world > 1 && stat_ptrs_dev != nullptr && local_stat_slot != nullptr;
  double* scratch = strips ? strip_scratch(n, workspace) : infonce_fwd_scratch(n, 0, workspace);
  uint16_t* both = static_cast<uint16_t*>(both16);
  // ONE launch casts both operands (and pushes idx): into every rank's peer-mapped slot (world > 1, NVLink
  // stores) or straight into the private buffers (world == 1, a "world" of this rank's own two pointers)
  const int wpb = 8;
  dim3 grid(static_cast<unsigned>((B + wpb - 1) / wpb), idx != nullptr ? 3 : 2);
  const unsigned long long* idx_u = reinterpret_cast<const unsigned long long*>(idx);
  if (world == 1) {
    unsigned long long* own_idx = reinterpret_cast<unsigned long long*>(idx_all);
    if (fmt == LECCR_FMT_F16)
      itc_push_kernel<0><<<grid, wpb * 32, 0, stream>>>(image_feat, ld_img, text_feat, ld_txt, idx_u, (int)B, D, nullptr,
                                                       nullptr, both, own_idx, 1, 0, scratch, 8);
    else
      itc_push_kernel<1><<<grid, wpb * 32, 0, stream>>>(image_feat, ld_img, text_feat, ld_txt, idx_u, (int)B, D, nullptr,
                                                       nullptr, both, own_idx, 1, 0, scratch, 8);
    LAUNCH_CHECK("itc_push_kernel");
    return infonce_fwd_impl(both, both + D, 2 * D, idx != nullptr ? idx_all : nullptr, n, D, fmt, temp, out, lse2, rcnt,
                            0, workspace, workspace_bytes, true, stream_);
  }
  if (rows_ptrs_dev == nullptr || flag_ptrs_dev == nullptr || local_slot == nullptr ||
      (idx != nullptr && idx_ptrs_dev == nullptr))
    return LECCR_ERR_ARG;
  {
    uint16_t* const* dsts = reinterpret_cast<uint16_t* const*>(rows_ptrs_dev);
    unsigned long long* const* idsts = reinterpret_cast<unsigned long long* const*>(idx_ptrs_dev);
    if (fmt == LECCR_FMT_F16)
      itc_push_kernel<0><<<grid, wpb * 32, 0, stream>>>(image_feat, ld_img, text_feat, ld_txt, idx_u, (int)B, D, dsts,
                                                       idsts, nullptr, nullptr, world,
                                                       static_cast<long long>(rank) * B, scratch, 8);
    else
      itc_push_kernel<1><<<grid, wpb * 32, 0, stream>>>(image_feat, ld_img, text_feat, ld_txt, idx_u, (int)B, D, dsts,
                                                       idsts, nullptr, nullptr, world,
                                                       static_cast<long long>(rank) * B, scratch, 8);
    LAUNCH_CHECK("itc_push_kernel");
  }
  int rc = leccr_peer_barrier(flag_ptrs_dev, world, rank, epoch, stream_);
  if (rc != LECCR_OK) return rc;
  // private copy (one memcpy: the slot and the private buffer share the layout [rows | idx]): the peer-mapped
  // slot is reused two calls later, the backward runs after that
  CUDA_TRY(cudaMemcpyAsync(both, local_slot, local_slot_bytes, cudaMemcpyDeviceToDevice, stream));
  if (strips)  // this rank's row strips + statistics exchange
    return infonce_fwd_strips(both, both + D, 2 * D, idx != nullptr ? idx_all : nullptr, n, D, fmt, temp, out, lse2,
                              rcnt, static_cast<int64_t>(rank) * B, B, reinterpret_cast<float* const*>(stat_ptrs_dev),
                              static_cast<const float*>(local_stat_slot), world, rank, flag_ptrs_dev, epoch == 0u ? 0u : epoch + 1u,
                              workspace, workspace_bytes, stream_);
  return infonce_fwd_impl(both, both + D, 2 * D, idx != nullptr ? idx_all : nullptr, n, D, fmt, temp, out, lse2, rcnt,
                          0, workspace, workspace_bytes, true, stream_);
}

size_t leccr_itc_bwd_workspace(int64_t n, int64_t row_count, int D) {
  if (n <= 0 || row_count <= 0 || D <= 0) return 0;
  const size_t tr = align256(static_cast<size_t>(D) * round_up8(n) * 2);
  return 2 * tr + leccr_infonce_bwd_workspace(n, row_count, D);
}

int leccr_itc_backward(const void* both16, const int64_t* idx_all, int64_t n, int D, int fmt, const float* temp,
                       const float* lse2, const float* rcnt, const float* out, int64_t row_begin, int64_t row_count,
                       const float* grad_out, float* dA, float* dB, float* dtemp, int one_directional, void* workspace,
                       size_t workspace_bytes, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (both16 == nullptr || workspace == nullptr || n <= 0 || D <= 0 || out == nullptr || grad_out == nullptr)
    return LECCR_ERR_ARG;
  if (workspace_bytes < leccr_itc_bwd_workspace(n, row_count, D)) return LECCR_ERR_WORKSPACE;
  const uint16_t* both = static_cast<const uint16_t*>(both16);
  const int64_t ldT = round_up8(n);
  const size_t tr = align256(static_cast<size_t>(D) * ldT * 2);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  void* aT = ws;
  void* bT = ws + tr;
  if (!bwd_mn_major()) {
    dim3 grid(static_cast<unsigned>((ldT + 31) / 32), static_cast<unsigned>((D + 31) / 32), 2);
    dim3 block(32, 8);
    transpose16_pair_kernel<<<grid, block, 0, stream>>>(both, both + D, 2 * D, (int)n, D, static_cast<uint16_t*>(aT),
                                                       static_cast<uint16_t*>(bT), ldT);
    LAUNCH_CHECK("transpose16_pair_kernel");
  }
  // dL/dtemp = grad_out * (d loss / d temp from the forward), written by the backward's last launch
  return infonce_bwd_impl(both, both + D, 2 * D, aT, bT, ldT, idx_all, n, D, fmt, temp, lse2, rcnt, row_begin,
                          row_count, grad_out, dA, dB, ws + 2 * tr, workspace_bytes - 2 * tr,
                          dtemp != nullptr ? out + (one_directional ? 4 : 1) : nullptr, dtemp, one_directional, stream_);
}

// ------------------------------------------------------------------------------------ caption contrastive loss
size_t leccr_caploss_fwd_workspace(int n_cap, int64_t B) {
  if (n_cap < 1 || B <= 0) return 0;
  return align256(static_cast<size_t>(n_cap) * B * B * 4);
}

int leccr_caploss_fwd(const void* cap16, int64_t ld_cap, const void* txt16, int64_t ld_txt, int n_cap, int64_t B, int K,
                      int fmt, const float* temp, float* out, float* L, uint8_t* amax, float* stats, void* workspace,
                      size_t workspace_bytes, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (cap16 == nullptr || txt16 == nullptr || temp == nullptr || out == nullptr || L == nullptr || amax == nullptr ||
      stats == nullptr || n_cap < 1 || n_cap > 255 || B <= 0 || K <= 0 || bad_fmt(fmt))
    return LECCR_ERR_ARG;
  int rc = leccr_check_device();
  if (rc != LECCR_OK) return rc;
  if (workspace == nullptr || workspace_bytes < leccr_caploss_fwd_workspace(n_cap, B)) return LECCR_ERR_WORKSPACE;
  float* S = static_cast<float*>(workspace);
  StoreProblem sp = {cap16, txt16, ld_cap, ld_txt, static_cast<int64_t>(n_cap) * B, B, S, B, nullptr};
  rc = launch_store(&sp, 1, K, fmt, 1.0f, nullptr, nullptr, 1, stream);
  if (rc != LECCR_OK) return rc;
  const int b = static_cast<int>(B);
  capmax_rows_kernel<<<b, 256, 0, stream>>>(S, n_cap, b, temp, L, amax, stats, stats + 2 * B);
  LAUNCH_CHECK("capmax_rows_kernel");
  capmax_cols_kernel<<<(b + 31) / 32, dim3(32, 8), 0, stream>>>(L, b, temp, stats + B, stats + 3 * B);
  LAUNCH_CHECK("capmax_cols_kernel");
  caploss_finalize_kernel<<<1, 256, 0, stream>>>(L, b, temp, stats, stats + B, stats + 2 * B, stats + 3 * B, out);
  LAUNCH_CHECK("caploss_finalize_kernel");
  return LECCR_OK;
}

size_t leccr_caploss_bwd_workspace(int n_cap, int64_t B, int D) {
  if (n_cap < 1 || B <= 0 || D <= 0) return 0;
  const int64_t nb = static_cast<int64_t>(n_cap) * B;
  const size_t g = align256(static_cast<size_t>(nb) * round_up8(B) * 2);   // G'  [nB][B8]
  const size_t gt = align256(static_cast<size_t>(B) * round_up8(nb) * 2);  // G'^T [B][nB8]
  const size_t tt = align256(static_cast<size_t>(D) * round_up8(B) * 2);   // text^T [D][B8]
  const size_t ct = align256(static_cast<size_t>(D) * round_up8(nb) * 2);  // caption^T [D][nB8]
  return g + gt + tt + ct + 256;
}

int leccr_caploss_bwd(const float* L, const uint8_t* amax, const float* stats, const void* cap16, int64_t ld_cap,
                      const void* txt16, int64_t ld_txt, int n_cap, int64_t B, int D, int fmt, const float* temp,
                      const float* out, const float* grad_out, float* dcap, float* dtxt, float* dtemp, void* workspace,
                      size_t workspace_bytes, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (L == nullptr || amax == nullptr || stats == nullptr || cap16 == nullptr || txt16 == nullptr || temp == nullptr ||
      out == nullptr || grad_out == nullptr || dcap == nullptr || dtxt == nullptr || n_cap < 1 || B <= 0 || D <= 0 ||
      bad_fmt(fmt))
    return LECCR_ERR_ARG;
  if (workspace == nullptr || workspace_bytes < leccr_caploss_bwd_workspace(n_cap, B, D)) return LECCR_ERR_WORKSPACE;
  const int64_t nb = static_cast<int64_t>(n_cap) * B;
  const int64_t b8 = round_up8(B), nb8 = round_up8(nb);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint16_t* G = reinterpret_cast<uint16_t*>(ws);
  ws += align256(static_cast<size_t>(nb) * b8 * 2);
  uint16_t* GT = reinterpret_cast<uint16_t*>(ws);
  ws += align256(static_cast<size_t>(B) * nb8 * 2);
  uint16_t* TT = reinterpret_cast<uint16_t*>(ws);
  ws += align256(static_cast<size_t>(D) * b8 * 2);
  uint16_t* CT = reinterpret_cast<uint16_t*>(ws);
  ws += align256(static_cast<size_t>(D) * nb8 * 2);
  float* scale = reinterpret_cast<float*>(ws);
  const int b = static_cast<int>(B);
  if (fmt == LECCR_FMT_F16)
    capgrad_kernel<0><<<b, 256, 0, stream>>>(L, amax, n_cap, b, (int)b8, temp, stats, stats + B, grad_out, G, scale);
  else
    capgrad_kernel<1><<<b, 256, 0, stream>>>(L, amax, n_cap, b, (int)b8, temp, stats, stats + B, grad_out, G, scale);
  LAUNCH_CHECK("capgrad_kernel");
  int rc = leccr_transpose16(G, nb, b, b8, GT, nb8, stream_);      // [nB][B] -> [B][nB8]
  if (rc != LECCR_OK) return rc;
  const bool mn = bwd_mn_major();  // text and caption rows are read as MN-major operands: no transposed copies
  if (!mn) {
    rc = leccr_transpose16(txt16, B, D, ld_txt, TT, b8, stream_);   // [B][D]  -> [D][B8]
    if (rc != LECCR_OK) return rc;
    rc = leccr_transpose16(cap16, nb, D, ld_cap, CT, nb8, stream_); // [nB][D] -> [D][nB8]
    if (rc != LECCR_OK) return rc;
  }
  // d caption = G' text * s ;  d text = G'^T caption * s   (s = grad_out / (2 B temp), fp32, in the epilogue)
  StoreProblem p0 = {G, mn ? txt16 : TT, b8, mn ? ld_txt : b8, nb, D, dcap, D, nullptr, mn};
  rc = launch_store(&p0, 1, b, fmt, 1.0f, scale, nullptr, 1, stream);
  if (rc != LECCR_OK) return rc;
  StoreProblem p1 = {GT, mn ? cap16 : CT, nb8, mn ? ld_cap : nb8, B, D, dtxt, D, nullptr, mn};
  rc = launch_store(&p1, 1, static_cast<int>(nb), fmt, 1.0f, scale, nullptr, 1, stream, grad_out, out + 1, dtemp);
  return rc;
}

int leccr_topk_dense(const float* S, int64_t ld, int64_t R, int64_t C, int by_columns, int k, float* out_val,
                     int32_t* out_idx, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (S == nullptr || out_val == nullptr || out_idx == nullptr || R <= 0 || C <= 0 || ld < C || k < 1 ||
      k > kDenseTopkMax)
    return LECCR_ERR_ARG;
  // by_columns: rank the columns of S (the reference's t2i matrix is the transpose VIEW of i2t, :152 / :179)
  const long long ld_r = by_columns ? 1 : ld, ld_c = by_columns ? ld : 1;
  const int n_rank = static_cast<int>(by_columns ? C : R), n_scan = static_cast<int>(by_columns ? R : C);
  const int wpb = 8;
  topk_dense_kernel<<<(n_rank + wpb - 1) / wpb, wpb * 32, 0, stream>>>(S, ld_r, ld_c, n_rank, n_scan, k, out_val, out_idx);
  LAUNCH_CHECK("topk_dense_kernel");
  return LECCR_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------ dstl_loss
extern "C" {

size_t leccr_dstl_fwd_workspace(int n_cap, int64_t N) {
  if (n_cap < 1 || N <= 0) return 0;
  const size_t nn = static_cast<size_t>(N) * N * 4;
  return align256(nn * n_cap) + align256(nn) + align256(static_cast<size_t>(N) * 4) + 256;  // SIM, Cmax, row_loss, mm
}

int leccr_dstl_fwd(const void* tt16, int64_t ld_tt, const void* ts16_rows, int64_t ld_tsr, const void* ts16_cols,
                   int64_t ld_tsc, const void* img16, int64_t ld_img, const void* cap16, int64_t ld_cap, int n_cap,
                   int64_t N, int K, int fmt, float alpha, float* out, float* Fm, float* TV, float* lse, void* workspace,
                   size_t workspace_bytes, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (tt16 == nullptr || ts16_rows == nullptr || ts16_cols == nullptr || img16 == nullptr || cap16 == nullptr ||
      out == nullptr || Fm == nullptr || TV == nullptr || lse == nullptr || n_cap < 1 || N <= 0 || K <= 0 || bad_fmt(fmt))
    return LECCR_ERR_ARG;
  int rc = leccr_check_device();
  if (rc != LECCR_OK) return rc;
  if (workspace == nullptr || workspace_bytes < leccr_dstl_fwd_workspace(n_cap, N)) return LECCR_ERR_WORKSPACE;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const size_t nn = static_cast<size_t>(N) * N * 4;
  float* SIM = reinterpret_cast<float*>(ws);
  ws += align256(nn * n_cap);
  float* Cmax = reinterpret_cast<float*>(ws);
  ws += align256(nn);
  float* row_loss = reinterpret_cast<float*>(ws);
  ws += align256(static_cast<size_t>(N) * 4);
  uint32_t* mm = reinterpret_cast<uint32_t*>(ws);
  // logits_tv and logits_sv share one tensor-core launch; the caption similarities take a second one
  StoreProblem sp[2] = {{tt16, img16, ld_tt, ld_img, N, N, TV, N, nullptr}, {ts16_rows, img16, ld_tsr, ld_img, N, N, Fm, N, nullptr}};
  rc = launch_store(sp, 2, K, fmt, 1.0f, nullptr, nullptr, 1, stream);
  if (rc != LECCR_OK) return rc;
  StoreProblem sc = {cap16, ts16_cols, ld_cap, ld_tsc, static_cast<int64_t>(n_cap) * N, N, SIM, N, nullptr};
  rc = launch_store(&sc, 1, K, fmt, 1.0f, nullptr, nullptr, 1, stream);
  if (rc != LECCR_OK) return rc;
  rc = leccr_double_sim_fuse(Fm, SIM, n_cap, static_cast<int64_t>(N) * N, Cmax, mm, alpha, 1.0f - alpha, LECCR_FUSE_NORM,
                             stream_);
  if (rc != LECCR_OK) return rc;
  const int n = static_cast<int>(N);
  dstl_rows_kernel<<<n, 256, 0, stream>>>(Fm, TV, n, lse, lse + N, row_loss);
  LAUNCH_CHECK("dstl_rows_kernel");
  dstl_finalize_kernel<<<1, 256, 0, stream>>>(row_loss, n, out);
  LAUNCH_CHECK("dstl_finalize_kernel");
  return LECCR_OK;
}

size_t leccr_dstl_bwd_workspace(int64_t N, int64_t row_count, int D) {
  if (N <= 0 || row_count <= 0 || D <= 0) return 0;
  const int k_chunks = static_cast<int>((N + BK - 1) / BK);
  const int row_blocks = static_cast<int>((row_count + BM - 1) / BM);
  const size_t parts = 2 * static_cast<size_t>(bwd_splits(k_chunks, row_blocks)) * row_count * D * 4;
  return 2 * align256(static_cast<size_t>(row_count) * round_up8(N) * 2) +
         2 * align256(static_cast<size_t>(D) * round_up8(N) * 2) + align256(parts) + 256;
}

int leccr_dstl_bwd(const float* Fm, const float* TV, const float* lse, const void* img16, int64_t ld_img, const void* tt16,
                   int64_t ld_tt, int64_t N, int D, int fmt, int64_t row_begin, int64_t row_count, const float* grad_out,
                   float* dimg, float* dtt, void* workspace, size_t workspace_bytes, leccr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (Fm == nullptr || TV == nullptr || lse == nullptr || img16 == nullptr || tt16 == nullptr || grad_out == nullptr ||
      dimg == nullptr || dtt == nullptr || N <= 0 || D <= 0 || bad_fmt(fmt) || row_begin < 0 || row_count <= 0 ||
      row_begin + row_count > N)
    return LECCR_ERR_ARG;
  if (workspace == nullptr || workspace_bytes < leccr_dstl_bwd_workspace(N, row_count, D)) return LECCR_ERR_WORKSPACE;
  const int64_t n8 = round_up8(N);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint16_t* Gr = reinterpret_cast<uint16_t*>(ws);
  ws += align256(static_cast<size_t>(row_count) * n8 * 2);
  uint16_t* GcT = reinterpret_cast<uint16_t*>(ws);
  ws += align256(static_cast<size_t>(row_count) * n8 * 2);
  uint16_t* imgT = reinterpret_cast<uint16_t*>(ws);
  ws += align256(static_cast<size_t>(D) * n8 * 2);
  uint16_t* ttT = reinterpret_cast<uint16_t*>(ws);
  ws += align256(static_cast<size_t>(D) * n8 * 2);
  const int k_chunks = static_cast<int>((N + BK - 1) / BK);
  const int row_blocks = static_cast<int>((row_count + BM - 1) / BM);
  const int splits = bwd_splits(k_chunks, row_blocks);
  float* parts0 = reinterpret_cast<float*>(ws);
  float* parts1 = parts0 + static_cast<size_t>(splits) * row_count * D;
  ws += align256(2 * static_cast<size_t>(splits) * row_count * D * 4);
  float* scale = reinterpret_cast<float*>(ws);
  const int n = static_cast<int>(N);
  if (fmt == LECCR_FMT_F16)
    dstl_grad_kernel<0><<<n, 256, 0, stream>>>(Fm, TV, n, (int)n8, lse, lse + N, (int)row_begin, (int)row_count, grad_out,
                                              Gr, GcT, scale);
  else
    dstl_grad_kernel<1><<<n, 256, 0, stream>>>(Fm, TV, n, (int)n8, lse, lse + N, (int)row_begin, (int)row_count, grad_out,
                                              Gr, GcT, scale);
  LAUNCH_CHECK("dstl_grad_kernel");
  const bool mn = bwd_mn_major();  // image and text rows are read as MN-major operands: no transposed copies
  if (!mn) {
    dim3 grid(static_cast<unsigned>((n8 + 31) / 32), static_cast<unsigned>((D + 31) / 32), 2);
    dim3 block(32, 8);
    if (ld_img != ld_tt) return LECCR_ERR_ARG;
    transpose16_pair_kernel<<<grid, block, 0, stream>>>(static_cast<const uint16_t*>(img16), static_cast<const uint16_t*>(tt16),
                                                       ld_img, n, D, imgT, ttT, n8);
    LAUNCH_CHECK("transpose16_pair_kernel");
  }
  // d text_t[loc] = Gr image * s ;  d image[loc] = GcT text_t * s    (s = grad_out / N^2)
  StoreProblem sp[2] = {
      {Gr, mn ? img16 : imgT, n8, mn ? ld_img : n8, row_count, D, dtt, D, parts0, mn},
      {GcT, mn ? tt16 : ttT, n8, mn ? ld_tt : n8, row_count, D, dimg, D, parts1, mn},
  };
  return launch_store(sp, 2, n, fmt, 1.0f, scale, nullptr, splits, stream);
}

}  // extern "C"

// ------------------------------------------------------------------------------------ NCCL collectives
// For hosts without torch.distributed (and for node-crossing groups, where peer memory does not reach): the two
// collectives of the path as thin C entries over NCCL.  libnccl.so.2 is resolved at run time (dlopen: the copy
// the process already loaded, e.g. torch's, is reused), so the library has no link-time dependency on it.
namespace {
struct NcclApi {
  int (*get_unique_id)(void*);
  int (*comm_init_rank)(void**, int, leccr_nccl_id, int);
  int (*comm_destroy)(void*);
  int (*all_gather)(const void*, void*, size_t, int, void*, cudaStream_t);
  const char* (*get_error_string)(int);
  bool ok;
};
NcclApi& nccl_api() {
  static NcclApi api = {};
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) return;
    api.get_unique_id = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclGetUniqueId"));
    api.comm_init_rank = reinterpret_cast<int (*)(void**, int, leccr_nccl_id, int)>(dlsym(h, "ncclCommInitRank"));
    api.comm_destroy = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclCommDestroy"));
    api.all_gather =
        reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(dlsym(h, "ncclAllGather"));
    api.get_error_string = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_gather;
  });
  return api;
}
int nccl_fail(int r, const char* what) {
  NcclApi& a = nccl_api();
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: NCCL error %d (%s)", what, r,
           a.get_error_string != nullptr ? a.get_error_string(r) : "?");
  return LECCR_ERR_NCCL;
}
}  // namespace

extern "C" {

int leccr_comm_unique_id(leccr_nccl_id* id_host) {
  if (id_host == nullptr) return LECCR_ERR_ARG;
  NcclApi& a = nccl_api();
  if (!a.ok) return LECCR_ERR_NCCL;
  const int r = a.get_unique_id(id_host);
  return r == 0 ? LECCR_OK : nccl_fail(r, "ncclGetUniqueId");
}

int leccr_comm_init(const leccr_nccl_id* id_host, int rank, int world, void** comm) {
  if (id_host == nullptr || comm == nullptr || world < 1 || rank < 0 || rank >= world) return LECCR_ERR_ARG;
  NcclApi& a = nccl_api();
  if (!a.ok) return LECCR_ERR_NCCL;
  const int r = a.comm_init_rank(comm, world, *id_host, rank);
  return r == 0 ? LECCR_OK : nccl_fail(r, "ncclCommInitRank");
}

int leccr_comm_destroy(void* comm) {
  if (comm == nullptr) return LECCR_ERR_ARG;
  NcclApi& a = nccl_api();
  if (!a.ok) return LECCR_ERR_NCCL;
  const int r = a.comm_destroy(comm);
  return r == 0 ? LECCR_OK : nccl_fail(r, "ncclCommDestroy");
}

int leccr_allgather(void* comm, const void* send, void* recv, size_t bytes_per_rank, leccr_stream_t stream_) {
  if (comm == nullptr || send == nullptr || recv == nullptr || bytes_per_rank == 0) return LECCR_ERR_ARG;
  NcclApi& a = nccl_api();
  if (!a.ok) return LECCR_ERR_NCCL;
  const int r = a.all_gather(send, recv, bytes_per_rank, /*ncclUint8*/ 1, comm, static_cast<cudaStream_t>(stream_));
  return r == 0 ? LECCR_OK : nccl_fail(r, "ncclAllGather");
}

}  // extern "C"
