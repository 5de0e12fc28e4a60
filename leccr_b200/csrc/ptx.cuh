// Thin inline-PTX wrappers for the sm_100a features the similarity kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and fences.
// Nothing here is specific to LECCR; the kernels in gemm_sm100.cuh build on it.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace leccr {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of the (converged) warp gets true.  Unlike `lane == 0`, ptxas knows that the code this predicate guards
// is executed by a single thread, so values feeding uniform-register operands (tcgen05.mma descriptors, TMA
// coordinates) need no per-instruction "elect + broadcast + loop over distinct values" sequence.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// try_wait with a suspend-time hint: the thread sleeps in hardware (no issue slots burnt) until the
// phase completes or ~hint_ns elapse.
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must not hang the GPU (a hung box is a lost round), so after a few
// seconds of waiting the kernel reports where it was stuck and traps.  backoff_ns > 0 sleeps between
// polls: the waiting role warps share their schedulers with epilogue warps and must not steal issue
// slots from them (use 0 only where wake-up latency matters more).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag, uint32_t backoff_ns = 0) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  // backoff_ns > 0: the wait is expected to be long (an epilogue warp waiting for the next accumulator, the producer
  // for a free stage): park the thread in hardware with a suspend-time hint instead of polling -- the polling loop
  // of the epilogue warps alone executed 300 M try_waits per cfg5 launch (issue slots and power on a part that runs
  // this kernel under its power cap)
  while (!(backoff_ns ? mbar_try_wait_hint(bar, parity, 2000u) : mbar_try_wait(bar, parity))) {
    if (++spins > (backoff_ns ? 4000000u : 100000000u)) {
      printf("leccr: mbarrier wait timed out (tag %d, block %d, thread %d, parity %u)\n", tag,
             (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 2-D tiled load global -> shared, completion signalled on an mbarrier (bytes).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* tmap,
                                                 uint64_t* bar, int c_inner, int c_outer,
                                                 uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer),
        "l"(hint)
      : "memory");
}
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// ----------------------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::f16 covers fp16 and bf16 operands.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// Shared-memory matrix descriptor for a K-major tile stored as rows of 128 bytes with the
// 128-byte swizzle TMA produces (8-row x 128 B atoms, 1024 B apart). Field layout follows the
// PTX ISA "matrix descriptor" for tcgen05: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64) with 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;          // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;  // stride byte offset: next 8-row group
  d |= static_cast<uint64_t>(1) << 46;          // descriptor version for sm_100
  d |= static_cast<uint64_t>(2) << 61;          // SWIZZLE_128B
  return d;
}

// MN-major operand (the contraction index is the SLOW axis of the row-major source, e.g. the gathered [n][D]
// matrix in  dA = G' * B): TMA lands {64 elements of MN (128 bytes), kBK rows of K} boxes with the 128-byte swizzle;
// a box is kBK / 8 atoms of 8 K-rows x 128 B (1024 bytes apart = the stride byte offset), the boxes of a tile follow
// each other along MN `mn_block_bytes` apart (the leading byte offset).  A K step of 16 is 16 rows = +2048 bytes.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t mn_block_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(mn_block_bytes >> 4) << 16;  // leading byte offset: next 64-element block along MN
  d |= static_cast<uint64_t>(1024 >> 4) << 32;            // stride byte offset: next group of 8 K-rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;                    // SWIZZLE_128B
  return d;
}

// The same for rows of kRowBytes = 2 * kBK bytes: 128 (SWIZZLE_128B, 1024-byte atoms) or 64 (SWIZZLE_64B,
// layout type 4, 8-row x 64 B atoms 512 B apart).  A K step of 16 elements is +32 bytes inside the row either way.
template <int kBK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  if (kBK == 64) return make_sw128_kmajor_desc(smem_addr);
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;  // SWIZZLE_64B
  return d;
}

// Instruction descriptor for kind::f16, fp32 accumulate, A K-major, B K-major or (b_mn) MN-major.
// fmt: 0 = fp16, 1 = bf16.
__host__ __device__ constexpr uint32_t make_idesc_f16(int fmt, int m, int n, bool b_mn = false) {
  return (1u << 4)                          // D format: f32
         | (static_cast<uint32_t>(fmt) << 7)   // A format
         | (static_cast<uint32_t>(fmt) << 10)  // B format
         | (b_mn ? (1u << 16) : 0u)            // B major: 0 = K, 1 = MN
         | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// 32 lanes x 32 consecutive columns of fp32 accumulators: thread t of the warp gets TMEM lane
// (lane_base + t), columns [col, col + 32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// tcgen05.wait::ld that also names the destination registers, so the compiler cannot schedule
// their consumers above the wait.
__device__ __forceinline__ void tmem_wait_ld_regs(float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                 "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]),
                 "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]),
                 "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]),
                 "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// Shared-memory accesses through 32-bit shared-window addresses (keeps LDS/STS instead of generic LD/ST).
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ long long lds_s64(uint32_t a) {
  long long v;
  asm volatile("ld.shared.s64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_s32(uint32_t a, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_s64(uint32_t a, long long v) {
  asm volatile("st.shared.s64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
// Barrier among the 128 threads of one epilogue warpgroup (barrier 0 is __syncthreads).
__device__ __forceinline__ void epi_bar_sync(int wg) {
  asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
}
// exp2 on the SFU without the denormal fix-up sequence (arguments here are <= 0 or modest).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// if (x > thr) { *addr = x; *(addr + IDX_OFF) = idx_base + E; addr += 4; }  -- 5 predicated
// instructions, no branch; the running address replaces a separate counter.
template <int IDX_OFF, int E>
__device__ __forceinline__ void append_if_gt(float x, float thr, uint32_t& addr, int idx_base) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .s32 t;\n"
      "setp.gt.f32 p, %1, %2;\n"
      "add.s32 t, %3, %5;\n"
      "@p st.shared.f32 [%0], %1;\n"
      "@p st.shared.s32 [%0+%4], t;\n"
      "@p add.u32 %0, %0, 4;\n"
      "}\n"
      : "+r"(addr)
      : "f"(x), "f"(thr), "r"(idx_base), "n"(IDX_OFF), "n"(E)
      : "memory");
}
// if (x > thr) { list_v[slot] = x; list_i[slot] = idx; } as predicated stores (no branch).
__device__ __forceinline__ void sts_pair_if_gt(float x, float thr, uint32_t addr_v, uint32_t addr_i, int idx) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.gt.f32 p, %0, %1;\n"
      "@p st.shared.f32 [%2], %0;\n"
      "@p st.shared.s32 [%3], %4;\n"
      "}\n"
      :
      : "f"(x), "f"(thr), "r"(addr_v), "r"(addr_i), "r"(idx)
      : "memory");
}

// Append to a list of interleaved (score, column) entries whose cursor is a shared-window ADDRESS:
// if (x > thr) *cursor = {x, idx} as ONE predicated 64-bit store; returns the cursor advance (8 or 0) so the
// caller forms the next cursor in a fresh register (overwriting the address register of a store still in
// flight stalls on its operand read).  Shared-store issue slots, not ALU work, bound the unfiltered loop.
__device__ __forceinline__ uint32_t append_if_gt(float x, float thr, uint32_t cursor, int idx) {
  uint32_t inc;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.gt.f32 p, %1, %2;\n"
      "selp.u32 %0, 8, 0, p;\n"
      "@p st.shared.v2.b32 [%3], {%4, %5};\n"
      "}\n"
      : "=r"(inc)
      : "f"(x), "f"(thr), "r"(cursor), "r"(__float_as_uint(x)), "r"(idx)
      : "memory");
  return inc;
}

// Four appends at once: the cursor offsets of the four entries are a prefix sum of the four hit flags that
// does not depend on the cursor, so the loop-carried chain is ONE add per four elements (the one-at-a-time
// form above is bound by its add -> select chain, ~14 cycles per element with a single warp per scheduler).
__device__ __forceinline__ uint32_t append4_if_gt(float x0, float x1, float x2, float x3, float thr, uint32_t cursor,
                                                  int idx0) {
  uint32_t next;
  asm volatile(
      "{\n"
      ".reg .pred p0, p1, p2, p3;\n"
      ".reg .u32 i0, i1, i2, i3, s01, s23, a1, a2, a3, j1, j2, j3;\n"
      "setp.gt.f32 p0, %2, %6;\n"
      "setp.gt.f32 p1, %3, %6;\n"
      "setp.gt.f32 p2, %4, %6;\n"
      "setp.gt.f32 p3, %5, %6;\n"
      "selp.u32 i0, 8, 0, p0;\n"
      "selp.u32 i1, 8, 0, p1;\n"
      "selp.u32 i2, 8, 0, p2;\n"
      "selp.u32 i3, 8, 0, p3;\n"
      "add.u32 s01, i0, i1;\n"
      "add.u32 s23, i2, i3;\n"
      "add.u32 a1, %1, i0;\n"
      "add.u32 a2, %1, s01;\n"
      "add.u32 a3, a2, i2;\n"
      "add.u32 j1, %11, 1;\n"
      "add.u32 j2, %11, 2;\n"
      "add.u32 j3, %11, 3;\n"
      "@p0 st.shared.v2.b32 [%1], {%7, %11};\n"
      "@p1 st.shared.v2.b32 [a1], {%8, j1};\n"
      "@p2 st.shared.v2.b32 [a2], {%9, j2};\n"
      "@p3 st.shared.v2.b32 [a3], {%10, j3};\n"
      "add.u32 s01, s01, s23;\n"
      "add.u32 %0, %1, s01;\n"
      "}\n"
      : "=r"(next)
      : "r"(cursor), "f"(x0), "f"(x1), "f"(x2), "f"(x3), "f"(thr), "r"(__float_as_uint(x0)),
        "r"(__float_as_uint(x1)), "r"(__float_as_uint(x2)), "r"(__float_as_uint(x3)), "r"(idx0)
      : "memory");
  return next;
}

// if (x > hi) mask_hi |= BIT; if (x > lo) mask_lo |= BIT: two compares and two predicated ORs -- the counting
// epilogue's whole per-element work.  Masks instead of counters: popc gives the counts, mask_lo & ~mask_hi names the
// elements inside the band without a second pass over the registers (which cannot be indexed per lane).
template <unsigned BIT>
__device__ __forceinline__ void mask_gt2(float x, float hi, float lo, unsigned& mask_hi, unsigned& mask_lo) {
  asm("{\n"
      ".reg .pred p, q;\n"
      "setp.gt.f32 p, %2, %3;\n"
      "setp.gt.f32 q, %2, %4;\n"
      "@p or.b32 %0, %0, %5;\n"
      "@q or.b32 %1, %1, %5;\n"
      "}\n"
      : "+r"(mask_hi), "+r"(mask_lo)
      : "f"(x), "f"(hi), "f"(lo), "n"(BIT));
}

// Order-preserving float <-> uint32 key (any sign, +-inf included).
__device__ __forceinline__ uint32_t f32_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_f32(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace leccr
