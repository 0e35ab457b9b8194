// tcgen05 (5th-gen tensor core) building blocks of the level sweep: fp16 x 3 split precision on kind::f16, K-major
// SWIZZLE_128B operand tiles in shared memory, fp32 accumulators in TMEM, and (the cluster sweep) the weight operand
// resident in TMEM.
//
// Why a split: single-pass TF32 / FP16 / BF16 operands give 1e-3 max-abs state error after the level recurrence
// (SURVEY.md §10-P5), outside the 1e-4 parity bar. Each fp32 operand x is split into hi = rn_f16(x), lo = rn_f16(x - hi):
// 22 significant bits for |x| in [6e-5, 65504), an absolute error <= 2^-25 below that (fp16 subnormals). The product is
// accumulated in fp32 as hi*hi + lo*hi + hi*lo (+ lo*lo where it comes for free); measured against an fp64 GEMM the split
// costs ~2e-6 max-abs on the node states after the full level recurrence, the same order as plain fp32.
//
// Shared-memory operand tile layout: canonical K-major SWIZZLE_128B — a row is 128 bytes (64 fp16 = one swizzle atom
// wide), rows are packed in groups of 8 (1024 B, the descriptor's stride-byte-offset), and the 16-byte chunk index of a
// row is XORed with (row % 8). One tcgen05.mma kind::f16 consumes K = 16 (32 bytes); stepping K inside the atom adds 32 B
// to the descriptor's start address (the hardware swizzles on absolute smem address bits, hence the 1024-byte alignment
// of every tile). Bit layouts follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor / UMMA::InstrDescriptor).
//
// TMEM-resident A operand (tcgen05.mma with [a_tmem]): row m of A sits in TMEM lane m, its K values are packed two fp16
// per 32-bit column (k = 2 * column + half); one K = 16 step reads 8 consecutive columns. Written with tcgen05.st 32x32b
// (thread t of warp w owns lane 32 * (w % 4) + t).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dagnn {
namespace tc {

constexpr int ROW_BYTES = 128;
constexpr int GROUP_BYTES = 1024;      // 8 rows
constexpr int KC16 = 64;               // k per chunk = one 128-byte swizzle atom of fp16

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of the 16-byte chunk c8 (0..7, 8 consecutive k) of row r inside a SWIZZLE_128B K-major tile
__device__ __forceinline__ uint32_t tile_off(int r, int c8) {
  return (uint32_t)((r >> 3) * GROUP_BYTES + (r & 7) * ROW_BYTES + ((c8 ^ (r & 7)) << 4));
}

// ---- descriptors
// group_stride = bytes between consecutive groups of 8 rows (the descriptor's stride-byte-offset): 1024 for a dense tile
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t group_stride = GROUP_BYTES) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);                     // start address, LBO (unused for SW128 K-major)
  const uint32_t hi = (group_stride >> 4) | (1u << 14) | (2u << 29);             // SBO, version = 1, SWIZZLE_128B
  return ((uint64_t)hi << 32) | lo;
}
__host__ __device__ constexpr uint32_t instr_desc_f16(int M, int N) {
  return (1u << 4) /* D = f32 */ | (0u << 7) /* A = f16 */ | (0u << 10) /* B = f16 */ | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);   // a_major = b_major = 0: K-major
}

// ---- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot_in_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {          // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// mbarrier arrives when every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// 32 lanes x 8 consecutive columns: thread t of warp w reads lane 32*(w%4)+t
__device__ __forceinline__ void ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 32 consecutive columns
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 8 consecutive columns, registers -> TMEM (thread t of warp w writes lane 32*(w%4)+t)
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- fp16 x 3 split ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 hh = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(hh);
  const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}
// 8 consecutive k of one row -> one 16-byte chunk of the hi tile and one of the lo tile
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t* h = reinterpret_cast<uint32_t*>(&hi);
  uint32_t* l = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
  for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], h[j], l[j]);
}
__device__ __forceinline__ void store_split8(unsigned char* hi_tile, unsigned char* lo_tile, int r, int c8, const float (&v)[8]) {
  uint4 hi, lo;
  split8(v, hi, lo);
  const uint32_t off = tile_off(r, c8);
  *reinterpret_cast<uint4*>(hi_tile + off) = hi;
  *reinterpret_cast<uint4*>(lo_tile + off) = lo;
}
// D[tmem] (+)= A[smem] * B[smem]^T, M x N x 16, issued by ONE thread
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (rows = TMEM lanes, 8 columns per K = 16 step) stays in tensor memory
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the three products of one K = 16 step: D (+)= Ahi*Bhi + Alo*Bhi + Ahi*Blo. `first`: the first product overwrites D
__device__ __forceinline__ void mma3_f16(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc,
                                         bool first) {
  mma_f16(tmem_d, a_hi, b_hi, idesc, first ? 0u : 1u);
  mma_f16(tmem_d, a_lo, b_hi, idesc, 1u);
  mma_f16(tmem_d, a_hi, b_lo, idesc, 1u);
}

}  // namespace tc
}  // namespace dagnn
