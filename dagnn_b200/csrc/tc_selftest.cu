// Stand-alone checks of the tcgen05 building blocks in tc.cuh with the same operand tile layout, descriptors, MMA issue and
// TMEM read-back the level kernels use:
//   k_tc_selftest     C[M,N] = A[M,K] * B[N,K]^T, both operands from shared memory, fp16 x 3 split. One CTA per 128 rows.
//   k_tc_selftest_ts  C[N,R] = X[N,K] * W[R,K]^T with W as the TMEM-resident A operand ([hi ; lo] stacked on the lanes) and
//                     the rows of X as the shared-memory B operand (hi and lo tiles) — the cluster sweep's arrangement.
#include "common.cuh"
#include "tc.cuh"

namespace dagnn {

__device__ __forceinline__ void st_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void st_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(tc::smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__global__ void __launch_bounds__(256, 1) k_tc_selftest(const float* __restrict__ A, const float* __restrict__ B,
                                                        float* __restrict__ C, int M, int N, int K) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* A_hi = base;
  unsigned char* A_lo = A_hi + 128 * tc::ROW_BYTES;
  unsigned char* B_hi = A_lo + 128 * tc::ROW_BYTES;
  unsigned char* B_lo = B_hi + 256 * tc::ROW_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(B_lo + 256 * tc::ROW_BYTES);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(slot, 256);
  if (tid == 0) {
    st_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  const int row0 = blockIdx.x * 128;
  const uint32_t idesc = tc::instr_desc_f16(128, N);
  const int nchunks = K / tc::KC16;
  for (int c = 0; c < nchunks; ++c) {
    for (int it = tid; it < 128 * 8; it += 256) {
      const int r = it >> 3, c8 = it & 7;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (row0 + r < M) load8(A + (size_t)(row0 + r) * K + c * tc::KC16 + 8 * c8, v);
      tc::store_split8(A_hi, A_lo, r, c8, v);
    }
    for (int it = tid; it < N * 8; it += 256) {
      const int r = it >> 3, c8 = it & 7;
      float v[8];
      load8(B + (size_t)r * K + c * tc::KC16 + 8 * c8, v);
      tc::store_split8(B_hi, B_lo, r, c8, v);
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint64_t ah = tc::smem_desc(tc::smem_addr(A_hi)), al = tc::smem_desc(tc::smem_addr(A_lo));
      const uint64_t bh = tc::smem_desc(tc::smem_addr(B_hi)), bl = tc::smem_desc(tc::smem_addr(B_lo));
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) tc::mma3_f16(tmem, ah + 2 * ks, al + 2 * ks, bh + 2 * ks, bl + 2 * ks, idesc, c == 0 && ks == 0);
      tc::commit(bar);
    }
    st_mbar_wait(bar, (uint32_t)(c & 1));   // MMAs of this chunk done: operand tiles may be overwritten
  }
  tc::fence_after_sync();
  const int q = warp & 3, half = warp >> 2;
  const int row = row0 + 32 * q + lane;
  const int ncb = N / 8;
  for (int cb = half; cb < ncb; cb += 2) {
    float v[8];
    tc::ld8(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(cb * 8), v);
    tc::wait_ld();
    if (row < M) {
      float4* o = reinterpret_cast<float4*>(C + (size_t)row * N + cb * 8);
      o[0] = make_float4(v[0], v[1], v[2], v[3]);
      o[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// W [R <= 64, K <= 256] lives in TMEM: lane r = hi half of row r, lane 64 + r = lo half, two fp16 k per column (columns
// [0, K/2)). X [N, K] goes through shared memory in 64-k chunks (hi and lo tile). Accumulators: columns [256, 256 + N).
// Both MMAs of a k step accumulate into the same columns: lane r ends up with W_hi . (x_hi + x_lo), lane 64 + r with
// W_lo . (x_hi + x_lo); their sum is the product in the full split precision (all four hi/lo terms).
__global__ void __launch_bounds__(256, 1) k_tc_selftest_ts(const float* __restrict__ W, const float* __restrict__ X,
                                                           float* __restrict__ C, int R, int N, int K) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* B_hi = base;
  unsigned char* B_lo = B_hi + 256 * tc::ROW_BYTES;
  float* red = reinterpret_cast<float*>(B_lo + 256 * tc::ROW_BYTES);        // [64][257] lo-half accumulators
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 64 * 258);          // 8-byte aligned
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(slot, 512);
  if (tid == 0) {
    st_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  if (warp < 4) {                                   // weights -> TMEM: thread = lane of tensor memory
    const int r = tid & 63, part = tid >> 6;        // lanes 0..63 hi, 64..127 lo
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t hi = 0u, lo = 0u;
        if (r < R) tc::split2(W[(size_t)r * K + 16 * ks + 2 * j], W[(size_t)r * K + 16 * ks + 2 * j + 1], hi, lo);
        w[j] = part ? lo : hi;
      }
      tc::st8(taddr + (uint32_t)(8 * ks), w);
    }
    tc::wait_st();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t idesc = tc::instr_desc_f16(128, N);
  const uint32_t dcol = tmem + 256u;
  const int nchunks = (K + tc::KC16 - 1) / tc::KC16;
  for (int c = 0; c < nchunks; ++c) {
    for (int it = tid; it < N * 8; it += 256) {
      const int r = it >> 3, c8 = it & 7;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (c * tc::KC16 + 8 * c8 < K) load8(X + (size_t)r * K + c * tc::KC16 + 8 * c8, v);
      tc::store_split8(B_hi, B_lo, r, c8, v);
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint64_t bh = tc::smem_desc(tc::smem_addr(B_hi)), bl = tc::smem_desc(tc::smem_addr(B_lo));
      const int nks = min(4, (K - c * tc::KC16) / 16);
      for (int ks = 0; ks < nks; ++ks) {
        const uint32_t a = tmem + (uint32_t)(32 * c + 8 * ks);
        tc::mma_f16_ts(dcol, a, bh + 2 * ks, idesc, (c == 0 && ks == 0) ? 0u : 1u);
        tc::mma_f16_ts(dcol, a, bl + 2 * ks, idesc, 1u);
      }
      tc::commit(bar);
    }
    st_mbar_wait(bar, (uint32_t)(c & 1));
  }
  tc::fence_after_sync();
  // lanes 64..127 (lo halves) park their accumulators in shared memory, lanes 0..63 add them and store
  if (warp < 4) {
    const int q = warp;
    for (int cb = 0; cb < N / 8; ++cb) {
      float v[8];
      tc::ld8(dcol + ((uint32_t)(32 * q) << 16) + (uint32_t)(cb * 8), v);
      tc::wait_ld();
      if (q >= 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[(32 * (q - 2) + lane) * 257 + cb * 8 + j] = v[j];
      }
    }
  }
  __syncthreads();
  if (warp < 2) {
    const int r = 32 * warp + lane;
    for (int cb = 0; cb < N / 8; ++cb) {
      float v[8];
      tc::ld8(dcol + ((uint32_t)(32 * warp) << 16) + (uint32_t)(cb * 8), v);
      tc::wait_ld();
      if (r < R) {
#pragma unroll
        for (int j = 0; j < 8; ++j) C[(size_t)(cb * 8 + j) * R + r] = v[j] + red[r * 257 + cb * 8 + j];
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace dagnn

using namespace dagnn;

extern "C" int dagnn_tc_selftest_f16x3(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, void* stream_) {
  DAGNN_REQUIRE(A && B && C, "tc_selftest: null pointer");
  DAGNN_REQUIRE(M > 0 && N >= 16 && N <= 256 && N % 16 == 0 && K >= tc::KC16 && K % tc::KC16 == 0,
                "tc_selftest: M>0, N%16==0 in [16,256], K%64==0");
  DAGNN_REQUIRE(((((uintptr_t)A) | ((uintptr_t)B) | ((uintptr_t)C)) & 15) == 0, "tc_selftest: 16-byte alignment");
  const size_t smem = 1024 + (size_t)(2 * 128 + 2 * 256) * tc::ROW_BYTES + 64;
  static PerDeviceOnce once;
  if (int rc = per_device_once(once, nullptr, [&](int) {
        DAGNN_CUDA_OK(cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return (int)DAGNN_OK;
      }))
    return rc;
  k_tc_selftest<<<(M + 127) / 128, 256, smem, static_cast<cudaStream_t>(stream_)>>>(A, B, C, M, N, K);
  return check_launch("k_tc_selftest");
}

extern "C" int dagnn_tc_selftest_ts(const float* W, const float* X, float* C, int32_t R, int32_t N, int32_t K, void* stream_) {
  DAGNN_REQUIRE(W && X && C, "tc_selftest_ts: null pointer");
  DAGNN_REQUIRE(R > 0 && R <= 64 && N >= 16 && N <= 256 && N % 16 == 0 && K >= 16 && K <= 256 && K % 16 == 0,
                "tc_selftest_ts: R in [1,64], N%16==0 in [16,256], K%16==0 in [16,256]");
  DAGNN_REQUIRE(((((uintptr_t)W) | ((uintptr_t)X) | ((uintptr_t)C)) & 15) == 0, "tc_selftest_ts: 16-byte alignment");
  const size_t smem = 1024 + (size_t)(2 * 256) * tc::ROW_BYTES + (64 * 258) * sizeof(float) + 64;
  static PerDeviceOnce once;
  if (int rc = per_device_once(once, nullptr, [&](int) {
        DAGNN_CUDA_OK(cudaFuncSetAttribute(k_tc_selftest_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return (int)DAGNN_OK;
      }))
    return rc;
  k_tc_selftest_ts<<<1, 256, smem, static_cast<cudaStream_t>(stream_)>>>(W, X, C, R, N, K);
  return check_launch("k_tc_selftest_ts");
}
