// Stand-alone check of the tcgen05 building blocks in tc.cuh: C[M,N] = A[M,K] * B[N,K]^T in 3xTF32 with the same
// operand tile layout, descriptors, MMA issue and TMEM read-back the level kernel uses. One CTA per 128 rows.
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

template <bool F16>
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
  const uint32_t idesc = F16 ? tc::instr_desc_f16(128, N) : tc::instr_desc_tf32(128, N);
  const int KCH = F16 ? tc::KC16 : tc::KC;
  const int nchunks = K / KCH;
  for (int c = 0; c < nchunks; ++c) {
    for (int it = tid; it < 128 * 8; it += 256) {
      const int r = it >> 3, c4 = it & 7;
      if (F16) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (row0 + r < M) {
          const float4 a = *reinterpret_cast<const float4*>(A + (size_t)(row0 + r) * K + c * KCH + 8 * c4);
          const float4 b = *reinterpret_cast<const float4*>(A + (size_t)(row0 + r) * K + c * KCH + 8 * c4 + 4);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
        tc::store_split8(A_hi, A_lo, r, c4, v);
      } else {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < M) v = *reinterpret_cast<const float4*>(A + (size_t)(row0 + r) * K + c * KCH + 4 * c4);
        tc::store_split(A_hi, A_lo, r, c4, v);
      }
    }
    for (int it = tid; it < N * 8; it += 256) {
      const int r = it >> 3, c4 = it & 7;
      if (F16) {
        const float4 a = *reinterpret_cast<const float4*>(B + (size_t)r * K + c * KCH + 8 * c4);
        const float4 b = *reinterpret_cast<const float4*>(B + (size_t)r * K + c * KCH + 8 * c4 + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        tc::store_split8(B_hi, B_lo, r, c4, v);
      } else {
        const float4 v = *reinterpret_cast<const float4*>(B + (size_t)r * K + c * KCH + 4 * c4);
        tc::store_split(B_hi, B_lo, r, c4, v);
      }
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint64_t ah = tc::smem_desc(tc::smem_addr(A_hi)), al = tc::smem_desc(tc::smem_addr(A_lo));
      const uint64_t bh = tc::smem_desc(tc::smem_addr(B_hi)), bl = tc::smem_desc(tc::smem_addr(B_lo));
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (F16) tc::mma3_f16(tmem, ah + 2 * ks, al + 2 * ks, bh + 2 * ks, bl + 2 * ks, idesc, c == 0 && ks == 0);
        else tc::mma3(tmem, ah + 2 * ks, al + 2 * ks, bh + 2 * ks, bl + 2 * ks, idesc, c == 0 && ks == 0);
      }
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

}  // namespace dagnn

using namespace dagnn;

template <bool F16>
static int tc_selftest(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, void* stream_) {
  DAGNN_REQUIRE(A && B && C, "tc_selftest: null pointer");
  const int kc = F16 ? tc::KC16 : tc::KC;
  DAGNN_REQUIRE(M > 0 && N >= 16 && N <= 256 && N % 16 == 0 && K >= kc && K % kc == 0, "tc_selftest: M>0, N%16==0 in [16,256], K%32==0 (tf32) / K%64==0 (f16)");
  DAGNN_REQUIRE(((((uintptr_t)A) | ((uintptr_t)B) | ((uintptr_t)C)) & 15) == 0, "tc_selftest: 16-byte alignment");
  const size_t smem = 1024 + (size_t)(2 * 128 + 2 * 256) * tc::ROW_BYTES + 64;
  static bool configured = false;
  if (!configured) {
    DAGNN_CUDA_OK(cudaFuncSetAttribute(k_tc_selftest<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  k_tc_selftest<F16><<<(M + 127) / 128, 256, smem, static_cast<cudaStream_t>(stream_)>>>(A, B, C, M, N, K);
  return check_launch("k_tc_selftest");
}

extern "C" int dagnn_tc_selftest_f32(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, void* stream) {
  return tc_selftest<false>(A, B, C, M, N, K, stream);
}
extern "C" int dagnn_tc_selftest_f16x3(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, void* stream) {
  return tc_selftest<true>(A, B, C, M, N, K, stream);
}
