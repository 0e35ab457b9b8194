// Generic GEMM on the 5th-generation tensor cores in the fp16 x 3 split (fp32-grade accuracy, tc.cuh):
//     C[M, N] (+)= A[M, K] * B[N, K]^T (+ bias[N])
// with each operand either K-contiguous (row-major [rows, K], leading dimension ld) or ROW-contiguous ([K, rows]: element
// (r, k) at base[k * ld + r] — a transposed view, which is how the weight-gradient GEMMs dW = dG^T * X and the data-gradient
// GEMMs dX = dG * W read their operands without a transpose pass). Used by the backward of the level sweep (sweep_bwd.cu) and
// by the small dense layers of the modules (heads, out_linear / hg_unify, fc1 / fc2): replaces the cuBLAS calls behind
// nn.Linear / nn.GRUCell's backward (ogbg-code/model/dagnn.py:181,209-215; dvae/dagnn.py:156,161,183).
//
// One CTA per 128 x 128 tile of C (x one slice of K), 8 warps: all of them convert the fp32 operand chunks (64 k) into fp16
// hi / lo halves in K-major SWIZZLE_128B tiles, one elected lane of warp 0 issues tcgen05.mma kind::f16 (M = 128, N = 128:
// hi*hi + lo*hi + hi*lo) into a TMEM accumulator. One 64 KB stage per CTA: up to three CTAs share an SM and overlap each
// other's conversion, MMAs and epilogue. Epilogue: TMEM -> registers -> shared memory (rotated columns, conflict-free) ->
// row-contiguous global stores.
#include "common.cuh"
#include "sync.cuh"
#include "tc.cuh"

namespace dagnn {

constexpr int kGT = 128;                                   // tile rows (M) and columns (N)
constexpr int kGStage = 4 * kGT * tc::ROW_BYTES;           // A hi, A lo, B hi, B lo tiles of one 64-k chunk = 64 KB
constexpr size_t kGSmem = 1024 + (size_t)kGStage + 64;     // ONE stage: three CTAs fit an SM (3 x 128 TMEM columns) and overlap each other

struct GemmP {
  const float* A; const float* B; float* C; const float* bias;
  long long lda, ldb, ldc;
  int M, N, K;
  int a_kmajor, b_kmajor;      // 1: [rows, K] row-major; 0: [K, rows]
  int beta;                    // 1: accumulate into C
  int a_vec, b_vec;            // K-contiguous operand may be read with 16-byte loads
  int ksplit, chunks_per;      // blockIdx.z takes 64-k chunks [z * chunks_per, (z + 1) * chunks_per); ksplit > 1: C is accumulated atomically
};

// one 64-k chunk of one operand: rows [r0, r0 + 128) x k [k0, k0 + 64) -> hi / lo tiles
__device__ __forceinline__ void stage_operand(const float* __restrict__ X, long long ld, int rows, int K, int r0, int k0, int kmajor, int vec,
                                              unsigned char* hi, unsigned char* lo, int tid) {
  if (kmajor) {
    // thread -> (row, 8-k granule): 8 consecutive threads read 256 contiguous bytes of a row
    for (int it = tid; it < kGT * 8; it += 256) {
      const int r = it >> 3, c8 = it & 7;
      const int row = r0 + r, k = k0 + 8 * c8;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (row < rows && k < K) {
        const float* src = X + (size_t)row * ld + k;
        if (vec && k + 8 <= K) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (k + j < K) ? __ldg(src + j) : 0.f;
        }
      }
      tc::store_split8(hi, lo, r, c8, v);
    }
  } else {
    // element (r, k) at X[k * ld + r]: consecutive threads take consecutive rows (coalesced for every k), 8 k each
    for (int it = tid; it < kGT * 8; it += 256) {
      const int r = it & (kGT - 1), c8 = it >> 7;
      const int row = r0 + r, k = k0 + 8 * c8;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (row < rows && k + j < K) ? __ldg(X + (size_t)(k + j) * ld + row) : 0.f;
      tc::store_split8(hi, lo, r, c8, v);
    }
  }
}

__global__ void __launch_bounds__(256, 3) k_gemm_f16x3(const __grid_constant__ GemmP P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + (size_t)kGStage);           // [0]: the MMAs issued so far are done
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 3);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(slot, 128);
  if (tid == 0) {
    mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  const int m0 = blockIdx.x * kGT, n0 = blockIdx.y * kGT;
  const int all_chunks = (P.K + 63) / 64;
  const int c_begin = blockIdx.z * P.chunks_per;
  const int nchunks = min(P.chunks_per, all_chunks - c_begin);           // >= 1 by construction of the grid
  const uint32_t idesc = tc::instr_desc_f16(128, kGT);
  unsigned char* st = base;
  for (int c = 0; c < nchunks; ++c) {
    if (c >= 1) mbar_wait(&bar[0], (uint32_t)(c - 1) & 1u);    // the MMAs that read the stage are done
    const int k0 = (c_begin + c) * 64;
    stage_operand(P.A, P.lda, P.M, P.K, m0, k0, P.a_kmajor, P.a_vec, st, st + kGT * tc::ROW_BYTES, tid);
    stage_operand(P.B, P.ldb, P.N, P.K, n0, k0, P.b_kmajor, P.b_vec, st + 2 * kGT * tc::ROW_BYTES, st + 3 * kGT * tc::ROW_BYTES, tid);
    tc::fence_async_smem();
    __syncthreads();
    if (warp == 0) {
      tc::fence_after_sync();
      const uint32_t sa = uni(smem_u32(st));
      const uint64_t ah = tc::smem_desc(sa), al = tc::smem_desc(sa + kGT * tc::ROW_BYTES);
      const uint64_t bh = tc::smem_desc(sa + 2 * kGT * tc::ROW_BYTES), bl = tc::smem_desc(sa + 3 * kGT * tc::ROW_BYTES);
      const uint32_t fresh = uni(c == 0 ? 1u : 0u);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) tc::mma3_f16(tmem, ah + 2 * ks, al + 2 * ks, bh + 2 * ks, bl + 2 * ks, idesc, fresh && ks == 0);
        tc::commit(&bar[0]);
      }
      __syncwarp();
    }
  }
  mbar_wait(&bar[0], (uint32_t)(nchunks - 1) & 1u);
  tc::fence_after_sync();
  // epilogue: TMEM lane = row of the tile (warps w and w + 4 split the columns) -> the stage memory as a [128][128] fp32 tile
  // with the columns of row r rotated by r (a warp storing one column of 32 rows, and a warp reading one row, both touch 32
  // different banks) -> every warp stores whole rows: 128 contiguous floats per row
  float* tile = reinterpret_cast<float*>(st);
  {
    const int q = warp & 3, half = warp >> 2;
    const int r = 32 * q + lane;
    for (int cb = half * 8; cb < half * 8 + 8; ++cb) {
      float v[8];
      tc::ld8(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(cb * 8), v);
      tc::wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) tile[r * kGT + ((cb * 8 + j + r) & (kGT - 1))] = v[j];
    }
  }
  __syncthreads();
  for (int r = warp; r < kGT; r += 8) {
    const int row = m0 + r;
    if (row >= P.M) break;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cl = lane + 32 * j, col = n0 + cl;
      if (col < P.N) {
        float* o = P.C + (size_t)row * P.ldc + col;
        float x = tile[r * kGT + ((cl + r) & (kGT - 1))] + ((P.bias && blockIdx.z == 0) ? __ldg(P.bias + col) : 0.f);
        if (P.ksplit > 1) atomicAdd(o, x);                     // C was zeroed by the launcher unless it accumulates anyway
        else { if (P.beta) x += *o; *o = x; }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

int gemm_f16x3(const float* A, long long lda, int a_kmajor, const float* B, long long ldb, int b_kmajor, float* C, long long ldc,
               const float* bias, int M, int N, int K, int beta, cudaStream_t st) {
  if (M <= 0 || N <= 0) return DAGNN_OK;
  DAGNN_REQUIRE(A && B && C && K > 0, "gemm: arguments");
  static PerDeviceOnce once;
  if (int rc = per_device_once(once, nullptr, [&](int) {
        DAGNN_CUDA_OK(cudaFuncSetAttribute(k_gemm_f16x3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGSmem));
        return (int)DAGNN_OK;
      }))
    return rc;
  GemmP P;
  P.A = A; P.B = B; P.C = C; P.bias = bias; P.lda = lda; P.ldb = ldb; P.ldc = ldc; P.M = M; P.N = N; P.K = K;
  P.a_kmajor = a_kmajor; P.b_kmajor = b_kmajor; P.beta = beta;
  P.a_vec = (a_kmajor && (lda & 3) == 0 && ((uintptr_t)A & 15) == 0) ? 1 : 0;
  P.b_vec = (b_kmajor && (ldb & 3) == 0 && ((uintptr_t)B & 15) == 0) ? 1 : 0;
  // few output tiles and a long contraction (weight gradients: K = number of nodes; the per-level data gradients): split K
  // over blockIdx.z so that the grid fills the SMs; partial tiles are accumulated with atomicAdd
  const int tiles = ceil_div(M, kGT) * ceil_div(N, kGT), chunks = ceil_div(K, 64);
  int ksplit = 1;
  if (tiles < 148 && chunks >= 4) ksplit = min(chunks / 2, ceil_div(296, tiles));
  P.chunks_per = ceil_div(chunks, ksplit);
  ksplit = ceil_div(chunks, P.chunks_per);
  P.ksplit = ksplit;
  if (ksplit > 1 && !beta) DAGNN_CUDA_OK(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st));
  dim3 grid((unsigned)ceil_div(M, kGT), (unsigned)ceil_div(N, kGT), (unsigned)ksplit);
  k_gemm_f16x3<<<grid, 256, kGSmem, st>>>(P);
  return check_launch("k_gemm_f16x3");
}

}  // namespace dagnn

using namespace dagnn;

// C ABI: y[M, N] = x[M, K] * w[N, K]^T + bias[N] — nn.Linear.forward (ogbg-code/model/dagnn.py:209-215 heads; dvae/dagnn.py:156,161
// hg_unify / out_linear; :183 fc1 / fc2), and the general form with transposed views for its backward.
extern "C" int dagnn_linear_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y, int64_t ldy,
                                int32_t M, int32_t N, int32_t K, void* stream) {
  return gemm_f16x3(x, ldx, 1, w, ldw, 1, y, ldy, bias, M, N, K, 0, static_cast<cudaStream_t>(stream));
}
extern "C" int dagnn_gemm_f32(const float* A, int64_t lda, int32_t a_kmajor, const float* B, int64_t ldb, int32_t b_kmajor, float* C,
                              int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t accumulate, void* stream) {
  return gemm_f16x3(A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, nullptr, M, N, K, accumulate, static_cast<cudaStream_t>(stream));
}
