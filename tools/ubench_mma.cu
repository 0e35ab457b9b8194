// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, both operands in shared memory) as a function of
// N, issued back to back by one lane of a converged warp with warp-uniform operands, one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I dagnn_b200/csrc -o /tmp/ubench_mma tools/ubench_mma.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
#include "tc.cuh"
using namespace dagnn;

__device__ __forceinline__ void mb_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(tc::smem_addr(bar)), "r"(parity) : "memory");
  } while (!done);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_f16_pred(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t on) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(on)
      : "memory");
}
// mode 0: one accumulator, operands at fixed addresses; 1: accumulators rotate over 4 column ranges; 2: A walks over 8 tiles
__global__ void __launch_bounds__(128, 1) k(int N, int reps, int mode, long long* out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 8 * 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < (8 * 16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;   // halves = 1.0
  if (warp == 0) tc::tmem_alloc(slot, 512);
  if (tid == 0) { mb_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    const uint32_t idesc = __shfl_sync(0xffffffffu, tc::instr_desc_f16(128, N), 0);
    const uint32_t sa = __shfl_sync(0xffffffffu, tc::smem_addr(base), 0);
    const uint32_t sb = __shfl_sync(0xffffffffu, tc::smem_addr(base + 8 * 16384), 0);
    for (int round = 0; round < 3; ++round) {
      long long t0 = clock64();
      // the whole warp runs the loop (operands stay warp-uniform), lane 0 issues; mode 3: 12 MMAs per iteration, unrolled
      if (mode == 4) {
        for (int r = 0; r < reps; r += 12) {
          const uint64_t ad = tc::smem_desc(sa), bd = tc::smem_desc(sb);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc::mma3_f16(tmem, ad + 2 * ks, ad + 2 * ks + 1024, bd + 2 * ks, bd + 2 * ks + 64, idesc, r == 0 && ks == 0);
          }
        }
      } else if (mode == 5) {
        const uint32_t on = lane == 0 ? 1u : 0u;
        for (int r = 0; r < reps; r += 12) {
          const uint64_t ad = tc::smem_desc(sa), bd = tc::smem_desc(sb);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            mma_f16_pred(tmem, ad + 2 * ks, bd + 2 * ks, idesc, (r == 0 && ks == 0) ? 0u : 1u, on);
            mma_f16_pred(tmem, ad + 2 * ks + 1024, bd + 2 * ks, idesc, 1u, on);
            mma_f16_pred(tmem, ad + 2 * ks, bd + 2 * ks + 64, idesc, 1u, on);
          }
        }
      } else if (mode == 3) {
        for (int r = 0; r < reps; r += 12) {
          const uint64_t ad = tc::smem_desc(sa), bd = tc::smem_desc(sb);
          if (lane == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc::mma3_f16(tmem, ad + 2 * ks, ad + 2 * ks + 1024, bd + 2 * ks, bd + 2 * ks + 64, idesc, r == 0 && ks == 0);
          }
        }
      } else {
        for (int r = 0; r < reps; ++r) {
          const uint32_t tm = tmem + (mode == 1 ? (uint32_t)((r & 3) * 64) : 0u);
          const uint64_t ad = tc::smem_desc(sa + (mode == 2 ? (uint32_t)((r & 7) * 16384) : 0u)) + 2 * (r & 3);
          const uint64_t bd = tc::smem_desc(sb) + 2 * (r & 3);
          if (lane == 0) tc::mma_f16(tm, ad, bd, idesc, r >= 4 ? 1u : 0u);
        }
      }
      if (lane == 0) tc::commit(bar);
      __syncwarp();
      long long t1 = clock64();
      mb_wait(bar, (uint32_t)(round & 1));
      long long t2 = clock64();
      if (lane == 0 && blockIdx.x == 0 && round == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  const size_t smem = 1024 + 8 * 16384 + 32768 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int reps = 96;
  for (int grid : {148})
    for (int mode = 3; mode < 6; ++mode)
      for (int N : {16, 32, 64, 128, 256}) {
        if (N > 128) continue;
        k<<<grid, 128, smem>>>(N, reps, mode, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("grid %3d mode %d N %3d: issue %6.1f cyc/mma, complete %6.1f cyc/mma\n", grid, mode, N, (double)out[0] / reps,
               (double)out[1] / reps);
      }
  return 0;
}
