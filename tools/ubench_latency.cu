// Micro-benchmark (tooling, not product): dependent ld.global.cg latency inside a persistent 148-CTA kernel, alone and
// next to the kinds of spinning the sweep kernel does (mbarrier try_wait in the same CTA, ld.acquire polling from other
// CTAs). Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ub tools/ubench_latency.cu && /tmp/ub
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include <numeric>
#include <random>
#include <algorithm>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode bit0: other warps of the CTA spin on an mbarrier; bit1: CTAs >= 74 poll a global flag with ld.acquire instead
__global__ void __launch_bounds__(288, 1) k_chase(const uint32_t* __restrict__ next, int n, int hops, int mode, unsigned int* flag,
                                                 long long* out) {
  __shared__ uint64_t bar;
  __shared__ volatile int done;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    done = 0;
  }
  __syncthreads();
  const bool poller = (mode & 2) && blockIdx.x >= 74;
  if (threadIdx.x == 0) {
    if (poller) {
      unsigned int v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); } while (v < 74u);
      out[blockIdx.x] = 0;
    } else {
      uint32_t idx = (blockIdx.x * 7919u) % (uint32_t)n;
      // warm: none. timed chain
      const long long t0 = clock64();
      for (int h = 0; h < hops; ++h) idx = __ldcg(next + idx);
      const long long t1 = clock64();
      out[blockIdx.x] = (t1 - t0) / hops + (idx == 0xffffffffu);
      if (mode & 2) atomicAdd(flag, 1u);
    }
    done = 1;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
  } else if (mode & 1) {
    uint32_t ok;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    } while (!ok);
  }
}

__global__ void k_touch(uint32_t* next, const uint32_t* src, int n) {   // device-side write so the lines are L2 resident & dirty
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) next[i] = src[i];
}

int main() {
  for (int mb : {8, 64, 512}) {
    const int n = mb * 1024 * 1024 / 4;
    std::vector<uint32_t> perm(n), nxt(n);
    std::iota(perm.begin(), perm.end(), 0u);
    std::mt19937 rng(1);
    std::shuffle(perm.begin(), perm.end(), rng);
    for (int i = 0; i < n; ++i) nxt[perm[i]] = perm[(i + 1) % n];
    uint32_t *d_src, *d_next; unsigned int* d_flag; long long* d_out;
    cudaMalloc(&d_src, n * 4); cudaMalloc(&d_next, n * 4); cudaMalloc(&d_flag, 4); cudaMalloc(&d_out, 148 * 8);
    cudaMemcpy(d_src, nxt.data(), n * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 4; ++mode) {
      k_touch<<<148 * 4, 256>>>(d_next, d_src, n);
      cudaMemset(d_flag, 0, 4);
      k_chase<<<148, 288>>>(d_next, n, 256, mode, d_flag, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mn = 1 << 30, mx = 0, sum = 0; int cnt = 0;
      for (int i = 0; i < 148; ++i) if (h[i] > 0) { mn = std::min(mn, h[i]); mx = std::max(mx, h[i]); sum += h[i]; ++cnt; }
      printf("buffer %4d MB mode %d (%s%s): cycles/hop min %lld avg %lld max %lld over %d CTAs  [%s]\n", mb, mode,
             (mode & 1) ? "mbarrier-spin " : "", (mode & 2) ? "acquire-pollers" : "", mn, cnt ? sum / cnt : 0, mx, cnt, cudaGetErrorString(e));
    }
    cudaFree(d_src); cudaFree(d_next); cudaFree(d_flag); cudaFree(d_out);
  }
  return 0;
}
