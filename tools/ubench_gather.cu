// Micro-benchmark (tooling, not product): throughput of the operand-build access pattern of the sweep kernel —
// 256 threads per CTA, thread (r0 = tid>>3, c8 = tid&7) loads 8 consecutive floats (2 x ld.global.cg.v4) of NR random rows of a
// [rows x 256] fp32 matrix per item — as a function of the number of active CTAs and of the software-pipelining depth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ug tools/ubench_gather.cu && /tmp/ug
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include <random>

template <int NR, int DEPTH>
__global__ void __launch_bounds__(256, 1) k_gather(const float* __restrict__ M, const int* __restrict__ rows, int nrows_total, int items,
                                                   long long* cyc, float* sink) {
  const int tid = threadIdx.x, r0 = tid >> 3, c8 = tid & 7;
  float4 buf[DEPTH][NR][2];
  float acc = 0.f;
  auto issue = [&](int it, int slot) {
#pragma unroll
    for (int x = 0; x < NR; ++x) {
      const int row = rows[(blockIdx.x * 977 + it * 128 + r0 + 32 * x) % nrows_total];
      const float* p = M + (size_t)row * 256 + ((it & 3) * 64 + 8 * c8);
      buf[slot][x][0] = __ldcg(reinterpret_cast<const float4*>(p));
      buf[slot][x][1] = __ldcg(reinterpret_cast<const float4*>(p + 4));
    }
  };
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll
  for (int d = 0; d < DEPTH - 1; ++d) issue(d, d);
#pragma unroll 1
  for (int it = 0; it < items; it += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      issue(it + d + DEPTH - 1, (d + DEPTH - 1) % DEPTH);
#pragma unroll
      for (int x = 0; x < NR; ++x) acc += buf[d][x][0].x + buf[d][x][0].w + buf[d][x][1].y + buf[d][x][1].z;
    }
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = (t1 - t0) / items;
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  const int R = 16384 * 4;   // 64 MB matrix: L2 resident
  float* M; int* rows; long long* cyc; float* sink;
  cudaMalloc(&M, (size_t)R * 256 * 4); cudaMalloc(&rows, R * 4); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  cudaMemset(M, 0, (size_t)R * 256 * 4);
  std::vector<int> h(R); std::mt19937 rng(1); for (int i = 0; i < R; ++i) h[i] = rng() % R;
  cudaMemcpy(rows, h.data(), R * 4, cudaMemcpyHostToDevice);
  auto run = [&](auto kern, const char* name, int grid) {
    kern<<<grid, 256>>>(M, rows, R, 256, cyc, sink);   // warm
    kern<<<grid, 256>>>(M, rows, R, 256, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long hc[148]; cudaMemcpy(hc, cyc, grid * 8, cudaMemcpyDeviceToHost);
    long long s = 0; for (int i = 0; i < grid; ++i) s += hc[i];
    printf("%-22s grid %3d: %6lld cycles/item  [%s]\n", name, grid, s / grid, cudaGetErrorString(e));
  };
  for (int grid : {1, 16, 64, 148}) {
    run(k_gather<4, 1>, "4 rows, depth 1", grid);   // 32 KB per item, no lookahead
    run(k_gather<4, 2>, "4 rows, depth 2", grid);
    run(k_gather<4, 4>, "4 rows, depth 4", grid);
    run(k_gather<1, 4>, "1 row, depth 4", grid);    // 8 KB per item
  }
  return 0;
}
