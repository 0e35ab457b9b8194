// Probe of the cluster facts the cluster sweep relies on (run on the B200 box: tools/probe_cluster):
//   1. how many 16-CTA clusters (non-portable size) with ~180 KB of dynamic shared memory are co-resident;
//   2. cost of barrier.cluster (arrive.release + wait.acquire) in a 16-CTA cluster;
//   3. the exchange pattern of a level: every CTA stores its slice to GLOBAL memory (generic proxy), fence.proxy.async,
//      cluster barrier, then every CTA bulk-copies (cp.async.bulk, async proxy) the slices of all CTAs into shared memory
//      and checks the values — latency of the round trip and correctness of the visibility chain.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_cluster tools/probe_cluster.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cluster_sync_() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }

// slice_bytes per CTA per round; buf [clusters][2][16 * slice_bytes]
__global__ void __launch_bounds__(256, 1) k_probe(unsigned char* buf, int slice_bytes, int iters, long long* out, int* errs, int* smids) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base);
  unsigned char* stage = base + 1024;
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_rank(), cid = cluster_id();
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    smids[blockIdx.x] = (int)smid();
  }
  __syncthreads();
  cluster_sync_();
  // (2) bare barrier cost
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) cluster_sync_();
  long long t1 = clock64();
  // (3) exchange through global memory + bulk copy
  const int total = 16 * slice_bytes;
  int bad = 0;
  long long t2 = clock64();
  for (int i = 0; i < iters; ++i) {
    unsigned char* region = buf + ((size_t)cid * 2 + (i & 1)) * total;
    uint32_t* mine = reinterpret_cast<uint32_t*>(region + (size_t)rank * slice_bytes);
    for (int w = tid; w < slice_bytes / 4; w += 256) mine[w] = (uint32_t)(i * 1000003 + rank * 4099 + w);
    asm volatile("fence.proxy.async;" ::: "memory");
    cluster_sync_();
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(total) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(stage)),
                   "l"(region), "r"(total), "r"(smem_u32(bar))
                   : "memory");
    }
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"((uint32_t)(i & 1)) : "memory");
    } while (!done);
    const uint32_t* st = reinterpret_cast<const uint32_t*>(stage);
    for (int w = tid; w < total / 4; w += 256) {
      const int r = w / (slice_bytes / 4), k = w % (slice_bytes / 4);
      if (st[w] != (uint32_t)(i * 1000003 + r * 4099 + k)) ++bad;
    }
    __syncthreads();
  }
  long long t3 = clock64();
  if (bad) atomicAdd(errs, bad);
  if (tid == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t3 - t2; }
  cluster_sync_();
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  printf("SMs %d\n", sms);
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (int smem_kb : {180}) {
    for (int cs : {8, 10, 12, 13, 14, 15, 16}) {
      CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 32); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = (size_t)smem_kb * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nc = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, k_probe, &cfg);
      printf("smem %3d KB cluster %2d: max active clusters %d (%s)\n", smem_kb, cs, nc, cudaGetErrorString(e));
    }
  }
  const int cs = 16, smem_kb = 180, iters = 2000;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024));
  for (int nclusters : {1, 8}) {
    for (int slice : {256, 2560, 8192}) {
      unsigned char* buf; long long* out; int* errs; int* smids;
      CK(cudaMalloc(&buf, (size_t)nclusters * 2 * 16 * slice));
      CK(cudaMalloc(&out, sizeof(long long) * 2 * cs * nclusters));
      CK(cudaMalloc(&errs, sizeof(int))); CK(cudaMemset(errs, 0, sizeof(int)));
      CK(cudaMalloc(&smids, sizeof(int) * cs * nclusters));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * nclusters); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = (size_t)smem_kb * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, k_probe, buf, slice, iters, out, errs, smids));
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(2 * cs * nclusters);
      std::vector<int> hs(cs * nclusters);
      int herr = 0;
      CK(cudaMemcpy(h.data(), out, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hs.data(), smids, sizeof(int) * hs.size(), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&herr, errs, sizeof(int), cudaMemcpyDeviceToHost));
      printf("clusters %d slice %5d B: barrier %.0f cyc, exchange round (store %d B + fence + barrier + bulk %d B + check) %.0f cyc, errors %d\n",
             nclusters, slice, (double)h[0] / iters, slice, 16 * slice, (double)h[1] / iters, herr);
      if (slice == 256) { printf("  smids of cluster 0:"); for (int i = 0; i < cs; ++i) printf(" %d", hs[i]); printf("\n"); }
      cudaFree(buf); cudaFree(out); cudaFree(errs); cudaFree(smids);
    }
  }
  return 0;
}
