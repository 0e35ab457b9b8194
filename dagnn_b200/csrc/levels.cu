// Longest-path level of every node of a batch of DAGs, on the edges as given and on the reversed edges: what the reference
// computes per graph on the host when a data set is built (`top_sort` src/utils_dag.py:8-35 — frontier peeling — and
// `add_order_info_01` :39-52) and stores as `_bi_layer_idx0/1`. Integer work, bit-exact by construction: the level of v
// is the fixed point of lvl[v] = max(0, max over edges u -> v of lvl[u] + 1).
//
// Chaotic relaxation: every pass visits every edge once (both directions in the same visit) and raises levels with
// atomicMax; values only grow and are bounded by the true level, so the order of updates inside a pass does not matter
// and a pass often moves more than one level. A pass that changed nothing proves the fixed point; it clears its flag and
// the remaining passes (launched up front, no host round trip) return at once. A batch deeper than the number of passes —
// or a cycle — leaves the last flag set: status 1, the caller retries with more passes.
#include "common.cuh"

namespace dagnn {

__global__ void __launch_bounds__(256) k_levels_init(int* __restrict__ lf, int* __restrict__ lb, int* __restrict__ flags, int N,
                                                    int nflags, int* __restrict__ summary) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;
  for (int v = i; v < N; v += stride) { lf[v] = 0; lb[v] = 0; }
  for (int f = i; f < nflags; f += stride) flags[f] = 0;
  if (i < 4) summary[i] = 0;
}

// flags[p] = pass p raised a level. Pass p runs iff pass p - 1 did (pass 0 always).
__global__ void __launch_bounds__(256) k_levels_relax(const int64_t* __restrict__ ei, int64_t E, int N, int* lf, int* lb,
                                                     int* flags, int pass, int* summary) {
  if (pass > 0 && flags[pass - 1] == 0) return;
  int changed = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = ei[e], v = ei[E + e];
    if (u < 0 || u >= N || v < 0 || v >= N) { summary[1] = 2; continue; }      // edge endpoint out of range
    const int a = __ldcg(lf + u) + 1;               // forward: the target sits below its source
    if (a > __ldcg(lf + v)) { atomicMax(lf + v, a); changed = 1; }
    const int b = __ldcg(lb + v) + 1;               // reversed edges: the source sits below its target
    if (b > __ldcg(lb + u)) { atomicMax(lb + u, b); changed = 1; }
  }
  if (__syncthreads_or(changed) && threadIdx.x == 0) flags[pass] = 1;
}

__global__ void __launch_bounds__(256) k_levels_finish(const int* __restrict__ lf, const int* __restrict__ lb, int N,
                                                      int64_t* __restrict__ out_f, int64_t* __restrict__ out_b,
                                                      const int* __restrict__ flags, int passes, int* summary) {
  __shared__ int smax[2];
  if (threadIdx.x < 2) smax[threadIdx.x] = 0;
  __syncthreads();
  int mf = 0, mb = 0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
    const int a = lf[v], b = lb[v];
    out_f[v] = a; out_b[v] = b;
    mf = max(mf, a); mb = max(mb, b);
  }
  atomicMax(&smax[0], mf); atomicMax(&smax[1], mb);
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicMax(&summary[0], smax[0] + 1);            // number of levels, forward
    atomicMax(&summary[2], smax[1] + 1);            // ... on the reversed edges (equal for a DAG)
    if (blockIdx.x == 0 && flags[passes - 1] != 0) atomicMax(&summary[1], 1);   // the last pass still raised a level
  }
}

}  // namespace dagnn

using namespace dagnn;

extern "C" size_t dagnn_levels_workspace_bytes(int64_t N, int32_t max_passes) {
  if (N < 0 || max_passes < 1) return 0;
  return (size_t)round_up64(4 * (2 * N + (int64_t)max_passes + 8), 256);
}

extern "C" int dagnn_levels_build(const int64_t* edge_index, int64_t N, int64_t E, int32_t max_passes, int64_t* lvl_fwd,
                                  int64_t* lvl_bwd, int32_t* summary, void* workspace, size_t workspace_bytes, void* stream_) {
  DAGNN_REQUIRE(lvl_fwd && lvl_bwd && summary && workspace, "levels: null pointer");
  DAGNN_REQUIRE(N > 0 && N < (1ll << 31) && E >= 0 && (E == 0 || edge_index), "levels: sizes");
  DAGNN_REQUIRE(max_passes >= 1 && max_passes <= (1 << 22), "levels: max_passes");
  DAGNN_REQUIRE(workspace_bytes >= dagnn_levels_workspace_bytes(N, max_passes), "levels: workspace too small");
  DAGNN_REQUIRE((((uintptr_t)workspace) & 3) == 0, "levels: workspace alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  int* lf = static_cast<int*>(workspace);
  int* lb = lf + N;
  int* flags = lb + N;
  const int nb = (int)((N + 255) / 256 < 148 * 8 ? (N + 255) / 256 : 148 * 8);
  k_levels_init<<<nb, 256, 0, st>>>(lf, lb, flags, (int)N, max_passes, summary);
  int rc = check_launch("k_levels_init");
  if (rc) return rc;
  if (E > 0) {
    const int eb = (int)((E + 255) / 256 < 148 * 8 ? (E + 255) / 256 : 148 * 8);
    for (int p = 0; p < max_passes; ++p) {
      k_levels_relax<<<eb, 256, 0, st>>>(edge_index, E, (int)N, lf, lb, flags, p, summary);
      rc = check_launch("k_levels_relax");
      if (rc) return rc;
    }
  }
  k_levels_finish<<<nb, 256, 0, st>>>(lf, lb, (int)N, lvl_fwd, lvl_bwd, flags, max_passes, summary);
  return check_launch("k_levels_finish");
}
