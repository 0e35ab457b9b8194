// Input side: the D-VAE text-row decoders and their collation on the device.
// Replaces, per batch, decode_ENAS_to_pygraph / decode_BN_to_pygraph (dvae/util.py:343-385 / :290-339: adjacency fill, one-hot
// node types, edge list in the row-major order of the adjacency's non-zeros, add_order_info's level arrays) and the collation
// of dvae/batch.py:26-145 (edge_index and node ids offset by the running node count, levels not) — one thread per graph
// (graphs have <= 32 nodes), an exclusive scan of the edge counts in between. Integer work: bit-exact.
#include "common.cuh"

namespace dagnn {

constexpr int kRowsMaxNodes = 32;

// adjacency of graph g as one bit mask per source node (bit v of adj[u] = edge u -> v); returns the number of nodes
__device__ __forceinline__ int row_adjacency(const int32_t* __restrict__ rows, int g, int n, int kind, uint32_t (&adj)[kRowsMaxNodes]) {
  const int nn = n + 2;
  for (int u = 0; u < nn; ++u) adj[u] = 0u;
  const int32_t* r = rows + (size_t)g * n * n;
  if (kind == 0) {                                       // ENAS: node i + 1 hangs off node i, plus one edge per set flag j -> i + 1
    for (int i = 0; i < n; ++i) {
      adj[i] |= 1u << (i + 1);
      for (int j = 0; j < i; ++j)
        if (r[i * n + 1 + j] == 1) adj[j] |= 1u << (i + 1);
    }
    adj[n] |= 1u << (n + 1);
  } else {                                               // BN: parent-less variables hang off the start node, childless ones feed the end node
    uint32_t has_child = 0u;
    for (int i = 0; i < n; ++i) {
      int s = 0;
      for (int j = 0; j < i; ++j) s += r[i * n + 1 + j];
      if (s == 0) adj[0] |= 1u << (i + 1);
      else
        for (int j = 0; j < i; ++j)
          if (r[i * n + 1 + j] == 1) { adj[j + 1] |= 1u << (i + 1); has_child |= 1u << j; }
    }
    for (int j = 0; j < n; ++j)
      if (!(has_child >> j & 1u)) adj[j + 1] |= 1u << (n + 1);
  }
  return nn;
}

__global__ void k_rows_count(const int32_t* __restrict__ rows, int B, int n, int kind, int* __restrict__ ecount) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B) return;
  uint32_t adj[kRowsMaxNodes];
  const int nn = row_adjacency(rows, g, n, kind, adj);
  int e = 0;
  for (int u = 0; u < nn; ++u) e += __popc(adj[u]);
  ecount[g] = e;
}
// single block: exclusive scan of ecount -> eoff[B + 1]
__global__ void __launch_bounds__(1024) k_rows_scan(const int* __restrict__ ecount, int B, int* __restrict__ eoff) {
  __shared__ int carry;
  __shared__ int wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int i = base + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int v = i < B ? ecount[i] : 0;
    const int incl = warp_incl_scan(v, lane);
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int w = wsum[lane];
      wsum[lane] = warp_incl_scan(w, lane) - w;
    }
    __syncthreads();
    const int excl = carry + wsum[wid] + incl - v;
    if (i < B) eoff[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) eoff[B] = carry;
}
__global__ void k_rows_fill(const int32_t* __restrict__ rows, int B, int n, int kind, int nvt, const int* __restrict__ eoff, float* __restrict__ x,
                            int64_t* __restrict__ edge_index, int64_t ecap, int64_t* __restrict__ bi, int64_t* __restrict__ batch,
                            int* __restrict__ status) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B) return;
  uint32_t adj[kRowsMaxNodes];
  const int nn = row_adjacency(rows, g, n, kind, adj);
  const int64_t N = (int64_t)B * nn, v0 = (int64_t)g * nn;
  // node types -> one-hot rows (start = 0, end = 1, variables = raw type + 2)
  for (int v = 0; v < nn; ++v) {
    const int t = v == 0 ? 0 : (v == nn - 1 ? 1 : rows[(size_t)g * n * n + (size_t)(v - 1) * n] + 2);
    if (t < 0 || t >= nvt) { *status = 2; continue; }
    float* xr = x + (v0 + v) * nvt;
    for (int c = 0; c < nvt; ++c) xr[c] = c == t ? 1.f : 0.f;
  }
  // edges in the row-major order of the adjacency's non-zeros (nx.DiGraph(adj).edges, dvae/util.py:321-330)
  int64_t e = eoff[g];
  for (int u = 0; u < nn; ++u) {
    uint32_t m = adj[u];
    while (m) {
      const int v = __ffs(m) - 1;
      m &= m - 1;
      if (e < ecap) { edge_index[e] = v0 + u; edge_index[ecap + e] = v0 + v; }
      ++e;
    }
  }
  // longest-path levels (src/utils_dag.py:8-35 on a graph whose node order is topological), both directions
  int l0[kRowsMaxNodes], l1[kRowsMaxNodes];
  for (int v = 0; v < nn; ++v) {
    int l = 0;
    for (int u = 0; u < v; ++u)
      if (adj[u] >> v & 1u) l = max(l, l0[u] + 1);
    l0[v] = l;
  }
  for (int u = nn - 1; u >= 0; --u) {
    int l = 0;
    for (int v = u + 1; v < nn; ++v)
      if (adj[u] >> v & 1u) l = max(l, l1[v] + 1);
    l1[u] = l;
  }
  for (int v = 0; v < nn; ++v) {            // bi_layer_index int64 [2][2][N]: [d][0] = level, [d][1] = node id (offset, dvae/batch.py:54-59)
    bi[0 * N + v0 + v] = l0[v];
    bi[1 * N + v0 + v] = v0 + v;
    bi[2 * N + v0 + v] = l1[v];
    bi[3 * N + v0 + v] = v0 + v;
    batch[v0 + v] = g;
  }
}

}  // namespace dagnn

using namespace dagnn;

extern "C" size_t dagnn_dvae_rows_workspace_bytes(int64_t B) { return B < 0 ? 0 : (size_t)(2 * B + 2) * sizeof(int); }

extern "C" int dagnn_dvae_rows_build(const int32_t* rows, int64_t B, int32_t n, int32_t kind, int32_t nvt, float* x, int64_t* edge_index, int64_t ecap,
                                     int64_t* bi_layer_index, int64_t* batch, int32_t* counts, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  DAGNN_REQUIRE(rows && x && edge_index && bi_layer_index && batch && counts && workspace, "rows_build: null pointer");
  DAGNN_REQUIRE(B > 0 && B < (1ll << 24) && n >= 1 && n + 2 <= kRowsMaxNodes && (kind == 0 || kind == 1) && nvt >= 3, "rows_build: sizes");
  DAGNN_REQUIRE(ecap >= 0 && workspace_bytes >= dagnn_dvae_rows_workspace_bytes(B), "rows_build: capacity / workspace");
  int* ecount = static_cast<int*>(workspace);
  int* eoff = ecount + B;
  const int blocks = (int)((B + 127) / 128);
  DAGNN_CUDA_OK(cudaMemsetAsync(counts, 0, 4 * sizeof(int32_t), st));
  k_rows_count<<<blocks, 128, 0, st>>>(rows, (int)B, n, kind, ecount);
  if (int rc = check_launch("k_rows_count")) return rc;
  k_rows_scan<<<1, 1024, 0, st>>>(ecount, (int)B, eoff);
  if (int rc = check_launch("k_rows_scan")) return rc;
  k_rows_fill<<<blocks, 128, 0, st>>>(rows, (int)B, n, kind, nvt, eoff, x, edge_index, ecap, bi_layer_index, batch, counts + 1);
  if (int rc = check_launch("k_rows_fill")) return rc;
  DAGNN_CUDA_OK(cudaMemcpyAsync(counts, eoff + B, sizeof(int), cudaMemcpyDeviceToDevice, st));     // counts[0] = number of edges
  return DAGNN_OK;
}
