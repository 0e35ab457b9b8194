// Integer pre-pass: level-sorted node order + in-edge CSR per direction (bit-exact part of the path).
// Replaces ogbg-code/model/dagnn.py:130-137,146-147,151-157 (see include/dagnn_b200.h).
//
// Node order: stable counting sort of the level array. A chunk of 512 consecutive entries is owned by one
// warp; per (level, chunk) counts are laid out level-major so that ONE exclusive scan over the whole table
// yields, for every (level, chunk), the first position of that chunk's nodes of that level — stability
// (ascending index inside a level, the order of the reference's boolean-mask select) falls out of the layout.
// Inside a chunk the warp walks 32 entries at a time and ranks equal levels with __match_any_sync.
//
// CSR: degree histogram -> exclusive scan -> atomic fill -> per-row rank sort by edge id (rows are short),
// so the edge order inside a row is ascending edge position, exactly the reference's `le_idx` concatenation.
#include "common.cuh"

namespace dagnn {

constexpr int kChunkBig = 512;        // entries of the level arrays per warp: 512 for big batches (a smaller count table),
constexpr int kChunkSmall = 128;      // 128 below 64 k nodes (4 x the blocks: the two counting-sort passes are latency-bound)
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(32) k_sched_count(const int64_t* __restrict__ lvl0, const int64_t* __restrict__ lvl1,
                                                    int N, int nchunks, int kChunk, int max_levels, int* __restrict__ cnt0,
                                                    int* __restrict__ cnt1, int* __restrict__ summary) {
  const int d = blockIdx.y, b = blockIdx.x, lane = threadIdx.x;
  const int64_t* lvl = d ? lvl1 : lvl0;
  int* cnt = d ? cnt1 : cnt0;
  int mx = 0;
  for (int g = 0; g < kChunk / 32; ++g) {
    const int k = b * kChunk + g * 32 + lane;
    bool act = k < N;
    long long l = act ? lvl[k] : 0;
    if (act && (l < 0 || l >= max_levels)) { summary[2] = 1; act = false; }
    const unsigned am = __ballot_sync(0xffffffffu, act);
    if (act) {
      const unsigned peers = __match_any_sync(am, (int)l);
      if ((peers & ((1u << lane) - 1u)) == 0) cnt[(size_t)l * nchunks + b] += __popc(peers);
      mx = max(mx, (int)l + 1);
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0 && mx > 0) atomicMax(&summary[d], mx);
}

// exclusive scan of `n` ints by one block (in -> out, may alias); returns the total in *total_out (optional)
__device__ void block_excl_scan(const int* in, int* out, int n, int* smem /*[33]*/, int* total_out) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per = (n + kScanThreads - 1) / kScanThreads;
  const int beg = min(n, tid * per), end = min(n, beg + per);
  int s = 0;
  for (int i = beg; i < end; ++i) s += in[i];
  int incl = warp_incl_scan(s, lane);
  if (lane == 31) smem[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = smem[lane];
    int wi = warp_incl_scan(w, lane);
    smem[lane] = wi - w;
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  int run = smem[wid] + incl - s;
  for (int i = beg; i < end; ++i) {
    const int v = in[i];
    out[i] = run;
    run += v;
  }
  if (total_out && tid == 0) *total_out = smem[32];
  __syncthreads();
}

__global__ void __launch_bounds__(kScanThreads) k_sched_scan_levels(int* cnt0, int* cnt1, int nchunks, int max_levels, int N,
                                                                    int* lvl_off0, int* lvl_off1, const int* summary) {
  __shared__ int smem[33];
  const int d = blockIdx.x;
  int* cnt = d ? cnt1 : cnt0;
  int* lvl_off = d ? lvl_off1 : lvl_off0;
  const int L = min(summary[d], max_levels);
  block_excl_scan(cnt, cnt, L * nchunks, smem, nullptr);
  for (int l = threadIdx.x; l <= max_levels; l += blockDim.x) lvl_off[l] = (l < L) ? cnt[(size_t)l * nchunks] : N;
}

__global__ void __launch_bounds__(32) k_sched_place(const int64_t* __restrict__ lvl0, const int64_t* __restrict__ lvl1,
                                                    const int64_t* __restrict__ nid0, const int64_t* __restrict__ nid1, int N,
                                                    int nchunks, int kChunk, int max_levels, int* __restrict__ cnt0, int* __restrict__ cnt1,
                                                    int* __restrict__ perm0, int* __restrict__ perm1, int* __restrict__ pos0,
                                                    int* __restrict__ pos1, int* __restrict__ summary) {
  const int d = blockIdx.y, b = blockIdx.x, lane = threadIdx.x;
  const int64_t* lvl = d ? lvl1 : lvl0;
  const int64_t* nid = d ? nid1 : nid0;
  int* cnt = d ? cnt1 : cnt0;
  int* perm = d ? perm1 : perm0;
  int* pos = d ? pos1 : pos0;
  for (int g = 0; g < kChunk / 32; ++g) {
    const int k = b * kChunk + g * 32 + lane;
    bool act = k < N;
    long long l = act ? lvl[k] : 0;
    if (act && (l < 0 || l >= max_levels)) act = false;
    const unsigned am = __ballot_sync(0xffffffffu, act);
    if (act) {
      const unsigned peers = __match_any_sync(am, (int)l);
      const int rank = __popc(peers & ((1u << lane) - 1u));
      int base = 0;
      if (rank == 0) {
        int* c = &cnt[(size_t)l * nchunks + b];
        base = *c;
        *c = base + __popc(peers);
      }
      base = __shfl_sync(peers, base, __ffs(peers) - 1);
      const long long node = nid ? nid[k] : (long long)k;
      if (node != k) summary[3] = 1;        // positions inside a level are not in node-id order (the cluster sweep cuts groups by it)
      if (node < 0 || node >= N) {
        summary[2] = 2;
      } else {
        perm[base + rank] = (int)node;
        pos[node] = base + rank;
      }
    }
    __syncwarp();
  }
}

__global__ void k_sched_degree(const int64_t* __restrict__ ei, int E, int N, int dirs, const int* __restrict__ pos0,
                               const int* __restrict__ pos1, int* __restrict__ deg0, int* __restrict__ deg1,
                               int* __restrict__ summary) {
  // a level outside the table (status 1) or a bad node id (2) leaves pos[] entries unset: no CSR is built then (the sweep and
  // the readout skip such a schedule too; the host raises or rebuilds with a larger level table)
  if (*reinterpret_cast<volatile int*>(summary + 2) == 1) return;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < E * dirs; idx += gridDim.x * blockDim.x) {
    const int d = idx / E, e = idx - d * E;
    const long long t = ei[(size_t)(1 - d) * E + e], o = ei[(size_t)d * E + e];
    if (t < 0 || t >= N || o < 0 || o >= N) { summary[2] = 2; continue; }
    atomicAdd(&(d ? deg1 : deg0)[(d ? pos1 : pos0)[t]], 1);
  }
}

__global__ void __launch_bounds__(kScanThreads) k_sched_rowptr(int* deg0, int* deg1, int N, int* rowptr0, int* rowptr1) {
  __shared__ int smem[33];
  const int d = blockIdx.x;
  int* deg = d ? deg1 : deg0;
  int* rowptr = d ? rowptr1 : rowptr0;
  block_excl_scan(deg, rowptr, N, smem, &rowptr[N]);
  for (int i = threadIdx.x; i < N; i += blockDim.x) deg[i] = 0;   // reused as the fill cursor
}

__global__ void k_sched_fill(const int64_t* __restrict__ ei, int E, int N, int dirs, const int* __restrict__ pos0,
                             const int* __restrict__ pos1, const int* __restrict__ rowptr0, const int* __restrict__ rowptr1,
                             int* __restrict__ cur0, int* __restrict__ cur1, int* __restrict__ tmp0, int* __restrict__ tmp1,
                             const int* __restrict__ summary) {
  if (summary[2] != 0) return;          // set by earlier launches only (k_sched_degree is complete)
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < E * dirs; idx += gridDim.x * blockDim.x) {
    const int d = idx / E, e = idx - d * E;
    const long long t = ei[(size_t)(1 - d) * E + e], o = ei[(size_t)d * E + e];
    if (t < 0 || t >= N || o < 0 || o >= N) continue;
    const int r = (d ? pos1 : pos0)[t];
    const int slot = atomicAdd(&(d ? cur1 : cur0)[r], 1);
    (d ? tmp1 : tmp0)[(d ? rowptr1 : rowptr0)[r] + slot] = e;
  }
}

// one warp per (row, direction): rank-sort the row's edge ids ascending, emit neighbour positions + edge attrs
__global__ void __launch_bounds__(256) k_sched_rows(const int64_t* __restrict__ ei, const float* __restrict__ edge_attr, int E, int N,
                                                    int dirs, const int* __restrict__ pos0, const int* __restrict__ pos1,
                                                    const int* __restrict__ rowptr0, const int* __restrict__ rowptr1,
                                                    const int* __restrict__ tmp0, const int* __restrict__ tmp1, int* __restrict__ eid0,
                                                    int* __restrict__ eid1, int* __restrict__ col0, int* __restrict__ col1,
                                                    float* __restrict__ ea0, float* __restrict__ ea1, const int* __restrict__ summary) {
  if (summary[2] != 0) return;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int w = blockIdx.x * warps_per_block + (threadIdx.x >> 5); w < N * dirs; w += gridDim.x * warps_per_block) {
    const int d = w / N, r = w - d * N;
    const int* rowptr = d ? rowptr1 : rowptr0;
    const int* tmp = d ? tmp1 : tmp0;
    const int* pos = d ? pos1 : pos0;
    int* eid = d ? eid1 : eid0;
    int* col = d ? col1 : col0;
    float* ea = d ? ea1 : ea0;
    const int e0 = rowptr[r], deg = rowptr[r + 1] - e0;
    if (deg <= 0) continue;
    if (deg <= 32) {
      const int v = lane < deg ? tmp[e0 + lane] : 0x7fffffff;
      int rank = 0;
      for (int i = 0; i < deg; ++i) rank += (__shfl_sync(0xffffffffu, v, i) < v) ? 1 : 0;
      if (lane < deg) {
        const int j = e0 + rank;
        eid[j] = v;
        col[j] = pos[ei[(size_t)d * E + v]];
        if (ea) { ea[2 * (size_t)j] = edge_attr[2 * (size_t)v]; ea[2 * (size_t)j + 1] = edge_attr[2 * (size_t)v + 1]; }
      }
    } else {
      for (int q = lane; q < deg; q += 32) {
        const int v = tmp[e0 + q];
        int rank = 0;
        for (int i = 0; i < deg; ++i) rank += (tmp[e0 + i] < v) ? 1 : 0;
        const int j = e0 + rank;
        eid[j] = v;
        col[j] = pos[ei[(size_t)d * E + v]];
        if (ea) { ea[2 * (size_t)j] = edge_attr[2 * (size_t)v]; ea[2 * (size_t)j + 1] = edge_attr[2 * (size_t)v + 1]; }
      }
    }
  }
}

// graph pointers (first node id of every graph) and, optionally, the number of levels of every graph (gdepth, zeroed by the
// launcher: max forward level + 1 — the cluster sweep balances its graph groups with it)
__global__ void k_sched_gptr(const int64_t* __restrict__ batch, const int64_t* __restrict__ lvl0, int N, int B, int max_levels,
                             int* __restrict__ gptr, int* __restrict__ gdepth, int* __restrict__ summary) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v <= N; v += gridDim.x * blockDim.x) {
    long long cur = (v < N) ? batch[v] : B;
    long long prev = (v > 0) ? batch[v - 1] : -1;
    if (cur < prev || cur > B || (v < N && cur >= B)) { summary[2] = 2; continue; }
    for (long long g = prev + 1; g <= cur; ++g) gptr[g] = v;
    if (gdepth && v < N) {
      const long long l = lvl0[v];
      if (l >= 0 && l < max_levels) atomicMax(&gdepth[cur], (int)l + 1);
    }
  }
}

__global__ void k_states_to_node_order(const int* __restrict__ pos, const float* __restrict__ src, int64_t lds, int H,
                                       float* __restrict__ dst, int64_t ldd, int N) {
  const int v = blockIdx.x;
  const float* s = src + (size_t)pos[v] * lds;
  float* o = dst + (size_t)v * ldd;
  for (int c = threadIdx.x; c < H; c += blockDim.x) o[c] = s[c];
}

struct SchedWs {
  int* cnt[2];
  int* deg[2];
  int* tmp[2];
  size_t bytes;
};

static int chunk_for(int64_t N) { return N <= 65536 ? kChunkSmall : kChunkBig; }

static SchedWs carve(void* ws, int64_t N, int64_t E, int max_levels) {
  SchedWs w;
  const int64_t nchunks = (N + chunk_for(N) - 1) / chunk_for(N);
  size_t off = 0;
  auto take = [&](int64_t n) {
    int* p = ws ? reinterpret_cast<int*>(static_cast<char*>(ws) + off) : nullptr;
    off += (size_t)round_up64(n * 4, 256);
    return p;
  };
  for (int d = 0; d < 2; ++d) w.cnt[d] = take((int64_t)max_levels * nchunks);
  for (int d = 0; d < 2; ++d) w.deg[d] = take(N + 1);
  for (int d = 0; d < 2; ++d) w.tmp[d] = take(E);
  w.bytes = off;
  return w;
}

}  // namespace dagnn

using namespace dagnn;

extern "C" size_t dagnn_schedule_workspace_bytes(int64_t N, int64_t E, int32_t max_levels) {
  if (N < 0 || E < 0 || max_levels < 1) return 0;
  return carve(nullptr, N, E, max_levels).bytes;
}

extern "C" int dagnn_schedule_build(const int64_t* edge_index, const int64_t* lvl0, const int64_t* lvl1, const int64_t* nid0,
                                    const int64_t* nid1, const float* edge_attr, const int64_t* batch,
                                    const DagnnSchedule* s, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  DAGNN_REQUIRE(s != nullptr, "schedule is null");
  DAGNN_REQUIRE(s->dirs == 1 || s->dirs == 2, "dirs must be 1 or 2");
  DAGNN_REQUIRE(s->N > 0 && s->N < (1ll << 31) - 1024 && s->E >= 0 && s->E < (1ll << 30), "N/E out of range");
  DAGNN_REQUIRE(s->max_levels >= 1, "max_levels");
  DAGNN_REQUIRE(lvl0 && (s->dirs == 1 || lvl1), "level arrays");
  DAGNN_REQUIRE(s->E == 0 || edge_index, "edge_index");
  DAGNN_REQUIRE(s->summary, "summary");
  for (int d = 0; d < s->dirs; ++d) {
    DAGNN_REQUIRE(s->perm[d] && s->pos[d] && s->lvl_off[d] && s->rowptr[d], "schedule arrays");
    DAGNN_REQUIRE(s->E == 0 || (s->col[d] && s->eid[d]), "schedule edge arrays");
    DAGNN_REQUIRE(!edge_attr || s->E == 0 || s->eattr[d], "eattr array");
  }
  DAGNN_REQUIRE(s->B == 0 || (batch && s->gptr), "batch / gptr");
  const int N = (int)s->N, E = (int)s->E, dirs = s->dirs, ML = s->max_levels;
  const int kChunk = chunk_for(N);
  const int nchunks = (N + kChunk - 1) / kChunk;
  SchedWs w = carve(workspace, N, E, ML);
  if (!workspace || workspace_bytes < w.bytes) return set_err(DAGNN_E_WORKSPACE, "workspace %zu < %zu", workspace_bytes, w.bytes);

  DAGNN_CUDA_OK(cudaMemsetAsync(workspace, 0, (size_t)((char*)w.tmp[0] - (char*)workspace), st));  // cnt + deg
  DAGNN_CUDA_OK(cudaMemsetAsync(s->summary, 0, 8 * sizeof(int), st));
  for (int d = 0; d < s->dirs; ++d) DAGNN_CUDA_OK(cudaMemsetAsync(s->pos[d], 0, (size_t)s->N * sizeof(int), st));   // never garbage
  int* lo1 = dirs == 2 ? s->lvl_off[1] : s->lvl_off[0];
  int *perm1 = dirs == 2 ? s->perm[1] : s->perm[0], *pos1 = dirs == 2 ? s->pos[1] : s->pos[0];
  int* rp1 = dirs == 2 ? s->rowptr[1] : s->rowptr[0];

  k_sched_count<<<dim3(nchunks, dirs), 32, 0, st>>>(lvl0, lvl1, N, nchunks, kChunk, ML, w.cnt[0], w.cnt[1], s->summary);
  if (int rc = check_launch("k_sched_count")) return rc;
  k_sched_scan_levels<<<dirs, kScanThreads, 0, st>>>(w.cnt[0], w.cnt[1], nchunks, ML, N, s->lvl_off[0], lo1, s->summary);
  if (int rc = check_launch("k_sched_scan_levels")) return rc;
  k_sched_place<<<dim3(nchunks, dirs), 32, 0, st>>>(lvl0, lvl1, nid0, nid1, N, nchunks, kChunk, ML, w.cnt[0], w.cnt[1], s->perm[0], perm1,
                                                     s->pos[0], pos1, s->summary);
  if (int rc = check_launch("k_sched_place")) return rc;
  if (E > 0) {
    const int eb = min(1184, ceil_div(E * dirs, 256));
    k_sched_degree<<<eb, 256, 0, st>>>(edge_index, E, N, dirs, s->pos[0], pos1, w.deg[0], w.deg[1], s->summary);
    if (int rc = check_launch("k_sched_degree")) return rc;
  }
  k_sched_rowptr<<<dirs, kScanThreads, 0, st>>>(w.deg[0], w.deg[1], N, s->rowptr[0], rp1);
  if (int rc = check_launch("k_sched_rowptr")) return rc;
  if (E > 0) {
    const int eb = min(1184, ceil_div(E * dirs, 256));
    k_sched_fill<<<eb, 256, 0, st>>>(edge_index, E, N, dirs, s->pos[0], pos1, s->rowptr[0], rp1, w.deg[0], w.deg[1], w.tmp[0],
                                     w.tmp[1], s->summary);
    if (int rc = check_launch("k_sched_fill")) return rc;
    const int rb = min(148 * 8, ceil_div(N * dirs, 8));
    k_sched_rows<<<rb, 256, 0, st>>>(edge_index, edge_attr, E, N, dirs, s->pos[0], pos1, s->rowptr[0], rp1, w.tmp[0], w.tmp[1],
                                     s->eid[0], dirs == 2 ? s->eid[1] : s->eid[0], s->col[0], dirs == 2 ? s->col[1] : s->col[0],
                                     edge_attr ? s->eattr[0] : nullptr, edge_attr ? (dirs == 2 ? s->eattr[1] : s->eattr[0]) : nullptr,
                                     s->summary);
    if (int rc = check_launch("k_sched_rows")) return rc;
  }
  if (s->B > 0) {
    if (s->gdepth) DAGNN_CUDA_OK(cudaMemsetAsync(s->gdepth, 0, (size_t)s->B * sizeof(int), st));
    k_sched_gptr<<<min(1184, ceil_div(N + 1, 256)), 256, 0, st>>>(batch, lvl0, N, (int)s->B, ML, s->gptr, s->gdepth, s->summary);
    if (int rc = check_launch("k_sched_gptr")) return rc;
  }
  return DAGNN_OK;
}

extern "C" int dagnn_states_to_node_order_f32(const DagnnSchedule* s, int32_t dir, const float* src, int64_t lds, int32_t H,
                                              float* dst, int64_t ldd, void* stream_) {
  DAGNN_REQUIRE(s && src && dst && dir >= 0 && dir < s->dirs && H > 0, "states_to_node_order args");
  k_states_to_node_order<<<(int)s->N, 128, 0, static_cast<cudaStream_t>(stream_)>>>(s->pos[dir], src, lds, H, dst, ldd, (int)s->N);
  return check_launch("k_states_to_node_order");
}
