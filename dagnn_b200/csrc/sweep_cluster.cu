// The DAGNN level sweep, cluster-resident: one thread-block cluster of 8 CTAs per (layer, direction, graph group) walks its
// levels with the GRU weights pinned on chip, "aggregate, then project".
//
// Why: the grid-wide sweep (sweep.cu) pays two grid barriers, a weight stream and a tile set-up per wavefront step — ~20 us
// per step on a chain of L + layers - 1 steps, whatever the step holds (profiles/r1h_summary.md). Graphs are independent and
// a level of one graph group is small, so the whole recurrence of a (layer, direction, group) fits one cluster:
//   * the CTA of cluster rank c owns U = H/8 hidden units (all three gates): its slice of W_hh lives in TENSOR MEMORY as the
//     A operand of tcgen05.mma (fp16 hi tile + lo tile, 96 lanes x K/2 columns each), its slice of W_ih in SHARED MEMORY
//     (K-major SWIZZLE_128B hi / lo tiles) — loaded once, never streamed again;
//   * per level:  gather phase — the rows of the level are dealt to the 64 worker warps of the cluster; a warp gathers the
//                 full predecessor rows h_j of its node (CSR, one 16-byte pair per lane), scores them on the fly
//                 (wk . h_j + edge terms), softmax, m_v = sum_e alpha_e h_j, and stores m_v as an fp16 hi/lo operand row (+ fp32);
//                 layer 0 also converts the node's input row x_v into an operand row;
//                 cluster barrier;
//                 projection — every CTA bulk-copies the operand rows of the level (cp.async.bulk, 128 rows x 64 k per
//                 stage) and issues D[gate unit, node] = W_ih x_v (SS) and W_hh m_v (TS, weights from TMEM), three split
//                 products each, fp32 accumulators in TMEM;
//                 epilogue — TMEM -> shared-memory transpose -> GRU pointwise for the CTA's units, h_v stored as fp32
//                 (position order, the output) and, for the next layer, as an operand row;
//                 cluster barrier.
//   * stacked layers pipeline across clusters: cluster (i, d, g) waits for cluster (i - 1, d, g) to publish level l through a
//     release/acquire counter in global memory; work items are ordered layer-major so that a cluster only ever waits for
//     lower-numbered clusters (no co-residency assumption needed for progress).
// HBM traffic is the algorithmic one (input row, predecessor rows, state row) plus the operand rows (fp16 hi/lo copies).
// Nothing of P = W_hh h / Gi = W_ih x is materialised.
//
// Replaces ogbg-code/model/dagnn.py:144-182 incl. AttnConv (:362-373), PyG propagate / softmax / scatter-add, nn.GRUCell (:181)
// and the index_put at :182; D-VAE variants dvae/dagnn.py:109-145, dvae/dagnn_bn.py:108-136 — same contract as sweep.cu.
#include "common.cuh"
#include "sync.cuh"
#include "tc.cuh"

namespace dagnn {

constexpr int kCWorkWarps = 8;
constexpr int kCWorkers = kCWorkWarps * 32;        // gather phase, epilogue
constexpr int kCThreads = kCWorkers + 64;          // + one warp that issues the MMAs (warp 8) and one that issues the bulk copies (warp 9)
constexpr int kCS = 8;                             // CTAs per cluster
constexpr int kCMaxH = 256;                        // K of an operand row: 4 chunks of 64
constexpr int kCU = kCMaxH / kCS;                  // hidden units per CTA at most (3 * 32 = 96 TMEM lanes)
constexpr int kRC = 64;                            // rows per projection chunk (N of the MMAs): two accumulator sets fit next to the weights
constexpr int kCSlots = 12;                        // operand slots of 8 KB (hi + lo tile of 32 rows x 64 k)
constexpr int kCRingBytes = kCSlots * 8192;        // 96 KB operand region
constexpr int kWRows = 96;                         // W_ih slice rows: gate * 32 + unit
constexpr int kWChunkBytes = kWRows * tc::ROW_BYTES;       // 12 KB per (plane, k chunk)
constexpr int kWBytes = 2 * 4 * kWChunkBytes;              // 96 KB
constexpr int kSub = 32;                           // rows per epilogue pass
constexpr int kSLd = 33;                           // stage row pitch (words): conflict-free both ways
// TMEM column map (512 allocated): W_hh hi / lo tiles, then two accumulator sets {W_ih x (64 columns), W_hh m (64 columns)}
constexpr int kColWhi = 0, kColWlo = 128, kColAcc = 256, kColSet = 128, kColAccH = 64;
constexpr int kCMaxItems = 15;                     // 8-CTA clusters of this footprint resident on a B200 (tools/probe_cluster.cu)

struct CDir {
  const int* perm;      // position -> node id
  const int* rowptr;    // [N+1] CSR rows by position
  const int* col;       // [E] neighbour position
  const float* eattr;   // [E,2] in CSR order or nullptr
  const int* lvl_off;   // [max_levels+1]
};
struct CLay {
  float* Hs;                   // H[d][i], [N, ldh] position order
  unsigned char* himg;         // operand rows of H[d][i] (written when a next layer exists)
  unsigned char* mimg;         // operand rows of the aggregates m
  float* m32;                  // [N, ldh] the aggregates in fp32 (z * m term of the cell)
  const float* bias;           // [4][HP]
  const float* wk;             // [HP]
  const float* attnc;          // [4]
  const float* vidk;           // [nvid]
  const __half* w_hh;          // packed image holding W_hh^i (columns [0, Mc))
  const __half* w_ih;          // packed image holding W_ih^i, first column ih_col0
  int ih_col0, ih_nck;         // ... and its k chunks per column block
};
struct ClusterP {
  int dirs, layers, G, H, Hq, Mc, HP, nvid, use_ea, Din, N, B;
  int U;                       // hidden units per CTA (multiple of 4, <= 32)
  int Kh, Kx;                  // operand widths padded to 16 (state / layer-0 input)
  int nckh, nckx;              // 64-k chunks of those
  int vec_x, max_levels;
  long long ldh, ldx, Q;       // Q: rows of an operand-row plane
  const float* X;              // [N, ldx] node order
  unsigned char* ximg[DAGNN_MAX_DIRS];   // operand rows of X in the position order of each direction
  const int* summary;          // [0] levels, [2] status, [3] node ids are not the identity
  const int* gptr;             // [B+1]
  const int* gdepth;           // [B] levels of every graph, or nullptr
  int4* tab;                   // [items][max_levels + 1] per level: first position, rows, first operand row, level start
  unsigned int* flags;         // [2][kCMaxItems][8] per CTA of every cluster: projection chunks whose states it has published; chunks
                               // whose aggregates it has gathered (pipelined levels)
  long long* trace;            // optional [levels][256][16] clock64 stamps per (level, CTA): 0 start, 1 gathered, 2 exchanged,
                               // 3 projected + cells done, 4 level closed, 5 copy warp: every stage started, 7 MMA warp: every
                               // stage issued, 8 workers: accumulators of the first chunk complete
  CDir dir[DAGNN_MAX_DIRS];
  CLay lay[DAGNN_MAX_DIRS][DAGNN_MAX_LAYERS];
};

struct CSmemTail {
  float stage[2][kWRows][kSLd];     // epilogue transpose: [accumulator][gate * 32 + unit][row of the pass]
  float bias[4][kCU];
  uint64_t full[kCSlots], empty[kCSlots], acc_full[2], tmem_free[2];
  uint32_t tmem_slot;
  int nlong;                        // gather phase: rows of this CTA with long in-edge lists (index into the level)
  int longrow[32];
};
static_assert(sizeof(float) * 2 * kWRows * kSLd >= sizeof(float) * 16 * 260, "the half-warp partials alias the epilogue stage");
constexpr size_t kCSmemBytes = 1024 + (size_t)kWBytes + (size_t)kCRingBytes + sizeof(CSmemTail);
static_assert(kCSmemBytes <= 232448, "shared memory plan exceeds the 227 KB opt-in limit");

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_idx() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kCWorkers) : "memory"); }
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// address of the 16-byte granule holding k = [8 gk, 8 gk + 8) of operand row q (plane 0 = hi, 1 = lo). Layout: per 64-k chunk,
// per group of 8 rows: the hi atom (8 rows x 128 bytes, granules XOR-swizzled with the row like the shared-memory tile it is
// copied into) then the lo atom — the rows [q0, q0 + 8 g) of one k chunk, both halves, are ONE contiguous run of 2 KB per
// group: one bulk copy per stage, and in shared memory the hi / lo tiles are read with a 2 KB group stride.
__device__ __forceinline__ unsigned char* oprow_ptr(unsigned char* img, int plane, int nck, long long Q, long long q, int gk) {
  (void)nck;
  return img + ((((size_t)(gk >> 3) * (size_t)(Q >> 3) + (size_t)(q >> 3)) * 2 + plane) << 10) + (((int)q & 7) << 7) +
         (((gk & 7) ^ ((int)q & 7)) << 4);
}
// 8 fp16 weights W[col, 8 gk .. 8 gk + 8) from a packed projection image (pack.cu: 64-column blocks x 64-k chunks, hi then lo tile)
__device__ __forceinline__ uint4 packed_w8(const __half* img, int nck, int col, int gk, int plane) {
  const int cb = col >> 6, j = col & 63;
  const size_t off = (((size_t)cb * nck + (gk >> 3)) * 2 + plane) * 4096 + (size_t)(j >> 3) * 512 + (size_t)(j & 7) * 64 +
                     (size_t)(((gk & 7) ^ (j & 7)) << 3);
  return __ldg(reinterpret_cast<const uint4*>(img + off));
}

// 16 bytes into the shared memory of CTA `rank` of the cluster, at the address `saddr` has in this CTA (DSMEM)
__device__ __forceinline__ void st_cluster_v4(uint32_t saddr, uint32_t rank, const uint4& v) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(saddr), "r"(rank));
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// Levels of <= 32 rows skip global memory for the aggregate operand: the gathering half-warp writes the hi / lo granules of its
// row straight into the operand slots of ALL CTAs of the cluster (slot = 64-k chunk, 8 KB: groups of 8 rows, hi atom then lo
// atom — the layout a bulk copy of the operand rows would have produced). `mslot0_s` = shared address of the first such slot.
__device__ __forceinline__ void store_oprow16_direct(uint32_t mslot0_s, int r, int hl, int K, const float4 (&v)[4]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int gk = 2 * hl + h;
    if (8 * gk >= K) continue;
    const float f[8] = {v[2 * h].x, v[2 * h].y, v[2 * h].z, v[2 * h].w, v[2 * h + 1].x, v[2 * h + 1].y, v[2 * h + 1].z, v[2 * h + 1].w};
    uint4 hi, lo;
    tc::split8(f, hi, lo);
    const uint32_t a = mslot0_s + (uint32_t)(gk >> 3) * 8192u + (uint32_t)(r >> 3) * 2048u + (uint32_t)(r & 7) * 128u +
                       (uint32_t)(((gk & 7) ^ (r & 7)) << 4);
#pragma unroll
    for (uint32_t c = 0; c < (uint32_t)kCS; ++c) {
      st_cluster_v4(a, c, hi);
      st_cluster_v4(a + 1024u, c, lo);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// gather phase: HALF a warp per node (two nodes per warp in lockstep). Lane hl = lane & 15 owns k = [16 hl, 16 hl + 16) of a
// row. A node costs its chain of dependent L2 round trips (row pointers -> in-edge list -> predecessor rows), so up to four
// predecessor rows per node are in flight at once and the softmax is ONLINE (running max / denominator / weighted sum,
// rescaled when the max moves): one pass over the in-edge list whatever its length. Lists longer than kLongEdges are split
// over the 16 half-warps of the CTA and merged through shared memory.
// ------------------------------------------------------------------------------------------------------------
constexpr int kLongEdges = 16;
constexpr int kMaxLong = 32;       // long lists per CTA and level that get the cooperative treatment (more: a half-warp alone)
constexpr int kPartLd = 260;       // floats per half-warp partial: 256 weighted sums, running max, denominator

struct Row16 { float4 v[4]; };
struct GAcc {
  float mx, den;
  float4 m[4];
};

// max that keeps a NaN (fmaxf drops it): a non-finite predecessor state must poison the aggregate, not vanish from it
__device__ __forceinline__ float nanmax(float a, float b) { return (b > a || b != b) ? b : a; }
__device__ __forceinline__ float half_sum(float v) {             // sum over the 16 lanes of a half-warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void load_row16(Row16& R, const float* __restrict__ row, int k0, int width4, bool on) {
  // width4 = valid floats of the row rounded up to 4 (rows are zero padded to it)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    R.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on && k0 + 4 * j + 4 <= width4) R.v[j] = ldcg4(row + k0 + 4 * j);
  }
}
__device__ __forceinline__ float dot16(const Row16& R, const float4 (&w)[4]) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s = fmaf(R.v[j].x, w[j].x, s); s = fmaf(R.v[j].y, w[j].y, s); s = fmaf(R.v[j].z, w[j].z, s); s = fmaf(R.v[j].w, w[j].w, s);
  }
  return s;
}
// 16 consecutive k of operand row q, owned by half-warp lane hl: two granules per plane
__device__ __forceinline__ void store_oprow16(unsigned char* img, int nck, long long Q, long long q, int hl, int K, const float4 (&v)[4]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int gk = 2 * hl + h;
    if (8 * gk >= K) continue;
    const float f[8] = {v[2 * h].x, v[2 * h].y, v[2 * h].z, v[2 * h].w, v[2 * h + 1].x, v[2 * h + 1].y, v[2 * h + 1].z, v[2 * h + 1].w};
    uint4 hi, lo;
    tc::split8(f, hi, lo);
    *reinterpret_cast<uint4*>(oprow_ptr(img, 0, nck, Q, q, gk)) = hi;
    *reinterpret_cast<uint4*>(oprow_ptr(img, 1, nck, Q, q, gk)) = lo;
  }
}

// x_v -> operand rows (layer 0): rows r = first, first + stride, ... of a level, NR of them in flight per half-warp
template <int NR>
__device__ __forceinline__ void convert_x_rows_n(const ClusterP& P, const CDir& D, int d, int pos0, int q0, int n, int first, int stride,
                                                 int hl) {
  const int k0 = 16 * hl;
  if (k0 >= P.Kx) return;
  const int w4 = (P.Din + 3) & ~3;
  for (int r = first; r < n; r += NR * stride) {
    Row16 R[NR];
    bool on[NR];
#pragma unroll
    for (int t = 0; t < NR; ++t) {
      const int rr = r + t * stride;
      on[t] = rr < n;
      const float* src = P.X + (size_t)(on[t] ? D.perm[pos0 + rr] : 0) * P.ldx;
      if (P.vec_x) {
        load_row16(R[t], src, k0, w4, on[t]);
      } else {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = (on[t] && k0 + j < P.Din) ? __ldcg(src + k0 + j) : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) R[t].v[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
      }
    }
#pragma unroll
    for (int t = 0; t < NR; ++t)
      if (on[t]) store_oprow16(P.ximg[d], P.nckx, P.Q, (long long)q0 + r + t * stride, hl, P.Kx, R[t].v);
  }
}

__device__ __forceinline__ void convert_x_rows(const ClusterP& P, const CDir& D, int d, int pos0, int q0, int n, int first, int stride, int hl) {
  convert_x_rows_n<2>(P, D, d, pos0, q0, n, first, stride, hl);
}

// score of one in-edge apart from the key term: edge type + vertex id (constants per destination cancel)
__device__ __forceinline__ float edge_terms(const ClusterP& P, const CDir& D, const CLay& Lp, int e, int sp, float ca0, float ca1) {
  float sc = 0.f;
  if (D.eattr) {
    const float2 ea = __ldg(reinterpret_cast<const float2*>(D.eattr) + e);
    sc = ca0 * ea.x + ca1 * ea.y;
  }
  if (P.nvid > 0) sc += __ldg(Lp.vidk + (D.perm[sp] % P.nvid));
  return sc;
}

// In-edges e = e_first + (t << ls), t = 0, 1, ..., below e_end, folded into the running softmax state A of this half-warp's
// node. A predecessor that is not in an earlier level (col >= first position of this level) scores without its key term,
// keeps its softmax mass and adds a zero row. Both half-warps of a warp run the same number of rounds.
__device__ __forceinline__ void gather_edges(const ClusterP& P, const CDir& D, const CLay& Lp, int e_first, int e_end, int ls, int lstart,
                                             int hl, const float4 (&wk)[4], float ca0, float ca1, GAcc& A) {
  const int k0 = 16 * hl, w4 = P.Hq;
  const float* __restrict__ Hs = Lp.Hs;
  int nb = e_end > e_first ? (((e_end - e_first - 1) >> ls) >> 2) + 1 : 0;
  nb = max(nb, __shfl_xor_sync(0xffffffffu, nb, 16));
#pragma unroll 1
  for (int b = 0; b < nb; ++b) {
    Row16 R[4];
    float sc[4], dt[4];
    bool val[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = e_first + ((4 * b + t) << ls);
      const bool live = e < e_end;
      const int sp = live ? D.col[e] : 0;
      val[t] = live && sp < lstart;
      load_row16(R[t], Hs + (size_t)sp * P.ldh, k0, w4, val[t]);
      sc[t] = live ? edge_terms(P, D, Lp, e, sp, ca0, ca1) : -INFINITY;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) dt[t] = dot16(R[t], wk);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
      for (int t = 0; t < 4; ++t) dt[t] += __shfl_xor_sync(0xffffffffu, dt[t], o);
    }
    float bm = -INFINITY;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (val[t]) sc[t] += dt[t];
      bm = nanmax(bm, sc[t]);
    }
    if (bm == -INFINITY) continue;                       // nothing of this node in this round (the other half-warp is still going)
    const float mnew = nanmax(A.mx, bm);
    const float scale = expf(A.mx - mnew);               // 0 on the first round (running max = -inf)
    float s = 0.f, w[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) { w[t] = expf(sc[t] - mnew); s += w[t]; }       // dead edges: exp(-inf) = 0
    A.den = fmaf(A.den, scale, s);
    A.mx = mnew;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 m = A.m[j];
      m.x *= scale; m.y *= scale; m.z *= scale; m.w *= scale;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float a = val[t] ? w[t] : 0.f;
        m.x = fmaf(a, R[t].v[j].x, m.x); m.y = fmaf(a, R[t].v[j].y, m.y); m.z = fmaf(a, R[t].v[j].z, m.z); m.w = fmaf(a, R[t].v[j].w, m.w);
      }
      A.m[j] = m;
    }
  }
}
__device__ __forceinline__ void gacc_init(GAcc& A) {
  A.mx = -INFINITY; A.den = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) A.m[j] = make_float4(0.f, 0.f, 0.f, 0.f);
}
// m_v = (weighted sum) / (denominator + 1e-16) (PyG softmax), stored as an operand row and in fp32
// (direct_s != 0: into the operand slots of every CTA of the cluster instead of the global operand rows; r = row of the level)
__device__ __forceinline__ void finish_row(const ClusterP& P, const CLay& Lp, int p, long long q, int hl, GAcc& A, uint32_t direct_s, int r) {
  const float inv = 1.f / (A.den + 1e-16f);
  const int k0 = 16 * hl;
#pragma unroll
  for (int j = 0; j < 4; ++j) { A.m[j].x *= inv; A.m[j].y *= inv; A.m[j].z *= inv; A.m[j].w *= inv; }
  if (direct_s) store_oprow16_direct(direct_s, r, hl, P.Kh, A.m);
  else store_oprow16(Lp.mimg, P.nckh, P.Q, q, hl, P.Kh, A.m);
  float* mo = Lp.m32 + (size_t)p * P.ldh + k0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (k0 + 4 * j + 4 <= P.Hq) *reinterpret_cast<float4*>(mo + 4 * j) = A.m[j];
}

// ------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lower_bound_dev(const int* a, int n, int key) {      // first index with a[idx] >= key
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kCThreads, 1) k_sweep_cluster(const __grid_constant__ ClusterP P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* Wih = base;                                    // [plane][k chunk][96 rows x 128 B]
  unsigned char* Ring = base + kWBytes;
  CSmemTail& S = *reinterpret_cast<CSmemTail*>(Ring + (size_t)kCRingBytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long t_entry = clock64();
  const int rank = (int)cluster_ctarank();
  const int item = (int)cluster_idx();
  const int g = item % P.G, di = item / P.G;
  const int i = di / P.dirs, d = di - i * P.dirs;
  const CDir& D = P.dir[d];
  const CLay& Lp = P.lay[d][i];
  const bool first_layer = i == 0, has_next = i + 1 < P.layers;
  const int u0 = rank * P.U;
  const int Kin = first_layer ? P.Kx : P.Kh;                   // padded width of the cell's input operand
  const int nck_in = first_layer ? P.nckx : P.nckh;
  unsigned char* inimg = first_layer ? P.ximg[d] : P.lay[d][i - 1].himg;

  if (tid == 0) {
    for (int s = 0; s < kCSlots; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&S.acc_full[s], 1); mbar_init(&S.tmem_free[s], kCWorkWarps); }
    S.nlong = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tc::tmem_alloc(&S.tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = S.tmem_slot;

  const bool ok = P.summary[2] == 0;           // schedule build flagged bad input: do nothing, the host raises
  const int L = ok ? P.summary[0] : 0;
  int4* tab = P.tab + (size_t)item * (P.max_levels + 1);

  // ---- level table of this (direction, group): CTA 0 of the cluster builds it, the cluster barrier below publishes it
  if (rank == 0 && L > 0) {
    // Graph groups = G consecutive ranges of graphs. A cluster's time is ~ levels x t_level + nodes x t_node (measured: 6.7 us
    // per level, 0.081 us per node: one level costs what kLevelNodes nodes cost), so the cuts minimise the largest
    // kLevelNodes x (deepest graph of the group) + (nodes of the group): every cluster derives the same cuts from the graph
    // pointers and depths (32-ary search on the bound, greedy feasibility scan per lane). Without depths: equal node counts.
    constexpr int kLevelNodes = 80, kMaxBalanced = 3000;
    int gb_lo = 0, gb_hi = P.N;
    if (P.G > 1) {
      if (P.summary[3] != 0) {                  // level arrays with their own node ids: no order to cut groups by
        if (g > 0) gb_lo = P.N;
      } else if (P.gdepth != nullptr && P.B <= kMaxBalanced) {
        int* sg = reinterpret_cast<int*>(&S.stage[0][0][0]);      // [B + 1] graph pointers, [B] depths (the stage is idle here)
        int* sd = sg + P.B + 1;
        int* sb = sd + P.B;                                        // [G + 1] cuts (graph indices)
        for (int k = tid; k <= P.B; k += kCThreads) sg[k] = P.gptr[k];
        for (int k = tid; k < P.B; k += kCThreads) sd[k] = P.gdepth[k];
        __syncthreads();
        if (warp == 0) {
          auto groups_needed = [&](int T, bool emit) {
            int groups = 1, nn = 0, dd = 0;
            if (emit) sb[0] = 0;
#pragma unroll 8
            for (int k = 0; k < P.B; ++k) {                        // (the loads do not depend on the running state: they pipeline)
              const int nk = sg[k + 1] - sg[k], dk = sd[k];
              const int n2 = nn + nk, d2 = max(dd, dk);
              if (nn > 0 && kLevelNodes * d2 + n2 > T) {
                if (emit && groups <= P.G) sb[groups] = k;
                ++groups; nn = nk; dd = dk;
              } else { nn = n2; dd = d2; }
            }
            if (emit) for (int q = groups; q <= P.G; ++q) sb[q] = P.B;
            return groups;
          };
          int lo = 0, hi = kLevelNodes * L + P.N;                  // feasible at hi (one group), infeasible below every single graph's cost
          for (int round = 0; round < 3 && hi - lo > 8; ++round) {      // to within 8 nodes' worth of cost
            const int step = max(1, (hi - lo + 31) / 32);
            const int T = min(hi, lo + step * (lane + 1));
            const bool okT = groups_needed(T, false) <= P.G;
            const unsigned m = __ballot_sync(0xffffffffu, okT);       // monotone: feasible from some lane on
            const int first = m ? __ffs(m) - 1 : 31;
            const int nhi = min(hi, lo + step * (first + 1)), nlo = first == 0 ? lo : min(hi, lo + step * first);
            hi = nhi; lo = nlo;
          }
          if (lane == 0) groups_needed(hi, true);
        }
        __syncthreads();
        gb_lo = sg[sb[g]];
        gb_hi = g + 1 == P.G ? P.N : sg[sb[g + 1]];
        __syncthreads();
      } else {
        const long long t0 = (long long)P.N * g / P.G, t1 = (long long)P.N * (g + 1) / P.G;
        gb_lo = g == 0 ? 0 : P.gptr[lower_bound_dev(P.gptr, P.B + 1, (int)t0)];
        gb_hi = g + 1 == P.G ? P.N : P.gptr[lower_bound_dev(P.gptr, P.B + 1, (int)t1)];
      }
    }
    for (int l = tid; l < L; l += kCThreads) {
      const int a = D.lvl_off[l], b = D.lvl_off[l + 1];
      const int lo = a + lower_bound_dev(D.perm + a, b - a, gb_lo);
      const int hi = a + lower_bound_dev(D.perm + a, b - a, gb_hi);
      tab[l] = make_int4(lo, hi - lo, 0, a);
    }
    __syncthreads();
    if (warp == 0) {
      int run = ((gb_lo + 7) & ~7) + 8 * L * g;                // operand rows of a group start past every earlier group's
      for (int l0 = 0; l0 < L; l0 += 32) {
        const int l = l0 + lane;
        const int v = l < L ? (tab[l].y + 7) & ~7 : 0;
        const int incl = warp_incl_scan(v, lane);
        if (l < L) tab[l].z = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
  }

  // ---- weights on chip: W_hh slice -> TMEM (hi tile, lo tile), W_ih slice -> shared memory, biases. The loads are issued in
  // batches (16 independent 16-byte loads per thread in flight): done one at a time this prologue took 35 us of a 1 ms kernel.
  if (warp < 4) {
    const int gate = tid >> 5, uu = tid & 31, u = u0 + uu;
    const bool live = gate < 3 && uu < P.U && u < P.H;
    const int col = gate * P.Hq + u;
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
    const int nckw = P.HP >> 6;                                 // k chunks of the packed W_hh image (Kh64 / 64)
    const int nks = P.Kh / 16;
    for (int ks0 = 0; ks0 < nks; ks0 += 4) {
      uint4 wv[4][4];                                           // [k step][hi granule 0, hi granule 1, lo granule 0, lo granule 1]
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[q][j] = make_uint4(0u, 0u, 0u, 0u);
        if (live && ks0 + q < nks) {
          wv[q][0] = packed_w8(Lp.w_hh, nckw, col, 2 * (ks0 + q), 0); wv[q][1] = packed_w8(Lp.w_hh, nckw, col, 2 * (ks0 + q) + 1, 0);
          wv[q][2] = packed_w8(Lp.w_hh, nckw, col, 2 * (ks0 + q), 1); wv[q][3] = packed_w8(Lp.w_hh, nckw, col, 2 * (ks0 + q) + 1, 1);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (ks0 + q < nks) {                                    // warp-uniform
          const uint32_t wh[8] = {wv[q][0].x, wv[q][0].y, wv[q][0].z, wv[q][0].w, wv[q][1].x, wv[q][1].y, wv[q][1].z, wv[q][1].w};
          const uint32_t wl[8] = {wv[q][2].x, wv[q][2].y, wv[q][2].z, wv[q][2].w, wv[q][3].x, wv[q][3].y, wv[q][3].z, wv[q][3].w};
          tc::st8(taddr + (uint32_t)(kColWhi + 8 * (ks0 + q)), wh);
          tc::st8(taddr + (uint32_t)(kColWlo + 8 * (ks0 + q)), wl);
        }
      }
    }
    tc::wait_st();
  } else {
    // 192 threads = 2 planes x 96 rows of the W_ih slice: a thread copies its row, 8 granules in flight
    const int t = tid - 128;
    if (t < 2 * kWRows) {
      const int plane = t / kWRows, r = t - plane * kWRows;
      const int gate = r >> 5, uu = r & 31, u = u0 + uu;
      const bool live = uu < P.U && u < P.H;
      const int colw = Lp.ih_col0 + gate * P.Hq + u;
      const int ngr = nck_in * 8;                               // 16-byte granules per row
      for (int g0 = 0; g0 < ngr; g0 += 8) {
        uint4 w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          w[j] = (live && 8 * (g0 + j) < Kin) ? packed_w8(Lp.w_ih, Lp.ih_nck, colw, g0 + j, plane) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(Wih + ((size_t)plane * nck_in + ((g0 + j) >> 3)) * kWChunkBytes + tc::tile_off(r, (g0 + j) & 7)) = w[j];
      }
    }
    for (int idx = t; idx < 4 * kCU; idx += kCThreads - 128) {
      const int b = idx / kCU, uu = idx % kCU, u = u0 + uu;
      S.bias[b][uu] = (uu < P.U && u < P.HP) ? __ldg(Lp.bias + (size_t)b * P.HP + u) : 0.f;
    }
  }
  const int hl = lane & 15;
  float4 wk[4];                                                 // key weights of the half-warp lane's k slice (gather phase)
#pragma unroll
  for (int j = 0; j < 4; ++j)
    wk[j] = (16 * hl + 4 * j + 4 <= P.HP) ? __ldg(reinterpret_cast<const float4*>(Lp.wk + 16 * hl + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float ca0 = D.eattr ? __ldg(Lp.attnc) : 0.f, ca1 = D.eattr ? __ldg(Lp.attnc + 1) : 0.f;
  tc::fence_async_smem();                // W_ih tiles were written with ordinary stores, the tensor core reads them
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  cluster_sync_all();

  const uint32_t wih_s = smem_u32(Wih), ring_s = smem_u32(Ring);
  uint32_t ct = 0;                       // projection chunks done so far (every role keeps its own count; they agree per level)
  uint32_t full_par = 0u, load_par = 0u; // MMA warp: bit s = parity of its next wait on full[s]; copy warp: bit s = copies into slot s so far & 1
  // half-warp index in the cluster, rank-minor: a short level spreads over the CTAs
  const int ghw = ((lane >> 4) * kCWorkWarps + (warp & (kCWorkWarps - 1))) * kCS + rank;
  constexpr int kHalfWarps = kCS * kCWorkWarps * 2;            // 128
  constexpr int kMmaWarp = kCWorkWarps, kCopyWarp = kCWorkWarps + 1;

  // layer 0: the operand rows of level 0's inputs (those of level l + 1 are converted during level l)
  if (first_layer && L > 0) {
    const int4 T0 = __ldcg(tab);
    if (warp < kCWorkWarps) convert_x_rows_n<4>(P, D, d, T0.x, T0.z, T0.y, ghw, kHalfWarps, hl);
    asm volatile("fence.proxy.async.global;" ::: "memory");
    cluster_sync_all();
  }

  if (P.trace && tid == 0 && L < P.max_levels) {       // kernel-level stamps in the entry behind the last level: entry, prologue done
    long long* tk = P.trace + (((size_t)L * 256 + blockIdx.x) << 4);
    tk[0] = t_entry; tk[1] = clock64();
  }
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    const int4 T = __ldcg(tab + l);
    const int4 Tn = (first_layer && l + 1 < L) ? __ldcg(tab + l + 1) : make_int4(0, 0, 0, 0);
    const int pos0 = T.x, n = T.y, q0 = T.z, lstart = T.w;
    long long* tr = (P.trace && l < P.max_levels) ? P.trace + (((size_t)l * 256 + blockIdx.x) << 4) : nullptr;
    if (tr && tid == 0) tr[0] = clock64();
    if (n == 0) break;                     // a group without nodes at this level has none at deeper ones
    const int nchunks = (n + kRC - 1) / kRC;
    const bool exchange = l > 0;           // the aggregates of this level travel through global memory to every CTA
    const bool pipelined = exchange && n > kCS * kCWorkWarps * 2;               // more rows than half-warps in the cluster (see below)
    unsigned int* mflag = P.flags + (size_t)(kCMaxItems + item) * kCS;          // chunks of aggregates gathered, per CTA of this cluster
    const int ns_gate = pipelined ? nck_in : -1;                                // copy warp: stage k == ns_gate waits for the counters
    // levels of <= 32 rows: the aggregate operand goes straight into every CTA's operand slots (DSMEM), no copies for it
    const bool direct = exchange && n <= 32;
    const uint32_t direct_s = direct ? ring_s + (uint32_t)nck_in * 8192u : 0u;

    // The projection of a level is a linear sequence of stages j = chunk * ns + k (k < nck_in: a 64-k chunk of the input
    // operand, then the 64-k chunks of the aggregate operand). The COPY warp starts one bulk copy per stage (hi + lo tile of
    // the chunk's rows) into slot j % nslots as soon as the MMAs that read the slot before have completed; the MMA warp waits
    // for the bytes, issues the 12 MMAs of the stage into the accumulator set of the chunk and commits (frees the slot; the
    // last stage of a chunk also signals the accumulators). Chunks alternate between two accumulator sets: the epilogue of one
    // overlaps the MMAs of the next. Levels of <= 32 rows cut the operand region into 12 slots (the whole chunk in flight at
    // once), the others into 6. Stage cursors advance without a runtime division (walked ~8 times per level by each role).
    const int ns = nck_in + (l > 0 ? P.nckh : 0);
    const int rp = n <= 32 ? 32 : kRC;                         // rows of a stage tile
    const int nslots = n <= 32 ? 12 : 6;
    const uint32_t stage_bytes = 2u * (uint32_t)rp * tc::ROW_BYTES;
    const int total_stages = nchunks * ns;
    struct Cur { int j, c, k, slot; };
    auto advance = [&](Cur& q) {
      ++q.j;
      if (++q.k == ns) { q.k = 0; ++q.c; }
      if (++q.slot == nslots) q.slot = 0;
    };
    Cur cur = {0, 0, 0, 0};
    // ---- copy warp: stages [cur.j, j_hi)
    auto copy_stages = [&](int j_hi) {
#pragma unroll 1
      for (; cur.j < j_hi; advance(cur)) {
        const int c = cur.c, k = cur.k, slot = cur.slot;
        if (direct && k >= nck_in) continue;                   // written by the gathering half-warps of the cluster
        const int r0 = c * kRC;
        const uint32_t bytes = (uint32_t)((min(kRC, n - r0) + 7) >> 3) * 2048u;     // hi + lo atom of every 8-row group
        // MMAs of the stage that used this slot before (this level): committed by the MMA warp, wait for them to finish
        if (cur.j >= nslots) mbar_wait(&S.empty[slot], ((load_par >> slot) & 1u) ^ 1u);
        if (k == 0 && !first_layer) {
          // the input rows of this chunk are the states the cluster of the layer below has published: every one of its CTAs
          // counts the chunks whose unit slice it has stored (same level tables on both sides, so chunk numbers agree)
          const unsigned int* fl = P.flags + (size_t)(item - P.dirs * P.G) * kCS + (lane & (kCS - 1));
          const unsigned int want = ct + (unsigned int)c + 1u;
          while (!__all_sync(0xffffffffu, ld_acquire_u32(fl) >= want)) {}
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        if (k == ns_gate) {
          // pipelined level: the aggregate rows of this chunk are there once every CTA of the cluster has counted it
          const unsigned int* fl = mflag + (lane & (kCS - 1));
          const unsigned int want = ct + (unsigned int)c + 1u;
          while (!__all_sync(0xffffffffu, ld_acquire_u32(fl) >= want)) {}
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        if (elect_one()) {
          const bool is_in = k < nck_in;
          const int kc = is_in ? k : k - nck_in;
          const unsigned char* src = (is_in ? inimg : Lp.mimg) + (((size_t)kc * (size_t)(P.Q >> 3) + (size_t)((q0 + r0) >> 3)) << 11);
          mbar_expect_tx(&S.full[slot], bytes);
          bulk_g2s(Ring + (size_t)slot * stage_bytes, src, bytes, &S.full[slot]);
        }
        __syncwarp();
        load_par ^= 1u << slot;
      }
    };
    // ---- MMA warp: stages [cur.j, j_hi)
    auto mma_stages = [&](int j_hi) {
#pragma unroll 1
      for (; cur.j < j_hi; advance(cur)) {
        const int c = cur.c, k = cur.k, slot = cur.slot;
        const uint32_t cc = ct + (uint32_t)c, set = cc & 1u;
        const bool in_slot = direct && k >= nck_in;            // operand already in the slot (DSMEM), behind the exchange barrier
        if (!in_slot) {
          mbar_wait(&S.full[slot], (full_par >> slot) & 1u);
          full_par ^= 1u << slot;
        }
        if (k == 0 && cc >= 2) mbar_wait(&S.tmem_free[set], ((cc >> 1) - 1u) & 1u);   // this set's previous chunk has been read
        tc::fence_after_sync();
        const bool is_in = k < nck_in;
        const int kc = is_in ? k : k - nck_in;
        const int Kop = is_in ? Kin : P.Kh;
        const int nks = min(4, (Kop - 64 * kc) >> 4);
        const int rows = min(kRC, n - c * kRC);
        const uint32_t idesc = uni(tc::instr_desc_f16(128, (rows + 15) & ~15));
        const uint32_t sb = uni(ring_s + (uint32_t)slot * stage_bytes);
        const uint64_t bh = tc::smem_desc(sb, 2048u), bl = tc::smem_desc(sb + 1024u, 2048u);
        const uint32_t fresh = uni(kc == 0 ? 0u : 1u);
        // 12 MMAs per full stage, issued back to back by one elected lane from warp-uniform operands and fully unrolled:
        // anything else (a counted loop, lane == 0 instead of elect.sync) costs 65-85 cycles per issue instead of 42-50
        // (tools/ubench_mma.cu, profiles/r1h_ubench_mma.txt)
        if (is_in) {
          const uint32_t a0 = uni(wih_s + (uint32_t)kc * kWChunkBytes), a1 = uni(wih_s + (uint32_t)(nck_in + kc) * kWChunkBytes);
          const uint64_t ah = tc::smem_desc(a0), al = tc::smem_desc(a1);
          const uint32_t acc = uni(tmem + (uint32_t)kColAcc + set * (uint32_t)kColSet);
          if (nks == 4) {
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                tc::mma_f16(acc, ah + 2 * ks, bh + 2 * ks, idesc, ks == 0 ? fresh : 1u);
                tc::mma_f16(acc, ah + 2 * ks, bl + 2 * ks, idesc, 1u);
                tc::mma_f16(acc, al + 2 * ks, bh + 2 * ks, idesc, 1u);
              }
            }
          } else if (elect_one()) {
            for (int ks = 0; ks < nks; ++ks) {
              tc::mma_f16(acc, ah + 2 * ks, bh + 2 * ks, idesc, ks == 0 ? fresh : 1u);
              tc::mma_f16(acc, ah + 2 * ks, bl + 2 * ks, idesc, 1u);
              tc::mma_f16(acc, al + 2 * ks, bh + 2 * ks, idesc, 1u);
            }
          }
        } else {
          const uint32_t acc = uni(tmem + (uint32_t)(kColAcc + kColAccH) + set * (uint32_t)kColSet);
          const uint32_t ah = uni(tmem + (uint32_t)(kColWhi + 32 * kc)), al = uni(tmem + (uint32_t)(kColWlo + 32 * kc));
          if (nks == 4) {
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                tc::mma_f16_ts(acc, ah + 8 * ks, bh + 2 * ks, idesc, ks == 0 ? fresh : 1u);
                tc::mma_f16_ts(acc, ah + 8 * ks, bl + 2 * ks, idesc, 1u);
                tc::mma_f16_ts(acc, al + 8 * ks, bh + 2 * ks, idesc, 1u);
              }
            }
          } else if (elect_one()) {
            for (int ks = 0; ks < nks; ++ks) {
              tc::mma_f16_ts(acc, ah + 8 * ks, bh + 2 * ks, idesc, ks == 0 ? fresh : 1u);
              tc::mma_f16_ts(acc, ah + 8 * ks, bl + 2 * ks, idesc, 1u);
              tc::mma_f16_ts(acc, al + 8 * ks, bh + 2 * ks, idesc, 1u);
            }
          }
        }
        if (elect_one()) {
          if (!in_slot) tc::commit(&S.empty[slot]);
          if (k == ns - 1) tc::commit(&S.acc_full[set]);
        }
        __syncwarp();
      }
    };
    // Levels of more than 128 rows run PIPELINED: the gather proceeds in rounds of 128 rows (one per half-warp of the cluster =
    // two projection chunks); after a round every CTA publishes "chunks gathered" through a release counter in global memory,
    // the copy warps poll the eight counters before copying the aggregate stages of a chunk — no cluster barrier in between —
    // and the worker warps run the cells of round t - 1 behind the gather of round t: gather, copies, MMAs and cells of
    // different chunks overlap. Short levels keep the single exchange barrier (one round trip less on the critical path).
    // stages that wait for nothing of this level: the input part of the first chunk (everything at level 0; in pipelined levels
    // the copy warp gates the aggregate stages itself)
    const int early = (exchange && !pipelined) ? nck_in : total_stages;

    // ---- worker pieces
    // rows r = first, first + stride, ... < r_end of this level: aggregates (and softmax) of the nodes, long in-edge lists
    // split over the CTA's half-warps afterwards
    auto gather_rows = [&](int first, int stride, int r_end) {
      for (int r = first; __any_sync(0xffffffffu, r < r_end); r += stride) {
        const bool on = r < r_end;
        const int p = pos0 + (on ? r : 0);
        int e0 = 0, e1 = 0;
        if (on) { e0 = D.rowptr[p]; e1 = D.rowptr[p + 1]; }
        bool lng = e1 - e0 > kLongEdges;                     // split over the CTA's half-warps below, if the list has room
        int slot = 0;
        if (lng && hl == 0) slot = atomicAdd(&S.nlong, 1);
        slot = __shfl_sync(0xffffffffu, slot, lane & 16);
        if (lng) {
          if (slot < kMaxLong) { if (hl == 0) S.longrow[slot] = r; e1 = e0; } else lng = false;
        }
        GAcc A;
        gacc_init(A);
        gather_edges(P, D, Lp, e0, e1, 0, lstart, hl, wk, ca0, ca1, A);
        if (on && !lng) finish_row(P, Lp, p, (long long)q0 + r, hl, A, direct_s, r);
      }
      workers_sync();
      const int nl = min(S.nlong, kMaxLong);
      float* part = &S.stage[0][0][0];                       // [16 half-warps][kPartLd]
      for (int j = 0; j < nl; ++j) {
        const int r = S.longrow[j];
        const int p = pos0 + r;
        const int e0 = D.rowptr[p], e1 = D.rowptr[p + 1];
        const int hw = 2 * warp + (lane >> 4);
        GAcc A;
        gacc_init(A);
        gather_edges(P, D, Lp, e0 + hw, e1, 4, lstart, hl, wk, ca0, ca1, A);
        float* mine = part + hw * kPartLd;
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(mine + 16 * hl + 4 * q) = A.m[q];
        if (hl == 0) { mine[256] = A.mx; mine[257] = A.den; }
        workers_sync();
        {                                                    // thread = k: merge the 16 partial softmax states
          float M = -INFINITY;
#pragma unroll
          for (int h = 0; h < 16; ++h) M = nanmax(M, part[h * kPartLd + 256]);
          float den = 0.f, acc = 0.f;
#pragma unroll
          for (int h = 0; h < 16; ++h) {
            const float w = expf(part[h * kPartLd + 256] - M);
            den = fmaf(part[h * kPartLd + 257], w, den);
            acc = fmaf(part[h * kPartLd + tid], w, acc);
          }
          const float v = acc / (den + 1e-16f);
          if (tid < P.Hq) Lp.m32[(size_t)p * P.ldh + tid] = v;
          float f[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) f[q] = __shfl_sync(0xffffffffu, v, (lane & ~7) + q);
          if ((lane & 7) == 0 && tid < P.Kh) {
            uint4 hi, lo;
            tc::split8(f, hi, lo);
            if (direct_s) {
              const int gk = tid >> 3;
              const uint32_t a = direct_s + (uint32_t)(gk >> 3) * 8192u + (uint32_t)(r >> 3) * 2048u + (uint32_t)(r & 7) * 128u +
                                 (uint32_t)(((gk & 7) ^ (r & 7)) << 4);
              for (uint32_t c = 0; c < (uint32_t)kCS; ++c) { st_cluster_v4(a, c, hi); st_cluster_v4(a + 1024u, c, lo); }
            } else {
              *reinterpret_cast<uint4*>(oprow_ptr(Lp.mimg, 0, P.nckh, P.Q, (long long)q0 + r, tid >> 3)) = hi;
              *reinterpret_cast<uint4*>(oprow_ptr(Lp.mimg, 1, P.nckh, P.Q, (long long)q0 + r, tid >> 3)) = lo;
            }
          }
        }
        workers_sync();
      }
      if (nl > 0 && tid == 0) S.nlong = 0;
      if (nl > 0) workers_sync();
    };
    // cells of projection chunk c of this level: TMEM lane = gate * 32 + unit, column = row of the chunk
    auto cells_chunk = [&](int c) {
      const int r0 = c * kRC;
      const int rows = min(kRC, n - r0);
      const uint32_t set = ct & 1u;
      const int row = tid >> 3, ug = tid & 7;                // cell math: 32 rows x 8 groups of 4 units per pass
      const int u = u0 + 4 * ug;
      const bool in_row = u + 4 <= P.Hq;
      // the aggregate's slice of the first pass. Short levels: every aggregate is visible behind the exchange barrier, the load
      // flies while the MMAs run. Pipelined levels: other CTAs' rows are only known to be there once the accumulators are (the
      // copy warp acquired the gather counters before the copies the MMAs consumed) — load behind that wait.
      float4 mv_next = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool mv_on = l > 0 && in_row && row < rows && 4 * ug < P.U;
      if (mv_on && !pipelined) mv_next = ldcg4(Lp.m32 + (size_t)(pos0 + r0 + row) * P.ldh + u);
      mbar_wait(&S.acc_full[set], (ct >> 1) & 1u);
      tc::fence_after_sync();
      if (mv_on && pipelined) mv_next = ldcg4(Lp.m32 + (size_t)(pos0 + r0 + row) * P.ldh + u);
      if (tr && tid == 0 && c == 0) tr[8] = clock64();
      const int qd = warp & 3, acc_id = warp >> 2;           // warps 0..3 read W_ih x, warps 4..7 W_hh m
#pragma unroll 1
      for (int s0 = 0; s0 < rows; s0 += kSub) {
        if (qd < 3) {
          float v[32];
          if (acc_id == 0 || l > 0) {
            tc::ld32(tmem + ((uint32_t)(32 * qd) << 16) + (uint32_t)(kColAcc + (acc_id ? kColAccH : 0) + s0) + set * (uint32_t)kColSet, v);
            tc::wait_ld();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;         // level 0: hidden state 0, nothing was projected
          }
          float* dst = &S.stage[acc_id][32 * qd + lane][0];
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[j] = v[j];
        }
        workers_sync();
        const int rr = s0 + row;
        if (rr < rows && 4 * ug < P.U) {
          const int p = pos0 + r0 + rr;
          const float4 mv = mv_next;
          if (l > 0 && in_row && rr + kSub < rows) mv_next = ldcg4(Lp.m32 + (size_t)(p + kSub) * P.ldh + u);
          float o[4];
          const float mm[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int uu = 4 * ug + k;
            const float xr = S.stage[0][uu][row], xz = S.stage[0][32 + uu][row], xn = S.stage[0][64 + uu][row];
            const float hr = S.stage[1][uu][row], hz = S.stage[1][32 + uu][row], hn = S.stage[1][64 + uu][row];
            const float rg = fast_sigmoid(xr + hr + S.bias[0][uu]);
            const float zg = fast_sigmoid(xz + hz + S.bias[1][uu]);
            const float ng = fast_tanh(xn + S.bias[2][uu] + rg * (hn + S.bias[3][uu]));
            o[k] = ng + zg * (mm[k] - ng);
          }
          if (in_row) *reinterpret_cast<float4*>(Lp.Hs + (size_t)p * P.ldh + u) = make_float4(o[0], o[1], o[2], o[3]);
          if (has_next && u < P.Kh) {
            uint32_t h0, h1, l0, l1;
            tc::split2(o[0], o[1], h0, l0);
            tc::split2(o[2], o[3], h1, l1);
            const long long q = (long long)q0 + r0 + rr;
            unsigned char* ph = oprow_ptr(Lp.himg, 0, P.nckh, P.Q, q, u >> 3) + (u & 4) * 2;
            unsigned char* pl = oprow_ptr(Lp.himg, 1, P.nckh, P.Q, q, u >> 3) + (u & 4) * 2;
            *reinterpret_cast<uint2*>(ph) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(pl) = make_uint2(l0, l1);
          }
        }
        if (has_next && s0 + kSub >= rows) asm volatile("fence.proxy.async.global;" ::: "memory");   // the next layer bulk-copies these rows
        workers_sync();
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.tmem_free[set]);
      // behind the last pass's closing workers_sync: the whole CTA has stored its slice of the chunk's states
      if (has_next && tid == 0) st_release_u32(P.flags + (size_t)item * kCS + rank, ct + 1u);
      ct += 1;
    };

    if (warp < kCWorkWarps) {
      if (pipelined) {
        // ---------------- pipelined level: rounds of 128 rows ----------------
        const uint32_t ct0 = ct;
        const int rounds = (n + kHalfWarps - 1) / kHalfWarps;
        int cells_done = 0;
#pragma unroll 1
        for (int t = 0; t < rounds; ++t) {
          const int r_end = min(n, (t + 1) * kHalfWarps);
          gather_rows(t * kHalfWarps + ghw, kHalfWarps, r_end);
          asm volatile("fence.proxy.async.global;" ::: "memory");
          workers_sync();
          if (tid == 0) st_release_u32(mflag + rank, ct0 + (uint32_t)((r_end + kRC - 1) / kRC));
          if (first_layer && Tn.y > 0) convert_x_rows(P, D, d, Tn.x, Tn.z, min(Tn.y, (t + 1) * kHalfWarps), t * kHalfWarps + ghw, kHalfWarps, hl);
          const int upto = t == 0 ? 0 : (t * kHalfWarps) / kRC;     // chunks whose rows were gathered a round ago
          for (; cells_done < upto; ++cells_done) cells_chunk(cells_done);
        }
        if (first_layer && Tn.y > rounds * kHalfWarps) convert_x_rows(P, D, d, Tn.x, Tn.z, Tn.y, rounds * kHalfWarps + ghw, kHalfWarps, hl);
        if (tr && tid == 0) tr[1] = tr[2] = clock64();
        for (; cells_done < nchunks; ++cells_done) cells_chunk(cells_done);
      } else {
        // ---------------- short level: gather, one exchange barrier, cells ----------------
        // the dependent-load chains of the two jobs would add up in a warp, so warps 0..3 gather and warps 4..7 convert the next
        // level's input rows at the same time
        const bool split = first_layer && exchange && n <= kHalfWarps / 2 && Tn.y <= kHalfWarps / 2;
        const int ghw_s = (((lane >> 4) * (kCWorkWarps / 2) + (warp & 3)) * kCS + rank);      // half-warp index among 64
        if (split && warp >= kCWorkWarps / 2) convert_x_rows(P, D, d, Tn.x, Tn.z, Tn.y, ghw_s, kHalfWarps / 2, hl);
        if (exchange) gather_rows(split ? (warp < kCWorkWarps / 2 ? ghw_s : n) : ghw, split ? kHalfWarps / 2 : kHalfWarps, n);
        if (first_layer && !split && Tn.y > 0) convert_x_rows_n<4>(P, D, d, Tn.x, Tn.z, Tn.y, ghw, kHalfWarps, hl);   // next level's input rows
      }
    } else if (warp == kCopyWarp) {
      copy_stages(early);
    } else {
      mma_stages(early);
    }
    if (!pipelined) {
      if (tr && tid == 0) tr[1] = clock64();
      if (exchange) {
        if (direct) asm volatile("fence.proxy.async.shared::cluster;" ::: "memory"); else asm volatile("fence.proxy.async.global;" ::: "memory");          // operand rows: ordinary stores here, bulk copies (async proxy) there
        cluster_sync_all();
      }
      if (tr && tid == 0) tr[2] = clock64();
      // ---------------- projection + cells ----------------
      if (warp == kCopyWarp) {
        copy_stages(total_stages);
        if (tr && lane == 0) tr[5] = clock64();
      } else if (warp == kMmaWarp) {
        mma_stages(total_stages);
        if (tr && lane == 0) tr[7] = clock64();
      } else {
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) cells_chunk(c);
      }
    }
    if (warp >= kCWorkWarps) ct += (uint32_t)nchunks;
    if (tr && tid == 0) tr[3] = clock64();
    // the states of this level: read by this cluster's next gather phase (other CTAs), by the next layer's bulk copies
    if (first_layer) asm volatile("fence.proxy.async.global;" ::: "memory");   // next level's input rows (bulk-copied); the states are read by plain loads
    cluster_sync_all();
    if (tr && tid == 0) tr[4] = clock64();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace dagnn

using namespace dagnn;

static size_t calign256(size_t x) { return (x + 255) / 256 * 256; }

namespace dagnn {

// rows of an operand-row plane: every (group, level) segment starts on a multiple of 8 rows, groups are spaced by 8 L rows
static int64_t cluster_q_rows(int64_t N, int32_t max_levels, int dirs, int layers) {
  const int gcap = kCMaxItems / (dirs * layers) > 0 ? kCMaxItems / (dirs * layers) : 1;
  return (N + 8 * (int64_t)max_levels * gcap + 64 + 7) / 8 * 8;
}

bool cluster_path_supported(int dirs, int layers, int Din, int H, int nvid) {
  (void)nvid;
  return H >= 1 && H <= kCMaxH && Din >= 1 && Din <= kCMaxH && dirs * layers <= kCMaxItems;
}

size_t cluster_workspace_bytes(int dirs, int layers, int Din, int H, int64_t N, int32_t max_levels) {
  if (!cluster_path_supported(dirs, layers, Din, H, 0)) return 0;
  const int64_t Q = cluster_q_rows(N, max_levels, dirs, layers);
  const int nckh = (H + 63) / 64, nckx = (Din + 63) / 64;
  const size_t plane_h = calign256((size_t)2 * nckh * Q * 128), plane_x = calign256((size_t)2 * nckx * Q * 128);
  const size_t ldh = (size_t)round_up(H, 4);
  size_t b = 1024;                                                          // flags
  b += calign256((size_t)kCMaxItems * (max_levels + 1) * sizeof(int4));     // level tables
  b += (size_t)dirs * plane_x;                                              // X operand rows per direction
  b += (size_t)dirs * layers * (2 * plane_h + calign256((size_t)N * ldh * sizeof(float)));   // h rows, m rows, m fp32
  return b + 1024;
}

static int g_cluster_max[kMaxDevices] = {0};       // resident 8-CTA clusters of k_sweep_cluster per device (0 = path unavailable)

int cluster_forward(const DagnnSweepArgs* A, cudaStream_t st, bool* handled) {
  *handled = false;
  const DagnnSchedule* S = A->sched;
  const int dirs = S->dirs, layers = A->num_layers, H = A->H;
  if (!cluster_path_supported(dirs, layers, A->Din, H, A->nvid)) return DAGNN_OK;
  static PerDeviceOnce once;
  int dev = 0;
  if (int rc = per_device_once(once, &dev, [&](int dv) {
        cudaError_t e = cudaFuncSetAttribute(k_sweep_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCSmemBytes);
        if (e != cudaSuccess) { cudaGetLastError(); g_cluster_max[dv] = 0; return (int)DAGNN_OK; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kCS * 32); cfg.blockDim = dim3(kCThreads); cfg.dynamicSmemBytes = kCSmemBytes;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        e = cudaOccupancyMaxActiveClusters(&nc, k_sweep_cluster, &cfg);
        if (e != cudaSuccess) { cudaGetLastError(); nc = 0; }
        g_cluster_max[dv] = nc > kCMaxItems ? kCMaxItems : nc;
        return (int)DAGNN_OK;
      }))
    return rc;
  const int items = dirs * layers;
  const int cmax = g_cluster_max[dev];
  if (cmax < items) return DAGNN_OK;                 // not enough resident clusters for the layer pipeline: grid-wide sweep
  int G = cmax / items;
  if (S->B < 1 || !S->gptr) G = 1;
  if (G > S->B && S->B >= 1) G = (int)S->B;
  if (G < 1) G = 1;

  ClusterP P;
  memset(&P, 0, sizeof(P));
  DagnnPackLayout lay[DAGNN_MAX_LAYERS];
  for (int i = 0; i < layers; ++i)
    if (int rc = dagnn_pack_layout(i == 0 ? A->Din : H, H, A->nvid, i == 0, i + 1 == layers, &lay[i])) return rc;
  P.dirs = dirs; P.layers = layers; P.G = G; P.H = H; P.Hq = lay[0].Hq; P.Mc = lay[0].Mc; P.HP = lay[0].HP; P.nvid = A->nvid;
  P.use_ea = A->use_edge_attr; P.Din = A->Din; P.N = (int)S->N; P.B = (int)S->B;
  P.U = round_up(ceil_div(H, kCS), 4);
  P.Kh = round_up(H, 16); P.Kx = round_up(A->Din, 16);
  P.nckh = ceil_div(P.Kh, 64); P.nckx = ceil_div(P.Kx, 64);
  P.vec_x = ((A->ldx & 3) == 0 && (A->Din & 3) == 0 && ((uintptr_t)A->X & 15) == 0) ? 1 : 0;
  P.max_levels = S->max_levels;
  P.ldh = A->ldh; P.ldx = A->ldx; P.Q = cluster_q_rows(S->N, S->max_levels, dirs, layers);
  P.X = A->X; P.summary = S->summary; P.gptr = S->gptr; P.gdepth = S->gdepth;
  P.trace = static_cast<long long*>(A->trace);
  char* ws = static_cast<char*>(A->workspace);
  P.flags = reinterpret_cast<unsigned int*>(ws); ws += 1024;
  P.tab = reinterpret_cast<int4*>(ws); ws += calign256((size_t)kCMaxItems * (S->max_levels + 1) * sizeof(int4));
  const size_t plane_h = calign256((size_t)2 * P.nckh * P.Q * 128), plane_x = calign256((size_t)2 * P.nckx * P.Q * 128);
  const size_t m32_bytes = calign256((size_t)S->N * A->ldh * sizeof(float));
  for (int d = 0; d < dirs; ++d) { P.ximg[d] = reinterpret_cast<unsigned char*>(ws); ws += plane_x; }
  for (int d = 0; d < dirs; ++d) {
    P.dir[d].perm = S->perm[d]; P.dir[d].rowptr = S->rowptr[d]; P.dir[d].col = S->col[d];
    P.dir[d].eattr = A->use_edge_attr ? S->eattr[d] : nullptr; P.dir[d].lvl_off = S->lvl_off[d];
    for (int i = 0; i < layers; ++i) {
      const DagnnPackLayout& Lz = lay[i];
      const float* pk = A->packed[d][i];
      CLay& q = P.lay[d][i];
      q.Hs = A->Hs[d][i]; q.bias = pk + Lz.bias_off; q.wk = pk + Lz.wk_off; q.attnc = pk + Lz.attnc_off; q.vidk = pk + Lz.vidk_off;
      q.w_hh = reinterpret_cast<const __half*>(pk + Lz.imgh_off);
      if (i == 0) {
        q.w_ih = reinterpret_cast<const __half*>(pk + Lz.imgx_off); q.ih_col0 = 0; q.ih_nck = Lz.Kin64 / 64;
      } else {                                       // W_ih of layer i rides behind W_hh of layer i - 1 (pack.cu)
        q.w_ih = reinterpret_cast<const __half*>(A->packed[d][i - 1] + lay[i - 1].imgh_off); q.ih_col0 = Lz.Mc; q.ih_nck = Lz.Kh64 / 64;
      }
      q.himg = reinterpret_cast<unsigned char*>(ws); ws += plane_h;
      q.mimg = reinterpret_cast<unsigned char*>(ws); ws += plane_h;
      q.m32 = reinterpret_cast<float*>(ws); ws += m32_bytes;
    }
  }
  if ((size_t)(ws - static_cast<char*>(A->workspace)) > A->workspace_bytes)
    return set_err(DAGNN_E_WORKSPACE, "sweep: workspace too small (dagnn_sweep_workspace_bytes)");
  DAGNN_CUDA_OK(cudaMemsetAsync(A->workspace, 0, 1024, st));       // level counters
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCS * items * G); cfg.blockDim = dim3(kCThreads); cfg.dynamicSmemBytes = kCSmemBytes; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  DAGNN_CUDA_OK(cudaLaunchKernelEx(&cfg, k_sweep_cluster, P));
  *handled = true;
  return check_launch("k_sweep_cluster");
}

}  // namespace dagnn
