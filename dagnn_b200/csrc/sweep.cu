// The DAGNN level sweep as ONE persistent cooperative kernel (one CTA per SM), "project, then aggregate".
//
// Algebra (exact; only the summation order changes — DESIGN.md §3.2): the GRU cell of layer i at node v needs
//   W_ih^i inp_v   and   W_hh^i m_v,   m_v = sum_e alpha_e h^i_{j(e)}   (alpha = softmax of the separable attention score).
// Both are linear, so  W_hh^i m_v = sum_e alpha_e (W_hh^i h^i_j)  and  inp_v = h^{i-1}_v  (or x_v for i = 0):
//   projection  P^i_j  = W_hh^i h^i_j      — once per node, when its state is produced, for ALL its future successors;
//   projection  Gi^i_v = W_ih^i h^{i-1}_v  — once per node (Gi^0 = W_ih^0 x_v for all nodes before the first level).
// h^i_j is the operand of both P^i and Gi^{i+1}: ONE dense GEMM over the contiguous, just-written rows of a level
// ([rows, H] x [H, 6H], no gather, no attention inside the GEMM) on the 5th-generation tensor cores, and the irregular
// part (CSR gather, softmax, weighted sum, gate math) becomes a bandwidth-bound pass with one warp per node that
// never touches a weight matrix.
//
// Wavefront step s runs every (direction d, layer i, level l) with l + i == s in two phases separated by grid barriers:
//   gate phase : one warp per node of the step's segments: softmax weights of its in-edges from the key scores sk[j]
//                (scalar gathers), a = Gi^i_v + sum_e alpha_e P^i_j, m_v = sum_e alpha_e h^i_j, GRU pointwise,
//                h^i_v and sk[v] = wk . h^i_v stored. Level 0: no in-edges are read (hidden = 0).
//   proj phase : tiles of <= 256 of the rows just produced x 64..256 output columns, dealt round-robin to the CTAs:
//                8 builder warps load the fp32 state rows (coalesced, one stage ahead), split them into fp16 hi/lo and
//                store them straight into swizzled shared memory; the issuer thread streams the pre-swizzled hi/lo
//                weight image with cp.async.bulk (UBLKCP) into a ring and issues tcgen05.mma kind::f16 (M = 128,
//                N = 64..256, fp16 x 3 split, fp32 accumulators in TMEM); the builder warps drain TMEM into P / Gi rows.
// States written in one phase are read in later phases by OTHER CTAs: all such reads use ld.global.cg (L2), the
// barrier is the cooperative-groups pattern (bar.sync; fence; atomic; spin on ld.acquire; bar.sync).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "sync.cuh"
#include "tc.cuh"

namespace dagnn {

// sweep_cluster.cu: the cluster-resident sweep (H, Din <= 256); `handled` = false leaves the launch to the grid-wide kernel below
bool cluster_path_supported(int dirs, int layers, int Din, int H, int nvid);
size_t cluster_workspace_bytes(int dirs, int layers, int Din, int H, int64_t N, int32_t max_levels);
int cluster_forward(const DagnnSweepArgs* A, cudaStream_t st, bool* handled);

constexpr int kBuilderWarps = 8;
constexpr int kBuilders = kBuilderWarps * 32;            // 256 threads: gate phase, operand build, epilogue
constexpr int kThreads = kBuilders + 32;                 // + one warp whose lane 0 issues bulk copies and MMAs
constexpr int kNR = 128 * 8 / kBuilders;                 // rows of a 128-row operand stage per builder thread
constexpr int kRStride = 128 / kNR;
constexpr int kAStageBytes = 2 * 128 * tc::ROW_BYTES;    // hi + lo tile of 128 rows x 64 k  = 32 KB
constexpr int kNAS = 2;                                  // operand stages
constexpr int kBlkBytes = 2 * 64 * tc::ROW_BYTES;        // hi + lo tile of one 64-column block x 64 k = 16 KB
constexpr int kBRegionBytes = 8 * kBlkBytes;             // weight ring: 2 x 4 blocks, 4 x 2 blocks or 8 x 1 block = 128 KB
constexpr int kNBBar = 8;
constexpr int kMaxSeg = DAGNN_MAX_DIRS * DAGNN_MAX_LAYERS;
constexpr int kTmemCols = 512;                           // 2 sub-tiles x 256 columns
constexpr int kMaxChunks = 16;                           // k chunks of a tile that may use compact operand stages
constexpr int kCoopEdges = 10;     // gate phase: nodes with more in-edges are aggregated by the whole CTA (split by columns)
constexpr int kMaxCoop = 48;       // ... per CTA and gate phase; beyond that a warp does the node alone
constexpr int kScanRows = 4096;    // steps with at most this many rows are latency-bound: their long in-edge lists are found a
constexpr int kMaxHeavy = 128;     // phase ahead (at most this many) and get a CTA of their own, from the start of the phase
constexpr int kMaxSmem = 232448;                         // 227 KB opt-in limit per CTA on sm_100

struct DirP {
  const int* perm;      // position -> node id
  const int* rowptr;    // [N+1] CSR rows by position
  const int* col;       // [E] neighbour position
  const float* eattr;   // [E,2] in CSR order or nullptr
  const int* lvl_off;   // [max_levels+1] first position of each level
};
struct LayP {
  float* Hs;            // H[d][i], [N, ldh] position order
  float* sk;            // [N] key score wk . h of every node
  float* Pm;            // [N, Mc] W_hh^i h^i_j  (columns gate * Hq + unit)
  float* Gi;            // [N, Mc] W_ih^i inp_v
  const int* gi_perm;   // non-null: Gi is in NODE order (layer 0 projected from the X image): row of position p = gi_perm[p]
  const float* bias;    // [4][HP]
  const float* wk;      // [HP]
  const float* attnc;   // [4]
  const float* vidk;    // [nvid]
  unsigned char* aimg;  // operand image of H[d][i]: fp16 hi/lo tiles in tcgen05 layout, written by the gate phase
  const __half* imgh;   // [W_hh^i ; W_ih^{i+1}] image (pack.cu)
  const __half* imgx;   // W_ih^0 image (layer 0 only)
};
struct SweepP {
  int dirs, layers, H, Hq, Mc, HP, nvid, use_ea;
  int Din0, nckx, nckh;       // layer-0 input width, 64-k chunks of the layer-0 input / of a state operand
  int vec_x, N, Kh64;
  long long ldh, ldx;
  const float* X;             // [N, ldx] node order (rows through perm)
  const unsigned char* ximg;  // optional operand image of X in node order: the first projection then runs in node order
  const int* summary;         // [0] number of levels of direction 0, [2] schedule status
  unsigned int* bar;          // grid barrier counter (zeroed by the launcher)
  long long* trace;           // optional [steps + 1][256][16] clock64 stamps, nullptr = off
  DirP dir[DAGNN_MAX_DIRS];
  LayP lay[DAGNN_MAX_DIRS][DAGNN_MAX_LAYERS];
};

struct Seg {
  int pos0, n, ntile, base;
  int bmod;                   // base % grid size: the tile loop of a CTA starts without a division
  int ncbt;                   // column tiles per row tile
  uint32_t rcbt;              // ceil(2^32 / ncbt), ncbt > 1: t / ncbt == __umulhi(t, rcbt) for t * ncbt < 2^32
};
// no runtime integer division in the per-step paths: every thread runs them, ~130 times per forward, and a division is
// ~50 instructions of dependent latency and of instruction-cache footprint
__device__ __forceinline__ void seg_di(const SweepP& P, int q, int& d, int& i) {     // q = d * layers + i, d < 2
  d = (q >= P.layers) ? 1 : 0;
  i = d ? q - P.layers : q;
}
__device__ __forceinline__ int blk_shift(int nbc) { return nbc >> 1; }               // log2 of 1, 2, 4
struct StepTab {
  Seg seg[kMaxSeg];           // projection of step s: rows to project (empty when nothing downstream needs them)
  int gpos0[kMaxSeg], gn[kMaxSeg];   // gate phase of step s: first position and rows of every segment's level
  int nbc, nst;               // 64-column blocks per tile (1, 2, 4), 128-row sub-tiles per tile
};
struct HeavyTab {
  int n;                      // gate phase of step s: nodes with more than kCoopEdges in-edges, -1 = not scanned
  int node[kMaxHeavy][2];     // ... their (segment, position), segment-major, positions ascending
};
constexpr int kEpiLd = 20;    // floats per scratch row: 16 columns + pad (16-byte aligned rows)
struct SmemTail {
  float epi[kBuilderWarps][32 * kEpiLd];   // projection epilogue: 32 rows x 16 columns per warp, transposed for row-contiguous stores
  StepTab tab[2];             // this proj phase's and the next one's segment tables
  int coop[kMaxCoop][2];      // gate phase: (segment, position) of the nodes this CTA aggregates cooperatively
  int ncoop;
  uint64_t a_full[kNAS], a_empty[kNAS], b_full[kNBBar], b_empty[kNBBar], acc_full;
  uint64_t c_full[kMaxChunks];      // small tiles: operand chunk c has landed
  uint64_t tmem_free;               // the worker warps have drained the accumulators of the previous tile
  uint32_t tmem_slot;
  HeavyTab heavy[2];                // long in-edge lists of this gate phase and the next one
};
constexpr size_t kSmemBytes = 1024 + (size_t)kNAS * kAStageBytes + kBRegionBytes + sizeof(SmemTail);
static_assert(kSmemBytes <= (size_t)kMaxSmem, "shared memory plan exceeds the 227 KB opt-in limit");

// all CTAs of the (cooperative, co-resident) grid; `target` = arrivals expected so far
// Only the builder warps write global memory, so they alone gate the arrival (named barrier 1); the issuer warp, which may
// still be queueing weight prefetches, joins at the closing CTA-wide barrier.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
  // state rows written with ordinary stores are read after the barrier by other CTAs' bulk copies (async proxy)
  asm volatile("fence.proxy.async;" ::: "memory");
  if (threadIdx.x < kBuilders) asm volatile("bar.sync 1, %0;" ::"n"(kBuilders) : "memory");
  if (threadIdx.x == 0) {
    // release: the CTA's global stores (ordered before this thread by the named barrier) are visible at GPU scope before
    // the arrival counts; no return value, so the first poll goes out right behind it instead of after a round trip
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
    while (ld_acquire_u32(bar) < target) {}
  }
  __syncthreads();
}
// ------------------------------------------------------------------------------------------------------------
// gate phase: one warp per node. Lane l owns the units 4 l + 128 j (+ 0..3), j < J, of a 128 J wide column pass.
// The chain of dependent L2 round trips is what a node costs (~0.3-0.5 us each), so everything that does not depend
// on the previous load is issued with it: row pointers + the node's own Gi row; then the in-edge list; then the key
// scores AND the projected rows of the first in-edges together; the softmax runs while those rows are in flight.
// ------------------------------------------------------------------------------------------------------------
struct GateAcc {                   // per lane: sum_e alpha_e * (P_r, P_z, P_n, h) for the lane's units
  float4 r, z, n, m;
};

template <int J>
struct GateRows {                  // projected + state rows of one in-edge, lane's units
  float4 r[J], z[J], n[J], h[J];
};
template <int J>
__device__ __forceinline__ void load_rows(GateRows<J>& R, const float* __restrict__ Pm, const float* __restrict__ Hs, int sp, bool on,
                                          int Mc, long long ldh, int Hq, int ub, int lane) {
  // one warp-uniform branch per in-edge (an in-edge that adds nothing is not read: its row may not have been written
  // yet), nothing predicated inside: 4 J independent 16-byte loads per lane that go out back to back. The units past Hq
  // of the last column pass re-read the first ones (gate_finish ignores them).
  const float* pr = Pm + (size_t)sp * Mc;
  const float* hr = Hs + (size_t)sp * ldh;
  if (on) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      int u = ub + 4 * lane + 128 * j;
      u = (u < Hq) ? u : 4 * lane;
      R.r[j] = ldcg4(pr + u);
      R.z[j] = ldcg4(pr + Hq + u);
      R.n[j] = ldcg4(pr + 2 * Hq + u);
      R.h[j] = ldcg4(hr + u);
    }
  } else {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < J; ++j) { R.r[j] = z4; R.z[j] = z4; R.n[j] = z4; R.h[j] = z4; }
  }
}

// Scheduling fence. ptxas sinks the loads of a row next to the FMAs that consume it to save registers; with in-order
// issue the warp then stalls on the first FMA before the next loads have gone out — three or four dependent L2 round
// trips per in-edge pair instead of one. rows_zero() is 0 computed from one word of every 16-byte load (x * 0, not
// foldable: x could be a NaN); adding it to the softmax weights makes every FMA wait for every load, so nothing is gained
// by delaying a load and they issue back to back. (A non-finite state poisons the sums either way.)
template <int J>
__device__ __forceinline__ float rows_zero(const GateRows<J>& R, float t) {
#ifdef DAGNN_NO_FENCE
  return t;
#endif
#pragma unroll
  for (int j = 0; j < J; ++j) {
    t = fmaf(R.r[j].x, 0.f, t); t = fmaf(R.z[j].x, 0.f, t); t = fmaf(R.n[j].x, 0.f, t); t = fmaf(R.h[j].x, 0.f, t);
  }
  return t;
}

template <int J>
__device__ __forceinline__ void add_rows(GateAcc (&A)[J], float a, const GateRows<J>& R) {
#pragma unroll
  for (int j = 0; j < J; ++j) { fma4(A[j].r, a, R.r[j]); fma4(A[j].z, a, R.z[j]); fma4(A[j].n, a, R.n[j]); fma4(A[j].m, a, R.h[j]); }
}

// score of in-edge e (lane-private): key score of the predecessor if it sits in an earlier level, else 0 — such an edge
// keeps its softmax mass and adds a zero row — plus the edge-type / vertex-id terms
__device__ __forceinline__ float edge_score(const SweepP& P, const DirP& D, const LayP& Lp, int e, int pos0, float ca0, float ca1,
                                            bool use_ea, int& sp) {
  sp = D.col[e];
  float sc = (sp < pos0) ? __ldcg(Lp.sk + sp) : 0.f;
  if (use_ea) {
    const float2 ea = __ldg(reinterpret_cast<const float2*>(D.eattr) + e);
    sc += ca0 * ea.x + ca1 * ea.y;
  }
  if (P.nvid > 0) sc += __ldg(Lp.vidk + (D.perm[sp] % P.nvid));
  return sc;
}

// GRU pointwise for the lane's units of one column pass; returns the lane's part of wk . h
// The new state row is stored twice: fp32 (successors' gathers, readout) and as fp16 hi / lo halves straight into the
// operand image of the projection GEMM (level-aligned 128-row tiles x 64-k chunks in the swizzled tcgen05 layout), so
// that the GEMM's operand side is a plain bulk copy.
template <int J>
struct GiRow {                     // the node's own input projection, lane's units
  float4 r[J], z[J], n[J];
};
template <int J>
__device__ __forceinline__ void load_gi(GiRow<J>& Gr, const SweepP& P, const LayP& Lp, int p, int ub, int lane) {
  const float* __restrict__ gi = Lp.Gi + (size_t)(Lp.gi_perm ? Lp.gi_perm[p] : p) * P.Mc;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int u = ub + 4 * lane + 128 * j;
    Gr.r[j] = z4; Gr.z[j] = z4; Gr.n[j] = z4;
    if (u < P.Hq) { Gr.r[j] = ldcg4(gi + u); Gr.z[j] = ldcg4(gi + P.Hq + u); Gr.n[j] = ldcg4(gi + 2 * P.Hq + u); }
  }
}
template <int J>
__device__ __forceinline__ float gate_finish(const SweepP& P, const LayP& Lp, int p, int pos0, int lvl, int ub, int lane,
                                             const GateAcc (&A)[J], const GiRow<J>& Gr) {
  const int Hq = P.Hq, HP = P.HP;
  const int rel = p - pos0;
  unsigned char* itile = Lp.aimg + (size_t)((pos0 >> 7) + lvl + (rel >> 7)) * P.nckh * kAStageBytes;
  const int rin = rel & 127;
  float skacc = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int u = ub + 4 * lane + 128 * j;
    if (u >= P.Kh64) continue;
    unsigned char* ihi = itile + (size_t)(u >> 6) * kAStageBytes + tc::tile_off(rin, (u & 63) >> 3) + (u & 4) * 2;
    if (u >= Hq) {                                      // k padding of the last chunk: zeros (NaN bit patterns x 0 would poison the GEMM)
      *reinterpret_cast<uint2*>(ihi) = make_uint2(0u, 0u);
      *reinterpret_cast<uint2*>(ihi + 128 * tc::ROW_BYTES) = make_uint2(0u, 0u);
      continue;
    }
    const float4 gr = Gr.r[j], gz = Gr.z[j], gn = Gr.n[j];
    const float4 br = __ldg(reinterpret_cast<const float4*>(Lp.bias + u));
    const float4 bz = __ldg(reinterpret_cast<const float4*>(Lp.bias + HP + u));
    const float4 bi = __ldg(reinterpret_cast<const float4*>(Lp.bias + 2 * HP + u));
    const float4 bh = __ldg(reinterpret_cast<const float4*>(Lp.bias + 3 * HP + u));
    const float4 wk = __ldg(reinterpret_cast<const float4*>(Lp.wk + u));
    float4 o;
#define DAGNN_GATE(c)                                                          \
  {                                                                            \
    const float rg = fast_sigmoid(gr.c + A[j].r.c + br.c);                     \
    const float zg = fast_sigmoid(gz.c + A[j].z.c + bz.c);                     \
    const float ng = fast_tanh(gn.c + bi.c + rg * (A[j].n.c + bh.c));          \
    o.c = ng + zg * (A[j].m.c - ng);                                           \
  }
    DAGNN_GATE(x) DAGNN_GATE(y) DAGNN_GATE(z) DAGNN_GATE(w)
#undef DAGNN_GATE
    *reinterpret_cast<float4*>(Lp.Hs + (size_t)p * P.ldh + u) = o;
    {
      const __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      const __half2 l0 = __floats2half2_rn(o.x - f0.x, o.y - f0.y), l1 = __floats2half2_rn(o.z - f1.x, o.w - f1.y);
      *reinterpret_cast<uint2*>(ihi) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      *reinterpret_cast<uint2*>(ihi + 128 * tc::ROW_BYTES) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    }
    skacc += o.x * wk.x + o.y * wk.y + o.z * wk.z + o.w * wk.w;       // units >= H: zero weights and biases -> o = 0
  }
  return skacc;
}

// -DDAGNN_GATE_TRACE (profiling builds only): lane 0 of warp 0 stamps the stages of the first node it handles in a step
// into trace slots 10..15; the stamp is ordered behind the value named as its dependency
#ifdef DAGNN_GATE_TRACE
#define DAGNN_GT(slot, dep)                                                                         \
  if (gt) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) : "r"(dep) : "memory"); gt[slot] = t_; }
#define DAGNN_GT_ARG , long long* gt
#else
#define DAGNN_GT(slot, dep)
#define DAGNN_GT_ARG
#endif

// one warp, one node
template <int J>
__device__ __forceinline__ void gate_row(const SweepP& P, const DirP& D, const LayP& Lp, int p, int pos0, int lvl, int lane DAGNN_GT_ARG) {
  const bool level0 = lvl == 0;
  const int Hq = P.Hq, Mc = P.Mc;
  const long long ldh = P.ldh;
  const float* __restrict__ Pm = Lp.Pm;
  const float* __restrict__ Hs = Lp.Hs;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool use_ea = P.use_ea && D.eattr != nullptr;
  const float ca0 = use_ea ? __ldg(Lp.attnc) : 0.f, ca1 = use_ea ? __ldg(Lp.attnc + 1) : 0.f;
  int e0 = 0, e1 = 0;
  if (!level0) { e0 = D.rowptr[p]; e1 = D.rowptr[p + 1]; }
  const int ne = e1 - e0;
  DAGNN_GT(11, ne)
  float skacc = 0.f;
  if (ne <= 32) {
    // ---- the common case: all in-edges in one round, one lane per edge
    int my_sp = 0;
    float my_sc = -INFINITY;
    if (lane < ne) my_sc = edge_score(P, D, Lp, e0 + lane, pos0, ca0, ca1, use_ea, my_sp);
    const bool my_valid = lane < ne && my_sp < pos0;
    DAGNN_GT(12, __float_as_int(my_sc))
#pragma unroll 1
    for (int ub = 0; ub < Hq; ub += 128 * J) {
      GateAcc A[J];
#pragma unroll
      for (int j = 0; j < J; ++j) { A[j].r = z4; A[j].z = z4; A[j].n = z4; A[j].m = z4; }
      GiRow<J> Gr;
      if constexpr (J <= 2) load_gi<J>(Gr, P, Lp, p, ub, lane);     // independent of the in-edges: in flight from the start
      if (ne > 0) {
        // rows of the first in-edges go in flight before the softmax (their address needs the edge list only)
        constexpr bool kTwo = J <= 2;                      // two in-edges in flight per lane when the registers allow it
        GateRows<J> R0;
        GateRows<kTwo ? J : 1> R1;
        const int sp0 = __shfl_sync(0xffffffffu, my_sp, 0), sp1 = __shfl_sync(0xffffffffu, my_sp, 1);
        const bool v0 = __shfl_sync(0xffffffffu, (int)my_valid, 0) != 0, v1 = __shfl_sync(0xffffffffu, (int)my_valid, 1) != 0;
        load_rows<J>(R0, Pm, Hs, sp0, v0, Mc, ldh, Hq, ub, lane);
        if constexpr (kTwo) load_rows<J>(R1, Pm, Hs, sp1, v1, Mc, ldh, Hq, ub, lane);
        const float mx = warp_max(my_sc);
        const float ex = (lane < ne) ? expf(my_sc - mx) : 0.f;
        const float inv = 1.f / (warp_sum(ex) + 1e-16f);
        const float my_a = my_valid ? ex * inv : 0.f;
        DAGNN_GT(13, __float_as_int(R0.r[0].x))
        add_rows<J>(A, __shfl_sync(0xffffffffu, my_a, 0), R0);
        if constexpr (kTwo) {
          add_rows<J>(A, __shfl_sync(0xffffffffu, my_a, 1), R1);
          for (int q = 2; q < ne; q += 2) {
            const float a0 = __shfl_sync(0xffffffffu, my_a, q), a1 = __shfl_sync(0xffffffffu, my_a, (q + 1) & 31);
            const int s0 = __shfl_sync(0xffffffffu, my_sp, q), s1 = __shfl_sync(0xffffffffu, my_sp, (q + 1) & 31);
            load_rows<J>(R0, Pm, Hs, s0, a0 != 0.f, Mc, ldh, Hq, ub, lane);
            load_rows<J>(R1, Pm, Hs, s1, a1 != 0.f && q + 1 < ne, Mc, ldh, Hq, ub, lane);
            const float t0 = rows_zero<J>(R1, rows_zero<J>(R0, 0.f));
            add_rows<J>(A, a0 + t0, R0);
            add_rows<J>(A, ((q + 1 < ne) ? a1 : 0.f) + t0, R1);
          }
        } else {
          (void)sp1; (void)v1;
          for (int q = 1; q < ne; ++q) {
            const float a0 = __shfl_sync(0xffffffffu, my_a, q);
            const int s0 = __shfl_sync(0xffffffffu, my_sp, q);
            if (a0 == 0.f) continue;
            load_rows<J>(R0, Pm, Hs, s0, true, Mc, ldh, Hq, ub, lane);
            add_rows<J>(A, a0 + rows_zero<J>(R0, 0.f), R0);
          }
        }
      }
      if constexpr (J > 2) load_gi<J>(Gr, P, Lp, p, ub, lane);
      DAGNN_GT(14, __float_as_int(Gr.r[0].x + A[0].r.x))
      skacc += gate_finish<J>(P, Lp, p, pos0, lvl, ub, lane, A, Gr);
    }
  } else {
    // ---- long edge lists (a warp alone): softmax statistics first, 32 in-edges per round, then the weighted rows
    float sum = 0.f, mx = -INFINITY;
    for (int eb = e0; eb < e1; eb += 32) {
      const int e = eb + lane;
      int sp;
      const float sc = (e < e1) ? edge_score(P, D, Lp, e, pos0, ca0, ca1, use_ea, sp) : -INFINITY;
      const float mnew = fmaxf(mx, warp_max(sc));
      sum = sum * expf(mx - mnew) + warp_sum((e < e1) ? expf(sc - mnew) : 0.f);
      mx = mnew;
    }
    const float inv = 1.f / (sum + 1e-16f);
#pragma unroll 1
    for (int ub = 0; ub < Hq; ub += 128 * J) {
      GateAcc A[J];
#pragma unroll
      for (int j = 0; j < J; ++j) { A[j].r = z4; A[j].z = z4; A[j].n = z4; A[j].m = z4; }
      for (int eb = e0; eb < e1; eb += 32) {
        const int e = eb + lane;
        int my_sp = 0;
        float my_a = 0.f;
        if (e < e1) {
          const float sc = edge_score(P, D, Lp, e, pos0, ca0, ca1, use_ea, my_sp);
          my_a = (my_sp < pos0) ? expf(sc - mx) * inv : 0.f;
        }
        const int nq = min(32, e1 - eb);
        for (int q = 0; q < nq; ++q) {
          const float a = __shfl_sync(0xffffffffu, my_a, q);
          const int sp = __shfl_sync(0xffffffffu, my_sp, q);
          if (a == 0.f) continue;                            // warp-uniform
          GateRows<J> R0;
          load_rows<J>(R0, Pm, Hs, sp, true, Mc, ldh, Hq, ub, lane);
          add_rows<J>(A, a + rows_zero<J>(R0, 0.f), R0);
        }
      }
      GiRow<J> Gr;
      load_gi<J>(Gr, P, Lp, p, ub, lane);
      skacc += gate_finish<J>(P, Lp, p, pos0, lvl, ub, lane, A, Gr);
    }
  }
  skacc = warp_sum(skacc);
  if (lane == 0) Lp.sk[p] = skacc;
  DAGNN_GT(15, __float_as_int(skacc))
}

// the whole CTA (kBuilderWarps warps), one node with a long in-edge list, split by COLUMNS: thread t owns the units
// t + 256 k, so a predecessor costs a lane four 4-byte loads (P_r, P_z, P_n, h of its unit; 128 contiguous bytes per warp),
// eight predecessors are in flight per lane, nothing is reduced across warps and the GRU pointwise work is one unit per
// thread. Every warp derives the softmax weights itself (one lane per in-edge, 32 consecutive in-edges per round; the
// scores of the only round stay in registers). `part` = kBuilderWarps floats of shared memory for the key score.
__device__ __noinline__ void gate_node_cta(const SweepP& P, const DirP& D, const LayP& Lp, int p, int pos0, int lvl, float* part,
                                           int warp, int lane DAGNN_GT_ARG) {
  DAGNN_GT(10, lane)
  const int Hq = P.Hq, HP = P.HP, Mc = P.Mc;
  const long long ldh = P.ldh;
  const bool use_ea = P.use_ea && D.eattr != nullptr;
  const float ca0 = use_ea ? __ldg(Lp.attnc) : 0.f, ca1 = use_ea ? __ldg(Lp.attnc + 1) : 0.f;
  const int e0 = D.rowptr[p], e1 = D.rowptr[p + 1];
  const bool one_round = e1 - e0 <= 32;
  float sum = 0.f, mx = -INFINITY, sc0 = -INFINITY;
  int sp0 = 0;
  for (int eb = e0; eb < e1; eb += 32) {
    const int e = eb + lane;
    sc0 = (e < e1) ? edge_score(P, D, Lp, e, pos0, ca0, ca1, use_ea, sp0) : -INFINITY;
    const float mnew = fmaxf(mx, warp_max(sc0));
    sum = sum * expf(mx - mnew) + warp_sum((e < e1) ? expf(sc0 - mnew) : 0.f);
    mx = mnew;
  }
  const float inv = 1.f / (sum + 1e-16f);
  DAGNN_GT(11, __float_as_int(inv))
  const int rel = p - pos0, rin = rel & 127;
  unsigned char* itile = Lp.aimg + (size_t)((pos0 >> 7) + lvl + (rel >> 7)) * P.nckh * kAStageBytes;
  const float* __restrict__ gi = Lp.Gi + (size_t)(Lp.gi_perm ? Lp.gi_perm[p] : p) * Mc;
  float skacc = 0.f;
#pragma unroll 1
  for (int ub = 0; ub < P.Kh64; ub += kBuilders) {
    const int u = ub + 32 * warp + lane;
    const bool live = u < Hq;
    const int uc = live ? u : 0;                        // idle threads of the last pass read unit 0 and drop it
    const float gr = __ldcg(gi + uc), gz = __ldcg(gi + Hq + uc), gn = __ldcg(gi + 2 * Hq + uc);   // in flight from the start
    float ar = 0.f, az = 0.f, an = 0.f, am = 0.f;
    for (int eb = e0; eb < e1; eb += 32) {
      const int e = eb + lane;
      int my_sp = sp0;
      float sc = sc0;
      if (!one_round) sc = (e < e1) ? edge_score(P, D, Lp, e, pos0, ca0, ca1, use_ea, my_sp) : -INFINITY;
      const float my_a = (e < e1 && my_sp < pos0) ? expf(sc - mx) * inv : 0.f;
      if (my_a == 0.f) my_sp = 0;                       // adds nothing: row 0 (written, finite) stands in, its own may not be
      const int nq = min(32, e1 - eb);
      for (int q = 0; q < nq; q += 8) {
        float a[8], vr[8], vz[8], vn[8], vh[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          a[k] = __shfl_sync(0xffffffffu, my_a, (q + k) & 31);
          const int sp = __shfl_sync(0xffffffffu, my_sp, (q + k) & 31);
          const float* pr = Lp.Pm + (size_t)sp * Mc + uc;
          vr[k] = __ldcg(pr); vz[k] = __ldcg(pr + Hq); vn[k] = __ldcg(pr + 2 * Hq);
          vh[k] = __ldcg(Lp.Hs + (size_t)sp * ldh + uc);
        }
        float t0 = 0.f;                                 // scheduling fence, see rows_zero()
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          t0 = fmaf(vr[k], 0.f, t0); t0 = fmaf(vz[k], 0.f, t0); t0 = fmaf(vn[k], 0.f, t0); t0 = fmaf(vh[k], 0.f, t0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float w = a[k] + t0;
          ar = fmaf(w, vr[k], ar); az = fmaf(w, vz[k], az); an = fmaf(w, vn[k], an); am = fmaf(w, vh[k], am);
        }
      }
    }
    DAGNN_GT(12, __float_as_int(ar))
    if (u < P.Kh64) {
      unsigned char* ihi = itile + (size_t)(u >> 6) * kAStageBytes + tc::tile_off(rin, (u & 63) >> 3) + (u & 7) * 2;
      __half hh = __float2half_rn(0.f), hl = hh;        // k padding of the last chunk: zeros
      if (live) {
        const float br = __ldg(Lp.bias + u), bz = __ldg(Lp.bias + HP + u), bi = __ldg(Lp.bias + 2 * HP + u),
                    bh = __ldg(Lp.bias + 3 * HP + u);
        const float rg = fast_sigmoid(gr + ar + br);
        const float zg = fast_sigmoid(gz + az + bz);
        const float ng = fast_tanh(gn + bi + rg * (an + bh));
        const float o = ng + zg * (am - ng);
        Lp.Hs[(size_t)p * ldh + u] = o;
        hh = __float2half_rn(o);
        hl = __float2half_rn(o - __half2float(hh));
        skacc += o * __ldg(Lp.wk + u);
      }
      *reinterpret_cast<__half*>(ihi) = hh;
      *reinterpret_cast<__half*>(ihi + 128 * tc::ROW_BYTES) = hl;
    }
  }
  DAGNN_GT(13, __float_as_int(skacc))
  skacc = warp_sum(skacc);
  if (lane == 0) part[warp] = skacc;
  asm volatile("bar.sync 1, %0;" ::"n"(kBuilders) : "memory");
  DAGNN_GT(14, lane)
  if (warp == 0 && lane == 0) {
    float sk = 0.f;
#pragma unroll
    for (int w = 0; w < kBuilderWarps; ++w) sk += part[w];
    Lp.sk[p] = sk;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kBuilders) : "memory");
  DAGNN_GT(15, lane)
}

// ------------------------------------------------------------------------------------------------------------
// proj phase
// ------------------------------------------------------------------------------------------------------------
struct Tile {                 // one projection work item, identical in every thread of the CTA
  const float* A;             // operand rows (fp32): states in position order, or X through perm
  long long lda;
  const int* perm;            // non-null: operand row of position p is A[perm[p]]
  const __half* img;          // weight image of the source
  float* out0;                // columns [0, Mc)
  float* out1;                // columns [Mc, 2 Mc) or nullptr
  const unsigned char* aimg;  // non-null: operand tiles come ready-made from this image (first 128-row tile of the work item)
  int K, nck, vec;            // valid operand width, 64-k chunks, rows are float4-loadable
  int p0, nrows, nst, cb0, ncb;   // first position, rows, 128-row sub-tiles, first 64-column block, blocks in this tile
  int small;                  // > 0: compact operand stages of `small` rows, every k chunk has its own stage (tail levels)
};
// Tail levels hold a handful of rows per tile. When all k chunks of such a tile fit into the operand region at once
// (stage = hi + lo tile of roundup(rows, 8) rows), nothing waits for a free stage: every operand chunk is in flight at
// once and the MMAs of a chunk start as soon as it has landed. The MMA still reads 128 rows from the stage base: the
// rows behind the stage are other stages' bytes — D rows nobody reads.

// raw operands of one work item of a builder thread, loaded one item ahead: kNR rows (r0 + x * kRStride of the sub-tile),
// 8 consecutive k
struct Pre {
  float4 v[kNR][2];
};

__device__ __forceinline__ void builders_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kBuilders) : "memory"); }

__device__ __forceinline__ void builder_tile(const SweepP& P, const Tile& T, unsigned char* As, SmemTail& S, uint32_t tmem, uint32_t ja,
                                             uint32_t ct, int tile_cols, long long* tr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool trc = tr != nullptr && tid == 0;
  const int r0 = tid >> 3, c8 = tid & 7;
  const int nitems = T.nck * T.nst;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // row groups (of kRStride rows) a sub-tile really has: the tail levels hold a handful of rows per tile
  auto groups_of = [&](int st) { return min(kNR, (min(128, T.nrows - st * 128) + kRStride - 1) / kRStride); };
  auto prefetch = [&](int it, Pre& R) {
    if (it >= nitems) return;
    const int c = it >> (T.nst - 1), st = it & (T.nst - 1);
    const int nx = groups_of(st);
    const int k0 = c * tc::KC16 + 8 * c8;
#pragma unroll
    for (int x = 0; x < kNR; ++x) {
      if (x >= nx) continue;
      R.v[x][0] = z4; R.v[x][1] = z4;
      const int r = st * 128 + r0 + kRStride * x;
      if (r >= T.nrows || k0 >= T.K) continue;
      const int p = T.p0 + r;
      const float* src = T.A + (size_t)(T.perm ? T.perm[p] : p) * T.lda + k0;
      if (T.vec) {
        R.v[x][0] = ldcg4(src);
        if (k0 + 4 < T.K) R.v[x][1] = ldcg4(src + 4);
      } else {
        float t[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) t[q] = (k0 + q < T.K) ? __ldcg(src + q) : 0.f;
        R.v[x][0] = make_float4(t[0], t[1], t[2], t[3]);
        R.v[x][1] = make_float4(t[4], t[5], t[6], t[7]);
      }
    }
  };

  Pre R;
  if (!T.aimg) prefetch(0, R);
#pragma unroll 1
  for (int it = 0; it < (T.aimg ? 0 : nitems); ++it) {
    const int st = it & (T.nst - 1);
    const uint32_t j = ja + (uint32_t)it;
    const uint32_t stage = j % kNAS, use = j / kNAS;
    unsigned char* A_hi = As + (size_t)stage * kAStageBytes;
    unsigned char* A_lo = A_hi + 128 * tc::ROW_BYTES;
    const int nx = groups_of(st);
    float x[kNR][8];
#pragma unroll
    for (int q = 0; q < kNR; ++q) {
      if (q >= nx) continue;
      const float4 a0 = R.v[q][0], a1 = R.v[q][1];
      x[q][0] = a0.x; x[q][1] = a0.y; x[q][2] = a0.z; x[q][3] = a0.w; x[q][4] = a1.x; x[q][5] = a1.y; x[q][6] = a1.z; x[q][7] = a1.w;
    }
    prefetch(it + 1, R);                                               // next item's loads in flight before anything else
    if (use >= 1) mbar_wait(&S.a_empty[stage], (use - 1) & 1u);        // MMAs that read this stage are done
    const int ra = st * 128 + r0;
#pragma unroll
    for (int q = 0; q < kNR; ++q)                                      // rows beyond the tile: D rows nobody reads
      if (q < nx && ra + kRStride * q < T.nrows) tc::store_split8(A_hi, A_lo, r0 + kRStride * q, c8, x[q]);
    tc::fence_async_smem();            // generic-proxy stores -> visible to the tensor core (async proxy)
    builders_sync();
    if (tid == 0) mbar_arrive(&S.a_full[stage]);
  }
  if (trc) tr[1] = clock64();

  // ---------------- epilogue: TMEM -> P / Gi rows ----------------
  mbar_wait(&S.acc_full, ct & 1u);
  tc::fence_after_sync();
  if (trc) tr[2] = clock64();
  {
    const int q = warp & 3, cg = warp >> 2;
    const int ngrp = T.ncb * 2;                           // 32-column groups of the tile (a 64-column block = 2 groups)
    const int Mc = P.Mc;
#pragma unroll 1
    for (int st = 0; st < T.nst; ++st) {
      const uint32_t tbase = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(st * tile_cols);   // TMEM lane = row of the sub-tile
#pragma unroll 1
      for (int g = cg; g < ngrp; g += kBuilderWarps / 4) {
        float v[32];
        __syncwarp();
        tc::ld32(tbase + (uint32_t)(32 * g), v);
        tc::wait_ld();
        // TMEM hands every lane one ROW (32 columns); stored like that, a warp instruction would scatter 16-byte pieces
        // over 32 rows (partial sectors). Transpose through shared memory: 8 lanes then write 128 contiguous bytes of a row.
        float* sc = S.epi[warp];
        const int col = T.cb0 * 64 + 32 * g;             // a 32-column group never straddles the two projected matrices
        float* dst0 = (col < Mc) ? T.out0 + col : T.out1 + (col - Mc);
        const int rbase = st * 128 + 32 * q;
#pragma unroll
        for (int h = 0; h < 2; ++h) {                     // two halves of 16 columns
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<float4*>(sc + lane * kEpiLd + 4 * k) =
                make_float4(v[16 * h + 4 * k], v[16 * h + 4 * k + 1], v[16 * h + 4 * k + 2], v[16 * h + 4 * k + 3]);
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int rr = 8 * k + (lane >> 2);          // 8 rows per instruction, 4 lanes x 16 bytes each (two full sectors)
            if (rbase + rr < T.nrows)
              *reinterpret_cast<float4*>(dst0 + (size_t)(T.p0 + rbase + rr) * Mc + 16 * h + 4 * (lane & 3)) =
                  *reinterpret_cast<const float4*>(sc + rr * kEpiLd + 4 * (lane & 3));
          }
          __syncwarp();
        }
      }
    }
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.tmem_free);           // the next tile's MMAs may overwrite the accumulators now
  }
  if (trc) tr[3] = clock64();
}

// ------------------------------------------------------------------------------------------------------------
// issuer (one thread): weight ring + MMA issue. Ring barriers are tracked per physical barrier (the ring geometry
// changes with the tile width between phases).
// ------------------------------------------------------------------------------------------------------------
struct RingState {
  uint32_t full_par, empty_par, pending;     // bit s: parity of the next wait on b_full[s] / b_empty[s]; commit outstanding
  uint32_t next;                             // next stage to fill
  uint32_t pre, pre_first;                   // chunks of the upcoming tile already in flight, stage of its chunk 0
  int nbc;                                   // blocks per stage the ring is currently laid out for (8 / nbc stages)
};
// the ring geometry changes with the tile width: every stage must be drained before the region is re-cut
__device__ __forceinline__ void ring_set_geometry(int nbc, SmemTail& S, RingState& R) {
  if (R.nbc == nbc) return;
  for (uint32_t q = 0; q < (uint32_t)kNBBar; ++q)
    if (R.pending >> q & 1u) {
      mbar_wait(&S.b_empty[q], R.empty_par >> q & 1u);
      R.empty_par ^= 1u << q;
      R.pending &= ~(1u << q);
    }
  R.next = 0;
  R.nbc = nbc;
}
// weight chunk c of a tile -> next ring stage: hi tiles of the tile's blocks, then their lo tiles
__device__ __forceinline__ void ring_load(const Tile& T, int nbc, int c, unsigned char* Bs, SmemTail& S, RingState& R, bool leader) {
  const uint32_t nbs = 8u >> blk_shift(nbc);
  const uint32_t s = R.next;
  R.next = (s + 1 == nbs) ? 0u : s + 1;
  if (R.pending >> s & 1u) {                         // MMAs that read this stage must be done
    mbar_wait(&S.b_empty[s], R.empty_par >> s & 1u);
    R.empty_par ^= 1u << s;
    R.pending &= ~(1u << s);
  }
  if (!leader) return;
  mbar_expect_tx(&S.b_full[s], (uint32_t)T.ncb * kBlkBytes);
  unsigned char* stg = Bs + (size_t)s * nbc * kBlkBytes;
  const unsigned char* img = reinterpret_cast<const unsigned char*>(T.img);
  for (int b = 0; b < T.ncb; ++b) {
    const unsigned char* src = img + ((size_t)(T.cb0 + b) * T.nck + c) * kBlkBytes;
    bulk_g2s(stg + (size_t)b * (kBlkBytes / 2), src, kBlkBytes / 2, &S.b_full[s]);
    bulk_g2s(stg + (size_t)nbc * (kBlkBytes / 2) + (size_t)b * (kBlkBytes / 2), src + kBlkBytes / 2, kBlkBytes / 2, &S.b_full[s]);
  }
}

// Run by the WHOLE issuer warp, converged: every lane keeps the same ring state and waits on the same barriers, lane 0
// (`leader`) alone starts copies, MMAs and commits. The operands of tcgen05.mma live in uniform registers; values the
// compiler cannot prove warp-uniform (anything derived from the shared-memory step table) pass through uni() first,
// otherwise every MMA is wrapped in a ~100-cycle elect / broadcast loop and a 64-column tile is issue-bound.
__device__ __forceinline__ void issuer_tile(const Tile& T, int nbc, unsigned char* As, unsigned char* Bs, SmemTail& S, uint32_t tmem,
                                            uint32_t ja, uint32_t cs, uint32_t ct, RingState& R, bool leader, bool prev_small,
                                            long long* tr) {
  const bool trc = tr != nullptr && leader;            // stamps of the first tile of a phase: 10 start, 11 first operands there,
  if (trc) tr[10] = clock64();                         // 13 / 14 chunk 0 / 1 issued, 12 everything issued
  const int nbs = 8 >> blk_shift(nbc);
  const uint32_t idesc = uni(tc::instr_desc_f16(128, 64 * T.ncb));
  // chunk c sits in stage (first + c) % nbs; the first chunks may already be in flight (issued while the previous tile
  // was still computing, or before a grid barrier)
  if (R.pre == 0) ring_set_geometry(nbc, S, R);
  const uint32_t first = R.pre ? R.pre_first : R.next;
  const int npre = min(nbs, T.nck);
  for (int c = (int)R.pre; c < npre; ++c) ring_load(T, nbc, c, Bs, S, R, leader);
  R.pre = 0;
  // operand side when the tiles come ready-made: item it = (chunk c, sub-tile st) -> stage (ja + it) % kNAS, two bulk copies
  // (hi, lo) of the rows the sub-tile really has, rounded up to the 8-row swizzle group
  const int nitems = T.nck * T.nst;
  auto load_A = [&](int it) {
    const int c = it >> (T.nst - 1), st = it & (T.nst - 1);
    const uint32_t j = ja + (uint32_t)it;
    const uint32_t stage = j % kNAS, use = j / kNAS;
    if (use >= 1) mbar_wait(&S.a_empty[stage], (use - 1) & 1u);        // MMAs that read this stage are done
    if (!leader) return;
    const uint32_t bytes = (uint32_t)((min(128, T.nrows - st * 128) + 7) & ~7) * tc::ROW_BYTES;
    const unsigned char* src = T.aimg + ((size_t)st * T.nck + c) * kAStageBytes;
    unsigned char* dst = As + (size_t)stage * kAStageBytes;
    mbar_expect_tx(&S.a_full[stage], 2 * bytes);
    bulk_g2s(dst, src, bytes, &S.a_full[stage]);
    bulk_g2s(dst + 128 * tc::ROW_BYTES, src + 128 * tc::ROW_BYTES, bytes, &S.a_full[stage]);
  };
  const uint32_t sbytes = 2u * (uint32_t)T.small * tc::ROW_BYTES;      // compact stage: hi + lo tile of T.small rows
  // The compact stages of a small tile share the operand region with the ring but are not part of it (no a_empty hand-over):
  // when this tile or the one before it is small, the previous tile's MMAs — possibly still reading the region — must be
  // done before anything is copied over it. (Across phases that barrier completed long ago.)
  if (ct >= 1 && (T.small || prev_small)) mbar_wait(&S.acc_full, (ct - 1) & 1u);
  if (T.small) {
    for (int c = 0; c < (leader ? T.nck : 0); ++c) {
      const unsigned char* src = T.aimg + (size_t)c * kAStageBytes;
      unsigned char* dst = As + (size_t)c * sbytes;
      mbar_expect_tx(&S.c_full[c], sbytes);
      bulk_g2s(dst, src, sbytes / 2, &S.c_full[c]);
      bulk_g2s(dst + sbytes / 2, src + 128 * tc::ROW_BYTES, sbytes / 2, &S.c_full[c]);
    }
  } else if (T.aimg) {
    for (int it = 0; it < min(kNAS, nitems); ++it) load_A(it);
  }
#pragma unroll 1
  for (int c = 0; c < T.nck; ++c) {
    const uint32_t s = (first + (uint32_t)c) & (uint32_t)(nbs - 1);
    mbar_wait(&S.b_full[s], R.full_par >> s & 1u);
    R.full_par ^= 1u << s;
    const uint32_t sb = uni(smem_u32(Bs + (size_t)s * nbc * kBlkBytes));
    const uint64_t bh = tc::smem_desc(sb), bl = tc::smem_desc(sb + uni((uint32_t)nbc * (kBlkBytes / 2)));
#pragma unroll 1
    for (int st = 0; st < T.nst; ++st) {
      const uint32_t j = ja + (uint32_t)(c * T.nst + st);
      const uint32_t stage = j % kNAS, use = j / kNAS;
      if (T.small) mbar_wait(&S.c_full[c], cs >> c & 1u);
      else mbar_wait(&S.a_full[stage], use & 1u);
      if (c == 0 && st == 0 && ct >= 1) mbar_wait(&S.tmem_free, (ct - 1) & 1u);   // operand tiles arrive by bulk copy: nothing
      tc::fence_after_sync();                                                      // else orders us behind the epilogue
      if (trc && c == 0 && st == 0) tr[11] = clock64();
      const uint32_t sa = uni(T.small ? smem_u32(As + (size_t)c * sbytes) : smem_u32(As + (size_t)stage * kAStageBytes));
      const uint64_t ah = tc::smem_desc(sa), al = tc::smem_desc(sa + uni(T.small ? sbytes / 2 : 128 * tc::ROW_BYTES));
      const uint32_t tm = uni(tmem + (uint32_t)(st * 64 * nbc));
      const bool fresh = uni(c == 0 ? 1u : 0u) != 0;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) tc::mma3_f16(tm, ah + 2 * ks, al + 2 * ks, bh + 2 * ks, bl + 2 * ks, idesc, fresh && ks == 0);
        if (!T.small) tc::commit(&S.a_empty[stage]);
      }
      if (!T.small && T.aimg && c * T.nst + st + kNAS < nitems) load_A(c * T.nst + st + kNAS);   // refill this stage once its MMAs are done
    }
    if (elect_one()) tc::commit(&S.b_empty[s]);
    if (trc && c < 2) tr[13 + c] = clock64();
    R.pending |= 1u << s;
    // refill: chunk c + nbs - 1 goes where chunk c - 1 was (its MMAs precede the ones just issued)
    if (c >= 1 && c + nbs - 1 < T.nck) ring_load(T, nbc, c + nbs - 1, Bs, S, R, leader);
  }
  if (trc) tr[12] = clock64();
  if (elect_one()) tc::commit(&S.acc_full);
}
// weights are constants: put the first chunks of this CTA's NEXT tile in flight right after the current tile's MMAs are
// queued — they land while the accumulators drain, the epilogue runs and (for the first tile of the next proj phase) the
// gate phase runs. The next tile is worked out here, behind the MMAs, not in front of them.
__device__ __forceinline__ void issuer_prefetch(const Tile& N, int next_nbc, unsigned char* Bs, SmemTail& S, RingState& R, bool leader) {
  ring_set_geometry(next_nbc, S, R);
  R.pre_first = R.next;
  const int n2 = min(8 >> blk_shift(next_nbc), N.nck);
  for (int c = 0; c < n2; ++c) ring_load(N, next_nbc, c, Bs, S, R, leader);
  R.pre = (uint32_t)n2;
}

// ------------------------------------------------------------------------------------------------------------
// step tables. Proj phase -1 projects X (segments = directions, all N rows); proj phase s >= 0 projects the rows the
// gate phase of step s has just produced (segments (d, i, l = s - i)).
// ------------------------------------------------------------------------------------------------------------
struct TileIt { int q, t; };

__device__ __forceinline__ int seg_blocks(const SweepP& P, int s, int q) {      // 64-column blocks of segment q's projection
  const int mb = P.Mc >> 6;
  if (s < 0) return mb;                                 // X -> Gi^0
  int d, i;
  seg_di(P, q, d, i);
  return (i + 1 < P.layers) ? 2 * mb : mb;              // H^i -> P^i [, Gi^{i+1}]
}
__device__ __forceinline__ bool tile_advance(const StepTab& tb, int nseg, int rank, int G, TileIt& it) {
  if (it.t >= 0) it.t += G;
  while (it.q < nseg) {
    const Seg g = tb.seg[it.q];
    if (it.t < 0) it.t = (rank >= g.bmod) ? rank - g.bmod : rank - g.bmod + G;     // my tiles of a segment: (base + t) % G == rank
    if (it.t < g.ntile) return true;
    ++it.q;
    it.t = -1;
  }
  return false;
}
__device__ __forceinline__ Tile make_tile(const SweepP& P, const StepTab& tb, int s, const TileIt& it) {
  const Seg g = tb.seg[it.q];
  const int nblk = seg_blocks(P, s, it.q);
  const int ncbt = g.ncbt;                              // column tiles per row tile
  const int rows_per = 128 * tb.nst;
  const int rt = (ncbt == 1) ? it.t : (int)__umulhi((uint32_t)it.t, g.rcbt), ctile = it.t - rt * ncbt;
  Tile T;
  if (s < 0) {
    const int d = it.q;
    T.A = P.X; T.lda = P.ldx; T.perm = P.ximg ? nullptr : P.dir[d].perm; T.img = P.lay[d][0].imgx; T.aimg = nullptr;
    T.out0 = P.lay[d][0].Gi; T.out1 = nullptr;
    T.K = P.Din0; T.nck = P.nckx; T.vec = P.vec_x;
  } else {
    int d, i;
    seg_di(P, it.q, d, i);
    const LayP& Lp = P.lay[d][i];
    T.A = Lp.Hs; T.lda = P.ldh; T.perm = nullptr; T.img = Lp.imgh;
    T.out0 = Lp.Pm; T.out1 = (i + 1 < P.layers) ? P.lay[d][i + 1].Gi : nullptr;
    T.K = P.Hq; T.nck = P.nckh; T.vec = 1;
  }
  T.p0 = g.pos0 + rt * rows_per;
  T.nrows = min(rows_per, g.n - rt * rows_per);
  T.nst = (T.nrows + 127) >> 7;
  T.cb0 = ctile * tb.nbc;
  T.ncb = min(tb.nbc, nblk - T.cb0);
  if (s >= 0) {
    int d, i;
    seg_di(P, it.q, d, i);
    T.aimg = P.lay[d][i].aimg + (size_t)((g.pos0 >> 7) + (s - i) + rt * tb.nst) * P.nckh * kAStageBytes;
  } else if (P.ximg) {
    T.aimg = P.ximg + (size_t)(rt * tb.nst) * P.nckx * kAStageBytes;       // rows = nodes, tiles of 128 nodes
  }
  T.small = 0;
  if (T.aimg && T.nst == 1 && T.nck <= kMaxChunks) {
    const int rp8 = (T.nrows + 7) & ~7;
    if (T.nck * 2 * rp8 * tc::ROW_BYTES <= kNAS * kAStageBytes) T.small = rp8;
  }
  return T;
}
// one warp (lanes = segments); the caller publishes the table with a CTA-wide barrier
__device__ __noinline__ void build_step_table(const SweepP& P, StepTab& tb, int s, int L, int nseg_max, int G) {
  const int tid = threadIdx.x & 31;
  const int nseg = (s < 0) ? P.dirs : nseg_max;
  if (tid < kMaxSeg) {
    Seg g = {0, 0, 0, 0, 0, 1, 0u};
    int gp = 0, gnn = 0;
    if (tid < nseg) {
      if (s < 0) g.n = P.N;
      else {
        int d, i;
        seg_di(P, tid, d, i);
        const int l = s - i;
        if (l >= 0 && l < L) {
          gp = P.dir[d].lvl_off[l];
          gnn = max(0, P.dir[d].lvl_off[l + 1] - gp);
          // the last level's states have no successors: only the next layer's input projection is still needed
          if (l + 1 < L || i + 1 < P.layers) { g.pos0 = gp; g.n = gnn; }
        }
      }
    }
    tb.seg[tid] = g;
    tb.gpos0[tid] = gp; tb.gn[tid] = gnn;
  }
  __syncwarp();
  if (tid == 0) {
    int t64 = 0;                                        // work in units of 128 rows x 64 columns
    for (int q = 0; q < nseg; ++q) t64 += ((tb.seg[q].n + 127) >> 7) * seg_blocks(P, s, q);
    const int nbc = (t64 >= 4 * G) ? 4 : (t64 >= 2 * G) ? 2 : 1;
    const int nst = (nbc == 4 && t64 >= 8 * G) ? 2 : 1;
    int base_ = 0, bmod = 0;
    const int rsh = 6 + nst, rmask = (64 << nst) - 1;   // rows per tile = 128 nst (nst = 1, 2)
#pragma unroll 1
    for (int q = 0; q < kMaxSeg; ++q) {
      const int ncbt = (q < nseg) ? (seg_blocks(P, s, q) + nbc - 1) >> blk_shift(nbc) : 1;
      tb.seg[q].ntile = (q < nseg) ? ((tb.seg[q].n + rmask) >> rsh) * ncbt : 0;
      tb.seg[q].base = base_;
      tb.seg[q].bmod = bmod;
      tb.seg[q].ncbt = ncbt;
      tb.seg[q].rcbt = 0xffffffffu / (uint32_t)ncbt + 1u;       // ncbt == 1 wraps to 0: make_tile does not divide then
      base_ += tb.seg[q].ntile;
      bmod = (bmod + tb.seg[q].ntile) % G;
    }
    tb.nbc = nbc; tb.nst = nst;
  }
  __syncwarp();
}

// one warp, after build_step_table: list the nodes of step s with long in-edge lists (row pointers are constants too).
// Every CTA derives the same list, so the split of the gate phase below needs no communication. Eight blocks of 32 row
// pointers are in flight at a time: the scan has to hide behind the gate phase of the step before.
__device__ __noinline__ void scan_heavy(const SweepP& P, const StepTab& tb, HeavyTab& hv, int s, int nseg) {
  const int lane = threadIdx.x & 31;
  int total = 0;
  for (int q = 0; q < nseg; ++q) total += tb.gn[q];
  int cnt = 0;
  if (s < 0 || total > kScanRows) cnt = -1;
  for (int q = 0; q < nseg && cnt >= 0; ++q) {
    int d, i;
    seg_di(P, q, d, i);
    const int l = s - i;
    const int n = tb.gn[q], pos0 = tb.gpos0[q];
    if (n <= 0 || l <= 0) continue;                     // level 0 ignores its in-edges
    const int* __restrict__ rp = P.dir[d].rowptr + pos0;
    for (int r0 = 0; r0 < n && cnt >= 0; r0 += 256) {
      int deg[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = r0 + 32 * k + lane;
        deg[k] = (r < n) ? rp[r + 1] - rp[r] : 0;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool h = deg[k] > kCoopEdges;
        const unsigned m = __ballot_sync(0xffffffffu, h);
        if (m && cnt >= 0) {
          const int slot = cnt + __popc(m & ((1u << lane) - 1u));
          if (h && slot < kMaxHeavy) { hv.node[slot][0] = q; hv.node[slot][1] = pos0 + r0 + 32 * k + lane; }
          cnt += __popc(m);
          if (cnt > kMaxHeavy) cnt = -1;
        }
      }
    }
  }
  if (lane == 0) hv.n = cnt;
  __syncwarp();
}

// J = float4 column groups per lane in the gate phase: 2 covers H <= 256 in one pass, 4 wider states (one instantiation
// each keeps the instruction footprint of a launch small — the kernel hops between phases ~130 times per forward)
template <int J>
__global__ void __launch_bounds__(kThreads, 1) k_sweep(const __grid_constant__ SweepP P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* As = base;
  unsigned char* Bs = base + (size_t)kNAS * kAStageBytes;
  SmemTail& S = *reinterpret_cast<SmemTail*>(Bs + kBRegionBytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kNAS; ++s) { mbar_init(&S.a_full[s], 1); mbar_init(&S.a_empty[s], 1); }
    for (int s = 0; s < kNBBar; ++s) { mbar_init(&S.b_full[s], 1); mbar_init(&S.b_empty[s], 1); }
    for (int s = 0; s < kMaxChunks; ++s) mbar_init(&S.c_full[s], 1);
    mbar_init(&S.acc_full, 1);
    mbar_init(&S.tmem_free, kBuilderWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) tc::tmem_alloc(&S.tmem_slot, kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = S.tmem_slot;

  const bool ok = P.summary[2] == 0;           // schedule build flagged bad input: do nothing, the host raises
  const int L = ok ? P.summary[0] : 0;
  const int nsteps = ok ? L + P.layers - 1 : 0;
  const int nseg = P.dirs * P.layers;
  const int G = (int)gridDim.x, rank = (int)blockIdx.x;
  bool prev_small = false;                     // the CTA's previous projection tile used the compact operand stages
  uint32_t ja = 0, ct = 0, cs = 0;             // operand-ring items and tiles processed so far, phase bits of the chunk barriers
  RingState R = {0u, 0u, 0u, 0u, 0u, 0u, 0};
  unsigned int nbar = 0;
  // proj phase s uses tab[s & 1] (-1 & 1 == 1). Level offsets are constants: the issuer warp builds the table of phase
  // s + 1 while the builder warps run the gate phase of step s, the grid barrier after it publishes the table.
  if (nsteps > 0 && warp == kBuilderWarps) {
    build_step_table(P, S.tab[1], -1, L, nseg, G);
    build_step_table(P, S.tab[0], 0, L, nseg, G);
    scan_heavy(P, S.tab[0], S.heavy[0], 0, nseg);
  }
  __syncthreads();

  // gate phase of step s (s >= 0), then proj phase s (s = -1: X)
  // (a schedule flagged unusable runs nothing at all — not even the projection of X: its tables were never built)
#pragma unroll 1
  for (int s = nsteps > 0 ? -1 : 0; s < nsteps; ++s) {
    long long* tr = P.trace ? P.trace + ((size_t)(s + 1) * 256 + blockIdx.x) * 16 : nullptr;
    if (tr && tid == 0) { tr[0] = clock64(); tr[1] = tr[2] = tr[3] = 0; tr[8] = tr[9] = 0; }
    if (s >= 0) {
      // ---------------- gate phase of step s ----------------
      if (warp < kBuilderWarps) {
        const StepTab& tg = S.tab[s & 1];
        if (tid == 0) S.ncoop = 0;
        builders_sync();
        int rbase = 0;
#ifdef DAGNN_GATE_TRACE
        bool gt_done = false;
        if (tr && tid == 0) tr[10] = tr[11] = tr[12] = tr[13] = tr[14] = tr[15] = 0;
#endif
        // nodes with long in-edge lists, found one phase ahead: CTA h takes the h-th one, all warps, first thing; while
        // they are few the ordinary rows go to the other CTAs only, so that no CTA does both and the phase is as long as
        // one node, not two
        const HeavyTab& hv = S.heavy[s & 1];
        const int nh = hv.n;
        int Gn = G, rank_n = rank;
        if (nh > 0 && nh <= G / 2) { Gn = G - nh; rank_n = rank - nh; }
        for (int h = rank; h < nh; h += G) {
          const int q = hv.node[h][0], p = hv.node[h][1];
          int d, i;
          seg_di(P, q, d, i);
#if defined(DAGNN_GATE_TRACE) && DAGNN_GATE_TRACE == 2
          gate_node_cta(P, P.dir[d], P.lay[d][i], p, tg.gpos0[q], s - i, reinterpret_cast<float*>(As), warp, lane,
                           (tr && tid == 0 && h == rank) ? tr : nullptr);
#elif defined(DAGNN_GATE_TRACE)
          gate_node_cta(P, P.dir[d], P.lay[d][i], p, tg.gpos0[q], s - i, reinterpret_cast<float*>(As), warp, lane, nullptr);
#else
          gate_node_cta(P, P.dir[d], P.lay[d][i], p, tg.gpos0[q], s - i, reinterpret_cast<float*>(As), warp, lane);
#endif
        }
        const int W = Gn * kBuilderWarps;
        for (int q = 0; q < nseg; ++q) {
          int d, i;
          seg_di(P, q, d, i);
          const int l = s - i;
          const int pos0 = tg.gpos0[q], n = tg.gn[q];     // level offsets: table built one phase ahead
          if (n <= 0) continue;
          const DirP& D = P.dir[d];
          // rows of the step are dealt to the warps of the (remaining) CTAs, CTA-minor, continuing across segments
          int r = warp * Gn + rank_n - rbase;           // rbase = rows dealt so far, mod W
          if (r < 0) r += W;
          if (rank_n < 0) r = n;
          for (; r < n; r += W) {
            const int p = pos0 + r;
            if (l > 0 && D.rowptr[p + 1] - D.rowptr[p] > kCoopEdges) {     // long in-edge list: leave it to the whole CTA
              if (nh >= 0) continue;                                        // ... done above
              int slot = 0;
              if (lane == 0) slot = atomicAdd(&S.ncoop, 1);
              slot = __shfl_sync(0xffffffffu, slot, 0);
              if (slot < kMaxCoop) {
                if (lane == 0) { S.coop[slot][0] = q; S.coop[slot][1] = p; }
                continue;
              }
            }
#ifdef DAGNN_GATE_TRACE
#if DAGNN_GATE_TRACE == 3      // a node with 8..kCoopEdges in-edges, any warp
            long long* gt = (tr && lane == 0 && l > 0 && D.rowptr[p + 1] - D.rowptr[p] >= 8) ? tr : nullptr;
#elif DAGNN_GATE_TRACE == 2    // cooperative nodes only
            long long* gt = nullptr;
#else                          // the first node of warp 0
            long long* gt = (tr && tid == 0 && !gt_done) ? tr : nullptr;
#endif
            if (gt) { gt[10] = clock64(); gt_done = true; }
            gate_row<J>(P, D, P.lay[d][i], p, pos0, l, lane, gt);
#else
            gate_row<J>(P, D, P.lay[d][i], p, pos0, l, lane);
#endif
          }
          rbase += n;
          if (rbase >= W) { rbase -= W; if (rbase >= W) rbase %= W; }
        }
        builders_sync();
        const int nc = min(S.ncoop, kMaxCoop);              // big steps: the nodes found while dealing the rows
        for (int c = 0; c < nc; ++c) {
          const int q = S.coop[c][0], p = S.coop[c][1];
          int d, i;
          seg_di(P, q, d, i);
          const int pos0 = tg.gpos0[q];
#if defined(DAGNN_GATE_TRACE)
          gate_node_cta(P, P.dir[d], P.lay[d][i], p, pos0, s - i, reinterpret_cast<float*>(As), warp, lane, nullptr);
#else
          gate_node_cta(P, P.dir[d], P.lay[d][i], p, pos0, s - i, reinterpret_cast<float*>(As), warp, lane);
#endif
        }
      }
      else if (s + 1 < nsteps) {
        build_step_table(P, S.tab[(s + 1) & 1], s + 1, L, nseg, G);
        scan_heavy(P, S.tab[(s + 1) & 1], S.heavy[(s + 1) & 1], s + 1, nseg);
      }
      if (tr && tid == 0) tr[8] = clock64();
      if (s + 1 < nsteps) grid_barrier(P.bar, ++nbar * (unsigned int)G);   // the last step's projection is empty
      if (tr && tid == 0) tr[9] = clock64();
    }
    // ---------------- proj phase s ----------------
    // the table of this phase was built one phase earlier (level offsets are constants); build the next one now so that
    // the issuer can look across the barriers
    const StepTab& tb = S.tab[s & 1];
    StepTab& tbn = S.tab[(s + 1) & 1];
    const bool has_next = s + 1 < nsteps;
    const int nseg_now = (s < 0) ? P.dirs : nseg;
    int my_tiles = 0;
    TileIt it = {0, -1};
    bool more = tile_advance(tb, nseg_now, rank, G, it);
#pragma unroll 1
    while (more) {
      const Tile T = make_tile(P, tb, s, it);
      more = tile_advance(tb, nseg_now, rank, G, it);
      if (warp < kBuilderWarps) {
        builder_tile(P, T, As, S, tmem, ja, ct, 64 * tb.nbc, my_tiles == 0 ? tr : nullptr);
      } else {                                       // the issuer warp, all lanes (see issuer_tile)
        issuer_tile(T, tb.nbc, As, Bs, S, tmem, ja, cs, ct, R, lane == 0, prev_small, my_tiles == 0 ? tr : nullptr);
        if (more) issuer_prefetch(make_tile(P, tb, s, it), tb.nbc, Bs, S, R, lane == 0);
        else if (has_next) {
          TileIt it2 = {0, -1};
          if (tile_advance(tbn, nseg, rank, G, it2)) issuer_prefetch(make_tile(P, tbn, s + 1, it2), tbn.nbc, Bs, S, R, lane == 0);
        }
      }
      __syncwarp();
      if (T.small) cs ^= (1u << T.nck) - 1u;         // tiles differ in their number of chunks: one phase bit per chunk barrier
      else ja += (uint32_t)(T.nck * T.nst);
      prev_small = T.small != 0;
      ct += 1;
      ++my_tiles;
    }
    if (tr && tid == 0) { tr[4] = clock64(); tr[6] = my_tiles; tr[7] = (64 * tb.nbc) | ((128 * tb.nst) << 12); }
    if (has_next) grid_barrier(P.bar, ++nbar * (unsigned int)G);
    if (tr && tid == 0) tr[5] = clock64();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, kTmemCols);
}

}  // namespace dagnn

using namespace dagnn;

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// operand image of one (direction, layer): 128-row tiles aligned to level starts — tile index of position p in level l
// (first position pos0) = pos0 / 128 + l + (p - pos0) / 128 — times the 64-k chunks of H, 32 KB (hi + lo) each
static size_t aimg_bytes(int H, int64_t N, int max_levels) {
  return align256((size_t)((N >> 7) + max_levels + 2) * (size_t)(round_up(H, 64) / 64) * kAStageBytes);
}
extern "C" size_t dagnn_sweep_workspace_bytes(int32_t dirs, int32_t layers, int32_t Din, int32_t H, int64_t N, int64_t E,
                                              int32_t max_levels) {
  if (dirs < 1 || dirs > DAGNN_MAX_DIRS || layers < 1 || layers > DAGNN_MAX_LAYERS || Din < 1 || H < 1 || N < 0 || E < 0 || max_levels < 1)
    return 0;
  const size_t Mc = (size_t)round_up(3 * round_up(H, 4), 64);
  // per (direction, layer): key scores [N], hidden projection P [N, Mc], input projection Gi [N, Mc], operand image of H
  const size_t grid_wide = 256 + (size_t)dirs * layers * (align256((size_t)N * sizeof(float)) + 2 * align256((size_t)N * Mc * sizeof(float)) +
                                                          aimg_bytes(H, N, max_levels));
  const size_t cluster = cluster_workspace_bytes(dirs, layers, Din, H, N, max_levels);   // either kernel may take the launch
  return grid_wide > cluster ? grid_wide : cluster;
}
extern "C" size_t dagnn_sweep_trace_bytes(int32_t max_steps) { return (size_t)(max_steps + 1) * 256 * 16 * sizeof(long long); }

extern "C" int dagnn_sweep_forward_f32(const DagnnSweepArgs* A, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  DAGNN_REQUIRE(A && A->sched, "sweep: null args");
  const DagnnSchedule* S = A->sched;
  const int dirs = S->dirs, layers = A->num_layers, H = A->H;
  DAGNN_REQUIRE(dirs >= 1 && dirs <= DAGNN_MAX_DIRS, "sweep: dirs");
  DAGNN_REQUIRE(layers >= 1 && layers <= DAGNN_MAX_LAYERS, "sweep: num_layers");
  DAGNN_REQUIRE(A->X && A->ldx >= A->Din && A->Din > 0, "sweep: X");
  DAGNN_REQUIRE(A->ldh % 4 == 0 && A->ldh >= round_up(H, 4), "sweep: ldh must be a multiple of 4 and >= roundup(H,4)");
  DAGNN_REQUIRE(A->workspace && ((uintptr_t)A->workspace & 255) == 0, "sweep: workspace must be 256-byte aligned");
  if (A->workspace_bytes < dagnn_sweep_workspace_bytes(dirs, layers, A->Din, H, S->N, S->E, S->max_levels))
    return set_err(DAGNN_E_WORKSPACE, "sweep: workspace too small (dagnn_sweep_workspace_bytes)");
  if (H < 1 || H > 4096) return set_err(DAGNN_E_UNSUPPORTED, "sweep: hidden size %d not in [1,4096]", H);
  if (A->nvid < 0) return set_err(DAGNN_E_INVALID, "sweep: nvid");
  if (S->N >= (1ll << 31)) return set_err(DAGNN_E_UNSUPPORTED, "sweep: more than 2^31 nodes");
  for (int d = 0; d < dirs; ++d) {
    DAGNN_REQUIRE(S->perm[d] && S->rowptr[d] && S->lvl_off[d] && (S->E == 0 || S->col[d]), "sweep: schedule arrays");
    DAGNN_REQUIRE(!A->use_edge_attr || S->E == 0 || S->eattr[d], "sweep: schedule carries no edge attributes");
    for (int i = 0; i < layers; ++i) {
      DAGNN_REQUIRE(A->Hs[d][i] && ((uintptr_t)A->Hs[d][i] & 15) == 0, "sweep: state buffers must be 16-byte aligned");
      DAGNN_REQUIRE(A->packed[d][i] && ((uintptr_t)A->packed[d][i] & 15) == 0, "sweep: packed params must be 16-byte aligned");
    }
  }
  // Which kernel: the cluster-resident sweep wins while levels are short (its cost is the chain of levels: ~7 us per level and
  // ~3.4 us per 64 rows of a level per cluster); the grid-wide sweep below spreads a big level over all 148 SMs and wins on
  // large batches (measured: config 2, 66 k node-steps: 1.07 ms vs 1.66 ms; batch 4096, 2.1 M node-steps: 18.4 ms vs 16.4 ms).
  // DAGNN_SWEEP_PATH=grid / cluster forces one of them (A/B measurements, tests of both kernels).
  const char* path_env = getenv("DAGNN_SWEEP_PATH");
  const bool force_grid = path_env && strcmp(path_env, "grid") == 0, force_cluster = path_env && strcmp(path_env, "cluster") == 0;
  const long long node_steps = (long long)S->N * dirs * layers;
  if (!force_grid && (force_cluster || node_steps <= 600000)) {
    bool handled = false;
    if (int rc = cluster_forward(A, st, &handled)) return rc;
    if (handled) return DAGNN_OK;
  }
  SweepP P;
  memset(&P, 0, sizeof(P));
  DagnnPackLayout lay[DAGNN_MAX_LAYERS];
  for (int i = 0; i < layers; ++i)
    if (int rc = dagnn_pack_layout(i == 0 ? A->Din : H, H, A->nvid, i == 0, i + 1 == layers, &lay[i])) return rc;
  P.dirs = dirs; P.layers = layers; P.H = H; P.Hq = lay[0].Hq; P.Mc = lay[0].Mc; P.HP = lay[0].HP; P.nvid = A->nvid;
  P.use_ea = A->use_edge_attr; P.Din0 = A->Din; P.nckx = lay[0].Kin64 / 64; P.nckh = lay[0].Kh64 / 64;
  P.vec_x = ((A->ldx & 3) == 0 && (A->Din & 3) == 0 && ((uintptr_t)A->X & 15) == 0) ? 1 : 0;
  P.N = (int)S->N; P.Kh64 = lay[0].Kh64;
  P.ldh = A->ldh; P.ldx = A->ldx; P.X = A->X; P.summary = S->summary; P.bar = static_cast<unsigned int*>(A->workspace);
  P.trace = static_cast<long long*>(A->trace);
  P.ximg = static_cast<const unsigned char*>(A->X_image);
  DAGNN_REQUIRE(!P.ximg || ((uintptr_t)P.ximg & 1023) == 0, "sweep: X_image must be 1024-byte aligned");
  char* ws = static_cast<char*>(A->workspace) + 256;
  const size_t sk_bytes = align256((size_t)S->N * sizeof(float)), pm_bytes = align256((size_t)S->N * P.Mc * sizeof(float));
  for (int d = 0; d < dirs; ++d) {
    DAGNN_REQUIRE(S->perm[d] && S->rowptr[d] && S->lvl_off[d] && (S->E == 0 || S->col[d]), "sweep: schedule arrays");
    DAGNN_REQUIRE(!A->use_edge_attr || S->E == 0 || S->eattr[d], "sweep: schedule carries no edge attributes");
    P.dir[d].perm = S->perm[d]; P.dir[d].rowptr = S->rowptr[d]; P.dir[d].col = S->col[d];
    P.dir[d].eattr = A->use_edge_attr ? S->eattr[d] : nullptr; P.dir[d].lvl_off = S->lvl_off[d];
    for (int i = 0; i < layers; ++i) {
      DAGNN_REQUIRE(A->Hs[d][i] && ((uintptr_t)A->Hs[d][i] & 15) == 0, "sweep: state buffers must be 16-byte aligned");
      DAGNN_REQUIRE(A->packed[d][i] && ((uintptr_t)A->packed[d][i] & 15) == 0, "sweep: packed params must be 16-byte aligned");
      const DagnnPackLayout& L = lay[i];
      const float* pk = A->packed[d][i];
      LayP& q = P.lay[d][i];
      q.Hs = A->Hs[d][i]; q.bias = pk + L.bias_off; q.wk = pk + L.wk_off; q.attnc = pk + L.attnc_off; q.vidk = pk + L.vidk_off;
      q.imgh = reinterpret_cast<const __half*>(pk + L.imgh_off);
      q.imgx = i == 0 ? reinterpret_cast<const __half*>(pk + L.imgx_off) : nullptr;
      q.gi_perm = (i == 0 && P.ximg) ? S->perm[d] : nullptr;
      q.sk = reinterpret_cast<float*>(ws); ws += sk_bytes;
      q.Pm = reinterpret_cast<float*>(ws); ws += pm_bytes;
      q.Gi = reinterpret_cast<float*>(ws); ws += pm_bytes;
      q.aimg = reinterpret_cast<unsigned char*>(ws); ws += aimg_bytes(H, S->N, S->max_levels);
    }
  }

  static PerDeviceOnce once;
  static int sm_count[kMaxDevices] = {0};
  int dev = 0;
  if (int rc = per_device_once(once, &dev, [&](int dv) {
        DAGNN_CUDA_OK(cudaFuncSetAttribute(k_sweep<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        DAGNN_CUDA_OK(cudaFuncSetAttribute(k_sweep<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        int n = 0, coop = 0;
        DAGNN_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dv));
        DAGNN_CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dv));
        if (!coop) return set_err(DAGNN_E_UNSUPPORTED, "sweep: device has no cooperative launch");
        sm_count[dv] = n;
        return (int)DAGNN_OK;
      }))
    return rc;
  const int G = sm_count[dev];
  DAGNN_CUDA_OK(cudaMemsetAsync(A->workspace, 0, 16, st));
  void* kargs[] = {(void*)&P};
  const void* kern = P.Hq <= 256 ? (const void*)k_sweep<2> : (const void*)k_sweep<4>;
  DAGNN_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3(G), dim3(kThreads), kargs, kSmemBytes, st));
  return check_launch("k_sweep");
}
