// The DAGNN level sweep as ONE persistent cooperative kernel (one CTA per SM, one grid barrier per wavefront step),
// gate GEMM on the 5th-generation tensor cores.
//
// Wavefront step s runs every (direction d, layer i, level l) with l + i == s: (l, i) depends on (l, i-1) [its
// input rows] and on (< l, i) [predecessor states], both finished in earlier steps. The sequential depth is
// L + layers - 1 grid barriers instead of L * layers * dirs kernel chains, and nothing on the host depends on the
// level sizes: level offsets and the level count are read from device memory (no host sync in a forward).
//
// Work unit = tile (d, i, l, up to 256 consecutive positions of the level, U = 16 or 64 hidden units); the tiles of a
// step are dealt round-robin to the CTAs; U = 64 when that still fills the grid, else 16 (latency-bound tail levels).
// Warp roles: 16 builder/epilogue warps + 1 issuer warp.
//   builders, pre-phase : one thread per row turns the in-edge scores into softmax weights alpha_e. Scores are scalar
//                         gathers: s_e = sum_j skp[nbr][j] (+ edge-type / vertex-id terms) — every producer of a state
//                         row also writes the partial key scores wk . h over 16-unit groups (separable attention score).
//   builders, main loop : per 64-wide k chunk and 128-row sub-tile, build the A operand [128, 64] (input rows, or
//                         m_v = sum_e alpha_e h_e for that k range) as fp16 hi / lo tiles straight into swizzled shared
//                         memory (the aggregate never goes back to HBM), 2-stage ring handed over by mbarriers.
//   issuer              : streams the pre-swizzled hi/lo weight image of the tile's units chunk by chunk with
//                         cp.async.bulk (UBLKCP) into a ring, issues 4 k-steps x 3 products of tcgen05.mma kind::f16
//                         (M = 128, N = 3U) per operand stage into TMEM, tcgen05.commit frees the stages.
//   builders, epilogue  : tcgen05.ld of [n_in | r | z | n_hid], sigmoid/tanh/blend in registers, state row and
//                         key-score partial stored.
// States written in one step are read in later steps by OTHER CTAs: all state reads use ld.global.cg (L2), the
// barrier is the cooperative-groups pattern (bar.sync; fence; atomic; spin on ld.acquire; bar.sync).
#include "common.cuh"
#include "tc.cuh"

namespace dagnn {

constexpr int kBuilderWarps = 8;
constexpr int kBuilders = kBuilderWarps * 32;            // 256 threads build operands and run the epilogue
constexpr int kNR = 128 * 8 / kBuilders;                 // rows of a 128-row operand stage per builder thread
constexpr int kRStride = 128 / kNR;
constexpr int kThreads = kBuilders + 32;                 // + one warp whose lane 0 issues bulk copies and MMAs
constexpr int kAStageBytes = 2 * 128 * tc::ROW_BYTES;    // hi + lo tile of 128 rows x 64 k  = 32 KB
constexpr int kNAS = 2;                                  // operand stages
constexpr int kBRegionBytes = 2 * 2 * 192 * tc::ROW_BYTES;   // weight ring: 2 stages of 48 KB (U = 64) or 8 of 12 KB (U = 16)
constexpr int kNBBar = 8;
constexpr int kEdgeCap = 3072;                           // in-edges of one tile cached in shared memory (else: global scratch)
constexpr int kMaxRows = 256;                            // rows per tile (two 128-row sub-tiles share every weight chunk)
constexpr int kMaxSeg = DAGNN_MAX_DIRS * DAGNN_MAX_LAYERS;
constexpr int kTmemCols = 512;                           // 2 sub-tiles x [n_in | r | z | n_hid] x 64 units
constexpr int kMaxSmem = 232448;                         // 227 KB opt-in limit per CTA on sm_100

struct DirP {
  const int* perm;      // position -> node id
  const int* rowptr;    // [N+1] CSR rows by position
  const int* col;       // [E] neighbour position
  const float* eattr;   // [E,2] in CSR order or nullptr
  const int* lvl_off;   // [max_levels+1] first position of each level
};
struct LayP {
  float* Hs;            // H[d][i], [N, ldh] position order: predecessor rows read, this level's rows written
  const float* bias;    // [4][HP]
  const float* wk;      // [HP]
  const float* attnc;   // [4]
  const float* vidk;    // [nvid]
  const __half* img16;  // weight images (pack.cu)
  const __half* img64;
  float* alpha;         // [E] scratch: softmax weights of tiles whose edge list exceeds the shared-memory cache
  float* skp;           // [N][nskp] partial key scores wk . h over 16-unit groups, written with every state row
};
constexpr int kHeavy = 2;                                // rows with more in-edges get their aggregate precomputed per tile
constexpr int kVeryHeavy = 24;                           // ... by the whole CTA instead of one warp
constexpr int kMaxVH = 16;
struct SweepP {
  int dirs, layers, H, Hq, nvid, use_ea;
  int Din0, nci0, ncih;       // layer-0 input width, 64-k chunks of the layer-0 input / of a hidden-width operand
  int NG, NT, nskp, vec_x;    // 16-unit groups, 64-unit tiles, row stride of skp, X rows are float4-loadable
  long long ldh, ldx;
  const float* X;             // [N, ldx] node order (rows through perm)
  const int* summary;         // [0] number of levels of direction 0, [2] schedule status
  unsigned int* bar;          // grid barrier counter (zeroed by the launcher)
  float* heavy;               // [grid][kMaxRows][Hq] per-CTA scratch: aggregates m_v of the tile's high-in-degree rows
  long long* trace;           // optional [steps][256][16] clock64 stamps, nullptr = off
  DirP dir[DAGNN_MAX_DIRS];
  LayP lay[DAGNN_MAX_DIRS][DAGNN_MAX_LAYERS];
};

struct Seg {
  int pos0, n, ntile, base;
};
struct StepTab {
  Seg seg[kMaxSeg];
  int U, nst;                 // units per tile, 128-row sub-tiles per tile
};
struct SmemTail {
  float alpha[kEdgeCap];
  int col[kEdgeCap];
  int rp[kMaxRows + 4];
  int nidx[kMaxRows];
  StepTab tab[2];             // this step's and the next step's segment tables
  float bias[5][64];          // b_r, b_z, b_in, b_hn, wk of the tile's units
  int vh[kMaxVH];             // rows of the tile whose edge list the whole CTA aggregates
  int nvh;
  uint64_t a_full[kNAS], a_empty[kNAS], b_full[kNBBar], b_empty[kNBBar], acc_full;
  uint32_t tmem_slot;
};
constexpr size_t kSmemBytes = 1024 + (size_t)kNAS * kAStageBytes + kBRegionBytes + sizeof(SmemTail);
static_assert(kSmemBytes <= (size_t)kMaxSmem, "shared memory plan exceeds the 227 KB opt-in limit");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void builders_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kBuilders) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// all CTAs of the (cooperative, co-resident) grid; `target` = arrivals expected so far
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (ld_acquire_u32(bar) < target) {}
    __threadfence();
  }
  __syncthreads();
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct Tile {                 // one work item, identical in every thread of the CTA
  int d, i, level0, pos0, p0, nrows, nst, ut, nci, nchunks;
};

// ------------------------------------------------------------------------------------------------------------
// builders: softmax weights, operand tiles, epilogue
// ------------------------------------------------------------------------------------------------------------
// position of hidden chunk h in the processing order: the chunk that holds the tile's own units goes LAST, so that its
// operand stage (m_v of exactly these units, as hi + lo) is still in shared memory when the epilogue needs h_prev
__device__ __forceinline__ int hperm(int h, int nch, int own) { return h == nch - 1 ? own : (h < own ? h : h + 1); }

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// raw operands of one work item of a builder thread, loaded one item ahead: kNR rows (r0 + x * kRStride of the sub-tile),
// up to two weighted source rows each (the input row with weight 1, or the first two in-edges), 8 consecutive k
struct Pre {
  float4 v[kNR][2][2];    // [row][source][half]
  float w[kNR][2];
};

__device__ __forceinline__ void builder_tile(const SweepP& P, const Tile& T, int U, unsigned char* As, SmemTail& S, uint32_t tmem,
                                             uint32_t ja, uint32_t ct, long long* tr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const DirP& D = P.dir[T.d];
  const LayP& Lp = P.lay[T.d][T.i];
  const float* __restrict__ Hcur = Lp.Hs;
  const long long ldh = P.ldh;
  const int Hq = P.Hq;
  const bool level0 = T.level0 != 0;
  const bool trc = tr != nullptr && tid == 0;
#ifdef DAGNN_TRACE_FINE
  long long t_pre = 0, t_wait = 0, t_build = 0, t_hand = 0, t_comb = 0, t_pref = 0, t0 = 0;
  if (trc) t0 = clock64();
#define TRC_ACC(var) if (trc) { const long long t1 = clock64(); var += t1 - t0; t0 = t1; }
#else
#define TRC_ACC(var)
#endif

  builders_sync();            // previous tile: epilogue reads of rp / alpha / col / bias are done
  for (int t = tid; t <= T.nrows; t += kBuilders) S.rp[t] = level0 ? 0 : D.rowptr[T.p0 + t];
  for (int t = tid; t < T.nrows; t += kBuilders) S.nidx[t] = (T.i == 0) ? D.perm[T.p0 + t] : T.p0 + t;
  {
    const int HP = P.NT * 64, ub = T.ut * U;
    for (int t = tid; t < 5 * U; t += kBuilders) {
      const int g = t / U, u = ub + (t - g * U);
      S.bias[g][t - g * U] = (g < 4) ? __ldg(Lp.bias + g * HP + u) : __ldg(Lp.wk + u);
    }
  }
  builders_sync();
  const int ebase = S.rp[0];
  const int ecount = S.rp[T.nrows] - ebase;
  const bool fits = ecount <= kEdgeCap;

  // ---------------- pre-phase: softmax weights of every in-edge of the tile's rows ----------------
  if (!level0) {
    const bool use_ea = P.use_ea && D.eattr != nullptr;
    const float ca0 = use_ea ? __ldg(Lp.attnc) : 0.f, ca1 = use_ea ? __ldg(Lp.attnc + 1) : 0.f;
    auto score = [&](int e, int& sp) {
      sp = D.col[e];
      float sc = 0.f;
      if (sp < T.pos0) {                               // predecessor state exists (earlier level), else a zero row
        const float* kp = Lp.skp + (size_t)sp * P.nskp;
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
        int j = 0;
        for (; j + 4 <= P.NG; j += 4) {
          const float4 v = ldcg4(kp + j);
          a4.x += v.x; a4.y += v.y; a4.z += v.z; a4.w += v.w;
        }
        for (; j < P.NG; ++j) a4.x += __ldcg(kp + j);
        sc = (a4.x + a4.y) + (a4.z + a4.w);
      }
      if (use_ea) {
        const float2 ea = __ldg(reinterpret_cast<const float2*>(D.eattr) + e);
        sc += ca0 * ea.x + ca1 * ea.y;
      }
      if (P.nvid > 0) sc += __ldg(Lp.vidk + (D.perm[sp] % P.nvid));
      return sc;
    };
    if (fits) {
      // edge-parallel scores (one dependent chain col -> key-score partials for the whole tile), then a per-row softmax
      // over shared memory. A not-yet-computed predecessor keeps its softmax mass and adds a zero row.
      for (int idx = tid; idx < ecount; idx += kBuilders) {
        int sp;
        S.alpha[idx] = score(ebase + idx, sp);
        S.col[idx] = sp;
      }
      builders_sync();
      if (tid < T.nrows) {
        const int i0 = S.rp[tid] - ebase, i1 = S.rp[tid + 1] - ebase;
        float mx = -INFINITY, sum = 0.f;
        for (int k = i0; k < i1; ++k) mx = fmaxf(mx, S.alpha[k]);
        for (int k = i0; k < i1; ++k) sum += expf(S.alpha[k] - mx);
        const float inv = 1.f / (sum + 1e-16f);
        for (int k = i0; k < i1; ++k) S.alpha[k] = (S.col[k] < T.pos0) ? expf(S.alpha[k] - mx) * inv : 0.f;
      }
      // rows with many in-edges: their aggregate m_v = sum_e alpha_e h_e is computed once per tile over the full width
      // and parked in this CTA's scratch; the row then looks like a single in-edge of weight 1 to the operand builders
      // (whose per-item gather is serial over edges). Up to kVeryHeavy in-edges: one warp per row, four predecessor
      // rows x 256 columns in flight; beyond: the whole CTA splits the edge list of the row, partials meet in shared
      // memory (the operand stages are idle during the pre-phase).
      if (tid == 0) S.nvh = 0;
      builders_sync();
      float* scr = P.heavy + (size_t)blockIdx.x * kMaxRows * Hq;
      auto agg_edges = [&](int kb, int k0, int k1, int kstep, float4& acc0, float4& acc1) {
        const bool two = kb + 128 < Hq;
        for (int k = k0; k < k1; k += 4 * kstep) {
          float w[4];
          float4 h0[4], h1[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int kk = k + t * kstep;
            w[t] = (kk < k1) ? S.alpha[kk] : 0.f;
            h0[t] = make_float4(0.f, 0.f, 0.f, 0.f); h1[t] = h0[t];
            if (w[t] != 0.f) {
              const float* hr = Hcur + (size_t)S.col[kk] * ldh + kb;
              h0[t] = ldcg4(hr);
              if (two) h1[t] = ldcg4(hr + 128);
            }
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            acc0.x = fmaf(w[t], h0[t].x, acc0.x); acc0.y = fmaf(w[t], h0[t].y, acc0.y);
            acc0.z = fmaf(w[t], h0[t].z, acc0.z); acc0.w = fmaf(w[t], h0[t].w, acc0.w);
            acc1.x = fmaf(w[t], h1[t].x, acc1.x); acc1.y = fmaf(w[t], h1[t].y, acc1.y);
            acc1.z = fmaf(w[t], h1[t].z, acc1.z); acc1.w = fmaf(w[t], h1[t].w, acc1.w);
          }
        }
      };
      auto mark_row = [&](int r, int i0, int i1, int first_lane_k) {       // one warp: row r now reads scratch row r
        for (int k = i0 + first_lane_k; k < i1; k += 32) {
          S.alpha[k] = (k == i0) ? 1.f : 0.f;
          if (k == i0) S.col[k] = ~r;                    // negative: row r of the scratch
        }
      };
      const bool vh_ok = (size_t)kBuilderWarps * Hq * sizeof(float) <= (size_t)kNAS * kAStageBytes;
      for (int r = warp; r < T.nrows; r += kBuilderWarps) {
        const int i0 = S.rp[r] - ebase, i1 = S.rp[r + 1] - ebase;
        if (i1 - i0 <= kHeavy) continue;                 // warp-uniform
        if (vh_ok && i1 - i0 > kVeryHeavy) {
          int slot = 0;
          if (lane == 0) slot = atomicAdd(&S.nvh, 1);
          slot = __shfl_sync(0xffffffffu, slot, 0);
          if (slot < kMaxVH) {
            if (lane == 0) S.vh[slot] = r;
            continue;
          }
        }
        for (int kb = 4 * lane; kb < Hq; kb += 256) {
          float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
          agg_edges(kb, i0, i1, 1, acc0, acc1);
          *reinterpret_cast<float4*>(scr + (size_t)r * Hq + kb) = acc0;
          if (kb + 128 < Hq) *reinterpret_cast<float4*>(scr + (size_t)r * Hq + kb + 128) = acc1;
        }
        __syncwarp();
        mark_row(r, i0, i1, lane);
      }
      builders_sync();
      const int nvh = min(S.nvh, kMaxVH);
      float* part = reinterpret_cast<float*>(As);        // [kBuilderWarps][Hq] partial aggregates
      for (int v = 0; v < nvh; ++v) {
        const int r = S.vh[v];
        const int i0 = S.rp[r] - ebase, i1 = S.rp[r + 1] - ebase;
        for (int kb = 4 * lane; kb < Hq; kb += 256) {
          float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
          agg_edges(kb, i0 + warp, i1, kBuilderWarps, acc0, acc1);
          *reinterpret_cast<float4*>(part + (size_t)warp * Hq + kb) = acc0;
          if (kb + 128 < Hq) *reinterpret_cast<float4*>(part + (size_t)warp * Hq + kb + 128) = acc1;
        }
        builders_sync();
        for (int kb = 4 * tid; kb < Hq; kb += 4 * kBuilders) {
          float4 a = *reinterpret_cast<const float4*>(part + kb);
#pragma unroll
          for (int w = 1; w < kBuilderWarps; ++w) {
            const float4 b = *reinterpret_cast<const float4*>(part + (size_t)w * Hq + kb);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
          }
          *reinterpret_cast<float4*>(scr + (size_t)r * Hq + kb) = a;
        }
        if (warp == 0) mark_row(r, i0, i1, lane);
        builders_sync();
      }
    } else if (tid < T.nrows) {
      // edge list larger than the cache: row-serial, softmax weights in global scratch. Only FINAL weights are stored —
      // the CTAs of the other unit tiles of these rows write the same values to the same addresses concurrently.
      const int e0 = S.rp[tid], e1 = S.rp[tid + 1];
      float mx = -INFINITY, sum = 0.f;
      int sp;
      for (int e = e0; e < e1; ++e) {
        const float sc = score(e, sp);
        const float mnew = fmaxf(mx, sc);
        sum = sum * expf(mx - mnew) + expf(sc - mnew);
        mx = mnew;
      }
      const float inv = 1.f / (sum + 1e-16f);
      for (int e = e0; e < e1; ++e) {
        const float sc = score(e, sp);
        Lp.alpha[e] = (sp < T.pos0) ? expf(sc - mx) * inv : 0.f;
      }
    }
    builders_sync();
  }
  const float* __restrict__ ga = Lp.alpha;
  const int* __restrict__ gc = D.col;
  auto alpha_of = [&](int e) { return fits ? S.alpha[e - ebase] : __ldcg(ga + e); };
  auto col_of = [&](int e) { return fits ? S.col[e - ebase] : gc[e]; };
  TRC_ACC(t_pre)

  // ---------------- main loop: operand tiles, raw loads one item ahead ----------------
  const int r0 = tid >> 3, c8 = tid & 7;
  const float* inp = (T.i == 0) ? P.X : P.lay[T.d][T.i - 1].Hs;
  const long long ld_inp = (T.i == 0) ? P.ldx : P.ldh;
  const int Din = (T.i == 0) ? P.Din0 : Hq;              // layers > 0: rows are zero-padded to Hq
  const bool vec_in = (T.i > 0) || P.vec_x;
  const int nch = T.nchunks - T.nci;
  const int own = (T.ut * U) >> 6;                       // hidden chunk that covers the tile's own units
  const int nitems = T.nchunks * T.nst;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* scr_ = P.heavy + (size_t)blockIdx.x * kMaxRows * Hq;

  // row groups (of kRStride rows) a sub-tile really has: the tail levels hold a handful of rows per tile
  auto groups_of = [&](int st) { return min(kNR, (min(128, T.nrows - st * 128) + kRStride - 1) / kRStride); };
  auto prefetch = [&](int it, Pre& R) {
    if (it >= nitems) return;
    const int c = it / T.nst, st = it - c * T.nst;
    const int nx = groups_of(st);
#pragma unroll
    for (int x = 0; x < kNR; ++x)
      if (x < nx) {
#pragma unroll
        for (int k = 0; k < 2; ++k) { R.v[x][k][0] = z4; R.v[x][k][1] = z4; R.w[x][k] = 0.f; }
      }
    const int ra = st * 128 + r0;
    if (c < T.nci) {
      const int k0 = c * tc::KC16 + 8 * c8;
      if (k0 >= Din) return;
      const bool two = k0 + 4 < Din;
#pragma unroll
      for (int x = 0; x < kNR; ++x) {
        const int r = ra + kRStride * x;
        if (x >= nx || r >= T.nrows) continue;
        const float* src = inp + (size_t)S.nidx[r] * ld_inp + k0;
        R.w[x][0] = 1.f;
        if (vec_in) {
          R.v[x][0][0] = ldcg4(src);
          if (two) R.v[x][0][1] = ldcg4(src + 4);
        } else {
          float t[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) t[q] = (k0 + q < Din) ? __ldcg(src + q) : 0.f;
          R.v[x][0][0] = make_float4(t[0], t[1], t[2], t[3]);
          R.v[x][0][1] = make_float4(t[4], t[5], t[6], t[7]);
        }
      }
    } else {
      const int k0 = hperm(c - T.nci, nch, own) * tc::KC16 + 8 * c8;
      if (k0 >= Hq) return;
      const bool two = k0 + 4 < Hq;
#pragma unroll
      for (int x = 0; x < kNR; ++x) {
        const int r = ra + kRStride * x;
        if (x >= nx || r >= T.nrows) continue;
        const int e0 = S.rp[r], e1 = S.rp[r + 1];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (e0 + k < e1) {
            const float w = alpha_of(e0 + k);
            R.w[x][k] = w;
            if (w != 0.f) {
              const int cc = col_of(e0 + k);
              const float* hr = (cc >= 0 ? Hcur + (size_t)cc * ldh : scr_ + (size_t)(~cc) * Hq) + k0;
              R.v[x][k][0] = ldcg4(hr);
              if (two) R.v[x][k][1] = ldcg4(hr + 4);
            }
          }
        }
      }
    }
  };

  Pre R;
  prefetch(0, R);
#pragma unroll 1
  for (int it = 0; it < nitems; ++it) {
    const int c = it / T.nst, st = it - c * T.nst;
    const uint32_t j = ja + (uint32_t)it;
    const uint32_t stage = j % kNAS, use = j / kNAS;
    unsigned char* A_hi = As + (size_t)stage * kAStageBytes;
    unsigned char* A_lo = A_hi + 128 * tc::ROW_BYTES;
    // combine the prefetched sources, then put the next item's loads in flight before anything else
    float x[kNR][8];
    const int nx = groups_of(st);
#pragma unroll
    for (int q = 0; q < kNR; ++q) {
      if (q >= nx) continue;
      const float w0 = R.w[q][0], w1 = R.w[q][1];
      const float4 a0 = R.v[q][0][0], a1 = R.v[q][0][1], b0 = R.v[q][1][0], b1 = R.v[q][1][1];
      x[q][0] = fmaf(w1, b0.x, w0 * a0.x); x[q][1] = fmaf(w1, b0.y, w0 * a0.y); x[q][2] = fmaf(w1, b0.z, w0 * a0.z); x[q][3] = fmaf(w1, b0.w, w0 * a0.w);
      x[q][4] = fmaf(w1, b1.x, w0 * a1.x); x[q][5] = fmaf(w1, b1.y, w0 * a1.y); x[q][6] = fmaf(w1, b1.z, w0 * a1.z); x[q][7] = fmaf(w1, b1.w, w0 * a1.w);
    }
    TRC_ACC(t_comb)
    prefetch(it + 1, R);
    TRC_ACC(t_pref)
    const int ra = st * 128 + r0;
    if (c >= T.nci) {                                    // rows with more than two in-edges: the rest, two edges in flight
      const int k0 = hperm(c - T.nci, nch, own) * tc::KC16 + 8 * c8;
      if (k0 < Hq) {
        const bool two = k0 + 4 < Hq;
#pragma unroll
        for (int q = 0; q < kNR; ++q) {
          const int r = ra + kRStride * q;
          if (q >= nx || r >= T.nrows) continue;
          const int e1 = S.rp[r + 1];
          for (int e = S.rp[r] + 2; e < e1; e += 2) {
            float w[2];
            float4 h0[2], h1[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              w[k] = (e + k < e1) ? alpha_of(e + k) : 0.f;
              h0[k] = z4; h1[k] = z4;
              if (w[k] != 0.f) {
                const float* hr = Hcur + (size_t)col_of(e + k) * ldh + k0;
                h0[k] = ldcg4(hr);
                if (two) h1[k] = ldcg4(hr + 4);
              }
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              x[q][0] = fmaf(w[k], h0[k].x, x[q][0]); x[q][1] = fmaf(w[k], h0[k].y, x[q][1]); x[q][2] = fmaf(w[k], h0[k].z, x[q][2]); x[q][3] = fmaf(w[k], h0[k].w, x[q][3]);
              x[q][4] = fmaf(w[k], h1[k].x, x[q][4]); x[q][5] = fmaf(w[k], h1[k].y, x[q][5]); x[q][6] = fmaf(w[k], h1[k].z, x[q][6]); x[q][7] = fmaf(w[k], h1[k].w, x[q][7]);
            }
          }
        }
      }
    }
    TRC_ACC(t_build)
    if (use >= 1) mbar_wait(&S.a_empty[stage], (use - 1) & 1u);     // MMAs that read this stage are done
    TRC_ACC(t_wait)
#pragma unroll
    for (int q = 0; q < kNR; ++q)                                           // rows beyond the tile: D rows nobody reads
      if (q < nx && ra + kRStride * q < T.nrows) tc::store_split8(A_hi, A_lo, r0 + kRStride * q, c8, x[q]);
    TRC_ACC(t_build)
    tc::fence_async_smem();            // generic-proxy stores -> visible to the tensor core (async proxy)
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.a_full[stage]);
    TRC_ACC(t_hand)
  }
  if (trc) tr[1] = clock64();
#ifdef DAGNN_TRACE_FINE
  if (trc) { tr[8] = t_pre; tr[9] = t_wait; tr[10] = t_build; tr[11] = t_hand; tr[12] = t_comb; tr[13] = t_pref; }
#endif

  // ---------------- epilogue ----------------
  mbar_wait(&S.acc_full, ct & 1u);
  tc::fence_after_sync();
  if (trc) tr[2] = clock64();
  {
    const int q = warp & 3, cg = warp >> 2;
    const int ng16 = U >> 4;
    const int ucol0 = (T.ut * U) & 63;                   // first own unit inside its 64-k chunk
#pragma unroll 1
    for (int st = 0; st < T.nst; ++st) {
      const int rr = 32 * q + lane;                      // row inside the sub-tile = TMEM lane
      const int r = st * 128 + rr;
      const bool rok = r < T.nrows;
      const uint32_t tbase = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(st * 4 * U);
      // the sub-tile's last operand stage still holds m_v of the own units (hi + lo)
      const uint32_t jl = ja + (uint32_t)((T.nchunks - 1) * T.nst + st);
      const unsigned char* H_hi = As + (size_t)(jl % kNAS) * kAStageBytes;
      const unsigned char* H_lo = H_hi + 128 * tc::ROW_BYTES;
#pragma unroll 1
      for (int g16 = cg; g16 < ng16; g16 += kBuilderWarps / 4) {
        const int u0 = T.ut * U + g16 * 16;
        if (u0 >= Hq) break;                              // warp-uniform
        float pk = 0.f;
#pragma unroll 1
        for (int h8 = 0; h8 < 2; ++h8) {
          const int cu = g16 * 16 + 8 * h8;               // column inside the tile's unit range
          const int uu = u0 + 8 * h8;
          if (uu >= Hq) break;                            // warp-uniform
          float an[8], ar[8], az[8], ah_[8];
          __syncwarp();
          tc::ld8(tbase + (uint32_t)cu, an);
          tc::ld8(tbase + (uint32_t)(U + cu), ar);
          tc::ld8(tbase + (uint32_t)(2 * U + cu), az);
          if (!level0) tc::ld8(tbase + (uint32_t)(3 * U + cu), ah_);
          float hp[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) hp[t] = 0.f;
          if (!level0) {
            const uint32_t off = tc::tile_off(rr, (ucol0 + cu) >> 3);
            const uint4 hh = *reinterpret_cast<const uint4*>(H_hi + off);
            const uint4 hl = *reinterpret_cast<const uint4*>(H_lo + off);
            const uint32_t* ph = reinterpret_cast<const uint32_t*>(&hh);
            const uint32_t* pl = reinterpret_cast<const uint32_t*>(&hl);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&ph[t]));
              const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&pl[t]));
              hp[2 * t] = fh.x + fl.x;
              hp[2 * t + 1] = fh.y + fl.y;
            }
          }
          tc::wait_ld();
          float o[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int ul = cu + t;
            const float rg = fast_sigmoid(ar[t] + S.bias[0][ul]);
            const float zg = fast_sigmoid(az[t] + S.bias[1][ul]);
            const float ng = fast_tanh(an[t] + S.bias[2][ul] + rg * ((level0 ? 0.f : ah_[t]) + S.bias[3][ul]));
            o[t] = ng + zg * (hp[t] - ng);
            pk = fmaf(o[t], S.bias[4][ul], pk);
          }
          if (rok) {
            float* dst = Lp.Hs + (size_t)(T.p0 + r) * ldh + uu;
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            if (uu + 4 < Hq) *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
          }
        }
        if (rok) Lp.skp[(size_t)(T.p0 + r) * P.nskp + (u0 >> 4)] = pk;
      }
    }
    tc::fence_before_sync();
  }
  if (trc) tr[3] = clock64();
}

// ------------------------------------------------------------------------------------------------------------
// issuer (one thread): weight ring + MMA issue. Ring barriers are tracked per physical barrier (the ring geometry
// changes with U between steps; at a step boundary every stage is drained).
// ------------------------------------------------------------------------------------------------------------
struct RingState {
  uint32_t full_par, empty_par, pending;     // bit s: parity of the next wait on b_full[s] / b_empty[s]; commit outstanding
  uint32_t next;                             // next stage to fill
  uint32_t pre, pre_first;                   // chunks of the upcoming tile already in flight, stage of its chunk 0
  int U;                                     // units per tile the ring is currently laid out for (2 x 48 KB or 8 x 12 KB stages)
};
struct BSrc {                                // where the weight chunks of a tile come from, and the ring geometry for its U
  const unsigned char* img;
  uint32_t bstage;
  int nbs, nci, nch, own;
};
__device__ __forceinline__ BSrc make_bsrc(const SweepP& P, const Tile& T, int U) {
  const LayP& Lp = P.lay[T.d][T.i];
  BSrc b;
  b.nbs = (U == 64) ? 2 : 8;
  b.bstage = 2u * 3u * (uint32_t)U * tc::ROW_BYTES;                         // hi + lo tile of 3U rows
  const int nc_all = ((T.i == 0) ? P.nci0 : P.ncih) + P.ncih;               // chunks per unit block in the image
  b.img = reinterpret_cast<const unsigned char*>(U == 64 ? Lp.img64 : Lp.img16) + (size_t)T.ut * nc_all * b.bstage;
  b.nci = T.nci; b.nch = T.nchunks - T.nci; b.own = (T.ut * U) >> 6;
  return b;
}
// weight chunk at processing position c of a tile -> next ring stage
__device__ __forceinline__ void ring_load(const BSrc& b, int c, unsigned char* Bs, SmemTail& S, RingState& R) {
  if (c >= b.nci) c = b.nci + hperm(c - b.nci, b.nch, b.own);
  const uint32_t s = R.next;
  R.next = (s + 1 == (uint32_t)b.nbs) ? 0u : s + 1;
  if (R.pending >> s & 1u) {                         // MMAs that read this stage must be done
    mbar_wait(&S.b_empty[s], R.empty_par >> s & 1u);
    R.empty_par ^= 1u << s;
    R.pending &= ~(1u << s);
  }
  mbar_expect_tx(&S.b_full[s], b.bstage);
  bulk_g2s(Bs + (size_t)s * b.bstage, b.img + (size_t)c * b.bstage, b.bstage, &S.b_full[s]);
}

// the ring geometry changes with U: every stage must be drained before the region is re-cut
__device__ __forceinline__ void ring_set_geometry(int U, SmemTail& S, RingState& R) {
  if (R.U == U) return;
  for (uint32_t q = 0; q < (uint32_t)kNBBar; ++q)
    if (R.pending >> q & 1u) {
      mbar_wait(&S.b_empty[q], R.empty_par >> q & 1u);
      R.empty_par ^= 1u << q;
      R.pending &= ~(1u << q);
    }
  R.next = 0;
  R.U = U;
}

__device__ __forceinline__ void issuer_tile(const SweepP& P, const Tile& T, int U, unsigned char* As, unsigned char* Bs, SmemTail& S,
                                            uint32_t tmem, uint32_t ja, RingState& R, long long* tr, const Tile* nextT, int nextU) {
  long long i_b = 0, i_a = 0, i_issue = 0, t0 = tr ? clock64() : 0;
  const BSrc b = make_bsrc(P, T, U);
  const int nbs = b.nbs;
  const uint32_t idesc3 = tc::instr_desc_f16(128, 3 * U), idesc2 = tc::instr_desc_f16(128, 2 * U), idesc1 = tc::instr_desc_f16(128, U);

  // chunk c sits in stage (first + c) % nbs; the first chunks may already be in flight (issued while the previous tile
  // was still computing, or before the grid barrier)
  if (R.pre == 0) ring_set_geometry(U, S, R);
  const uint32_t first = R.pre ? R.pre_first : R.next;
  const int npre = min(nbs, T.nchunks);
  for (int c = (int)R.pre; c < npre; ++c) ring_load(b, c, Bs, S, R);
  R.pre = 0;
#pragma unroll 1
  for (int c = 0; c < T.nchunks; ++c) {
    const uint32_t s = (first + (uint32_t)c) % (uint32_t)nbs;
    mbar_wait(&S.b_full[s], R.full_par >> s & 1u);
    R.full_par ^= 1u << s;
    if (tr) { const long long t1 = clock64(); i_b += t1 - t0; t0 = t1; }
    const uint32_t sb = smem_u32(Bs + (size_t)s * b.bstage);
    const uint64_t bh = tc::smem_desc(sb), bl = tc::smem_desc(sb + 3u * (uint32_t)U * tc::ROW_BYTES);
#pragma unroll 1
    for (int st = 0; st < T.nst; ++st) {
      const uint32_t j = ja + (uint32_t)(c * T.nst + st);
      const uint32_t stage = j % kNAS, use = j / kNAS;
      mbar_wait(&S.a_full[stage], use & 1u);
      tc::fence_after_sync();
      if (tr) { const long long t1 = clock64(); i_a += t1 - t0; t0 = t1; }
      const uint32_t sa = smem_u32(As + (size_t)stage * kAStageBytes);
      const uint64_t ah = tc::smem_desc(sa), al = tc::smem_desc(sa + 128 * tc::ROW_BYTES);
      const uint32_t tm = tmem + (uint32_t)(st * 4 * U);
      if (c < T.nci) {                                  // input part: columns [0, 3U) = n_in | r | z
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc::mma3_f16(tm, ah + 2 * ks, al + 2 * ks, bh + 2 * ks, bl + 2 * ks, idesc3, c == 0 && ks == 0);
      } else {                                          // hidden part: columns [U, 4U) = r | z | n_hid
        const uint32_t tmh = tm + (uint32_t)U;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (c == T.nci && ks == 0) {                  // first touch of n_hid: overwrite it, keep accumulating r | z
            const uint64_t nrow = (uint64_t)((2u * (uint32_t)U * tc::ROW_BYTES) >> 4);
            tc::mma_f16(tmh, ah, bh, idesc2, 1u);
            tc::mma_f16(tmh + 2u * (uint32_t)U, ah, bh + nrow, idesc1, 0u);
            tc::mma_f16(tmh, al, bh, idesc3, 1u);
            tc::mma_f16(tmh, ah, bl, idesc3, 1u);
          } else {
            tc::mma3_f16(tmh, ah + 2 * ks, al + 2 * ks, bh + 2 * ks, bl + 2 * ks, idesc3, false);
          }
        }
      }
      tc::commit(&S.a_empty[stage]);
      if (tr) { const long long t1 = clock64(); i_issue += t1 - t0; t0 = t1; }
    }
    tc::commit(&S.b_empty[s]);
    R.pending |= 1u << s;
    // refill: chunk c + nbs - 1 goes where chunk c - 1 was (its MMAs precede the ones just issued)
    if (c >= 1 && c + nbs - 1 < T.nchunks) ring_load(b, c + nbs - 1, Bs, S, R);
  }
  tc::commit(&S.acc_full);
  if (tr) { tr[14] = i_a; tr[15] = i_issue; (void)i_b; }
  // weights are constants: put the first chunks of this CTA's NEXT tile in flight now — they land while the current
  // accumulators drain, the epilogue runs and (for the first tile of the next step) the grid barrier is crossed
  if (nextT) {
    ring_set_geometry(nextU, S, R);
    const BSrc nb = make_bsrc(P, *nextT, nextU);
    R.pre_first = R.next;
    const int n2 = min(nb.nbs, nextT->nchunks);
    for (int c = 0; c < n2; ++c) ring_load(nb, c, Bs, S, R);
    R.pre = (uint32_t)n2;
  }
}

// ------------------------------------------------------------------------------------------------------------
// step tables: the segments (d, i, l = s - i) of a wavefront step, their tiling and this CTA's share of the tiles
// ------------------------------------------------------------------------------------------------------------
struct TileIt { int q, t; };

__device__ __forceinline__ bool tile_advance(const StepTab& tb, int nseg, int rank, int G, TileIt& it) {
  if (it.t >= 0) it.t += G;
  while (it.q < nseg) {
    const Seg g = tb.seg[it.q];
    if (it.t < 0) it.t = ((rank - g.base) % G + G) % G;     // my tiles of a segment: global ids base + t with (base + t) % G == rank
    if (it.t < g.ntile) return true;
    ++it.q;
    it.t = -1;
  }
  return false;
}
__device__ __forceinline__ Tile make_tile(const SweepP& P, const StepTab& tb, int s, const TileIt& it) {
  const Seg g = tb.seg[it.q];
  const int NU = (tb.U == 64) ? P.NT : P.NG;
  const int rows_per = 128 * tb.nst;
  Tile T;
  const int rt = it.t / NU;
  T.d = it.q / P.layers; T.i = it.q - T.d * P.layers;
  T.level0 = (s - T.i == 0); T.pos0 = g.pos0;
  T.ut = it.t - rt * NU;
  T.p0 = g.pos0 + rt * rows_per;
  T.nrows = min(rows_per, g.n - rt * rows_per);
  T.nst = (T.nrows + 127) >> 7;
  T.nci = (T.i == 0) ? P.nci0 : P.ncih;
  T.nchunks = T.nci + (T.level0 ? 0 : P.ncih);
  return T;
}
// all threads; two __syncthreads inside
__device__ __forceinline__ void build_step_table(const SweepP& P, StepTab& tb, int s, int L, int nseg, int G) {
  const int tid = threadIdx.x;
  if (tid < nseg) {
    const int d = tid / P.layers, i = tid - d * P.layers, l = s - i;
    Seg g = {0, 0, 0, 0};
    if (l >= 0 && l < L) {
      g.pos0 = P.dir[d].lvl_off[l];
      g.n = max(0, P.dir[d].lvl_off[l + 1] - g.pos0);
    }
    tb.seg[tid] = g;
  }
  __syncthreads();
  if (tid == 0) {
    int t64 = 0;
    for (int q = 0; q < nseg; ++q) t64 += ceil_div(tb.seg[q].n, 128) * P.NT;
    const int U = (4 * t64 >= G) ? 64 : 16;
    const int nst = (U == 64 && t64 >= 2 * G) ? 2 : 1;
    const int NU = (U == 64) ? P.NT : P.NG;
    int base_ = 0;
    for (int q = 0; q < nseg; ++q) {
      tb.seg[q].ntile = ceil_div(tb.seg[q].n, 128 * nst) * NU;
      tb.seg[q].base = base_;
      base_ += tb.seg[q].ntile;
    }
    tb.U = U; tb.nst = nst;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1) k_sweep(const __grid_constant__ SweepP P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* As = base;
  unsigned char* Bs = base + (size_t)kNAS * kAStageBytes;
  SmemTail& S = *reinterpret_cast<SmemTail*>(Bs + kBRegionBytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kNAS; ++s) { mbar_init(&S.a_full[s], kBuilderWarps); mbar_init(&S.a_empty[s], 1); }
    for (int s = 0; s < kNBBar; ++s) { mbar_init(&S.b_full[s], 1); mbar_init(&S.b_empty[s], 1); }
    mbar_init(&S.acc_full, 1);
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
  uint32_t ja = 0, ct = 0;
  RingState R = {0u, 0u, 0u, 0u, 0u, 0u, 0};
  unsigned int nbar = 0;
  if (nsteps > 0) build_step_table(P, S.tab[0], 0, L, nseg, G);

#pragma unroll 1
  for (int s = 0; s < nsteps; ++s) {
    long long* tr = P.trace ? P.trace + ((size_t)s * 256 + blockIdx.x) * 16 : nullptr;
    if (tr && tid == 0) { tr[0] = clock64(); tr[1] = tr[2] = tr[3] = 0; tr[8] = tr[9] = tr[10] = tr[11] = 0; }
    // the table of step s was built during step s - 1 (level offsets are constants); build the one of step s + 1 now so
    // that the issuer can look across the grid barrier
    const StepTab& tb = S.tab[s & 1];
    StepTab& tbn = S.tab[(s + 1) & 1];
    const bool has_next = s + 1 < nsteps;
    if (has_next) build_step_table(P, tbn, s + 1, L, nseg, G);
    const int U = tb.U;
    int my_tiles = 0;
    TileIt it = {0, -1};
    bool more = tile_advance(tb, nseg, rank, G, it);
#pragma unroll 1
    while (more) {
      const Tile T = make_tile(P, tb, s, it);
      more = tile_advance(tb, nseg, rank, G, it);
      if (warp < kBuilderWarps) {
        builder_tile(P, T, U, As, S, tmem, ja, ct, my_tiles == 0 ? tr : nullptr);
      } else if (lane == 0) {
        Tile N;
        int nU = U;
        bool hn = more;
        if (more) N = make_tile(P, tb, s, it);
        else if (has_next) {
          TileIt it2 = {0, -1};
          hn = tile_advance(tbn, nseg, rank, G, it2);
          if (hn) { N = make_tile(P, tbn, s + 1, it2); nU = tbn.U; }
        }
        issuer_tile(P, T, U, As, Bs, S, tmem, ja, R, my_tiles == 0 ? tr : nullptr, hn ? &N : nullptr, nU);
      }
      __syncwarp();
      ja += (uint32_t)(T.nchunks * T.nst);
      ct += 1;
      ++my_tiles;
    }
    if (tr && tid == 0) { tr[4] = clock64(); tr[6] = my_tiles; tr[7] = U | ((128 * tb.nst) << 8); }
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
constexpr int kMaxGrid = 160;     // CTAs (= SMs) the per-CTA scratch is sized for
static size_t heavy_bytes(int H) { return align256((size_t)kMaxGrid * kMaxRows * round_up(H, 4) * sizeof(float)); }

extern "C" size_t dagnn_sweep_workspace_bytes(int32_t dirs, int32_t layers, int32_t Din, int32_t H, int64_t N, int64_t E) {
  if (dirs < 1 || dirs > DAGNN_MAX_DIRS || layers < 1 || layers > DAGNN_MAX_LAYERS || Din < 1 || H < 1 || N < 0 || E < 0) return 0;
  const size_t nskp = (size_t)round_up(ceil_div(H, 16), 4);
  return 256 + heavy_bytes(H) + (size_t)dirs * layers * (align256((size_t)N * nskp * sizeof(float)) + align256((size_t)E * sizeof(float)));
}
extern "C" size_t dagnn_sweep_trace_bytes(int32_t max_steps) { return (size_t)max_steps * 256 * 16 * sizeof(long long); }

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
  if (A->workspace_bytes < dagnn_sweep_workspace_bytes(dirs, layers, A->Din, H, S->N, S->E))
    return set_err(DAGNN_E_WORKSPACE, "sweep: workspace too small (dagnn_sweep_workspace_bytes)");
  if (H < 1 || H > 4096) return set_err(DAGNN_E_UNSUPPORTED, "sweep: hidden size %d not in [1,4096]", H);
  if (A->nvid < 0) return set_err(DAGNN_E_INVALID, "sweep: nvid");
  DagnnPackLayout lay0, layL;
  if (int rc = dagnn_pack_layout(A->Din, H, A->nvid, &lay0)) return rc;
  if (int rc = dagnn_pack_layout(H, H, A->nvid, &layL)) return rc;
  SweepP P;
  memset(&P, 0, sizeof(P));
  P.dirs = dirs; P.layers = layers; P.H = H; P.Hq = round_up(H, 4); P.nvid = A->nvid;
  P.use_ea = A->use_edge_attr; P.Din0 = A->Din; P.nci0 = lay0.Kin64 / 64; P.ncih = lay0.Kh64 / 64;
  P.NG = lay0.NG; P.NT = lay0.NT; P.nskp = round_up(lay0.NG, 4);
  P.vec_x = ((A->ldx & 3) == 0 && (A->Din & 3) == 0 && ((uintptr_t)A->X & 15) == 0) ? 1 : 0;
  P.ldh = A->ldh; P.ldx = A->ldx; P.X = A->X; P.summary = S->summary; P.bar = static_cast<unsigned int*>(A->workspace);
  P.trace = static_cast<long long*>(A->trace);
  char* ws = static_cast<char*>(A->workspace) + 256;
  P.heavy = reinterpret_cast<float*>(ws);
  ws += heavy_bytes(H);
  const size_t skp_bytes = align256((size_t)S->N * P.nskp * sizeof(float)), alpha_bytes = align256((size_t)S->E * sizeof(float));
  for (int d = 0; d < dirs; ++d) {
    DAGNN_REQUIRE(S->perm[d] && S->rowptr[d] && S->lvl_off[d] && (S->E == 0 || S->col[d]), "sweep: schedule arrays");
    DAGNN_REQUIRE(!A->use_edge_attr || S->E == 0 || S->eattr[d], "sweep: schedule carries no edge attributes");
    P.dir[d].perm = S->perm[d]; P.dir[d].rowptr = S->rowptr[d]; P.dir[d].col = S->col[d];
    P.dir[d].eattr = A->use_edge_attr ? S->eattr[d] : nullptr; P.dir[d].lvl_off = S->lvl_off[d];
    for (int i = 0; i < layers; ++i) {
      DAGNN_REQUIRE(A->Hs[d][i] && ((uintptr_t)A->Hs[d][i] & 15) == 0, "sweep: state buffers must be 16-byte aligned");
      DAGNN_REQUIRE(A->packed[d][i] && ((uintptr_t)A->packed[d][i] & 15) == 0, "sweep: packed params must be 16-byte aligned");
      const DagnnPackLayout& L = i == 0 ? lay0 : layL;
      const float* pk = A->packed[d][i];
      LayP& q = P.lay[d][i];
      q.Hs = A->Hs[d][i]; q.bias = pk + L.bias_off; q.wk = pk + L.wk_off; q.attnc = pk + L.attnc_off; q.vidk = pk + L.vidk_off;
      q.img16 = reinterpret_cast<const __half*>(pk + L.img16_off);
      q.img64 = reinterpret_cast<const __half*>(pk + L.img64_off);
      q.skp = reinterpret_cast<float*>(ws); ws += skp_bytes;
      q.alpha = reinterpret_cast<float*>(ws); ws += alpha_bytes;
    }
  }

  int dev = 0;
  DAGNN_CUDA_OK(cudaGetDevice(&dev));
  static int sm_count[64] = {0};
  if (dev >= 64) return set_err(DAGNN_E_UNSUPPORTED, "sweep: device ordinal %d", dev);
  if (sm_count[dev] == 0) {
    DAGNN_CUDA_OK(cudaFuncSetAttribute(k_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    int n = 0, coop = 0;
    DAGNN_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    DAGNN_CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return set_err(DAGNN_E_UNSUPPORTED, "sweep: device has no cooperative launch");
    sm_count[dev] = n;
  }
  const int G = sm_count[dev] < kMaxGrid ? sm_count[dev] : kMaxGrid;
  DAGNN_CUDA_OK(cudaMemsetAsync(A->workspace, 0, 16, st));
  void* kargs[] = {(void*)&P};
  DAGNN_CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_sweep, dim3(G), dim3(kThreads), kargs, kSmemBytes, st));
  return check_launch("k_sweep");
}
