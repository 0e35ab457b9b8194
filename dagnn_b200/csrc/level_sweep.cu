// The DAGNN level sweep as ONE persistent cooperative kernel (one CTA per SM, grid barrier per wavefront step).
//
// Wavefront step s runs every (direction d, layer i, level l) with l + i == s: (l, i) depends on (l, i-1) [its
// input rows] and on (< l, i) [predecessor states], both finished in earlier steps. The sequential depth is
// L + layers - 1 grid barriers instead of L * layers * dirs kernel chains, and nothing on the host depends on the
// level sizes: level offsets and the level count are read from device memory (no host sync in a forward).
//
// Work unit = tile (d, i, l, BM consecutive positions of the level, one 32-unit slice of H); the tiles of a step
// are dealt round-robin to the CTAs. Per tile:
//   phase 1  gather     : per node, stream its in-edge CSR row, online-softmax the additive-attention scores
//                         (warp-reduced dot with the key vector + edge-type term [+ vertex-id term]) and accumulate
//                         the weighted predecessor rows -> m_v; copy the node's input row. Both land in shared
//                         memory as the A tile [BM, Kin | Kh] — the aggregate never goes back to HBM.
//   phase 2  gate GEMM  : [BM, Kin+Kh] x packed GRU weights [Kin+Kh, 3 x 32]; the weight stream of a slice is one
//                         contiguous run, brought in by cp.async.bulk (UBLKCP) into a multi-stage mbarrier ring that
//                         keeps running across tiles; FP32 FFMA (exact fp32: DESIGN.md §3.3). Thread = one hidden
//                         unit x 3 gates, looping over R rows of the tile: weights are read from shared memory once
//                         per thread as float4 over k, A values are warp-wide broadcasts -> 12 FMA per smem wavefront.
//   phase 3  epilogue   : sigmoid/tanh/blend in registers, 128-bit stores of the new state rows.
// States written in one step are read in later steps by OTHER CTAs: all state reads use ld.global.cg (L2), the
// barrier is the cooperative-groups pattern (bar.sync; fence; atomic; spin on ld.acquire; bar.sync).
#include "common.cuh"
#include "tc.cuh"

namespace dagnn {

constexpr int US = DAGNN_UNIT_SLICE;            // hidden units per slice (32)
constexpr int BK = DAGNN_K_BLOCK;               // K rows per weight stage (16)
constexpr int kThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kWSliceFloats = BK * 3 * US;      // 1536 floats = one k-block of one 32-unit slice
constexpr int kWSliceBytes = kWSliceFloats * 4; // 6144
constexpr int kWStageFloats = 2 * kWSliceFloats; // a ring stage holds up to two slices (64-unit tiles)
constexpr int kBarBytes = 128;                  // 8 ring barriers + tensor-core path: full_b[2], mma_done[2]
constexpr int kTcStageBytes = 2 * 128 * tc::ROW_BYTES + 2 * 192 * tc::ROW_BYTES;   // A hi/lo + B hi/lo = 80 KB
constexpr int kTcBytes = 2 * kTcStageBytes + 1024;                                 // two stages + row-pointer cache
constexpr int kTcCols = 256;                    // TMEM columns: [n_in | r | z | n_hid] x 64 units
constexpr int kMaxSmem = 232448;                // 227 KB opt-in limit per CTA on sm_100

struct DirP {
  const int* perm;      // position -> node id
  const int* rowptr;    // [N+1] CSR rows by position
  const int* col;       // [E] neighbour position
  const float* eattr;   // [E,2] in CSR order or nullptr
  const int* lvl_off;   // [max_levels+1] first position of each level
};
struct LayP {
  float* Hs;            // H[d][i], [N, ldh] position order: predecessor rows read, this level's rows written
  const float* w;       // packed weights [NS][Kin+Kh][3][US]
  const float* bias;    // [4][NS*US]
  const float* wk;      // [NS*US]
  const float* attnc;   // [4]
  const float* vidk;    // [nvid]
  const float* wtc;     // tensor-core image of the weights (pack.cu k_pack_tc)
  float* alpha;         // [E] scratch: softmax weight of every in-edge (CSR order) of the rows a TC tile owns
  float* skp;           // [N][nsk] partial key scores wk . h over 8-unit groups, written with every state row
};
struct SweepP {
  int dirs, layers, H, Hq, Kh, NS, nvid, use_ea;
  int Din0, Kin0;       // layer 0 input width and its padded K; layers > 0 take H / Kh
  int allow64, stages;
  int tc_on, nsk, nskv, nci0, ncih, UT, ubytes, pad2_;   // tensor-core path: on/off, key-score partials per row, 32-k chunks of the layer-0 input /
                        // hidden part, 64-unit tiles, bytes of the union smem region
  int upc, SP, nsmall, ldag;   // weight-stationary path: units per CTA (0 = off), CTAs per (d,i) pair, row threshold, scratch ld
  float* Ag;            // [pairs][nsmall][ldag] gathered A rows of the small segments of the current step
  long long ldh, ldx;
  const float* X;       // [N, ldx] node order (rows through perm)
  const int* summary;   // [0] number of levels of direction 0, [2] schedule status
  unsigned int* bar;    // grid barrier counter (zeroed by the launcher)
  long long* trace;     // optional [steps][grid][8] clock64 stamps (dagnn_sweep_trace_bytes), nullptr = off
  DirP dir[DAGNN_MAX_DIRS];
  LayP lay[DAGNN_MAX_DIRS][DAGNN_MAX_LAYERS];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
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
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ void fma4(float (&acc)[4], float a, const float4& w) {
  acc[0] = fmaf(a, w.x, acc[0]); acc[1] = fmaf(a, w.y, acc[1]); acc[2] = fmaf(a, w.z, acc[2]); acc[3] = fmaf(a, w.w, acc[3]);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// one BK-deep block of the gate GEMM for this thread's unit (3 gates) x R rows. W: this slice's stage
// [4 kq][3 gates][32 units][4 k] floats; A0: first row of the thread's row group at the block's first k column.
// IN: the k rows belong to the input part (n-gate -> acc_n) else to the hidden part (n-gate -> acc_h).
template <int R, bool IN>
__device__ __forceinline__ void mac_block(const float* __restrict__ W, const float* __restrict__ A0, int ldA, int lane,
                                          float (&acc_r)[R], float (&acc_z)[R], float (&acc_n)[R], float (&acc_h)[R]) {
  const float4* W4 = reinterpret_cast<const float4*>(W) + lane;
#pragma unroll
  for (int kq = 0; kq < BK / 4; ++kq) {
    const float4 wr = W4[(kq * 3 + 0) * US];
    const float4 wz = W4[(kq * 3 + 1) * US];
    const float4 wn = W4[(kq * 3 + 2) * US];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(A0 + r * ldA + kq * 4);
      acc_r[r] = fmaf(a.w, wr.w, fmaf(a.z, wr.z, fmaf(a.y, wr.y, fmaf(a.x, wr.x, acc_r[r]))));
      acc_z[r] = fmaf(a.w, wz.w, fmaf(a.z, wz.z, fmaf(a.y, wz.y, fmaf(a.x, wz.x, acc_z[r]))));
      if (IN) acc_n[r] = fmaf(a.w, wn.w, fmaf(a.z, wn.z, fmaf(a.y, wn.y, fmaf(a.x, wn.x, acc_n[r]))));
      else acc_h[r] = fmaf(a.w, wn.w, fmaf(a.z, wn.z, fmaf(a.y, wn.y, fmaf(a.x, wn.x, acc_h[r]))));
    }
  }
}

struct TileCtx {
  uint64_t* bars;
  float* Ws;
  float* As;
  long long* tr; // trace slot of the current (step, CTA) while its first tile runs, else nullptr
  unsigned char* U;  // 1024-aligned union region: tensor-core stages | FFMA ring + A tile
  uint64_t* full_b;  // [2] B chunk landed
  uint64_t* mma_done; // [2] MMAs reading a stage finished
  int* rp;           // [130] row pointers of the current TC tile
  uint32_t tmem;     // TMEM base address (kTcCols columns)
  uint32_t cc;       // cumulative TC chunk count (stage = cc & 1)
  uint32_t it;   // cumulative weight-ring iteration count of this CTA (stage = it % kStages, parity = (it / kStages) & 1)
};

// gather + attention for one node (one warp): writes the A row [inp | m_v]
__device__ __forceinline__ void gather_row(const SweepP& P, const DirP& D, const LayP& Lp, const float* __restrict__ inp,
                                           long long ld_inp, bool inp_via_perm, int Din, int Kin, bool level0, int p, int pos0,
                                           float* __restrict__ arow, const float4 (&wk4)[4], float ca0, float ca1, bool use_ea,
                                           int lane) {
  const int Hq = P.Hq, Kh = P.Kh;
  const float* src = inp + (size_t)(inp_via_perm ? D.perm[p] : p) * ld_inp;
  for (int c = lane; c < Kin; c += 32) arow[c] = (c < Din) ? __ldcg(src + c) : 0.f;
  if (level0) return;
  const int e0 = D.rowptr[p], e1 = D.rowptr[p + 1];
  float mx = -INFINITY, lsum = 0.f;
  float4 acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* Hcur = Lp.Hs;
  const long long ldh = P.ldh;
  for (int eb = e0; eb < e1; eb += 32) {
    // one coalesced load of up to 32 neighbour positions (+ edge attributes), then broadcast per edge
    const int ne = min(32, e1 - eb);
    int my_sp = 0;
    float my_bias = 0.f;
    if (lane < ne) {
      my_sp = D.col[eb + lane];
      if (use_ea) {
        const float2 ea = __ldg(reinterpret_cast<const float2*>(D.eattr) + eb + lane);
        my_bias = ca0 * ea.x + ca1 * ea.y;
      }
      if (P.nvid > 0) my_bias += __ldg(Lp.vidk + (D.perm[my_sp] % P.nvid));
    }
    float4 row[4], nrow[4];
    int sp = __shfl_sync(0xffffffffu, my_sp, 0);
    bool valid = sp < pos0;   // predecessor sits in an earlier level -> its state exists
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * lane + 128 * j;
      row[j] = (valid && c < Hq) ? ldcg4(Hcur + (size_t)sp * ldh + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int q = 0; q < ne; ++q) {
      // prefetch the next edge's row while this one is reduced
      if (q + 1 < ne) {
        const int spn = __shfl_sync(0xffffffffu, my_sp, q + 1);
        const bool vn = spn < pos0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 4 * lane + 128 * j;
          nrow[j] = (vn && c < Hq) ? ldcg4(Hcur + (size_t)spn * ldh + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) dot += dot4(row[j], wk4[j]);
      const float s = warp_sum(dot) + __shfl_sync(0xffffffffu, my_bias, q);
      const float mnew = fmaxf(mx, s);
      const float sc = expf(mx - mnew), pe = expf(s - mnew);
      lsum = lsum * sc + pe;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j].x = acc[j].x * sc + pe * row[j].x; acc[j].y = acc[j].y * sc + pe * row[j].y;
        acc[j].z = acc[j].z * sc + pe * row[j].z; acc[j].w = acc[j].w * sc + pe * row[j].w;
      }
      mx = mnew;
#pragma unroll
      for (int j = 0; j < 4; ++j) row[j] = nrow[j];
    }
  }
  const float inv = (e1 > e0) ? 1.f / (lsum + 1e-16f) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = 4 * lane + 128 * j;
    if (c < Kh) {
      float4 o = make_float4(acc[j].x * inv, acc[j].y * inv, acc[j].z * inv, acc[j].w * inv);
      if (c >= Hq) o = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(arow + Kin + c) = o;
    }
  }
}

// one tile: rows [p0, p0+nvalid) of level `l` (first position pos0) of (d, i), unit slices [sl0, sl0 + NSL)
// R rows per thread, NSL 32-unit slices per tile: BM = R * 8 / NSL rows.
template <int R, int NSL>
__device__ __forceinline__ void process_tile(const SweepP& P, int d, int i, bool level0, int pos0, int p0, int nvalid, int sl0,
                                             TileCtx& T) {
  constexpr int BM = R * 8 / NSL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const DirP& D = P.dir[d];
  const LayP& Lp = P.lay[d][i];
  const int Din = (i == 0) ? P.Din0 : P.H;
  const int Kin = (i == 0) ? P.Kin0 : P.Kh;
  const int Kh = P.Kh, Hq = P.Hq;
  const int ldA = Kin + Kh + 4;
  const int nkb_in = Kin / BK;
  const int nkb = nkb_in + (level0 ? 0 : Kh / BK);
  const int S = P.stages;
  const int nsl = min(NSL, P.NS - sl0);                          // slices of this tile that exist
  const size_t slice_stride = (size_t)(Kin + Kh) * (3 * US);     // floats per slice in the packed stream
  const float* wsrc = Lp.w + (size_t)sl0 * slice_stride;
  float* As = T.As;

  auto issue = [&](int kb) {                                     // tid 0: bring k-block kb of the tile's slices
    const uint32_t st = (T.it + kb) % S;
    const uint32_t bar = smem_u32(&T.bars[st]);
    mbar_expect_tx(bar, (uint32_t)(nsl * kWSliceBytes));
    for (int q = 0; q < nsl; ++q)
      bulk_g2s(smem_u32(T.Ws + st * kWStageFloats + q * kWSliceFloats), wsrc + q * slice_stride + (size_t)kb * kWSliceFloats,
               kWSliceBytes, bar);
  };

  tc::fence_async_smem();
  __syncthreads();   // previous tile: ring drained, A tile and epilogue reads done
  if (tid == 0) {
    const int npre = min(S - 1, nkb);
    for (int kb = 0; kb < npre; ++kb) issue(kb);
  }

  // ---------------- phase 1: gather + attention -> A tile ----------------
  {
    float4 wk4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * lane + 128 * j;
      wk4[j] = (c < Hq && !level0) ? __ldg(reinterpret_cast<const float4*>(Lp.wk + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const bool use_ea = P.use_ea && D.eattr != nullptr;
    const float ca0 = use_ea ? __ldg(Lp.attnc) : 0.f, ca1 = use_ea ? __ldg(Lp.attnc + 1) : 0.f;
    const float* inp = (i == 0) ? P.X : P.lay[d][i - 1].Hs;
    const long long ld_inp = (i == 0) ? P.ldx : P.ldh;
    for (int m = warp; m < BM; m += 8) {
      float* arow = As + m * ldA;
      if (m >= nvalid) {
        for (int c = lane; c < Kin + Kh; c += 32) arow[c] = 0.f;
        continue;
      }
      gather_row(P, D, Lp, inp, ld_inp, i == 0, Din, Kin, level0, p0 + m, pos0, arow, wk4, ca0, ca1, use_ea, lane);
    }
  }
  __syncthreads();

  // ---------------- phase 2: gate GEMM ----------------
  float acc_r[R], acc_z[R], acc_n[R], acc_h[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc_r[r] = acc_z[r] = acc_n[r] = acc_h[r] = 0.f;

  const int wu = warp % NSL;                  // which of the tile's slices this warp owns
  const int row0 = (warp / NSL) * R;          // first row of this warp's row group
  const bool active = wu < nsl;
  const float* Arow0 = As + row0 * ldA;
#pragma unroll 1
  for (int kb = 0; kb < nkb; ++kb) {
    if (tid == 0 && kb + S - 1 < nkb) issue(kb + S - 1);
    const uint32_t itk = T.it + kb;
    const uint32_t st = itk % S;
    mbar_wait(smem_u32(&T.bars[st]), (itk / S) & 1u);
    if (active) {
      const float* W = T.Ws + st * kWStageFloats + wu * kWSliceFloats;
      if (kb < nkb_in) mac_block<R, true>(W, Arow0 + kb * BK, ldA, lane, acc_r, acc_z, acc_n, acc_h);
      else mac_block<R, false>(W, Arow0 + kb * BK, ldA, lane, acc_r, acc_z, acc_n, acc_h);
    }
    __syncthreads();
  }
  T.it += nkb;

  // ---------------- phase 3: GRU pointwise + store (+ key-score partials over 8-unit groups) ----------------
  const int u = (sl0 + wu) * US + lane;
  const bool uok = active && u < Hq;
  const int HP = P.NS * US;
  const float br = uok ? __ldg(Lp.bias + u) : 0.f, bz = uok ? __ldg(Lp.bias + HP + u) : 0.f,
              bi = uok ? __ldg(Lp.bias + 2 * HP + u) : 0.f, bh = uok ? __ldg(Lp.bias + 3 * HP + u) : 0.f;
  const float wku = uok ? __ldg(Lp.wk + u) : 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int m = row0 + r;
    float o = 0.f;
    if (uok && m < nvalid) {
      const float hp = level0 ? 0.f : As[m * ldA + Kin + u];
      const float rg = sigmoidf_(acc_r[r] + br);
      const float zg = sigmoidf_(acc_z[r] + bz);
      const float ng = tanhf(acc_n[r] + bi + rg * (acc_h[r] + bh));
      o = ng + zg * (hp - ng);
      Lp.Hs[(size_t)(p0 + m) * P.ldh + u] = o;
    }
    float pk = o * wku;
    pk += __shfl_xor_sync(0xffffffffu, pk, 1);
    pk += __shfl_xor_sync(0xffffffffu, pk, 2);
    pk += __shfl_xor_sync(0xffffffffu, pk, 4);
    if (active && (lane & 7) == 0 && m < nvalid && (u >> 3) < P.nsk) Lp.skp[(size_t)(p0 + m) * P.nsk + (u >> 3)] = pk;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Weight-stationary path for SMALL segments (n <= nsmall rows): CTA b owns `upc` hidden units of pair q = b / SP for
// the whole sweep and keeps that slice of the GRU weights resident in shared memory ([K/4][3 gates][upc] float4 over
// k). A step then costs no weight traffic at all: phase G spreads the rows of all small segments over every warp of
// the grid (gather + attention -> A rows in an L2-resident scratch), grid barrier, phase M: every owner CTA pulls its
// segment's A rows into shared memory and computes its units for all rows.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_resident(const SweepP& P, int q, int slot, float4* __restrict__ Wres) {
  const int d = q / P.layers, i = q - d * P.layers;
  const int Kin = (i == 0) ? P.Kin0 : P.Kh;
  const int K = Kin + P.Kh, nkb = K / BK;
  const int unit0 = slot * P.upc, sl = unit0 / US, uo = unit0 - sl * US;
  const float4* src = reinterpret_cast<const float4*>(P.lay[d][i].w) + (size_t)sl * nkb * (BK / 4) * 3 * US;
  const int total = (K / 4) * 3 * P.upc;
  for (int idx = threadIdx.x; idx < total; idx += kThreads) {
    const int u = idx % P.upc, g = (idx / P.upc) % 3, kqg = idx / (3 * P.upc);
    Wres[idx] = __ldg(src + ((size_t)kqg * 3 + g) * US + uo + u);     // packed: [kb][kq][g][32 units] float4
  }
}

// phase M for one chunk of <= 32*T rows already staged in As. Thread = (unit lane u8, row slot); UPT units per thread.
template <int T, int UPT>
__device__ __forceinline__ void stationary_chunk(const SweepP& P, const LayP& Lp, const float4* __restrict__ Wres, const float* As,
                                                 int ldA, int Kin, bool level0, int rows, int prow0, int unit0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int u8 = lane & 7, slot = warp * 4 + (lane >> 3);
  const int upc = P.upc;
  const int nkq_in = Kin / 4, nkq = nkq_in + (level0 ? 0 : P.Kh / 4);
  float acc_r[T][UPT], acc_z[T][UPT], acc_n[T][UPT], acc_h[T][UPT];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int j = 0; j < UPT; ++j) acc_r[t][j] = acc_z[t][j] = acc_n[t][j] = acc_h[t][j] = 0.f;
  const float* A0 = As + slot * ldA;
  const bool live = slot < rows;      // rows beyond the chunk: whole quarter-warps idle
  if (live) {
#pragma unroll 4
    for (int kq = 0; kq < nkq; ++kq) {
      float4 a[T];
#pragma unroll
      for (int t = 0; t < T; ++t) a[t] = *reinterpret_cast<const float4*>(A0 + t * 32 * ldA + kq * 4);
      const float4* w = Wres + (size_t)kq * 3 * upc + u8;
      const bool in = kq < nkq_in;
#pragma unroll
      for (int j = 0; j < UPT; ++j) {
        const float4 wr = w[8 * j], wz = w[upc + 8 * j], wn = w[2 * upc + 8 * j];
#pragma unroll
        for (int t = 0; t < T; ++t) {
          acc_r[t][j] = fmaf(a[t].w, wr.w, fmaf(a[t].z, wr.z, fmaf(a[t].y, wr.y, fmaf(a[t].x, wr.x, acc_r[t][j]))));
          acc_z[t][j] = fmaf(a[t].w, wz.w, fmaf(a[t].z, wz.z, fmaf(a[t].y, wz.y, fmaf(a[t].x, wz.x, acc_z[t][j]))));
          const float nn = fmaf(a[t].w, wn.w, fmaf(a[t].z, wn.z, fmaf(a[t].y, wn.y, a[t].x * wn.x)));
          if (in) acc_n[t][j] += nn; else acc_h[t][j] += nn;
        }
      }
    }
  }
  const int HP = P.NS * US;
#pragma unroll
  for (int j = 0; j < UPT; ++j) {
    const int u = unit0 + u8 + 8 * j;
    const bool uok = live && u < P.Hq;
    const float br = uok ? __ldg(Lp.bias + u) : 0.f, bz = uok ? __ldg(Lp.bias + HP + u) : 0.f,
                bi = uok ? __ldg(Lp.bias + 2 * HP + u) : 0.f, bh = uok ? __ldg(Lp.bias + 3 * HP + u) : 0.f;
    const float wku = uok ? __ldg(Lp.wk + u) : 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int m = slot + 32 * t;
      float o = 0.f;
      if (uok && m < rows) {
        const float hp = level0 ? 0.f : As[m * ldA + Kin + u];
        const float rg = sigmoidf_(acc_r[t][j] + br);
        const float zg = sigmoidf_(acc_z[t][j] + bz);
        const float ng = tanhf(acc_n[t][j] + bi + rg * (acc_h[t][j] + bh));
        o = ng + zg * (hp - ng);
        Lp.Hs[(size_t)(prow0 + m) * P.ldh + u] = o;
      }
      float pk = o * wku;               // the 8 lanes of a row slot hold 8 consecutive units: one key-score partial
      pk += __shfl_xor_sync(0xffffffffu, pk, 1);
      pk += __shfl_xor_sync(0xffffffffu, pk, 2);
      pk += __shfl_xor_sync(0xffffffffu, pk, 4);
      const int part = (unit0 + 8 * j) >> 3;
      if (u8 == 0 && live && m < rows && part < P.nsk) Lp.skp[(size_t)(prow0 + m) * P.nsk + part] = pk;
    }
  }
}

template <int UPT>
__device__ __forceinline__ void stationary_segment(const SweepP& P, int q, int slot, bool level0, int pos0, int n,
                                                   const float4* __restrict__ Wres, float* As) {
  const int d = q / P.layers, i = q - d * P.layers;
  const LayP& Lp = P.lay[d][i];
  const int Kin = (i == 0) ? P.Kin0 : P.Kh;
  const int Kuse = Kin + (level0 ? 0 : P.Kh);
  const int ldA = Kin + P.Kh + 4;
  const int chunk = P.allow64 ? 64 : 32;
  const float* Aq = P.Ag + (size_t)q * P.nsmall * P.ldag;
  const int k4n = Kuse / 4;
  for (int c0 = 0; c0 < n; c0 += chunk) {
    const int rows = min(chunk, n - c0);
    __syncthreads();                       // As free (previous chunk / previous tile)
    for (int idx = threadIdx.x; idx < rows * k4n; idx += kThreads) {
      const int r = idx / k4n, c4 = idx - r * k4n;
      *reinterpret_cast<float4*>(As + r * ldA + 4 * c4) = ldcg4(Aq + (size_t)(c0 + r) * P.ldag + 4 * c4);
    }
    __syncthreads();
    if (rows > 32) stationary_chunk<2, UPT>(P, Lp, Wres, As, ldA, Kin, level0, rows, pos0 + c0, slot * P.upc);
    else stationary_chunk<1, UPT>(P, Lp, Wres, As, ldA, Kin, level0, rows, pos0 + c0, slot * P.upc);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Tensor-core path for BIG segments: tile = 128 rows x 64 hidden units, gate GEMM on tcgen05 in 3xTF32 (tc.cuh).
//   pre-phase : one thread per row turns the in-edge scores into softmax weights alpha_e. Scores are scalar gathers:
//               s_e = sum_j skp[nbr][j] (+ edge-type / vertex-id terms) — every producer of a state row also writes the
//               partial key scores wk . h over 8-unit groups (separable attention score).
//   main loop : per 32-wide k chunk, all threads build the A chunk [128, 32] (input rows, or m_v = sum_e alpha_e h_e
//               for that k range) as tf32 hi/lo tiles in the swizzled layout while the previous chunk's MMAs run; the
//               weight chunk arrives as one 48 KB bulk copy of the pre-swizzled hi/lo image; thread 0 issues
//               4 k-steps x 3 products of tcgen05.mma (M=128; N=192 input part, N=128+64 hidden part) into TMEM.
//   epilogue  : tcgen05.ld of [n_in | r | z | n_hid], gates in registers, state row + key-score partials stored.
// Two-stage ring; `cc` counts chunks across tiles so mbarrier parities stay consistent.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void process_tile_tc(const SweepP& P, int d, int i, bool level0, int pos0, int p0, int nvalid, int ut,
                                                TileCtx& T) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const DirP& D = P.dir[d];
  const LayP& Lp = P.lay[d][i];
  const int Din = (i == 0) ? P.Din0 : P.H;
  const int nci = (i == 0) ? P.nci0 : P.ncih;
  const int nch_full = P.ncih;
  const int nchunks = nci + (level0 ? 0 : nch_full);
  const float* wimg = Lp.wtc + (size_t)ut * (nci + nch_full) * 2 * (192 * 32);
  const float* inp = (i == 0) ? P.X : P.lay[d][i - 1].Hs;
  const long long ld_inp = (i == 0) ? P.ldx : P.ldh;
  const bool vec_in = (i > 0) || ((P.ldx & 3) == 0 && (Din & 3) == 0 && ((uintptr_t)P.X & 15) == 0);
  const float* Hcur = Lp.Hs;
  const long long ldh = P.ldh;
  const int Hq = P.Hq;
  int* rp = T.rp;

  auto stage_ptr = [&](uint32_t s) { return T.U + (size_t)s * kTcStageBytes; };
  auto issue_B = [&](uint32_t ccur, int c) {          // thread 0: weight chunk c -> stage ccur & 1
    const uint32_t s = ccur & 1u;
    const uint32_t bar = smem_u32(&T.full_b[s]);
    mbar_expect_tx(bar, 2 * 192 * tc::ROW_BYTES);
    bulk_g2s(smem_u32(stage_ptr(s) + 2 * 128 * tc::ROW_BYTES), wimg + (size_t)c * 2 * (192 * 32), 2 * 192 * tc::ROW_BYTES, bar);
  };

  tc::fence_async_smem();   // generic smem traffic of the previous work item before async-proxy writes to the same bytes
  __syncthreads();          // previous tile: operand stages, row-pointer cache and TMEM reads are done
  if (tid == 0) issue_B(T.cc, 0);
  for (int t = tid; t <= nvalid; t += kThreads) rp[t] = D.rowptr[p0 + t];
  __syncthreads();

  // ---------------- pre-phase: softmax weights of every in-edge of the tile's rows ----------------
  // (only FINAL weights are stored: the CTAs of the other unit tiles of these rows write the same values concurrently)
  if (!level0 && tid < nvalid) {
    const int e0 = rp[tid], e1 = rp[tid + 1];
    const bool use_ea = P.use_ea && D.eattr != nullptr;
    const float ca0 = use_ea ? __ldg(Lp.attnc) : 0.f, ca1 = use_ea ? __ldg(Lp.attnc + 1) : 0.f;
    auto score = [&](int e, bool& valid) {
      const int sp = D.col[e];
      valid = sp < pos0;                             // predecessor state exists (earlier level), else a zero row
      float sc = 0.f;
      if (valid) {
        const float* kp = Lp.skp + (size_t)sp * P.nsk;
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
        int j = 0;
        for (; j + 4 <= P.nskv; j += 4) {
          const float4 v = ldcg4(kp + j);
          a4.x += v.x; a4.y += v.y; a4.z += v.z; a4.w += v.w;
        }
        for (; j < P.nskv; ++j) a4.x += __ldcg(kp + j);
        sc = (a4.x + a4.y) + (a4.z + a4.w);
      }
      if (use_ea) {
        const float2 ea = __ldg(reinterpret_cast<const float2*>(D.eattr) + e);
        sc += ca0 * ea.x + ca1 * ea.y;
      }
      if (P.nvid > 0) sc += __ldg(Lp.vidk + (D.perm[sp] % P.nvid));
      return sc;
    };
    float mx = -INFINITY, sum = 0.f;
    bool valid;
    for (int e = e0; e < e1; ++e) {
      const float sc = score(e, valid);
      const float mnew = fmaxf(mx, sc);
      sum = sum * expf(mx - mnew) + expf(sc - mnew);
      mx = mnew;
    }
    const float inv = 1.f / (sum + 1e-16f);
    for (int e = e0; e < e1; ++e) {
      const float sc = score(e, valid);
      Lp.alpha[e] = valid ? expf(sc - mx) * inv : 0.f;   // a not-yet-computed predecessor keeps its softmax mass, adds a zero row
    }
  }
  __syncthreads();

  // ---------------- main loop over k chunks ----------------
  const uint32_t idesc192 = tc::instr_desc_tf32(128, 192), idesc128 = tc::instr_desc_tf32(128, 128),
                 idesc64 = tc::instr_desc_tf32(128, 64);
#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    const uint32_t ccur = T.cc + (uint32_t)c;
    const uint32_t s = ccur & 1u;
    if (ccur >= 2) mbar_wait(smem_u32(&T.mma_done[s]), ((ccur - 2) >> 1) & 1u);   // MMAs that read this stage are done
    unsigned char* A_hi = stage_ptr(s);
    unsigned char* A_lo = A_hi + 128 * tc::ROW_BYTES;
    if (c < nci) {
      const int k0 = c * tc::KC;
      for (int it = tid; it < 128 * 8; it += kThreads) {
        const int r = it >> 3, c4 = it & 7, k = k0 + 4 * c4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nvalid && k < Din) {
          const float* src = inp + (size_t)(i == 0 ? D.perm[p0 + r] : p0 + r) * ld_inp + k;
          if (vec_in) v = ldcg4(src);               // layers > 0: row padded to Hq >= k + 4 with zeros
          else {
            v.x = __ldcg(src);
            if (k + 1 < Din) v.y = __ldcg(src + 1);
            if (k + 2 < Din) v.z = __ldcg(src + 2);
            if (k + 3 < Din) v.w = __ldcg(src + 3);
          }
        }
        tc::store_split(A_hi, A_lo, r, c4, v);
      }
    } else {
      const int k0 = (c - nci) * tc::KC;
      for (int it = tid; it < 128 * 8; it += kThreads) {
        const int r = it >> 3, c4 = it & 7, k = k0 + 4 * c4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nvalid && k < Hq) {
          const int e1 = rp[r + 1];
          for (int e = rp[r]; e < e1; ++e) {
            const float a = __ldcg(Lp.alpha + e);
            if (a != 0.f) {
              const float4 h = ldcg4(Hcur + (size_t)D.col[e] * ldh + k);
              v.x = fmaf(a, h.x, v.x); v.y = fmaf(a, h.y, v.y); v.z = fmaf(a, h.z, v.z); v.w = fmaf(a, h.w, v.w);
            }
          }
        }
        tc::store_split(A_hi, A_lo, r, c4, v);
      }
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(smem_u32(&T.full_b[s]), (ccur >> 1) & 1u);
      tc::fence_after_sync();
      const uint32_t sa = smem_u32(A_hi);
      const uint64_t ah = tc::smem_desc(sa), al = tc::smem_desc(sa + 128 * tc::ROW_BYTES);
      const uint64_t bh = tc::smem_desc(sa + 2 * 128 * tc::ROW_BYTES), bl = tc::smem_desc(sa + 2 * 128 * tc::ROW_BYTES + 192 * tc::ROW_BYTES);
      if (c < nci) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc::mma3(T.tmem, ah + 2 * ks, al + 2 * ks, bh + 2 * ks, bl + 2 * ks, idesc192, c == 0 && ks == 0);
      } else {
        const uint64_t nrow = (128 * tc::ROW_BYTES) >> 4;    // B rows 128..191 = n gate
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          tc::mma3(T.tmem + 64, ah + 2 * ks, al + 2 * ks, bh + 2 * ks, bl + 2 * ks, idesc128, false);
          tc::mma3(T.tmem + 192, ah + 2 * ks, al + 2 * ks, bh + nrow + 2 * ks, bl + nrow + 2 * ks, idesc64, c == nci && ks == 0);
        }
      }
      tc::commit(&T.mma_done[s]);
      if (c + 1 < nchunks) {                         // weight chunk c+1 -> other stage once chunk c-1's MMAs are done
        if (ccur >= 1) mbar_wait(smem_u32(&T.mma_done[s ^ 1u]), ((ccur - 1) >> 1) & 1u);
        issue_B(ccur + 1, c + 1);
      }
    }
  }
  T.cc += (uint32_t)nchunks;

  // ---------------- epilogue ----------------
  {
    const uint32_t last = T.cc - 1;
    mbar_wait(smem_u32(&T.mma_done[last & 1u]), (last >> 1) & 1u);
    tc::fence_after_sync();
    const int q = warp & 3, hh = warp >> 2;
    const int r = 32 * q + lane;
    const bool rok = r < nvalid;
    const uint32_t tbase = T.tmem + ((uint32_t)(32 * q) << 16);
    const int HP = P.NS * US;
    const int e0 = rok ? rp[r] : 0, e1 = rok ? rp[r + 1] : 0;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const int cu = hh * 32 + 8 * j;                 // column inside the 64-unit tile
      const int u0 = ut * 64 + cu;
      if (u0 >= Hq) break;                            // warp-uniform
      float an[8], ar[8], az[8], ah_[8];
      tc::ld8(tbase + (uint32_t)cu, an);
      tc::ld8(tbase + 64u + (uint32_t)cu, ar);
      tc::ld8(tbase + 128u + (uint32_t)cu, az);
      if (!level0) tc::ld8(tbase + 192u + (uint32_t)cu, ah_);
      tc::wait_ld();
      float hp[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) hp[t] = 0.f;
      if (!level0) {
        for (int e = e0; e < e1; ++e) {
          const float a = __ldcg(Lp.alpha + e);
          if (a != 0.f) {
            const float* hr = Hcur + (size_t)D.col[e] * ldh + u0;
            const float4 h0 = ldcg4(hr);
            float4 h1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u0 + 4 < Hq) h1 = ldcg4(hr + 4);
            hp[0] = fmaf(a, h0.x, hp[0]); hp[1] = fmaf(a, h0.y, hp[1]); hp[2] = fmaf(a, h0.z, hp[2]); hp[3] = fmaf(a, h0.w, hp[3]);
            hp[4] = fmaf(a, h1.x, hp[4]); hp[5] = fmaf(a, h1.y, hp[5]); hp[6] = fmaf(a, h1.z, hp[6]); hp[7] = fmaf(a, h1.w, hp[7]);
          }
        }
      }
      float o[8], pk = 0.f;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int u = u0 + t;
        const float rg = sigmoidf_(ar[t] + __ldg(Lp.bias + u));
        const float zg = sigmoidf_(az[t] + __ldg(Lp.bias + HP + u));
        const float ng = tanhf(an[t] + __ldg(Lp.bias + 2 * HP + u) + rg * ((level0 ? 0.f : ah_[t]) + __ldg(Lp.bias + 3 * HP + u)));
        o[t] = ng + zg * (hp[t] - ng);
        pk = fmaf(o[t], __ldg(Lp.wk + u), pk);
      }
      if (rok) {
        float* dst = Lp.Hs + (size_t)(p0 + r) * ldh + u0;
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        if (u0 + 4 < Hq) *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
        Lp.skp[(size_t)(p0 + r) * P.nsk + (u0 >> 3)] = pk;
      }
    }
    tc::fence_before_sync();
  }
}

__global__ void __launch_bounds__(kThreads, 1) k_sweep_persistent(const __grid_constant__ SweepP P) {
  extern __shared__ __align__(128) unsigned char smem[];
  TileCtx T;
  T.bars = reinterpret_cast<uint64_t*>(smem);
  T.Ws = reinterpret_cast<float*>(smem + kBarBytes);
  T.As = T.Ws + P.stages * kWStageFloats;
  T.it = 0;
  T.tr = nullptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kMaxStages; ++s) mbar_init(smem_u32(&T.bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (P.summary[2] != 0) return;               // schedule build flagged bad input: the host raises (uniform exit)
  const int L = P.summary[0];
  const int nsteps = L + P.layers - 1;
  const int nseg = P.dirs * P.layers;
  const int G = (int)gridDim.x;
  const int ldAmax = max(P.Kin0, P.Kh) + P.Kh + 4;
  float4* Wres = reinterpret_cast<float4*>(T.As + (size_t)(P.allow64 ? 64 : 32) * ldAmax);
  // weight-stationary ownership
  int my_q = -1, my_slot = 0;
  if (P.upc > 0 && (int)blockIdx.x < nseg * P.SP) {
    my_q = (int)blockIdx.x / P.SP;
    my_slot = (int)blockIdx.x - my_q * P.SP;
    load_resident(P, my_q, my_slot, Wres);
  }
  unsigned int nbar = 0;

#pragma unroll 1
  for (int s = 0; s < nsteps; ++s) {
    long long* tr = P.trace ? P.trace + ((size_t)s * 256 + blockIdx.x) * 8 : nullptr;
    if (tr && tid == 0) { tr[0] = clock64(); tr[1] = tr[2] = tr[3] = 0; }
    // ---- classify the segments of this step
    const int NS2 = (P.NS + 1) / 2;
    int tbig = 0, small_rows = 0;
    unsigned small_mask = 0;
    for (int q = 0; q < nseg; ++q) {
      const int d = q / P.layers, i = q - d * P.layers, l = s - i;
      if (l < 0 || l >= L) continue;
      const int n = P.dir[d].lvl_off[l + 1] - P.dir[d].lvl_off[l];
      if (n <= 0) continue;
      if (P.upc > 0 && n <= P.nsmall) { small_mask |= 1u << q; small_rows += n; }
      else tbig += ceil_div(n, 64) * NS2;
    }
    // ---- small segments, phase G: one row per warp, rows dealt over all warps of the grid
    if (small_mask) {
      int rbase = 0;
      for (int q = 0; q < nseg; ++q) {
        if (!(small_mask >> q & 1)) continue;
        const int d = q / P.layers, i = q - d * P.layers, l = s - i;
        const DirP& D = P.dir[d];
        const LayP& Lp = P.lay[d][i];
        const int pos0 = D.lvl_off[l], n = D.lvl_off[l + 1] - pos0;
        const bool level0 = l == 0;
        const int Din = (i == 0) ? P.Din0 : P.H, Kin = (i == 0) ? P.Kin0 : P.Kh;
        const int W8 = G * 8;
        int r = (((int)blockIdx.x + warp * G) - rbase % W8 + W8) % W8;     // warp-global id (CTA-minor) minus offset
        if (r < n) {
          float4 wk4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = 4 * lane + 128 * j;
            wk4[j] = (c < P.Hq && !level0) ? __ldg(reinterpret_cast<const float4*>(Lp.wk + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          const bool use_ea = P.use_ea && D.eattr != nullptr;
          const float ca0 = use_ea ? __ldg(Lp.attnc) : 0.f, ca1 = use_ea ? __ldg(Lp.attnc + 1) : 0.f;
          const float* inp = (i == 0) ? P.X : P.lay[d][i - 1].Hs;
          const long long ld_inp = (i == 0) ? P.ldx : P.ldh;
          for (; r < n; r += W8)
            gather_row(P, D, Lp, inp, ld_inp, i == 0, Din, Kin, level0, pos0 + r, pos0,
                       P.Ag + ((size_t)q * P.nsmall + r) * P.ldag, wk4, ca0, ca1, use_ea, lane);
        }
        rbase += n;
      }
      if (tr && tid == 0) tr[1] = clock64();
      grid_barrier(P.bar, ++nbar * (unsigned int)G);
      if (tr && tid == 0) tr[2] = clock64();
      // ---- phase M: my pair's segment, my units, resident weights
      if (my_q >= 0 && (small_mask >> my_q & 1)) {
        const int d = my_q / P.layers, i = my_q - d * P.layers, l = s - i;
        const int pos0 = P.dir[d].lvl_off[l], n = P.dir[d].lvl_off[l + 1] - pos0;
        if (P.upc == 8) stationary_segment<1>(P, my_q, my_slot, l == 0, pos0, n, Wres, T.As);
        else if (P.upc == 16) stationary_segment<2>(P, my_q, my_slot, l == 0, pos0, n, Wres, T.As);
        else stationary_segment<4>(P, my_q, my_slot, l == 0, pos0, n, Wres, T.As);
      }
      if (tr && tid == 0) tr[3] = clock64();
    }
    // ---- big segments: streamed-weight tiles dealt round-robin
    const bool big = P.allow64 && tbig >= G;
    const int bm = big ? 64 : 32;
    int my_tiles = 0;
    const int nsl_tile = big ? NS2 : P.NS;       // unit-slice tiles per row tile
    int base = 0;
    for (int q = 0; q < nseg; ++q) {
      if (small_mask >> q & 1) continue;
      const int d = q / P.layers, i = q - d * P.layers, l = s - i;
      if (l < 0 || l >= L) continue;
      const int pos0 = P.dir[d].lvl_off[l];
      const int n = P.dir[d].lvl_off[l + 1] - pos0;
      if (n <= 0) continue;
      const int ntile = ceil_div(n, bm) * nsl_tile;
      // my tiles of this segment: global tile ids g = base + t, g % G == (G - 1 - blockIdx.x): the CTAs without a
      // stationary slice take tiles first
      int t = ((G - 1 - (int)blockIdx.x) - base % G + G) % G;
      for (; t < ntile; t += G) {
        const int rt = t / nsl_tile, sl = t - rt * nsl_tile;
        const int p0 = pos0 + rt * bm;
        const int nvalid = min(bm, n - rt * bm);
        if (big) process_tile<16, 2>(P, d, i, l == 0, pos0, p0, nvalid, 2 * sl, T);
        else process_tile<4, 1>(P, d, i, l == 0, pos0, p0, nvalid, sl, T);
        ++my_tiles;
      }
      base += ntile;
    }
    if (tr && tid == 0) { tr[4] = clock64(); tr[6] = my_tiles; tr[7] = (big ? 1 : 0) | (small_mask << 1); }
    if (s + 1 < nsteps) grid_barrier(P.bar, ++nbar * (unsigned int)G);
    if (tr && tid == 0) tr[5] = clock64();
  }
}

static size_t smem_for(int BM, int Kin, int Kh, int stages, int upc = 0) {
  return (size_t)kBarBytes + (size_t)stages * kWStageFloats * 4 + (size_t)BM * (Kin + Kh + 4) * 4 + (size_t)(Kin + Kh) * 3 * upc * 4;
}
constexpr int kNSmall = 256;      // segments with at most this many rows take the weight-stationary path

}  // namespace dagnn

using namespace dagnn;

extern "C" size_t dagnn_sweep_workspace_bytes(int32_t dirs, int32_t layers, int32_t Din, int32_t H) {
  if (dirs < 1 || dirs > DAGNN_MAX_DIRS || layers < 1 || layers > DAGNN_MAX_LAYERS || Din < 1 || H < 1) return 0;
  const int Kh = round_up(H, DAGNN_K_BLOCK), Kin0 = round_up(Din, DAGNN_K_BLOCK);
  const size_t ldag = (size_t)(Kin0 > Kh ? Kin0 : Kh) + Kh;
  return 256 + (size_t)dirs * layers * kNSmall * ldag * sizeof(float);
}
extern "C" size_t dagnn_sweep_trace_bytes(int32_t max_steps) { return (size_t)max_steps * 256 * 8 * sizeof(long long); }

extern "C" int dagnn_sweep_forward_f32(const DagnnSweepArgs* A, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  DAGNN_REQUIRE(A && A->sched, "sweep: null args");
  const DagnnSchedule* S = A->sched;
  const int dirs = S->dirs, layers = A->num_layers, H = A->H;
  DAGNN_REQUIRE(layers >= 1 && layers <= DAGNN_MAX_LAYERS, "sweep: num_layers");
  DAGNN_REQUIRE(A->X && A->ldx >= A->Din && A->Din > 0, "sweep: X");
  DAGNN_REQUIRE(A->ldh % 4 == 0 && A->ldh >= round_up(H, 4), "sweep: ldh must be a multiple of 4 and >= roundup(H,4)");
  DAGNN_REQUIRE(A->workspace && A->workspace_bytes >= dagnn_sweep_workspace_bytes(dirs, layers, A->Din, H) &&
                    ((uintptr_t)A->workspace & 255) == 0,
                "sweep: workspace (dagnn_sweep_workspace_bytes, 256-byte aligned)");
  if (H < 1 || H > 512) return set_err(DAGNN_E_UNSUPPORTED, "sweep: hidden size %d not in [1,512]", H);
  if (A->nvid < 0) return set_err(DAGNN_E_INVALID, "sweep: nvid");
  DagnnPackLayout lay0, layL;
  if (int rc = dagnn_pack_layout(A->Din, H, A->nvid, &lay0)) return rc;
  if (int rc = dagnn_pack_layout(H, H, A->nvid, &layL)) return rc;
  SweepP P;
  memset(&P, 0, sizeof(P));
  P.dirs = dirs; P.layers = layers; P.H = H; P.Hq = round_up(H, 4); P.Kh = lay0.Kh; P.NS = lay0.NS; P.nvid = A->nvid;
  P.use_ea = A->use_edge_attr; P.Din0 = A->Din; P.Kin0 = lay0.Kin;
  P.ldh = A->ldh; P.ldx = A->ldx; P.X = A->X; P.summary = S->summary; P.bar = static_cast<unsigned int*>(A->workspace);
  P.trace = static_cast<long long*>(A->trace);
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
      q.Hs = A->Hs[d][i]; q.w = pk + L.w_off; q.bias = pk + L.bias_off; q.wk = pk + L.wk_off; q.attnc = pk + L.attnc_off;
      q.vidk = pk + L.vidk_off;
    }
  }
  const int Kin_max = layers > 1 ? (lay0.Kin > layL.Kin ? lay0.Kin : layL.Kin) : lay0.Kin;
  if (smem_for(32, Kin_max, P.Kh, 3) > (size_t)kMaxSmem) return set_err(DAGNN_E_UNSUPPORTED, "sweep: Din=%d too wide", A->Din);

  int dev = 0;
  DAGNN_CUDA_OK(cudaGetDevice(&dev));
  static int sm_count[64] = {0};
  if (dev >= 64) return set_err(DAGNN_E_UNSUPPORTED, "sweep: device ordinal %d", dev);
  if (sm_count[dev] == 0) {
    DAGNN_CUDA_OK(cudaFuncSetAttribute(k_sweep_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    int n = 0, coop = 0;
    DAGNN_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    DAGNN_CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return set_err(DAGNN_E_UNSUPPORTED, "sweep: device has no cooperative launch");
    sm_count[dev] = n;
  }
  const int G = sm_count[dev];
  // shared-memory plan: A tile rows (64 or 32), weight ring depth, resident slice of the weight-stationary path
  P.allow64 = smem_for(64, Kin_max, P.Kh, 4) <= (size_t)kMaxSmem;
  P.upc = 0;
  const int pairs = dirs * layers;
  const bool no_stat = getenv("DAGNN_NO_STATIONARY") != nullptr;
  for (int a64 = P.allow64; a64 >= 0 && !P.upc && !no_stat; --a64)
    for (int upc = 8; upc <= 32; upc *= 2)
      if (pairs * ceil_div(H, upc) <= G && smem_for(a64 ? 64 : 32, Kin_max, P.Kh, 3, upc) <= (size_t)kMaxSmem) {
        P.upc = upc; P.SP = ceil_div(H, upc); P.allow64 = a64;
        break;
      }
  P.nsmall = kNSmall;
  P.ldag = Kin_max + P.Kh;
  P.Ag = reinterpret_cast<float*>(static_cast<char*>(A->workspace) + 256);
  P.stages = kMaxStages;
  while (smem_for(P.allow64 ? 64 : 32, Kin_max, P.Kh, P.stages, P.upc) > (size_t)kMaxSmem) --P.stages;
  const size_t smem = smem_for(P.allow64 ? 64 : 32, Kin_max, P.Kh, P.stages, P.upc);

  DAGNN_CUDA_OK(cudaMemsetAsync(A->workspace, 0, 16, st));
  void* kargs[] = {(void*)&P};
  DAGNN_CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_sweep_persistent, dim3(G), dim3(kThreads), kargs, smem, st));
  return check_launch("k_sweep_persistent");
}
