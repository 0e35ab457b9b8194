// The DAGNN level sweep: one fused kernel per wavefront step.
//
// A tile = (direction d, layer i, level l, BM consecutive positions of the level, one 128-unit slice of H).
//   phase 1  gather     : per node, stream its in-edge CSR row, online-softmax the additive-attention scores
//                         (warp-reduced dot with the key vector + edge-type term), accumulate the weighted
//                         predecessor rows -> m_v; copy the node's input row. Both land in shared memory as
//                         the A tile [BM, Kin | Kh] — the aggregate never goes back to HBM.
//   phase 2  gate GEMM  : [BM, Kin+Kh] x packed GRU weights [Kin+Kh, 3 x 128]; the weight stream is one
//                         contiguous run per slice, brought in by cp.async.bulk (UBLKCP) into a 3-stage
//                         mbarrier ring; FP32 FFMA register tiles (exact fp32 — see DESIGN.md for why not
//                         single-pass TF32).
//   phase 3  epilogue   : sigmoid/tanh/blend in registers, coalesced float4 store of the new state rows.
// Wavefront: step s runs every (d, i, l) with l + i == s in ONE grid, so the sequential depth is
// L + layers - 1 launches instead of L * layers * dirs.
#include "common.cuh"

namespace dagnn {

constexpr int BN = DAGNN_UNIT_SLICE;            // hidden units per slice
constexpr int BK = DAGNN_K_BLOCK;               // K rows per weight stage
constexpr int kThreads = 256;
constexpr int kStages = 3;
constexpr int kWStageFloats = BK * 3 * BN;      // 6144
constexpr int kWStageBytes = kWStageFloats * 4; // 24576
constexpr int kMaxSeg = DAGNN_MAX_DIRS * DAGNN_MAX_LAYERS;
constexpr int kBarBytes = 128;
constexpr int kMaxSmem = 232448;                // 227 KB opt-in limit per CTA on sm_100

struct Seg {
  const float* inp;      // X (node order, rows through perm) or H[d][i-1] (position order)
  const int* perm;       // position -> node id, or nullptr when inp is already in position order
  long long ld_inp;
  const float* Hcur;     // H[d][i]: predecessor rows are read here ...
  float* Hout;           // ... and this level's rows are written here (same buffer)
  const int* rowptr;
  const int* col;
  const float* eattr;    // [E,2] in CSR order or nullptr
  const int* perm_vid;   // position -> node id for the vertex-id term, or nullptr
  const float* w;        // packed weights  [NS][Kin+Kh][3][128]
  const float* bias;     // [4][NS*128]
  const float* wk;       // [NS*128]
  const float* attnc;    // [4]
  const float* vidk;     // [nvid]
  int pos0, n_nodes;     // positions [pos0, pos0 + n_nodes) = this level
  int Din, Kin;
  int level0;            // 1: hidden = 0, no aggregation, K = Kin
  int tile_begin;        // first tile (block) index of this segment inside the step's grid
};

struct StepArgs {
  int nseg, H, Hq, Kh, NS, nvid, use_ea, pad_;
  long long ldh;
  Seg seg[kMaxSeg];
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

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ void fma4(float (&acc)[4], float a, const float4& w) {
  acc[0] = fmaf(a, w.x, acc[0]); acc[1] = fmaf(a, w.y, acc[1]); acc[2] = fmaf(a, w.z, acc[2]); acc[3] = fmaf(a, w.w, acc[3]);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// one BK-deep block of the gate GEMM. IN: the k rows belong to the input part (n-gate -> acc_n) else to the
// hidden part (n-gate -> acc_h). A0 points at this thread's first row, column kb*BK.
template <int TM, bool IN>
__device__ __forceinline__ void mac_block(const float* __restrict__ W, const float* __restrict__ A0, int row_stride, int tx,
                                          float (&acc_r)[TM][4], float (&acc_z)[TM][4], float (&acc_n)[TM][4],
                                          float (&acc_h)[TM][4]) {
#pragma unroll
  for (int kk = 0; kk < BK; kk += 4) {
    float4 av[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) av[i] = *reinterpret_cast<const float4*>(A0 + i * row_stride + kk);
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      const float* wrow = W + (kk + k4) * (3 * BN) + tx * 4;
      const float4 wr = *reinterpret_cast<const float4*>(wrow);
      const float4 wz = *reinterpret_cast<const float4*>(wrow + BN);
      const float4 wn = *reinterpret_cast<const float4*>(wrow + 2 * BN);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float a = comp(av[i], k4);
        fma4(acc_r[i], a, wr);
        fma4(acc_z[i], a, wz);
        if (IN) fma4(acc_n[i], a, wn); else fma4(acc_h[i], a, wn);
      }
    }
  }
}

template <int BM>
__global__ void __launch_bounds__(kThreads, 1) k_level_step(const __grid_constant__ StepArgs a) {
  constexpr int TM = BM / 8;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  float* Ws = reinterpret_cast<float*>(smem + kBarBytes);
  float* As = Ws + kStages * kWStageFloats;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  int si = 0;
#pragma unroll 1
  while (si + 1 < a.nseg && (int)blockIdx.x >= a.seg[si + 1].tile_begin) ++si;
  const Seg& S = a.seg[si];
  const int t = (int)blockIdx.x - S.tile_begin;
  const int nt = t / a.NS, sl = t - nt * a.NS;
  const int p0 = S.pos0 + nt * BM;
  const int nvalid = min(BM, S.n_nodes - nt * BM);
  const int Kin = S.Kin, Kh = a.Kh, Hq = a.Hq;
  const int ldA = Kin + Kh + 4;
  const int nkb_in = Kin / BK;
  const int nkb = nkb_in + (S.level0 ? 0 : Kh / BK);
  const float* wsrc = S.w + (size_t)sl * (Kin + Kh) * (3 * BN);
  const long long ldh = a.ldh;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int kb = 0; kb < kStages - 1 && kb < nkb; ++kb) {
      mbar_expect_tx(smem_u32(&bars[kb]), kWStageBytes);
      bulk_g2s(smem_u32(Ws + kb * kWStageFloats), wsrc + (size_t)kb * kWStageFloats, kWStageBytes, smem_u32(&bars[kb]));
    }
  }

  // ---------------- phase 1: gather + attention -> A tile ----------------
  {
    float4 wk4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * lane + 128 * j;
      wk4[j] = (c < Hq) ? __ldg(reinterpret_cast<const float4*>(S.wk + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const bool use_ea = a.use_ea && S.eattr != nullptr;
    const float ca0 = use_ea ? __ldg(S.attnc) : 0.f, ca1 = use_ea ? __ldg(S.attnc + 1) : 0.f;
    for (int m = warp; m < BM; m += 8) {
      float* arow = As + m * ldA;
      if (m >= nvalid) {
        for (int c = lane; c < Kin + Kh; c += 32) arow[c] = 0.f;
        continue;
      }
      const int p = p0 + m;
      const float* src = S.inp + (size_t)(S.perm ? S.perm[p] : p) * S.ld_inp;
      for (int c = lane; c < Kin; c += 32) arow[c] = (c < S.Din) ? __ldcg(src + c) : 0.f;
      if (S.level0) continue;
      const int e0 = S.rowptr[p], e1 = S.rowptr[p + 1];
      float mx = -INFINITY, lsum = 0.f;
      float4 acc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      int sp_next = (e0 < e1) ? S.col[e0] : 0;
      for (int e = e0; e < e1; ++e) {
        const int sp = sp_next;
        if (e + 1 < e1) sp_next = S.col[e + 1];
        const bool valid = sp < S.pos0;   // predecessor sits in an earlier level -> its state exists
        float4 row[4];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 4 * lane + 128 * j;
          row[j] = (valid && c < Hq) ? ldcg4(S.Hcur + (size_t)sp * ldh + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          dot += row[j].x * wk4[j].x + row[j].y * wk4[j].y + row[j].z * wk4[j].z + row[j].w * wk4[j].w;
        }
        float s = warp_sum(dot);
        if (use_ea) {
          const float2 ea = __ldg(reinterpret_cast<const float2*>(S.eattr) + e);
          s += ca0 * ea.x + ca1 * ea.y;
        }
        if (a.nvid > 0) s += __ldg(S.vidk + (S.perm_vid[sp] % a.nvid));
        const float mnew = fmaxf(mx, s);
        const float sc = expf(mx - mnew), pe = expf(s - mnew);
        lsum = lsum * sc + pe;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[j].x = acc[j].x * sc + pe * row[j].x; acc[j].y = acc[j].y * sc + pe * row[j].y;
          acc[j].z = acc[j].z * sc + pe * row[j].z; acc[j].w = acc[j].w * sc + pe * row[j].w;
        }
        mx = mnew;
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
  }
  __syncthreads();

  // ---------------- phase 2: gate GEMM ----------------
  float acc_r[TM][4], acc_z[TM][4], acc_n[TM][4], acc_h[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc_r[i][j] = acc_z[i][j] = acc_n[i][j] = acc_h[i][j] = 0.f;

  const float* Arow0 = As + warp * ldA;
  const int row_stride = 8 * ldA;
#pragma unroll 1
  for (int kb = 0; kb < nkb; ++kb) {
    if (tid == 0) {
      const int nb = kb + kStages - 1;
      if (nb < nkb) {
        const int s = nb % kStages;
        mbar_expect_tx(smem_u32(&bars[s]), kWStageBytes);
        bulk_g2s(smem_u32(Ws + s * kWStageFloats), wsrc + (size_t)nb * kWStageFloats, kWStageBytes, smem_u32(&bars[s]));
      }
    }
    const int st = kb % kStages;
    mbar_wait(smem_u32(&bars[st]), (uint32_t)((kb / kStages) & 1));
    const float* W = Ws + st * kWStageFloats;
    if (kb < nkb_in) mac_block<TM, true>(W, Arow0 + kb * BK, row_stride, lane, acc_r, acc_z, acc_n, acc_h);
    else mac_block<TM, false>(W, Arow0 + kb * BK, row_stride, lane, acc_r, acc_z, acc_n, acc_h);
    __syncthreads();
  }

  // ---------------- phase 3: GRU pointwise + store ----------------
  const int u = sl * BN + lane * 4;
  if (u < Hq) {
    const int HP = a.NS * BN;
    const float4 br = __ldg(reinterpret_cast<const float4*>(S.bias + u));
    const float4 bz = __ldg(reinterpret_cast<const float4*>(S.bias + HP + u));
    const float4 bi = __ldg(reinterpret_cast<const float4*>(S.bias + 2 * HP + u));
    const float4 bh = __ldg(reinterpret_cast<const float4*>(S.bias + 3 * HP + u));
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = warp + 8 * i;
      if (m < nvalid) {
        float4 hp = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!S.level0) hp = *reinterpret_cast<const float4*>(As + m * ldA + Kin + u);
        float o[4];
        const float hpv[4] = {hp.x, hp.y, hp.z, hp.w};
        const float brv[4] = {br.x, br.y, br.z, br.w}, bzv[4] = {bz.x, bz.y, bz.z, bz.w};
        const float biv[4] = {bi.x, bi.y, bi.z, bi.w}, bhv[4] = {bh.x, bh.y, bh.z, bh.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float r = sigmoidf_(acc_r[i][j] + brv[j]);
          const float z = sigmoidf_(acc_z[i][j] + bzv[j]);
          const float n = tanhf(acc_n[i][j] + biv[j] + r * (acc_h[i][j] + bhv[j]));
          o[j] = n + z * (hpv[j] - n);
        }
        *reinterpret_cast<float4*>(S.Hout + (size_t)(p0 + m) * ldh + u) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

static size_t smem_for(int BM, int Kin, int Kh) {
  return (size_t)kBarBytes + (size_t)kStages * kWStageBytes + (size_t)BM * (Kin + Kh + 4) * 4;
}

template <int BM>
static int launch_step(const StepArgs& a, int tiles, size_t smem, cudaStream_t st) {
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {
    DAGNN_CUDA_OK(cudaFuncSetAttribute(k_level_step<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    configured[dev] = true;
  } else if (dev >= 64) {
    DAGNN_CUDA_OK(cudaFuncSetAttribute(k_level_step<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  }
  k_level_step<BM><<<tiles, kThreads, smem, st>>>(a);
  return check_launch("k_level_step");
}

}  // namespace dagnn

using namespace dagnn;

extern "C" int dagnn_sweep_forward_f32(const DagnnSweepArgs* A, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  DAGNN_REQUIRE(A && A->sched, "sweep: null args");
  const DagnnSchedule* S = A->sched;
  const int dirs = S->dirs, layers = A->num_layers, L = A->num_levels, H = A->H;
  DAGNN_REQUIRE(layers >= 1 && layers <= DAGNN_MAX_LAYERS, "sweep: num_layers");
  DAGNN_REQUIRE(L >= 1 && L <= S->max_levels, "sweep: num_levels");
  DAGNN_REQUIRE(A->X && A->ldx >= A->Din && A->Din > 0, "sweep: X");
  DAGNN_REQUIRE(A->ldh % 4 == 0 && A->ldh >= round_up(H, 4), "sweep: ldh must be a multiple of 4 and >= roundup(H,4)");
  if (H < 1 || H > 512) return set_err(DAGNN_E_UNSUPPORTED, "sweep: hidden size %d not in [1,512]", H);
  if (A->nvid < 0) return set_err(DAGNN_E_INVALID, "sweep: nvid");
  for (int d = 0; d < dirs; ++d) {
    DAGNN_REQUIRE(A->lvl_off_host[d], "sweep: lvl_off_host");
    for (int i = 0; i < layers; ++i) {
      DAGNN_REQUIRE(A->Hs[d][i] && ((uintptr_t)A->Hs[d][i] & 15) == 0, "sweep: state buffers must be 16-byte aligned");
      DAGNN_REQUIRE(A->packed[d][i] && ((uintptr_t)A->packed[d][i] & 15) == 0, "sweep: packed params must be 16-byte aligned");
    }
    DAGNN_REQUIRE(!A->use_edge_attr || S->E == 0 || S->eattr[d], "sweep: schedule carries no edge attributes");
  }
  DagnnPackLayout lay[2];
  if (int rc = dagnn_pack_layout(A->Din, H, A->nvid, &lay[0])) return rc;
  if (int rc = dagnn_pack_layout(H, H, A->nvid, &lay[1])) return rc;
  const int Kh = lay[0].Kh, NS = lay[0].NS;
  const int Kin_max = layers > 1 ? (lay[0].Kin > lay[1].Kin ? lay[0].Kin : lay[1].Kin) : lay[0].Kin;
  if (smem_for(16, Kin_max, Kh) > (size_t)kMaxSmem) return set_err(DAGNN_E_UNSUPPORTED, "sweep: Din=%d too wide", A->Din);

  const int nsteps = L + layers - 1;
  for (int s = 0; s < nsteps; ++s) {
    StepArgs a;
    a.nseg = 0; a.H = H; a.Hq = round_up(H, 4); a.Kh = Kh; a.NS = NS; a.nvid = A->nvid; a.use_ea = A->use_edge_attr; a.pad_ = 0;
    a.ldh = A->ldh;
    int max_nodes = 0, kin_step = 0;
    struct Pending { int d, i, l, n; } pend[kMaxSeg];
    int np = 0;
    for (int d = 0; d < dirs; ++d)
      for (int i = 0; i < layers; ++i) {
        const int l = s - i;
        if (l < 0 || l >= L) continue;
        const int n = A->lvl_off_host[d][l + 1] - A->lvl_off_host[d][l];
        if (n <= 0) continue;
        pend[np++] = {d, i, l, n};
        max_nodes = n > max_nodes ? n : max_nodes;
        const int kin = lay[i > 0].Kin;
        kin_step = kin > kin_step ? kin : kin_step;
      }
    if (np == 0) continue;
    int BM = 64;
    if (max_nodes <= 16) BM = 16; else if (max_nodes <= 32) BM = 32;
    while (BM > 16 && smem_for(BM, kin_step, Kh) > (size_t)kMaxSmem) BM >>= 1;
    int tiles = 0;
    for (int q = 0; q < np; ++q) {
      const int d = pend[q].d, i = pend[q].i, l = pend[q].l;
      const DagnnPackLayout& P = lay[i > 0];
      Seg& g = a.seg[a.nseg++];
      g.inp = (i == 0) ? A->X : A->Hs[d][i - 1];
      g.perm = (i == 0) ? S->perm[d] : nullptr;
      g.ld_inp = (i == 0) ? A->ldx : A->ldh;
      g.Hcur = A->Hs[d][i];
      g.Hout = A->Hs[d][i];
      g.rowptr = S->rowptr[d];
      g.col = S->col[d];
      g.eattr = A->use_edge_attr ? S->eattr[d] : nullptr;
      g.perm_vid = A->nvid > 0 ? S->perm[d] : nullptr;
      const float* pk = A->packed[d][i];
      g.w = pk + P.w_off; g.bias = pk + P.bias_off; g.wk = pk + P.wk_off; g.attnc = pk + P.attnc_off; g.vidk = pk + P.vidk_off;
      g.pos0 = A->lvl_off_host[d][l];
      g.n_nodes = pend[q].n;
      g.Din = P.Din; g.Kin = P.Kin;
      g.level0 = (l == 0);
      g.tile_begin = tiles;
      tiles += ceil_div(pend[q].n, BM) * NS;
    }
    const size_t smem = smem_for(BM, kin_step, Kh);
    int rc;
    if (BM == 64) rc = launch_step<64>(a, tiles, smem, st);
    else if (BM == 32) rc = launch_step<32>(a, tiles, smem, st);
    else rc = launch_step<16>(a, tiles, smem, st);
    if (rc) return rc;
  }
  return DAGNN_OK;
}
